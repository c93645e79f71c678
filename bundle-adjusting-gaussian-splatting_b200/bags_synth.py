"""Seeded synthetic cameras / Gaussians / upstream gradients for tests and bench.py.

Everything is generated on the CPU with explicit ``torch.Generator`` seeds so the same scene can be rebuilt on the
GPU box, in the CPU test-suite and by the golden-vector script.  Camera matrices follow the reference's conventions
(scene/cameras.py:108-113, utils/graphics_utils.py:83-107): ``viewmatrix = W2C^T``, ``projmatrix = viewmatrix @ P^T``,
``campos = inverse(viewmatrix)[3, :3]``.
"""
import math
from typing import NamedTuple

import torch

FOVX = 1.2
ZNEAR, ZFAR = 0.01, 100.0

# Named workloads (BASELINE.json configs).  `scale` is the median per-axis Gaussian scale in world units.
CONFIGS = {
    "C1": dict(P=100_000, W=800, H=800),
    "C2": dict(P=3_000_000, W=1297, H=840),
    "C3": dict(P=1_000_000, W=1920, H=1080),
    "C4": dict(P=2_000_000, W=3840, H=2160),
}


class Camera(NamedTuple):
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor   # (4,4)
    projmatrix: torch.Tensor   # (4,4)
    campos: torch.Tensor       # (3,)
    intrinsic: torch.Tensor    # (4,4) (passed through, never read by the kernels)
    c2w: torch.Tensor          # (4,4) float64, column-vector convention


def projection_matrix(znear, zfar, fovx, fovy):
    """utils/graphics_utils.py:83-107 (getProjectionMatrix)."""
    t = math.tan(fovy / 2) * znear
    r = math.tan(fovx / 2) * znear
    P = torch.zeros(4, 4, dtype=torch.float64)
    P[0, 0] = 2.0 * znear / (2 * r)
    P[1, 1] = 2.0 * znear / (2 * t)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def quat_to_rot(q):
    q = q / q.norm()
    r, x, y, z = q.tolist()
    return torch.tensor([
        [1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
        [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
        [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]], dtype=torch.float64)


def make_camera(W, H, seed=100, fovx=FOVX):
    g = torch.Generator().manual_seed(seed)
    fovy = 2 * math.atan(math.tan(fovx / 2) * H / W)
    q = torch.randn(4, generator=g, dtype=torch.float64) * 0.05 + torch.tensor([1.0, 0, 0, 0], dtype=torch.float64)
    Rm = quat_to_rot(q)
    t = torch.randn(3, generator=g, dtype=torch.float64) * 0.1
    W2C = torch.eye(4, dtype=torch.float64)
    W2C[:3, :3] = Rm
    W2C[:3, 3] = t
    view = W2C.t().contiguous()
    proj = view @ projection_matrix(ZNEAR, ZFAR, fovx, fovy).t()
    campos = torch.linalg.inv(view)[3, :3]
    fx = W / (2 * math.tan(fovx / 2))
    fy = H / (2 * math.tan(fovy / 2))
    K = torch.eye(4, dtype=torch.float64)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, W / 2, H / 2
    return Camera(W, H, math.tan(fovx / 2), math.tan(fovy / 2), view.float(), proj.float(), campos.float(),
                  K.float(), torch.linalg.inv(W2C))


class Gaussians(NamedTuple):
    means3D: torch.Tensor    # (P,3)
    scales: torch.Tensor     # (P,3) already exp()
    rotations: torch.Tensor  # (P,4) already normalised
    opacities: torch.Tensor  # (P,1) already sigmoid()
    shs: torch.Tensor        # (P,16,3)


def make_gaussians(P, cam, seed=0, scale=0.005, cull_frac=0.05):
    """Random Gaussians in front of `cam` (about 5% behind the 0.2 near plane, about 25% with off-screen centres)."""
    g = torch.Generator().manual_seed(seed)
    z = torch.rand(P, generator=g, dtype=torch.float64) * (8.0 - 0.3) + 0.3
    behind = torch.rand(P, generator=g, dtype=torch.float64) < cull_frac
    zb = torch.rand(P, generator=g, dtype=torch.float64) * 2.2 - 2.0
    z = torch.where(behind, zb, z)
    xy = (torch.rand(P, 2, generator=g, dtype=torch.float64) * 2 - 1) * 1.15
    x = xy[:, 0] * z * cam.tanfovx
    y = xy[:, 1] * z * cam.tanfovy
    pc = torch.stack([x, y, z, torch.ones_like(z)], dim=1)           # camera space
    pw = (cam.c2w @ pc.t()).t()[:, :3]
    scales = torch.exp(torch.randn(P, 3, generator=g, dtype=torch.float64) * 0.6 + math.log(scale))
    rot = torch.randn(P, 4, generator=g, dtype=torch.float64)
    rot = rot / rot.norm(dim=1, keepdim=True)
    opac = torch.sigmoid(torch.randn(P, 1, generator=g, dtype=torch.float64) * 1.5)
    shs = torch.randn(P, 16, 3, generator=g, dtype=torch.float64)
    shs[:, 1:, :] *= 0.15
    return Gaussians(pw.float().contiguous(), scales.float().contiguous(), rot.float().contiguous(),
                     opac.float().contiguous(), shs.float().contiguous())


def make_upstream(H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    N = H * W
    dc = torch.randn(3, H, W, generator=g) / N
    dd = torch.randn(1, H, W, generator=g) / N
    da = torch.randn(1, H, W, generator=g) / N
    return dc, dd, da


def estimate_tile_instances(gs, cam):
    """Cheap host estimate of (visible count, tile instances) for sizing reports; not used by any kernel."""
    vm = cam.viewmatrix.double()
    p = gs.means3D.double()
    pv = p @ vm[:3, :3] + vm[3, :3]
    zc = pv[:, 2]
    vis = zc > 0.2
    fx = cam.W / (2 * cam.tanfovx)
    fy = cam.H / (2 * cam.tanfovy)
    smax = gs.scales.double().max(dim=1).values
    rad = torch.ceil(3 * torch.sqrt((smax * fx / zc.clamp_min(0.2)) ** 2 + 0.3))
    px = (pv[:, 0] / zc.clamp_min(0.2)) * fx + cam.W / 2
    py = (pv[:, 1] / zc.clamp_min(0.2)) * fy + cam.H / 2
    gx, gy = (cam.W + 15) // 16, (cam.H + 15) // 16
    x0 = ((px - rad) / 16).floor().clamp(0, gx)
    x1 = ((px + rad + 15) / 16).floor().clamp(0, gx)
    y0 = ((py - rad) / 16).floor().clamp(0, gy)
    y1 = ((py + rad + 15) / 16).floor().clamp(0, gy)
    tiles = ((x1 - x0) * (y1 - y0)) * vis
    return int((tiles > 0).sum()), int(tiles.sum())
