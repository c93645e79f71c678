"""TEST INFRASTRUCTURE ONLY -- float64 host restatement of the reference rasterizer (forward and backward).

This file is the CPU oracle of the parity tests and the `cpu_baseline` leg of bench.py.  Nothing in the product
(`bundle-adjusting-gaussian-splatting_b200/`) imports it; the product path has no CPU fallback.

It restates, in PyTorch float64 on the host, the algorithm of the reference CUDA rasterizer
(paths relative to /root/reference/3dgs-pose/cuda_rasterizer):

  preprocess()      forward.cu:339-657  (in_frustum auxiliary.h:139-164, computeCov3D forward.cu:169-203,
                    computeCov2D :99-164, computeColorFromSH :20-96, getRect/ndc2Pix auxiliary.h:41-56)
  bin_tiles()       rasterizer_impl.cu:70-138,287-327  (duplicateWithKeys, stable sort of tile|depth keys,
                    identifyTileRanges)
  blend_forward()   forward.cu:662-798
  blend_backward()  backward.cu:925-1129
  preprocess_backward()  backward.cu:37-170,175-373,377-440,548-922 and the .mean(0)/+1e-4 of rasterize_points.cu:274

including the reference's quirks Q1-Q15.  The backward is the reference's hand-written
formulas, NOT autograd of the forward (the two differ: pupil shift ignored in backward, extra W/2,H/2 factors, ...).

Parity pinning: the reference ships no tests or golden vectors, so this oracle is pinned against outputs of the
reference itself: tests/golden/*.npz are produced on a B200 by oracle/make_golden.py from the unmodified reference
build (oracle/_ref) and tests/test_oracle_golden.py checks this file against them.  Integer quantities decided by
fp32 rounding on the GPU (radii, tile counts, n_contrib) can differ from a float64 evaluation on a handful of
elements; the golden test bounds that fraction, the GPU parity tests compare the CUDA path with the reference build
bit for bit.
"""
import math

import contextlib

import torch

F64 = torch.float64

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]
TILE = 16


# Where the oracle computes.  The CPU is the default (tests without a GPU, bench.py's cpu_baseline); the full-size
# pose-gradient checks run the same float64 code on the GPU (`with on_device("cuda"):`) because 10^6 Gaussians at 1080p
# take minutes on host cores.  Still test infrastructure either way: nothing here is reachable from the product.
DEVICE = "cpu"


@contextlib.contextmanager
def on_device(dev):
    """Runs the oracle on `dev`: inputs are moved there in float64 and every tensor the oracle creates is allocated
    there (torch's device context manager covers the factory calls)."""
    global DEVICE
    prev, DEVICE = DEVICE, str(dev)
    try:
        with torch.device(dev):
            yield
    finally:
        DEVICE = prev


def _d(x):
    return None if x is None else x.detach().to(DEVICE, F64)


def _xform4x3(p, m):
    x, y, z = p.unbind(1)
    return torch.stack([m[0] * x + m[4] * y + m[8] * z + m[12],
                        m[1] * x + m[5] * y + m[9] * z + m[13],
                        m[2] * x + m[6] * y + m[10] * z + m[14]], 1)


def _xform4x4(p, m):
    x, y, z = p.unbind(1)
    return torch.stack([m[0] * x + m[4] * y + m[8] * z + m[12],
                        m[1] * x + m[5] * y + m[9] * z + m[13],
                        m[2] * x + m[6] * y + m[10] * z + m[14],
                        m[3] * x + m[7] * y + m[11] * z + m[15]], 1)


def _rot_from_quat(q):
    """Standard rotation matrix Rq of the (un-normalised, forward.cu:178) quaternion (r,x,y,z); (P,3,3)."""
    r, x, y, z = q.unbind(1)
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], 1),
        torch.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], 1),
        torch.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)], 1)


def cov3d_from_scale_rot(scales, mod, rot):
    """Sigma = Rq S^2 Rq^T, upper triangle (forward.cu:169-203)."""
    Rq = _rot_from_quat(rot)
    S = mod * scales
    M = Rq * S[:, None, :]           # Rq @ diag(S)
    Sig = M @ M.transpose(1, 2)
    return torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], 1)


def _sym3(c6):
    return torch.stack([torch.stack([c6[:, 0], c6[:, 1], c6[:, 2]], 1),
                        torch.stack([c6[:, 1], c6[:, 3], c6[:, 4]], 1),
                        torch.stack([c6[:, 2], c6[:, 4], c6[:, 5]], 1)], 1)


def _proj_rows(t, fx, fy, view):
    """A = Jn @ Rw (rows 0,1): the 2x3 matrix with cov2D = A Sigma A^T (forward.cu:115-156)."""
    P = t.shape[0]
    Rw = torch.stack([torch.stack([view[0], view[4], view[8]]),
                      torch.stack([view[1], view[5], view[9]]),
                      torch.stack([view[2], view[6], view[10]])])      # W2C rotation, standard form
    tz = t[:, 2]
    Jn = torch.zeros(P, 2, 3, dtype=F64)
    Jn[:, 0, 0] = fx / tz
    Jn[:, 0, 2] = -(fx * t[:, 0]) / (tz * tz)
    Jn[:, 1, 1] = fy / tz
    Jn[:, 1, 2] = -(fy * t[:, 1]) / (tz * tz)
    return Jn @ Rw, Rw


def eval_sh(deg, sh, dirs):
    """sh (P,16,3), dirs (P,3) unit; returns (P,3) before the +0.5 (forward.cu:54-87)."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = SH_C0 * sh[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            res = (res + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] +
                   SH_C2[2] * (2.0 * zz - xx - yy) * sh[:, 6] + SH_C2[3] * xz * sh[:, 7] +
                   SH_C2[4] * (xx - yy) * sh[:, 8])
            if deg > 2:
                res = (res + SH_C3[0] * y * (3.0 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10] +
                       SH_C3[2] * y * (4.0 * zz - xx - yy) * sh[:, 11] +
                       SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * sh[:, 12] +
                       SH_C3[4] * x * (4.0 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14] +
                       SH_C3[6] * x * (xx - 3.0 * yy) * sh[:, 15])
    return res


class Scene:
    """Float64 copy of one rasterizer call's inputs."""

    def __init__(self, means3D, opacities, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy, bg,
                 shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None,
                 shift_factors=None, sh_degree=3, scale_modifier=1.0):
        self.means3D = _d(means3D)
        self.opacities = _d(opacities).reshape(-1)
        self.view = _d(viewmatrix).reshape(16)
        self.proj = _d(projmatrix).reshape(16)
        self.campos = _d(campos).reshape(3)
        self.W, self.H = int(W), int(H)
        self.tanfovx, self.tanfovy = float(tanfovx), float(tanfovy)
        # focal lengths are computed in fp32 by the reference (rasterizer_impl.cu:231-232); the fp64 value is within 1 ulp
        self.fx = self.W / (2.0 * self.tanfovx)
        self.fy = self.H / (2.0 * self.tanfovy)
        self.bg = _d(torch.as_tensor(bg, dtype=torch.float32)).reshape(3)
        self.shs = _d(shs) if shs is not None and shs.numel() else None
        self.colors_precomp = _d(colors_precomp) if colors_precomp is not None and colors_precomp.numel() else None
        self.scales = _d(scales) if scales is not None and scales.numel() else None
        self.rotations = _d(rotations) if rotations is not None and rotations.numel() else None
        self.cov3D_precomp = _d(cov3D_precomp) if cov3D_precomp is not None and cov3D_precomp.numel() else None
        self.shift = _d(shift_factors).reshape(3) if shift_factors is not None else torch.zeros(3, dtype=F64)
        self.D = int(sh_degree)
        self.mod = float(scale_modifier)
        self.P = self.means3D.shape[0]
        self.gx = (self.W + TILE - 1) // TILE
        self.gy = (self.H + TILE - 1) // TILE


def _theta_powers(p_hom):
    ab = p_hom[:, :2] / p_hom[:, 3:4]
    theta = torch.atan(torch.sqrt(ab[:, 0] ** 2 + ab[:, 1] ** 2))
    t2 = theta * theta
    t3 = t2 * theta
    t5 = t2 * t3
    t7 = t2 * t5
    return t3, t5, t7


def preprocess(sc):
    """forward.cu:339-657.  Returns a dict of per-Gaussian tensors (length P)."""
    P = sc.P
    p = sc.means3D
    p_view = _xform4x3(p, sc.view)
    alive = p_view[:, 2] > 0.2                                       # Q1: near plane only
    p_hom = _xform4x4(p, sc.proj)
    t3, t5, t7 = _theta_powers(p_hom)
    shift = sc.shift[0] * t3 + sc.shift[1] * t5 + sc.shift[2] * t7
    w_shift = p_hom[:, 3] + shift
    p_w = 1.0 / (w_shift + 0.0000001)                                # Q3
    p_proj = p_hom[:, :3] * p_w[:, None]

    cov3D = sc.cov3D_precomp if sc.cov3D_precomp is not None else cov3d_from_scale_rot(sc.scales, sc.mod, sc.rotations)

    t = p_view.clone()
    t[:, 2] = t[:, 2] + shift                                        # Q2: shift enters cov2D in forward only
    limx, limy = 1.3 * sc.tanfovx, 1.3 * sc.tanfovy
    txtz, tytz = t[:, 0] / t[:, 2], t[:, 1] / t[:, 2]
    t[:, 0] = torch.clamp(txtz, -limx, limx) * t[:, 2]
    t[:, 1] = torch.clamp(tytz, -limy, limy) * t[:, 2]
    A, _ = _proj_rows(t, sc.fx, sc.fy, sc.view)
    cov2 = A @ _sym3(cov3D) @ A.transpose(1, 2)
    cx = cov2[:, 0, 0] + 0.3
    cy = cov2[:, 0, 1]
    cz = cov2[:, 1, 1] + 0.3
    det = cx * cz - cy * cy
    ok = alive & (det != 0)
    det_inv = 1.0 / det
    conic = torch.stack([cz * det_inv, -cy * det_inv, cx * det_inv], 1)
    mid = 0.5 * (cx + cz)
    disc = torch.sqrt(torch.clamp_min(mid * mid - det, 0.1))         # Q5
    lam = torch.maximum(mid + disc, mid - disc)
    radius = torch.ceil(3.0 * torch.sqrt(lam))
    px = ((p_proj[:, 0] + 1.0) * sc.W - 1.0) * 0.5                   # Q6 (already double here)
    py = ((p_proj[:, 1] + 1.0) * sc.H - 1.0) * 0.5
    radius = torch.where(ok, radius, torch.zeros_like(radius))
    radius = torch.nan_to_num(radius, nan=0.0, posinf=0.0, neginf=0.0)
    ri = radius.to(torch.int64).to(F64)

    def tr(v, g):     # (int) cast truncates toward zero, then clamp to the grid (auxiliary.h:46-56)
        return torch.clamp(torch.trunc(torch.nan_to_num(v, nan=0.0)), 0, g).to(torch.int64)
    x0 = tr((px - ri) / TILE, sc.gx)
    y0 = tr((py - ri) / TILE, sc.gy)
    x1 = tr((px + ri + TILE - 1) / TILE, sc.gx)
    y1 = tr((py + ri + TILE - 1) / TILE, sc.gy)
    area = (x1 - x0) * (y1 - y0)
    visible = ok & (area > 0)
    area = torch.where(visible, area, torch.zeros_like(area))
    radii = torch.where(visible, ri.to(torch.int64), torch.zeros(P, dtype=torch.int64))

    if sc.colors_precomp is None:
        # forward.cu:26-39: direction from the pupil-shifted camera-space point mapped back to world space
        vm = sc.view.reshape(4, 4)                 # vm[c][r] = W2C[r][c]
        W2C = vm.t()
        C2W = torch.linalg.inv(W2C)
        pos = p + shift[:, None] * C2W[:3, 2][None, :]
        dirv = pos - sc.campos[None, :]
        dirv = dirv / dirv.norm(dim=1, keepdim=True)
        res = eval_sh(sc.D, sc.shs, dirv) + 0.5
        clamped = res < 0
        rgb = torch.clamp_min(res, 0.0)
    else:
        rgb = sc.colors_precomp
        clamped = torch.zeros(P, 3, dtype=torch.bool)
    z = torch.zeros(P, dtype=F64)
    return dict(visible=visible, radii=radii, tiles_touched=area, rect=(x0, y0, x1, y1),
                depth=p_view[:, 2], xy=torch.stack([px, py], 1), conic=conic, opacity=sc.opacities, rgb=rgb,
                clamped=clamped, cov3D=cov3D,
                mean2Dx=torch.where(visible, p_hom[:, 2], z),        # Q7
                mean2Dy=torch.where(visible, w_shift, z))


def bin_tiles(sc, pre):
    """duplicateWithKeys + stable sort on (tile, fp32 depth bits) + identifyTileRanges.  Returns point_list (R),
    keys (R, int64: tile<<32 | depth bits), ranges (tiles,2)."""
    vis = torch.nonzero(pre["visible"]).reshape(-1)
    x0, y0, x1, y1 = [r[vis] for r in pre["rect"]]
    w = x1 - x0
    cnt = w * (y1 - y0)
    R = int(cnt.sum())
    ids = torch.repeat_interleave(vis, cnt)
    start = torch.cumsum(cnt, 0) - cnt
    local = torch.arange(R, dtype=torch.int64) - torch.repeat_interleave(start, cnt)
    ww = torch.repeat_interleave(w, cnt)
    ty = torch.repeat_interleave(y0, cnt) + local // torch.clamp_min(ww, 1)
    tx = torch.repeat_interleave(x0, cnt) + local % torch.clamp_min(ww, 1)
    tile = ty * sc.gx + tx
    dbits = pre["depth"].to(torch.float32).view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = (tile << 32) | dbits[ids]
    order = torch.sort(keys, stable=True).indices        # emission order is (id, tile): ties keep ascending id
    keys = keys[order]
    point_list = ids[order]
    ntiles = sc.gx * sc.gy
    tile_sorted = keys >> 32
    counts = torch.bincount(tile_sorted, minlength=ntiles)
    ends = torch.cumsum(counts, 0)
    starts = ends - counts
    ranges = torch.stack([starts, ends], 1)
    ranges[counts == 0] = 0                               # untouched tiles keep the memset value
    return point_list, keys, ranges


def _tile_pixels(sc, t):
    ty, tx = divmod(t, sc.gx)
    ys = torch.arange(ty * TILE, min((ty + 1) * TILE, sc.H))
    xs = torch.arange(tx * TILE, min((tx + 1) * TILE, sc.W))
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    return xx.reshape(-1), yy.reshape(-1)


def _tile_alpha(sc, pre, ids, xx, yy):
    xy = pre["xy"][ids]
    con = pre["conic"][ids]
    o = pre["opacity"][ids]
    dx = xy[None, :, 0] - xx[:, None].to(F64)
    dy = xy[None, :, 1] - yy[:, None].to(F64)
    power = -0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) - con[None, :, 1] * dx * dy
    G = torch.exp(power)
    alpha = torch.clamp_max(o[None, :] * G, 0.99)
    valid = (power <= 0) & (alpha >= 1.0 / 255.0)
    a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
    T_incl = torch.cumprod(1 - a_eff, 1)
    term = valid & (T_incl < 0.0001)
    L = ids.shape[0]
    idx = torch.arange(L)
    has = term.any(1)
    first = torch.argmax(term.to(torch.int8), 1)
    stop = torch.where(has, first, torch.full_like(first, L))
    live = valid & (idx[None, :] < stop[:, None])
    T_before = torch.cat([torch.ones(T_incl.shape[0], 1, dtype=F64), T_incl[:, :-1]], 1)
    return dx, dy, G, alpha, live, T_before, T_incl, stop


def blend_forward(sc, pre, point_list, ranges):
    """forward.cu:662-798.  Returns color (3,H,W), depth (H,W), alpha (H,W), final_T (H,W), n_contrib (H,W)."""
    H, W = sc.H, sc.W
    color = torch.zeros(3, H, W, dtype=F64)
    depth = torch.zeros(H, W, dtype=F64)
    alpha_img = torch.zeros(H, W, dtype=F64)
    final_T = torch.ones(H, W, dtype=F64)
    n_contrib = torch.zeros(H, W, dtype=torch.int64)
    ranges_host = ranges.tolist()
    for t in range(sc.gx * sc.gy):
        xx, yy = _tile_pixels(sc, t)
        s, e = ranges_host[t]
        if e > s:
            ids = point_list[s:e]
            dx, dy, G, alpha, live, T_before, T_incl, stop = _tile_alpha(sc, pre, ids, xx, yy)
            w = torch.where(live, alpha * T_before, torch.zeros_like(alpha))
            C = w @ pre["rgb"][ids]
            Dp = w @ pre["depth"][ids]
            A = w.sum(1)
            L = ids.shape[0]
            idx = torch.arange(1, L + 1)
            last = (live * idx[None, :]).max(1).values
            Tf = torch.where(stop > 0, T_incl.gather(1, (stop - 1).clamp_min(0)[:, None])[:, 0],
                             torch.ones_like(A))
        else:
            n = xx.shape[0]
            C = torch.zeros(n, 3, dtype=F64)
            Dp = torch.zeros(n, dtype=F64)
            A = torch.zeros(n, dtype=F64)
            last = torch.zeros(n, dtype=torch.int64)
            Tf = torch.ones(n, dtype=F64)
        color[:, yy, xx] = (C + Tf[:, None] * sc.bg[None, :]).t()
        depth[yy, xx] = Dp                                            # Q8: un-normalised, no background
        alpha_img[yy, xx] = A
        final_T[yy, xx] = Tf
        n_contrib[yy, xx] = last
    return color, depth, alpha_img, final_T, n_contrib


def blend_backward(sc, pre, point_list, ranges, dL_dcolor, dL_ddepth, dL_dalpha):
    """backward.cu:925-1129 in closed form.  Returns per-Gaussian accumulators (P,...)."""
    P = sc.P
    g_color = _d(dL_dcolor).reshape(3, sc.H, sc.W)
    g_depth = _d(dL_ddepth).reshape(sc.H, sc.W)
    g_alpha = _d(dL_dalpha).reshape(sc.H, sc.W)
    acc = dict(dL_dcolors=torch.zeros(P, 3, dtype=F64), dL_ddepths=torch.zeros(P, dtype=F64),
               dL_dmean2D=torch.zeros(P, 2, dtype=F64), dL_dmean2D_abs=torch.zeros(P, 2, dtype=F64),
               dL_dconic=torch.zeros(P, 3, dtype=F64), covariance=torch.zeros(P, dtype=F64),
               dL_dopacity=torch.zeros(P, dtype=F64))
    ddelx_dx, ddely_dy = 0.5 * sc.W, 0.5 * sc.H
    ranges_host = ranges.tolist()
    for t in range(sc.gx * sc.gy):
        s, e = ranges_host[t]
        if e <= s:
            continue
        xx, yy = _tile_pixels(sc, t)
        ids = point_list[s:e]
        dx, dy, G, alpha, live, T_before, T_incl, stop = _tile_alpha(sc, pre, ids, xx, yy)
        gc = g_color[:, yy, xx].t()            # (n,3)
        gd = g_depth[yy, xx]
        ga = g_alpha[yy, xx]
        z = torch.zeros_like(alpha)
        w = torch.where(live, alpha * T_before, z)                    # alpha_j T_j
        T_after = T_before * (1 - torch.where(live, alpha, z))
        Tf = torch.where(stop > 0, T_incl.gather(1, (stop - 1).clamp_min(0)[:, None])[:, 0], torch.ones_like(gd))
        col = pre["rgb"][ids]
        dep = pre["depth"][ids]
        con = pre["conic"][ids]
        o = pre["opacity"][ids]

        def suffix_excl(v):        # sum over k>j of v_k
            return torch.flip(torch.cumsum(torch.flip(v, [1]), 1), [1]) - v
        # accum_rec of the reference = colour composited behind j, normalised by the transmittance after j
        dLa = torch.zeros_like(alpha)
        for ch in range(3):
            A_ch = suffix_excl(w * col[None, :, ch]) / T_after
            dLa = dLa + (col[None, :, ch] - A_ch) * gc[:, ch:ch + 1]
        A_d = suffix_excl(w * dep[None, :]) / T_after
        dLa = dLa + (dep[None, :] - A_d) * gd[:, None]
        A_a = suffix_excl(w) / T_after
        dLa = dLa + (1 - A_a) * ga[:, None]
        dLa = dLa * T_before
        bg_dot = (sc.bg[None, :] * gc).sum(1)
        dLa = dLa + (-Tf[:, None] / (1 - alpha)) * bg_dot[:, None]
        dLa = torch.where(live, dLa, z)
        dL_dG = o[None, :] * dLa
        gdx, gdy = G * dx, G * dy
        dG_ddelx = -gdx * con[None, :, 0] - gdy * con[None, :, 1]
        dG_ddely = -gdy * con[None, :, 2] - gdx * con[None, :, 1]
        mx = dL_dG * dG_ddelx * ddelx_dx
        my = dL_dG * dG_ddely * ddely_dy
        acc["dL_dcolors"].index_add_(0, ids, w.t() @ gc)
        acc["dL_ddepths"].index_add_(0, ids, w.t() @ gd)
        acc["dL_dmean2D"].index_add_(0, ids, torch.stack([mx.sum(0), my.sum(0)], 1))
        acc["dL_dmean2D_abs"].index_add_(0, ids, torch.stack([mx.abs().sum(0), my.abs().sum(0)], 1))
        acc["dL_dconic"].index_add_(0, ids, torch.stack([(-0.5 * gdx * dx * dL_dG).sum(0),
                                                        (-0.5 * gdx * dy * dL_dG).sum(0),
                                                        (-0.5 * gdy * dy * dL_dG).sum(0)], 1))
        acc["covariance"].index_add_(0, ids, live.sum(0).to(F64) * (con[:, 0] + con[:, 2]))
        acc["dL_dopacity"].index_add_(0, ids, (G * dLa).sum(0))
    return acc


def preprocess_backward(sc, pre, acc):
    """K8 + K9 + .mean(0) reductions.  Returns the 14 outputs of RasterizeGaussiansBackwardCUDA (as a dict) plus
    `viewmatrix_rows`: the per-visible-Gaussian candidates of the racy dL_dviewmatrix (Q13)."""
    P = sc.P
    vis = pre["visible"]
    vi = torch.nonzero(vis).reshape(-1)
    m = sc.means3D[vi]
    view, proj = sc.view, sc.proj
    cov3D = pre["cov3D"][vi]
    dconic = acc["dL_dconic"][vi]
    ddepth = acc["dL_ddepths"][vi]
    dm2 = acc["dL_dmean2D"][vi]
    fx, fy = sc.fx, sc.fy

    # ---- K8: computeCov2DCUDA (no pupil shift, Q2) ----
    t = _xform4x3(m, view)
    limx, limy = 1.3 * sc.tanfovx, 1.3 * sc.tanfovy
    txtz, tytz = t[:, 0] / t[:, 2], t[:, 1] / t[:, 2]
    t = t.clone()
    t[:, 0] = torch.clamp(txtz, -limx, limx) * t[:, 2]
    t[:, 1] = torch.clamp(tytz, -limy, limy) * t[:, 2]
    xg = ((txtz >= -limx) & (txtz <= limx)).to(F64)
    yg = ((tytz >= -limy) & (tytz <= limy)).to(F64)
    A, Rw = _proj_rows(t, fx, fy, view)           # A[:,0,:] = T[0][*], A[:,1,:] = T[1][*] of the reference
    V = _sym3(cov3D)
    cov2 = A @ V @ A.transpose(1, 2)
    a = cov2[:, 0, 0] + 0.3
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + 0.3
    denom = a * c - b * b
    d2i = 1.0 / (denom * denom + 0.0000001)
    nz = (d2i != 0).to(F64)
    dL_da = nz * d2i * (-c * c * dconic[:, 0] + 2 * b * c * dconic[:, 1] + (denom - a * c) * dconic[:, 2])
    dL_dc = nz * d2i * (-a * a * dconic[:, 2] + 2 * a * b * dconic[:, 1] + (denom - a * c) * dconic[:, 0])
    dL_db = nz * d2i * 2 * (b * c * dconic[:, 0] - (denom + 2 * b * b) * dconic[:, 1] + a * b * dconic[:, 2])
    T0, T1 = A[:, 0, :], A[:, 1, :]
    dcov = torch.zeros(vi.shape[0], 6, dtype=F64)
    dcov[:, 0] = T0[:, 0] * T0[:, 0] * dL_da + T0[:, 0] * T1[:, 0] * dL_db + T1[:, 0] * T1[:, 0] * dL_dc
    dcov[:, 3] = T0[:, 1] * T0[:, 1] * dL_da + T0[:, 1] * T1[:, 1] * dL_db + T1[:, 1] * T1[:, 1] * dL_dc
    dcov[:, 5] = T0[:, 2] * T0[:, 2] * dL_da + T0[:, 2] * T1[:, 2] * dL_db + T1[:, 2] * T1[:, 2] * dL_dc
    dcov[:, 1] = 2 * T0[:, 0] * T0[:, 1] * dL_da + (T0[:, 0] * T1[:, 1] + T0[:, 1] * T1[:, 0]) * dL_db + 2 * T1[:, 0] * T1[:, 1] * dL_dc
    dcov[:, 2] = 2 * T0[:, 0] * T0[:, 2] * dL_da + (T0[:, 0] * T1[:, 2] + T0[:, 2] * T1[:, 0]) * dL_db + 2 * T1[:, 0] * T1[:, 2] * dL_dc
    dcov[:, 4] = 2 * T0[:, 2] * T0[:, 1] * dL_da + (T0[:, 1] * T1[:, 2] + T0[:, 2] * T1[:, 1]) * dL_db + 2 * T1[:, 1] * T1[:, 2] * dL_dc
    dcov = dcov * nz[:, None]
    VT0 = (V @ T0[:, :, None])[:, :, 0]           # Vrk[k][*] . T[0][*]
    VT1 = (V @ T1[:, :, None])[:, :, 0]
    dT0 = 2 * VT0 * dL_da[:, None] + VT1 * dL_db[:, None]      # dL_dT00, dL_dT01, dL_dT02
    dT1 = 2 * VT1 * dL_dc[:, None] + VT0 * dL_db[:, None]      # dL_dT10, dL_dT11, dL_dT12
    # W[i][j] of the reference (GLM column i, row j) = view[i + 4j]
    Wg = lambda i, j: view[i + 4 * j]
    dJ00 = Wg(0, 0) * dT0[:, 0] + Wg(0, 1) * dT0[:, 1] + Wg(0, 2) * dT0[:, 2]
    dJ02 = Wg(2, 0) * dT0[:, 0] + Wg(2, 1) * dT0[:, 1] + Wg(2, 2) * dT0[:, 2]
    dJ11 = Wg(1, 0) * dT1[:, 0] + Wg(1, 1) * dT1[:, 1] + Wg(1, 2) * dT1[:, 2]
    dJ12 = Wg(2, 0) * dT1[:, 0] + Wg(2, 1) * dT1[:, 1] + Wg(2, 2) * dT1[:, 2]
    tz = 1.0 / t[:, 2]
    tz2 = tz * tz
    tz3 = tz2 * tz
    dtx = xg * -fx * tz2 * dJ02
    dty = yg * -fy * tz2 * dJ12
    dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * t[:, 0]) * tz3 * dJ02 + (2 * fy * t[:, 1]) * tz3 * dJ12

    p_hom = _xform4x4(m, proj)
    t3, t5, t7 = _theta_powers(p_hom)
    dshift = torch.stack([dtz * t3, dtz * t5, dtz * t7], 1)

    # racy view-matrix terms (backward.cu:351-363): one candidate row per visible Gaussian
    vrow = torch.zeros(vi.shape[0], 16, dtype=F64)
    v = view
    s0 = v[0] * dT0[:, 0] + v[1] * dT0[:, 1]
    s0f = s0 + v[2] * dT0[:, 2]
    s1 = v[4] * dT1[:, 0] + v[5] * dT1[:, 1]
    s1f = s1 + v[6] * dT1[:, 2]
    vrow[:, 0] = fx * tz * dT0[:, 0]
    vrow[:, 1] = fx * tz * dT0[:, 1]
    vrow[:, 4] = fy * tz * dT1[:, 0]
    vrow[:, 5] = fy * tz * dT1[:, 1]
    vrow[:, 2] = -s0 * m[:, 0] * fx * tz2 - s1f * m[:, 0] * fy * tz2 + (tz - v[2] * m[:, 0]) * fx * tz2 * dT0[:, 2]
    vrow[:, 6] = -s0f * m[:, 1] * fx * tz2 - s1 * m[:, 1] * fy * tz2 + (tz - v[6] * m[:, 1]) * fy * tz2 * dT1[:, 2]
    vrow[:, 10] = -(s0f * fx + s1f * fy) * m[:, 2] * tz2
    vrow[:, 14] = -(s0f * fx + s1f * fy) * tz2
    vrow[:, 2] += m[:, 0] * ddepth
    vrow[:, 6] += m[:, 1] * ddepth
    vrow[:, 10] += m[:, 2] * ddepth
    vrow[:, 14] += ddepth

    dmean = torch.stack([v[0] * dtx + v[1] * dty + v[2] * dtz,
                         v[4] * dtx + v[5] * dty + v[6] * dtz,
                         v[8] * dtx + v[9] * dty + v[10] * dtz], 1)

    # ---- K9: projection part ----
    m_w = 1.0 / (p_hom[:, 3] + 0.0000001)
    pr = proj
    mul1 = (pr[0] * m[:, 0] + pr[4] * m[:, 1] + pr[8] * m[:, 2] + pr[12]) * m_w * m_w
    mul2 = (pr[1] * m[:, 0] + pr[5] * m[:, 1] + pr[9] * m[:, 2] + pr[13]) * m_w * m_w
    hw, hh = float(sc.W // 2), float(sc.H // 2)                       # Q9: integer division
    dmx, dmy = dm2[:, 0], dm2[:, 1]
    pm = torch.stack([hw * (pr[0] * m_w - pr[3] * mul1) * dmx + hh * (pr[1] * m_w - pr[3] * mul2) * dmy,
                      hw * (pr[4] * m_w - pr[7] * mul1) * dmx + hh * (pr[5] * m_w - pr[7] * mul2) * dmy,
                      hw * (pr[8] * m_w - pr[11] * mul1) * dmx + hh * (pr[9] * m_w - pr[11] * mul2) * dmy], 1)
    k = -1.0 / (m_w * m_w)
    first = dmx * hw * p_hom[:, 0] * k * t3                           # Q11: theta3 in all three first summands
    dshift = dshift + torch.stack([first + dmy * hh * p_hom[:, 1] * k * t3,
                                   first + dmy * hh * p_hom[:, 1] * k * t5,
                                   first + dmy * hh * p_hom[:, 1] * k * t7], 1)
    dproj = torch.zeros(vi.shape[0], 16, dtype=F64)                   # Q10: rows 2,6,10,14 stay zero
    dproj[:, 0] = m_w * m[:, 0] * dmx
    dproj[:, 4] = m_w * m[:, 1] * dmx
    dproj[:, 8] = m_w * m[:, 2] * dmx
    dproj[:, 12] = m_w * dmx
    dproj[:, 1] = m_w * m[:, 0] * dmy
    dproj[:, 5] = m_w * m[:, 1] * dmy
    dproj[:, 9] = m_w * m[:, 2] * dmy
    dproj[:, 13] = m_w * dmy
    q = -(m_w * m_w) * (p_hom[:, 0] * dmx + p_hom[:, 1] * dmy)
    dproj[:, 3] = q * m[:, 0]
    dproj[:, 7] = q * m[:, 1]
    dproj[:, 11] = q * m[:, 2]
    dproj[:, 15] = q
    dmean = dmean + pm
    dmean = dmean + torch.stack([v[2] * ddepth, v[6] * ddepth, v[10] * ddepth], 1)

    out = {}
    M = sc.shs.shape[1] if sc.shs is not None else 0
    dsh_full = torch.zeros(P, M, 3, dtype=F64)
    dcampos = torch.zeros(vi.shape[0], 3, dtype=F64)
    if sc.shs is not None:
        # backward.cu:37-170 (direction WITHOUT pupil shift)
        sh = sc.shs[vi]
        dRGB = acc["dL_dcolors"][vi] * (~pre["clamped"][vi]).to(F64)
        dir_orig = m - sc.campos[None, :]
        dirn = dir_orig / dir_orig.norm(dim=1, keepdim=True)
        x, y, z = dirn[:, 0:1], dirn[:, 1:2], dirn[:, 2:3]
        dsh = torch.zeros(vi.shape[0], M, 3, dtype=F64)
        dRGBdx = torch.zeros(vi.shape[0], 3, dtype=F64)
        dRGBdy = torch.zeros_like(dRGBdx)
        dRGBdz = torch.zeros_like(dRGBdx)
        dsh[:, 0] = SH_C0 * dRGB
        if sc.D > 0:
            dsh[:, 1] = -SH_C1 * y * dRGB
            dsh[:, 2] = SH_C1 * z * dRGB
            dsh[:, 3] = -SH_C1 * x * dRGB
            dRGBdx = -SH_C1 * sh[:, 3]
            dRGBdy = -SH_C1 * sh[:, 1]
            dRGBdz = SH_C1 * sh[:, 2]
            if sc.D > 1:
                xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
                dsh[:, 4] = SH_C2[0] * xy * dRGB
                dsh[:, 5] = SH_C2[1] * yz * dRGB
                dsh[:, 6] = SH_C2[2] * (2.0 * zz - xx - yy) * dRGB
                dsh[:, 7] = SH_C2[3] * xz * dRGB
                dsh[:, 8] = SH_C2[4] * (xx - yy) * dRGB
                dRGBdx = dRGBdx + SH_C2[0] * y * sh[:, 4] + SH_C2[2] * 2.0 * -x * sh[:, 6] + SH_C2[3] * z * sh[:, 7] + SH_C2[4] * 2.0 * x * sh[:, 8]
                dRGBdy = dRGBdy + SH_C2[0] * x * sh[:, 4] + SH_C2[1] * z * sh[:, 5] + SH_C2[2] * 2.0 * -y * sh[:, 6] + SH_C2[4] * 2.0 * -y * sh[:, 8]
                dRGBdz = dRGBdz + SH_C2[1] * y * sh[:, 5] + SH_C2[2] * 2.0 * 2.0 * z * sh[:, 6] + SH_C2[3] * x * sh[:, 7]
                if sc.D > 2:
                    dsh[:, 9] = SH_C3[0] * y * (3.0 * xx - yy) * dRGB
                    dsh[:, 10] = SH_C3[1] * xy * z * dRGB
                    dsh[:, 11] = SH_C3[2] * y * (4.0 * zz - xx - yy) * dRGB
                    dsh[:, 12] = SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * dRGB
                    dsh[:, 13] = SH_C3[4] * x * (4.0 * zz - xx - yy) * dRGB
                    dsh[:, 14] = SH_C3[5] * z * (xx - yy) * dRGB
                    dsh[:, 15] = SH_C3[6] * x * (xx - 3.0 * yy) * dRGB
                    dRGBdx = dRGBdx + (SH_C3[0] * sh[:, 9] * 3.0 * 2.0 * xy + SH_C3[1] * sh[:, 10] * yz +
                                       SH_C3[2] * sh[:, 11] * -2.0 * xy + SH_C3[3] * sh[:, 12] * -3.0 * 2.0 * xz +
                                       SH_C3[4] * sh[:, 13] * (-3.0 * xx + 4.0 * zz - yy) +
                                       SH_C3[5] * sh[:, 14] * 2.0 * xz + SH_C3[6] * sh[:, 15] * 3.0 * (xx - yy))
                    dRGBdy = dRGBdy + (SH_C3[0] * sh[:, 9] * 3.0 * (xx - yy) + SH_C3[1] * sh[:, 10] * xz +
                                       SH_C3[2] * sh[:, 11] * (-3.0 * yy + 4.0 * zz - xx) +
                                       SH_C3[3] * sh[:, 12] * -3.0 * 2.0 * yz + SH_C3[4] * sh[:, 13] * -2.0 * xy +
                                       SH_C3[5] * sh[:, 14] * -2.0 * yz + SH_C3[6] * sh[:, 15] * -3.0 * 2.0 * xy)
                    dRGBdz = dRGBdz + (SH_C3[1] * sh[:, 10] * xy + SH_C3[2] * sh[:, 11] * 4.0 * 2.0 * yz +
                                       SH_C3[3] * sh[:, 12] * 3.0 * (2.0 * zz - xx - yy) +
                                       SH_C3[4] * sh[:, 13] * 4.0 * 2.0 * xz + SH_C3[5] * sh[:, 14] * (xx - yy))
        ddir = torch.stack([(dRGBdx * dRGB).sum(1), (dRGBdy * dRGB).sum(1), (dRGBdz * dRGB).sum(1)], 1)
        vv = dir_orig
        sum2 = (vv * vv).sum(1)
        inv32 = 1.0 / torch.sqrt(sum2 * sum2 * sum2)
        dmx3 = ((sum2 - vv[:, 0] ** 2) * ddir[:, 0] - vv[:, 1] * vv[:, 0] * ddir[:, 1] - vv[:, 2] * vv[:, 0] * ddir[:, 2]) * inv32
        dmy3 = (-vv[:, 0] * vv[:, 1] * ddir[:, 0] + (sum2 - vv[:, 1] ** 2) * ddir[:, 1] - vv[:, 2] * vv[:, 1] * ddir[:, 2]) * inv32
        dmz3 = (-vv[:, 0] * vv[:, 2] * ddir[:, 0] - vv[:, 1] * vv[:, 2] * ddir[:, 1] + (sum2 - vv[:, 2] ** 2) * ddir[:, 2]) * inv32
        dsh_mean = torch.stack([dmx3, dmy3, dmz3], 1)
        dmean = dmean + dsh_mean
        dcampos = -dsh_mean
        dsh_full[vi] = dsh

    dscale = torch.zeros(P, 3, dtype=F64)
    drot = torch.zeros(P, 4, dtype=F64)
    if sc.scales is not None:
        # backward.cu:377-440; gradient w.r.t. the given (not re-normalised) quaternion
        Rq = _rot_from_quat(sc.rotations[vi])
        s = sc.mod * sc.scales[vi]
        Mstd = s[:, :, None] * Rq.transpose(1, 2)        # S @ Rq^T  (the reference's M = S * R in standard form)
        dSig = torch.stack([torch.stack([dcov[:, 0], 0.5 * dcov[:, 1], 0.5 * dcov[:, 2]], 1),
                            torch.stack([0.5 * dcov[:, 1], dcov[:, 3], 0.5 * dcov[:, 4]], 1),
                            torch.stack([0.5 * dcov[:, 2], 0.5 * dcov[:, 4], dcov[:, 5]], 1)], 1)
        dM = 2.0 * Mstd @ dSig                           # standard form of dL_dM
        # dL_dMt[i][j] (GLM) of the reference == dM[i][j] in standard form; Rt[i][j] (GLM) == Rq^T[i][j]... see note
        RqT = Rq.transpose(1, 2)
        dscale_v = torch.stack([(RqT[:, 0, :] * dM[:, 0, :]).sum(1), (RqT[:, 1, :] * dM[:, 1, :]).sum(1),
                                (RqT[:, 2, :] * dM[:, 2, :]).sum(1)], 1)
        dMt = dM * s[:, :, None]
        g = lambda i, j: dMt[:, i, j]
        r, x, y, z = sc.rotations[vi].unbind(1)
        dq = torch.stack([
            2 * z * (g(0, 1) - g(1, 0)) + 2 * y * (g(2, 0) - g(0, 2)) + 2 * x * (g(1, 2) - g(2, 1)),
            2 * y * (g(1, 0) + g(0, 1)) + 2 * z * (g(2, 0) + g(0, 2)) + 2 * r * (g(1, 2) - g(2, 1)) - 4 * x * (g(2, 2) + g(1, 1)),
            2 * x * (g(1, 0) + g(0, 1)) + 2 * r * (g(2, 0) - g(0, 2)) + 2 * z * (g(1, 2) + g(2, 1)) - 4 * y * (g(2, 2) + g(0, 0)),
            2 * r * (g(0, 1) - g(1, 0)) + 2 * x * (g(2, 0) + g(0, 2)) + 2 * y * (g(1, 2) + g(2, 1)) - 4 * z * (g(1, 1) + g(0, 0))], 1)
        dscale[vi] = dscale_v
        drot[vi] = dq

    def scatter(x, width):
        full = torch.zeros(P, width, dtype=F64)
        full[vi] = x
        return full

    out["dL_dmeans3D"] = scatter(dmean, 3)
    out["dL_dmeans2D"] = torch.cat([acc["dL_dmean2D"], torch.zeros(P, 1, dtype=F64)], 1) * vis[:, None]
    out["dL_dmeans2D_densify"] = torch.cat([acc["dL_dmean2D_abs"], torch.zeros(P, 1, dtype=F64)], 1) * vis[:, None]
    out["dL_dcolors"] = acc["dL_dcolors"]
    out["dL_dopacity"] = acc["dL_dopacity"].reshape(P, 1)
    out["dL_dcov3D"] = scatter(dcov, 6)
    out["dL_dsh"] = dsh_full
    out["dL_dscales"] = dscale
    out["dL_drotations"] = drot
    out["dL_dcampos"] = dcampos.sum(0) / P                            # Q12: mean over all P rows
    out["dL_dprojmatrix"] = dproj.sum(0) / P
    out["dL_dshift"] = dshift.sum(0) / P
    # conditioning of the three reduced vectors: mean over P of |term| (fp32 summation order perturbs the reference's
    # own .mean(0) by ~1e-6 of this, which can exceed 1e-4 of the signed result when the terms cancel)
    out["pose_abs"] = {"dL_dcampos": dcampos.abs().sum(0) / P, "dL_dprojmatrix": dproj.abs().sum(0) / P,
                       "dL_dshift": dshift.abs().sum(0) / P}
    out["covariance"] = acc["covariance"].reshape(P, 1) + 1e-4        # Q14
    out["viewmatrix_rows"] = vrow / P                                  # Q13: admissible values (one row each)
    out["viewmatrix_row_ids"] = vi
    out["dL_dviewmatrix"] = vrow[-1] / P if vi.numel() else torch.zeros(16, dtype=F64)   # highest visible index
    return out


def rasterize_forward(sc):
    pre = preprocess(sc)
    point_list, keys, ranges = bin_tiles(sc, pre)
    color, depth, alpha, final_T, n_contrib = blend_forward(sc, pre, point_list, ranges)
    return dict(pre=pre, point_list=point_list, keys=keys, ranges=ranges, color=color, depth=depth, alpha=alpha,
                final_T=final_T, n_contrib=n_contrib, num_rendered=int(point_list.shape[0]),
                mean2D=torch.stack([pre["mean2Dx"], pre["mean2Dy"]], 1), radii=pre["radii"])


def rasterize_backward(sc, fwd, dL_dcolor, dL_ddepth, dL_dalpha):
    acc = blend_backward(sc, fwd["pre"], fwd["point_list"], fwd["ranges"], dL_dcolor, dL_ddepth, dL_dalpha)
    return preprocess_backward(sc, fwd["pre"], acc)
