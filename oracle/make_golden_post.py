#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's post-rasterizer steps (distortion, loss).

    python oracle/make_golden_post.py        (build container: needs /root/reference; writes tests/golden/post_*.npz)

Imports the UNMODIFIED reference modules utils/util_distortion.py and utils/loss_utils.py from /root/reference and runs
apply_distortion / center_crop / l1_loss / ssim on seeded host inputs.  The reference hard-codes `.cuda()` in
center_crop and imports packages that are irrelevant to these functions (imageio, matplotlib, the scene package); in
this GPU-less container `Tensor.cuda` is patched to the identity and those imports are stubbed -- none of the
reference's own code is modified.  Each file stores inputs, outputs and gradients (float32, computed by the reference
on the host)."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BAGS_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "scene", "scene.dataset_readers"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["scene.dataset_readers"].read_intrinsics_binary = None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    torch.Tensor.cuda = lambda self, *a, **k: self          # no GPU here; the functions are device-agnostic otherwise
    sys.path.insert(0, REF)
    import utils.util_distortion as ud
    import utils.loss_utils as lu
    return ud, lu


class FixedLens:
    """Stands in for the lens network: returns a fixed leaf tensor of (N,2) directions."""
    def __init__(self, pts):
        self.pts = pts

    def forward(self, P, sensor_to_frustum=False):
        return self.pts


def smooth_image(C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    y, x = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    img = torch.zeros(C, H, W)
    for c in range(C):
        for _ in range(4):
            fx, fy, ph = (torch.rand(3, generator=g) * torch.tensor([6.0, 6.0, 6.28])).tolist()
            img[c] += 0.125 * torch.sin(fx * x * 6.28 + fy * y * 6.28 + ph)
        img[c] += 0.5
    return img.contiguous()


def control_flow(Hs, Ws, seed, spread=1.0, noise=0.01):
    """Identity sampling grid at control-point resolution (scaled by `spread`: > 1 looks outside the image and
    produces zero padding) plus a smooth perturbation."""
    g = torch.Generator().manual_seed(seed)
    y, x = torch.meshgrid(torch.linspace(-1, 1, Hs), torch.linspace(-1, 1, Ws), indexing="ij")
    r2 = x * x + y * y
    k = 0.08
    fx = spread * x * (1 + k * r2) + noise * torch.randn(Hs, Ws, generator=g)
    fy = spread * y * (1 + k * r2) + noise * torch.randn(Hs, Ws, generator=g)
    return torch.stack([fx, fy], -1).contiguous()


DISTORT_CASES = {
    # name: image (H, W), control grid (Hs, Ws), canvas = fish_res * flow_scale, crop = fish_res
    "post_distort_a": dict(H=96, W=128, Hs=12, Ws=16, fish=(48, 64), flow_scale=(2.0, 2.0), seed=11, spread=1.0),
    "post_distort_b": dict(H=90, W=70, Hs=9, Ws=11, fish=(51, 37), flow_scale=(1.5, 1.7), seed=12, spread=1.25),
}


def run_distort(ud, c):
    H, W, Hs, Ws = c["H"], c["W"], c["Hs"], c["Ws"]
    image = smooth_image(3, H, W, c["seed"]).requires_grad_(True)
    ctrl = control_flow(Hs, Ws, c["seed"] + 100, c["spread"])
    proj = torch.tensor([[0.9, 0.05, 0, 0], [-0.03, 1.1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
    pts = (ctrl.reshape(-1, 2) @ torch.linalg.inv(proj[:2, :2])).detach().requires_grad_(True)   # flow = pts @ proj
    cam = types.SimpleNamespace(projection_matrix=proj, flow4gt=proj, image_height=H, image_width=W,
                                fish_gt_image_resolution=(3, c["fish"][0], c["fish"][1]), fish_gt_image=None)
    P_sensor = torch.zeros(Hs, Ws, 2)
    out, mask, flow = ud.apply_distortion(None, FixedLens(pts), None, P_sensor, cam, image, apply2gt=False,
                                          flow_scale=list(c["flow_scale"]))
    g = torch.Generator().manual_seed(c["seed"] + 200)
    up = torch.randn(out.shape, generator=g)
    (out * up).sum().backward()
    res = dict(image=image.detach(), pts=pts.detach(), proj=proj, up=up, out=out.detach(), mask=mask.detach(),
               flow=flow.detach(), d_image=image.grad, d_pts=pts.grad, fish=np.array(c["fish"]),
               flow_scale=np.array(c["flow_scale"]), ctrl_shape=np.array([Hs, Ws]))
    # second call of the training loop: the canvas-resolution flow is handed back (no interpolation)
    out2, mask2, _ = ud.apply_distortion(flow.detach(), None, None, P_sensor, cam, image.detach(), apply2gt=False,
                                         flow_scale=list(c["flow_scale"]))
    res["out_reuse"] = out2
    # apply2gt branch on the same data: the "fish-eye ground truth" is resampled to the render resolution
    cam.fish_gt_image = smooth_image(3, c["fish"][0], c["fish"][1], c["seed"] + 300)
    pts2 = pts.detach().clone().requires_grad_(True)
    gt, mgt, fgt = ud.apply_distortion(None, FixedLens(pts2), None, P_sensor, cam, None, apply2gt=True)
    res.update(gt_src=cam.fish_gt_image, gt_out=gt.detach(), gt_mask=mgt, gt_flow=fgt.detach())
    # center_crop alone
    cc_in = smooth_image(3, H, W, c["seed"] + 400).unsqueeze(0).requires_grad_(True)
    cc = ud.center_crop(cc_in, c["fish"][0], c["fish"][1])
    upc = torch.randn(cc.shape, generator=g)
    (cc * upc).sum().backward()
    res.update(cc_in=cc_in.detach(), cc_out=cc.detach(), cc_up=upc, cc_grad=cc_in.grad)
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in res.items()}


LOSS_CASES = {"post_loss_a": dict(H=64, W=80, seed=21), "post_loss_b": dict(H=37, W=53, seed=22)}


def run_loss(lu, c):
    g = torch.Generator().manual_seed(c["seed"])
    img = (smooth_image(3, c["H"], c["W"], c["seed"]) + 0.05 * torch.randn(3, c["H"], c["W"], generator=g)) \
        .clamp(0, 1).requires_grad_(True)
    gt = (smooth_image(3, c["H"], c["W"], c["seed"]) + 0.05 * torch.randn(3, c["H"], c["W"], generator=g)).clamp(0, 1)
    Ll1 = lu.l1_loss(img, gt)
    s = lu.ssim(img, gt)
    loss = (1.0 - 0.2) * Ll1 + 0.2 * (1.0 - s)           # train.py:325 with lambda_dssim = 0.2
    loss.backward()
    return dict(img=img.detach().numpy(), gt=gt.numpy(), l1=Ll1.item(), ssim=s.item(), loss=loss.item(),
                d_img=img.grad.numpy())


def main():
    ud, lu = import_reference()
    os.makedirs(OUT, exist_ok=True)
    for name, c in DISTORT_CASES.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **run_distort(ud, c))
        print("wrote", name)
    for name, c in LOSS_CASES.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **run_loss(lu, c))
        print("wrote", name)


if __name__ == "__main__":
    main()
