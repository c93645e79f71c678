"""CPU: the post-step oracle (oracle/post_oracle.py) against golden vectors produced by the UNMODIFIED reference
modules (oracle/make_golden_post.py imports utils/util_distortion.py and utils/loss_utils.py from /root/reference)."""
import glob
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys
sys.path.insert(0, ROOT)
from oracle import post_oracle as PO  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def _t(a, dtype=torch.float32):
    return torch.from_numpy(np.asarray(a)).to(dtype)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "post_distort_*.npz"))))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_distortion_oracle_matches_reference(path, dtype):
    z = np.load(path)
    image = _t(z["image"], dtype).requires_grad_(True)
    pts = _t(z["pts"], dtype).requires_grad_(True)
    proj = _t(z["proj"], dtype)
    Hs, Ws = [int(v) for v in z["ctrl_shape"]]
    fish = [int(v) for v in z["fish"]]
    fs = z["flow_scale"]
    ctrl = (torch.cat([pts, torch.ones(pts.shape[0], 1, dtype=dtype)], 1).reshape(Hs, Ws, 3)[:, :, :2]) @ proj[:2, :2]
    canvas = (int(fish[0] * fs[0]), int(fish[1] * fs[1]))
    out, mask, flow = PO.distort_render(image, ctrl, canvas, fish)
    (out * _t(z["up"], dtype)).sum().backward()
    # float32: same operator sequence as the reference -> rounding-level agreement; float64: bounded by the fp32
    # conditioning of the reference itself (flow rounding x image gradient)
    tol = 2e-6 if dtype == torch.float32 else 2e-5
    assert (out.detach().float() - _t(z["out"])).abs().max() < tol
    assert (flow.detach().float() - _t(z["flow"])).abs().max() < 1e-6
    assert (mask.float() != _t(z["mask"])).float().mean() < (0.0 if dtype == torch.float32 else 1e-3) + 1e-9
    for got, want in ((image.grad, z["d_image"]), (pts.grad, z["d_pts"])):
        want = _t(want)
        rel = (got.float() - want).norm() / want.norm()
        assert rel < (1e-5 if dtype == torch.float32 else 1e-4), rel
    # reuse of the canvas-resolution flow (second and later iterations)
    out2, _, _ = PO.distort_render(_t(z["image"], dtype), _t(z["flow"], dtype), None, fish)
    assert (out2.float() - _t(z["out_reuse"])).abs().max() < tol
    # apply2gt branch
    gt, mgt, fgt = PO.distort_gt(_t(z["gt_src"], dtype), ctrl.detach(), (image.shape[1], image.shape[2]))
    assert (gt.float() - _t(z["gt_out"])).abs().max() < tol
    assert (fgt.float() - _t(z["gt_flow"])).abs().max() < 1e-6
    # center_crop alone
    ci = _t(z["cc_in"], dtype).requires_grad_(True)
    cc = PO.center_crop(ci, fish[0], fish[1])
    (cc * _t(z["cc_up"], dtype)).sum().backward()
    assert (cc.detach().float() - _t(z["cc_out"])).abs().max() < (1e-6 if dtype == torch.float32 else 2e-5)
    # (float64 un-normalises the float32 grid without the float32 rounding: the ~1e-5 second-tap weights differ)
    assert (ci.grad.float() - _t(z["cc_grad"])).abs().max() < (1e-5 if dtype == torch.float32 else 5e-4)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "post_loss_*.npz"))))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_loss_oracle_matches_reference(path, dtype):
    z = np.load(path)
    img = _t(z["img"], dtype).requires_grad_(True)
    gt = _t(z["gt"], dtype)
    loss, l1, s = PO.training_loss(img, gt, 0.2)
    loss.backward()
    assert abs(float(l1) - float(z["l1"])) < 1e-6
    assert abs(float(s) - float(z["ssim"])) < 2e-6
    assert abs(float(loss) - float(z["loss"])) < 1e-6
    want = _t(z["d_img"])
    assert (img.grad.float() - want).norm() / want.norm() < 2e-5


def test_goldens_exist():
    assert len(glob.glob(os.path.join(GOLD, "post_distort_*.npz"))) >= 2
    assert len(glob.glob(os.path.join(GOLD, "post_loss_*.npz"))) >= 2
