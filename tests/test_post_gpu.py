"""GPU parity of the post-rasterizer steps: fused distortion (N1) and fused L1+SSIM loss (N2).

Checked against (i) the golden vectors produced by the unmodified reference modules and (ii) the reference's own
operator sequence (oracle/post_oracle.py = F.interpolate / F.grid_sample / F.conv2d) run live on the same device,
at small and at BASELINE config-4 sizes.  Tolerances (fp32): images <= 1e-5 max-abs on smooth inputs (the chain
amplifies the fp32 rounding of the flow by (W-1)/2 pixels x image gradient, for the reference as much as for us),
gradients <= 1e-4 relative L2, loss values <= 1e-6."""
import glob
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
from oracle import post_oracle as PO  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _dev():
    return torch.device("cuda:0")


def _t(a):
    return torch.from_numpy(np.asarray(a)).float().to(_dev())


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


class FixedLens:
    def __init__(self, pts):
        self.pts = pts

    def forward(self, P, sensor_to_frustum=False):
        return self.pts


def smooth(C, H, W, seed, dev):
    g = torch.Generator().manual_seed(seed)
    y, x = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    img = torch.zeros(C, H, W)
    for c in range(C):
        for _ in range(4):
            fx, fy, ph = (torch.rand(3, generator=g) * torch.tensor([6.0, 6.0, 6.28])).tolist()
            img[c] += 0.125 * torch.sin(fx * x * 6.28 + fy * y * 6.28 + ph)
        img[c] += 0.5
    return img.contiguous().to(dev)


def ctrl_flow(Hs, Ws, seed, spread, dev, noise=0.01):
    g = torch.Generator().manual_seed(seed)
    y, x = torch.meshgrid(torch.linspace(-1, 1, Hs), torch.linspace(-1, 1, Ws), indexing="ij")
    r2 = x * x + y * y
    fx = spread * x * (1 + 0.08 * r2) + noise * torch.randn(Hs, Ws, generator=g)
    fy = spread * y * (1 + 0.08 * r2) + noise * torch.randn(Hs, Ws, generator=g)
    return torch.stack([fx, fy], -1).contiguous().to(dev)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "post_distort_*.npz"))))
def test_distortion_against_reference_golden(path):
    import bags_distortion as BD
    z = np.load(path)
    image = _t(z["image"]).requires_grad_(True)
    pts = _t(z["pts"]).requires_grad_(True)
    proj = _t(z["proj"])
    Hs, Ws = [int(v) for v in z["ctrl_shape"]]
    fish = [int(v) for v in z["fish"]]
    cam = types.SimpleNamespace(projection_matrix=proj, flow4gt=proj, image_height=image.shape[1],
                                image_width=image.shape[2], fish_gt_image_resolution=(3, fish[0], fish[1]),
                                fish_gt_image=_t(z["gt_src"]))
    P_sensor = torch.zeros(Hs, Ws, 2)
    out, mask, flow = BD.apply_distortion(None, FixedLens(pts), None, P_sensor, cam, image, apply2gt=False,
                                          flow_scale=[float(v) for v in z["flow_scale"]])
    (out * _t(z["up"])).sum().backward()
    assert float((out.detach() - _t(z["out"])).abs().max()) <= 1e-5
    assert float((flow - _t(z["flow"])).abs().max()) <= 1e-6
    assert float((mask != _t(z["mask"])).float().mean()) == 0.0
    assert rel_l2(image.grad, _t(z["d_image"])) <= 1e-4
    assert rel_l2(pts.grad, _t(z["d_pts"])) <= 1e-4
    out2, mask2, flow2 = BD.apply_distortion(flow, None, None, P_sensor, cam, image.detach(), apply2gt=False,
                                             flow_scale=[float(v) for v in z["flow_scale"]])
    assert float((out2 - _t(z["out_reuse"])).abs().max()) <= 1e-5
    assert flow2 is flow
    gt, mgt, fgt = BD.apply_distortion(None, FixedLens(pts.detach()), None, P_sensor, cam, None, apply2gt=True)
    assert float((gt - _t(z["gt_out"])).abs().max()) <= 1e-5
    assert float((mgt != _t(z["gt_mask"])).float().mean()) == 0.0
    assert float((fgt - _t(z["gt_flow"])).abs().max()) <= 1e-6
    ci = _t(z["cc_in"]).requires_grad_(True)
    cc = BD.center_crop(ci, fish[0], fish[1])
    (cc * _t(z["cc_up"])).sum().backward()
    assert float((cc.detach() - _t(z["cc_out"])).abs().max()) <= 1e-6
    assert float((ci.grad - _t(z["cc_grad"])).abs().max()) <= 1e-5


CASES_LIVE = [
    # H, W (render canvas), Hs, Ws (control grid), canvas (Hf, Wf), crop (Hc, Wc), spread
    dict(H=270, W=480, Hs=34, Ws=60, canvas=(270, 480), crop=(135, 240), spread=1.0),
    dict(H=301, W=517, Hs=38, Ws=65, canvas=(402, 689), crop=(301, 517), spread=1.3),     # zero padding + odd sizes
    dict(H=200, W=200, Hs=25, Ws=25, canvas=(200, 200), crop=(200, 200), spread=0.9),     # crop == canvas
    dict(H=2160, W=3840, Hs=270, Ws=480, canvas=(2160, 3840), crop=(1080, 1920), spread=1.0),   # BASELINE config 4
]


@pytest.mark.parametrize("c", CASES_LIVE, ids=lambda c: f"{c['W']}x{c['H']}")
def test_distortion_against_live_operator_sequence(c):
    import bags_distortion as BD
    dev = _dev()
    image = smooth(3, c["H"], c["W"], 5, dev)
    ctrl = ctrl_flow(c["Hs"], c["Ws"], 6, c["spread"], dev)
    g = torch.Generator().manual_seed(7)
    up = torch.randn(3, *c["crop"], generator=g).to(dev)
    i1, f1 = image.clone().requires_grad_(True), ctrl.clone().requires_grad_(True)
    out, mask, flow = BD.distort_image(i1, f1, canvas_hw=c["canvas"], crop_hw=c["crop"], return_flow=True)
    (out * up).sum().backward()
    i2, f2 = image.clone().requires_grad_(True), ctrl.clone().requires_grad_(True)
    want, wmask, wflow = PO.distort_render(i2, f2, c["canvas"], c["crop"])
    (want * up).sum().backward()
    assert float((flow - wflow).abs().max()) <= 1e-6
    assert float((out - want).abs().max()) <= 1e-5
    assert float((mask != wmask).float().mean()) <= 1e-6
    assert rel_l2(i1.grad, i2.grad) <= 1e-4
    assert rel_l2(f1.grad, f2.grad) <= 1e-4
    # float64 evaluation of the same chain: both fp32 implementations sit within the chain's conditioning of it
    w64, _, _ = PO.distort_render(image.double(), ctrl.double(), c["canvas"], c["crop"])
    e_ours, e_ref = float((out - w64).abs().max()), float((want - w64).abs().max())
    assert e_ours <= max(2.0 * e_ref, 2e-5), (e_ours, e_ref)


def test_distortion_gt_branch_and_edge_cases():
    import bags_distortion as BD
    dev = _dev()
    gt = smooth(3, 120, 160, 8, dev)
    ctrl = ctrl_flow(15, 20, 9, 1.2, dev)
    out, mask = BD.distort_image(gt, ctrl, canvas_hw=(240, 320), mask_mode=1)
    want, wmask, _ = PO.distort_gt(gt, ctrl, (240, 320))
    assert float((out - want).abs().max()) <= 1e-5
    assert float((mask != wmask).float().mean()) <= 1e-6
    # single channel; flow entirely outside the image -> zeros and an all-zero mask; NaN flow -> zeros
    one = gt[:1].contiguous()
    o1, m1 = BD.distort_image(one, ctrl, canvas_hw=(60, 80), crop_hw=(30, 40))
    w1 = PO.center_crop(torch.nn.functional.grid_sample(one[None], PO.upsample_flow(ctrl, (60, 80))[None],
                                                        align_corners=True), 30, 40)[0]
    assert float((o1 - w1).abs().max()) <= 1e-5
    far = torch.full((4, 4, 2), 3.0, device=dev)
    o2, m2 = BD.distort_image(gt, far, canvas_hw=(32, 32), crop_hw=(16, 16))
    assert float(o2.abs().max()) == 0.0 and float(m2.max()) == 0.0
    nan = torch.full((4, 4, 2), float("nan"), device=dev)
    o3, _ = BD.distort_image(gt, nan, canvas_hw=(32, 32), crop_hw=(16, 16))
    assert float(o3.abs().max()) == 0.0
    with pytest.raises(RuntimeError):
        BD.distort_image(gt.cpu(), ctrl.cpu())
    with pytest.raises(Exception):
        BD.distort_image(gt, ctrl, canvas_hw=(10, 10), crop_hw=(20, 20))


# ---- fused L1 + SSIM loss (N2) ------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "post_loss_*.npz"))))
def test_loss_against_reference_golden(path):
    import bags_loss as BL
    z = np.load(path)
    img = _t(z["img"]).requires_grad_(True)
    gt = _t(z["gt"])
    loss, l1, s = BL.training_loss(img, gt, 0.2)
    loss.backward()
    assert abs(float(l1) - float(z["l1"])) <= 1e-6
    assert abs(float(s) - float(z["ssim"])) <= 5e-6
    assert abs(float(loss) - float(z["loss"])) <= 2e-6
    assert rel_l2(img.grad, _t(z["d_img"])) <= 1e-4
    # the reference's two separate calls
    i2 = _t(z["img"]).requires_grad_(True)
    l = (1.0 - 0.2) * BL.l1_loss(i2, gt) + 0.2 * (1.0 - BL.ssim(i2, gt))
    l.backward()
    assert abs(float(l) - float(z["loss"])) <= 2e-6
    assert rel_l2(i2.grad, _t(z["d_img"])) <= 1e-4


@pytest.mark.parametrize("hw", [(1080, 1920), (840, 1297), (33, 31), (5, 7)], ids=lambda h: f"{h[1]}x{h[0]}")
def test_loss_against_live_operator_sequence(hw):
    import bags_loss as BL
    dev = _dev()
    H, W = hw
    g = torch.Generator().manual_seed(31)
    img = (smooth(3, H, W, 32, dev) + 0.05 * torch.randn(3, H, W, generator=g).to(dev)).clamp(0, 1)
    gt = (smooth(3, H, W, 32, dev) + 0.05 * torch.randn(3, H, W, generator=g).to(dev)).clamp(0, 1)
    a = img.clone().requires_grad_(True)
    loss, l1, s = BL.training_loss(a, gt, 0.2)
    loss.backward()
    b = img.double().requires_grad_(True)            # float64 evaluation of the reference's operator sequence
    wloss, wl1, ws = PO.training_loss(b, gt.double(), 0.2)
    wloss.backward()
    assert abs(float(l1) - float(wl1)) <= 1e-6
    assert abs(float(s) - float(ws)) <= 5e-6
    assert abs(float(loss) - float(wloss)) <= 2e-6
    assert rel_l2(a.grad, b.grad) <= 1e-4
    c = img.clone().requires_grad_(True)             # and its float32 evaluation on the same device (cuDNN convs)
    floss, _, fs = PO.training_loss(c, gt, 0.2)
    floss.backward()
    assert abs(float(s) - float(fs)) <= 1e-5
    assert rel_l2(a.grad, c.grad) <= 1e-4


def test_loss_batched_and_edge_cases():
    import bags_loss as BL
    dev = _dev()
    g = torch.Generator().manual_seed(41)
    x = torch.rand(2, 3, 40, 56, generator=g).to(dev)
    y = torch.rand(2, 3, 40, 56, generator=g).to(dev)
    got = BL.ssim(x, y, size_average=False)
    want = PO.ssim(x.double(), y.double(), size_average=False)
    assert got.shape == (2,) and float((got.double() - want).abs().max()) <= 5e-6
    assert abs(float(BL.ssim(x, y)) - float(PO.ssim(x.double(), y.double()))) <= 5e-6
    assert abs(float(BL.l1_loss(x, y)) - float(PO.l1_loss(x.double(), y.double()))) <= 1e-6
    assert abs(float(BL.ssim(x, x)) - 1.0) <= 1e-6          # identical images
    z = x.clone().requires_grad_(True)                     # L1 gradient: sign / numel, exactly 0 where equal
    y2 = y.clone()
    y2[0, 0, :4] = x[0, 0, :4]
    BL.l1_loss(z, y2).backward()
    assert float(z.grad[0, 0, :4].abs().max()) == 0.0
    want_g = torch.sign(x - y2) / x.numel()
    assert float((z.grad - want_g).abs().max()) <= 1e-12
    with pytest.raises(RuntimeError):
        BL.l1_loss(x.cpu(), y.cpu())
    with pytest.raises(RuntimeError):
        BL.ssim(x, y, window_size=7)
