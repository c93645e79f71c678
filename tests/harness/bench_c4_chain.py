#!/usr/bin/env python
"""BASELINE.json config C4 as specified: 2 M Gaussians rasterized to a 4K (3840x2160) extended-FoV canvas, THEN the
lens-distortion post-step of --outside_rasterizer (utils/util_distortion.py:271-311 + 58-77: F.interpolate of the
control-point flow, F.grid_sample of the render, center_crop, mask), forward + backward, per view.

  arm "ours"      : this repo's rasterizer + bags_distortion.distort_image (one fused gather / scatter kernel pair)
  arm "reference" : the unmodified reference rasterizer (oracle/_ref) + the reference's torch operator sequence
                    (oracle/post_oracle.distort_render, pinned to the unmodified utils/util_distortion.py by
                    tests/test_post_oracle_golden.py)
Same harness for both: blocking forward, one stream, CUDA events, median of 5 regions of `--views` views.
Prints one JSON line per arm."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import parity_common as pc  # noqa: E402

S = pc.S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--arms", default="ours,reference")
    a = ap.parse_args()
    c = S.CONFIGS["C4"]
    P, W, H = c["P"], c["W"], c["H"]
    Hc, Wc = H // 2, W // 2                   # the fisheye ground-truth frame the canvas is cropped to
    dev = torch.device("cuda:0")
    cam0 = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam0, seed=0, scale=0.004)
    params = [t.to(dev).requires_grad_(True) for t in (gs.means3D, gs.shs, gs.opacities, gs.scales, gs.rotations)]
    cams = [S.make_camera(W, H, seed=100 + v) for v in range(4)]
    g = torch.Generator().manual_seed(3)
    # smooth random flow: control grid H/8 x W/8, N(0, 0.01^2) around the identity
    ys = torch.linspace(-1, 1, H // 8)
    xs = torch.linspace(-1, 1, W // 8)
    ident = torch.stack(torch.meshgrid(xs, ys, indexing="xy"), dim=-1)            # (Hs, Ws, 2) in (x, y)
    ctrl = (ident + 0.01 * torch.randn(ident.shape, generator=g)).to(dev).requires_grad_(True)
    up = (torch.randn(3, Hc, Wc, generator=g) / (Hc * Wc)).to(dev)
    bg = torch.zeros(3, device=dev)
    shift = torch.zeros(3, device=dev, requires_grad=True)

    for arm in a.arms.split(","):
        if arm == "ours":
            mod = pc.load_ours()
            import bags_distortion as BD

            def post(img):
                return BD.distort_image(img, ctrl, canvas_hw=(H, W), crop_hw=(Hc, Wc))[0]
        else:
            mod = pc.load_ref()
            from oracle import post_oracle as PO

            def post(img):
                return PO.distort_render(img, ctrl, (H, W), (Hc, Wc))[0]
        leaves = [(cm.viewmatrix.to(dev).requires_grad_(True), cm.projmatrix.to(dev).requires_grad_(True),
                   cm.campos.to(dev).requires_grad_(True), cm.intrinsic.to(dev)) for cm in cams]

        def one(k):
            cm, (view, proj, campos, intr) = cams[k % 4], leaves[k % 4]
            for p in params:
                p.grad = None
            st = mod.GaussianRasterizationSettings(H, W, cm.tanfovx, cm.tanfovy, bg, 1.0, view, proj, intr, 3, campos,
                                                   False, False, 0)
            m2 = torch.zeros_like(params[0], requires_grad=True)
            m2d = torch.zeros_like(params[0], requires_grad=True)
            color, radii, depth, weights, _ = mod.GaussianRasterizer(st)(
                means3D=params[0], means2D=m2, means2D_densify=m2d, shift_factors=shift, opacities=params[2],
                shs=params[1], scales=params[3], rotations=params[4])
            out = post(color)
            (out * up).sum().backward()
            return color.grad_fn

        for k in range(4):
            one(k)
        torch.cuda.synchronize()
        regions = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(a.views):
                one(k)
            e1.record()
            torch.cuda.synchronize()
            regions.append(e0.elapsed_time(e1) / a.views)
        # the post-step alone on a fixed canvas
        canvas = torch.rand(3, H, W, device=dev, requires_grad=True)
        for _ in range(3):
            (post(canvas) * up).sum().backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            canvas.grad = None
            (post(canvas) * up).sum().backward()
        e1.record()
        torch.cuda.synchronize()
        regions.sort()
        print(json.dumps({"workload": f"C4 chain: {P} Gaussians SH3 -> {W}x{H} canvas -> distortion flow "
                                      f"({H // 8}x{W // 8} control grid) -> center crop {Wc}x{Hc}, fwd+bwd",
                          "impl": arm, "ms_per_view": round(regions[len(regions) // 2], 3),
                          "views_per_s": round(1e3 / regions[len(regions) // 2], 2),
                          "ms_per_view_regions": [round(r, 3) for r in regions],
                          "post_step_alone_ms": round(e0.elapsed_time(e1) / 10, 3)}), flush=True)


if __name__ == "__main__":
    main()
