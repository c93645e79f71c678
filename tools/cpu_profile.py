#!/usr/bin/env python
"""Where does the HOST time of a view go?  cProfile of fwd+bwd views of a small scene (C1: the GPU side is 0.35 ms per
view there, so the loop is host bound).  Development tool.

    python tools/cpu_profile.py [--config C1] [--views 300]
"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
import torch  # noqa: E402

import bags_synth as S  # noqa: E402
import bags_view_parallel as VP  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C1")
    ap.add_argument("--views", type=int, default=300)
    a = ap.parse_args()
    cfg = S.CONFIGS[a.config]
    P, W, H = cfg["P"], cfg["W"], cfg["H"]
    dev = torch.device("cuda:0")
    cam = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam, seed=0, scale=0.004)
    params = VP.FlatGaussianParams(gs.means3D.to(dev), gs.shs.to(dev), gs.opacities.to(dev), gs.scales.to(dev),
                                   gs.rotations.to(dev), device=dev)
    params.enable_direct_accumulation(dgr)
    dgr.set_grad_sinks_active(True)
    bg = torch.zeros(3, device=dev)
    shift = torch.zeros(3, device=dev, requires_grad=True)
    vm, pm, cp = [t.to(dev).requires_grad_(True) for t in (cam.viewmatrix, cam.projmatrix, cam.campos)]
    intr = cam.intrinsic.to(dev)
    up = [u.to(dev) for u in S.make_upstream(H, W, seed=1)]

    def one():
        st = dgr.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, vm, pm, intr, 3, cp, False,
                                               False, 0)
        m2 = torch.zeros_like(params["means3D"], requires_grad=True)
        m2d = torch.zeros_like(params["means3D"], requires_grad=True)
        color, radii, depth, w, _ = dgr.GaussianRasterizer(st)(params["means3D"], m2, m2d, shift, params["opacities"],
                                                               shs=params["shs"], scales=params["scales"],
                                                               rotations=params["rotations"])
        torch.autograd.backward([color, depth, w], up)

    for _ in range(20):
        one()
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(a.views):
        one()
    torch.cuda.synchronize()
    print(f"unprofiled: {(time.time() - t0) / a.views * 1e3:.3f} ms per view (wall)")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(a.views):
        one()
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
    print(s.getvalue())
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(25)
    print(s.getvalue())


if __name__ == "__main__":
    main()
