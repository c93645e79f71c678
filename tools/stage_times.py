#!/usr/bin/env python
"""Per-stage CUDA-event times of one view (fwd+bwd) on the GPU box, for several pixels-per-thread settings of the blend kernels.

    python tools/stage_times.py [--config C3] [--scale 0.004] [--views 6] [--ppt 1:1,2:4]

Development tool (not part of the bench contract): prints one JSON line per group size."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
import torch  # noqa: E402

import bags_capi as capi  # noqa: E402
import bags_synth as S  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3")
    ap.add_argument("--scale", type=float, default=0.004)
    ap.add_argument("--views", type=int, default=6)
    ap.add_argument("--ppt", default="1:1,2:2,2:4,4:4,1:4", help="fwd:bwd pixels per thread")
    ap.add_argument("--P", type=int, default=0)
    ap.add_argument("--tune", default="", help="extra knob settings tried in turn, e.g. 3=8,3=32 (knob=value)")
    a = ap.parse_args()
    cfg = S.CONFIGS[a.config]
    P, W, H = a.P or cfg["P"], cfg["W"], cfg["H"]
    dev = torch.device("cuda:0")
    cam0 = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam0, seed=0, scale=a.scale)
    par = {k: getattr(gs, k).to(dev).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    shift = torch.zeros(3, device=dev, requires_grad=True)
    bg = torch.zeros(3, device=dev)
    views = []
    for v in range(a.views):
        cam = S.make_camera(W, H, seed=100 + v)
        views.append((cam, [t.to(dev).requires_grad_(True) for t in (cam.viewmatrix, cam.projmatrix, cam.campos)],
                      cam.intrinsic.to(dev), [u.to(dev) for u in S.make_upstream(H, W, seed=1 + v)]))

    def one(vw):
        cam, (vm, pm, cp), intr, up = vw
        st = dgr.GaussianRasterizationSettings(H, W, cam.tanfovx, cam.tanfovy, bg, 1.0, vm, pm, intr, 3, cp, False,
                                               False, 0)
        m2 = torch.zeros_like(par["means3D"], requires_grad=True)
        m2d = torch.zeros_like(par["means3D"], requires_grad=True)
        color, radii, depth, w, _ = dgr.GaussianRasterizer(st)(par["means3D"], m2, m2d, shift, par["opacities"],
                                                               shs=par["shs"], scales=par["scales"],
                                                               rotations=par["rotations"])
        R = color.grad_fn.num_rendered
        torch.autograd.backward([color, depth, w], up)
        return R, int((radii > 0).sum())

    L = capi.lib()
    groups = [(G, t) for t in (a.tune.split(",") if a.tune else [""]) for G in a.ppt.split(",")]
    for G, tune in groups:
        if tune:
            L.bags_set_tuning(int(tune.split("=")[0]), int(tune.split("=")[1]))
        L.bags_set_tuning(0, int(G.split(":")[0]))
        L.bags_set_tuning(1, int(G.split(":")[1]))
        for vw in views[:2]:
            R, V = one(vw)
        torch.cuda.synchronize()
        L.bags_profile_enable(1)
        capi.profile_read()
        for vw in views:
            one(vw)
        ms, calls = capi.profile_read()
        L.bags_profile_enable(0)
        torch.cuda.synchronize()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            for vw in views:
                one(vw)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.time() - t0) / (3 * len(views)) * 1e3
        dev_ms = e0.elapsed_time(e1) / (3 * len(views))
        out = {k: round(ms[k] / max(calls["preprocess"], 1), 4) for k in ms}
        print(json.dumps({"ppt_fwd:bwd": G, "tune": tune, "P": P, "V": V, "R": R, "scale": a.scale, "stage_ms": out,
                          "sum_ms": round(sum(out.values()), 4), "ms_per_view_events": round(dev_ms, 4),
                          "ms_per_view_wall": round(wall, 4)}))


if __name__ == "__main__":
    main()
