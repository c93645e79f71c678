#!/usr/bin/env python
"""Development tool: GPU timeline (kernel name / start / duration, via torch.profiler + CUPTI) of a few bench steps,
to see idle gaps and the non-library kernels between our launches.  Writes gpurun_out/timeline.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bags_synth as S  # noqa: E402
import bags_view_parallel as VP  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402

c = S.CONFIGS["C3"]
P, W, H = c["P"], c["W"], c["H"]
dev = torch.device("cuda:0")
cam0 = S.make_camera(W, H, seed=100)
gs = S.make_gaussians(P, cam0, seed=0, scale=0.004)
params = VP.FlatGaussianParams(gs.means3D.to(dev), gs.shs.to(dev), gs.opacities.to(dev), gs.scales.to(dev),
                               gs.rotations.to(dev), device=dev)
bg = torch.zeros(3, device=dev)
shift = torch.zeros(3, device=dev, requires_grad=True)
views = []
for v in range(4):
    cam = S.make_camera(W, H, seed=100 + v)
    views.append(dict(cam=cam, view=cam.viewmatrix.to(dev).requires_grad_(True),
                      proj=cam.projmatrix.to(dev).requires_grad_(True),
                      campos=cam.campos.to(dev).requires_grad_(True), intrinsic=cam.intrinsic.to(dev),
                      up=[u.to(dev) for u in S.make_upstream(H, W, seed=1 + v)]))


def step():
    params.zero_grad()
    for vw in views:
        st = dgr.GaussianRasterizationSettings(H, W, vw["cam"].tanfovx, vw["cam"].tanfovy, bg, 1.0, vw["view"],
                                               vw["proj"], vw["intrinsic"], 3, vw["campos"], False, False, 0)
        m2d = torch.zeros_like(params["means3D"], requires_grad=True)
        m2dd = torch.zeros_like(params["means3D"], requires_grad=True)
        color, radii, depth, weights, _ = dgr.GaussianRasterizer(st)(
            means3D=params["means3D"], means2D=m2d, means2D_densify=m2dd, shift_factors=shift,
            opacities=params["opacities"], shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        torch.autograd.backward([color, depth, weights], vw["up"])


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "trace.json")
prof.export_chrome_trace(path)
tr = json.load(open(path))
ev = [e for e in tr["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
out = [{"name": e["name"][:80], "ts": e["ts"], "dur": e["dur"], "cat": e["cat"]} for e in ev]
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "timeline.json"), "w"))
os.remove(path)
print("events", len(out))
