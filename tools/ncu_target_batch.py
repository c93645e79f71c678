#!/usr/bin/env python
"""Short workload for ncu captures of the batched launch set: the four bench cameras of C3 (1M Gaussians, 1080p) as ONE
GaussianBatchRasterizer forward + backward (gradient sinks on, like bench.py's default mode)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
import torch  # noqa: E402

import bags_synth as S  # noqa: E402
import bags_view_parallel as VP  # noqa: E402
import diff_gaussian_rasterization as mod  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
c = S.CONFIGS["C3"]
W, H, P = c["W"], c["H"], c["P"]
dev = torch.device("cuda:0")
cam0 = S.make_camera(W, H, seed=100)
gs = S.make_gaussians(P, cam0, scale=scale)
params = VP.FlatGaussianParams(gs.means3D.to(dev), gs.shs.to(dev), gs.opacities.to(dev), gs.scales.to(dev),
                               gs.rotations.to(dev), device=dev)
params.enable_direct_accumulation(mod)
bg = torch.zeros(3, device=dev)
cams = [S.make_camera(W, H, seed=100 + v) for v in range(B)]
sets = [mod.GaussianRasterizationSettings(H, W, cm.tanfovx, cm.tanfovy, bg, 1.0, cm.viewmatrix.to(dev).requires_grad_(True),
                                          cm.projmatrix.to(dev).requires_grad_(True), cm.intrinsic.to(dev), 3,
                                          cm.campos.to(dev).requires_grad_(True), False, False, 0) for cm in cams]
ups = [torch.stack([S.make_upstream(H, W, seed=1 + v)[k] for v in range(B)]).to(dev) for k in range(3)]
rast = mod.GaussianBatchRasterizer(sets, 10_000_000)
m2d = torch.zeros(B, P, 3, device=dev, requires_grad=True)
m2dd = torch.zeros(B, P, 3, device=dev, requires_grad=True)
shift = torch.zeros(3, device=dev, requires_grad=True)
for phase in ("warm", "profiled"):
    if phase == "profiled":
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    params.zero_grad()
    with params.accumulating():
        color, radii, depth, weights, _ = rast(params["means3D"], m2d, m2dd, shift, params["opacities"],
                                               shs=params["shs"], scales=params["scales"], rotations=params["rotations"])
        torch.autograd.backward([color, depth, weights], ups)
    torch.cuda.synchronize()
    print(phase, "counts", rast.counts.tolist())
torch.cuda.profiler.stop()
