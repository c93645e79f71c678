#!/usr/bin/env python
"""Who is closer to float64?  Runs one view of a named config through this repo's rasterizer, the unmodified
reference build (oracle/_ref) and the float64 oracle (oracle/raster_oracle.py, on the GPU) and prints, for every
gradient, the relative L2 distance of both fp32 implementations from the float64 result.  Evidence tool for the
pose / intrinsic gradient tolerance (tolerance notes); prints one JSON line per config.

    python tests/harness/pose_oracle_check.py C1 C3 [--scale 0.004]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import parity_common as pc  # noqa: E402
from oracle import raster_oracle as O  # noqa: E402

S = pc.S


def one(name, scale, shift=(0.0, 0.0, 0.0)):
    c = S.CONFIGS[name]
    dev = torch.device("cuda:0")
    cam = S.make_camera(c["W"], c["H"])
    gs = S.make_gaussians(c["P"], cam, scale=scale)
    up = S.make_upstream(c["H"], c["W"])
    o = pc.run_view(pc.load_ours(), cam, gs, up, dev, shift=shift)
    r = pc.run_view(pc.load_ref(), cam, gs, up, dev, shift=shift)
    t0 = time.time()
    with O.on_device("cuda:0"):
        sc = O.Scene(gs.means3D, gs.opacities, cam.viewmatrix, cam.projmatrix, cam.campos, cam.W, cam.H, cam.tanfovx,
                     cam.tanfovy, (0, 0, 0), shs=gs.shs, scales=gs.scales, rotations=gs.rotations,
                     shift_factors=torch.tensor(shift), sh_degree=3)
        f = O.rasterize_forward(sc)
        b = O.rasterize_backward(sc, f, *up)
    torch.cuda.synchronize()
    t_oracle = time.time() - t0
    keys = {"means3D": "dL_dmeans3D", "shs": "dL_dsh", "opacities": "dL_dopacity", "scales": "dL_dscales",
            "rotations": "dL_drotations", "campos": "dL_dcampos", "projmatrix": "dL_dprojmatrix",
            "shift_factors": "dL_dshift", "means2D": "dL_dmeans2D", "means2D_densify": "dL_dmeans2D_densify"}
    out = {"config": name, "scale": scale, "shift": list(shift), "R": o["num_rendered"],
           "R_oracle": int(f["num_rendered"]), "oracle_seconds": round(t_oracle, 1), "rel_l2_vs_float64": {}}
    for k, ok in keys.items():
        ref64 = b[ok].detach().cpu().reshape(-1)
        eo = pc.rel_l2(o["grads"][k].reshape(-1), ref64)
        er = pc.rel_l2(r["grads"][k].reshape(-1), ref64)
        out["rel_l2_vs_float64"][k] = {"ours": float(f"{eo:.3e}"), "reference": float(f"{er:.3e}"),
                                       "ours_vs_reference": float(f"{pc.rel_l2(o['grads'][k], r['grads'][k]):.3e}")}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    argv = sys.argv[1:]
    scale = 0.004
    if "--scale" in argv:
        k = argv.index("--scale")
        scale = float(argv[k + 1])
        del argv[k:k + 2]
    args = [a for a in argv if not a.startswith("--")]
    for n in args or ["C1"]:
        one(n, scale)
        if n == "C1":
            one(n, scale, shift=(1e-3, -5e-4, 2e-4))
