#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference build (oracle/_ref).

Run on a GPU box:   python oracle/make_golden.py            (writes gpurun_out/golden/*.npz; copy into tests/golden/)

Each file holds the seeded inputs of one small scene and everything the reference returns for it through its public
API (GaussianRasterizer forward + autograd backward), plus the integer internals parsed out of its opaque state
buffers (tiles_touched, sorted keys, point list, tile ranges, n_contrib, final_T).  The CPU test-suite checks the
float64 oracle (oracle/raster_oracle.py) against these files; the GPU suite checks the CUDA path against them and
against the live reference build.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc  # noqa: E402

S = pc.S

CASES = {
    # name: scene + call options
    "tiny_sh3": dict(P=3000, W=160, H=120, scale=0.02, sh_degree=3),
    "tiny_bg_shift_sh2": dict(P=3000, W=176, H=100, scale=0.02, sh_degree=2, bg=(1.0, 0.5, 0.25),
                              shift=(1e-3, -5e-4, 2e-4), cam_seed=101, gs_seed=3),
    "tiny_precomp": dict(P=2500, W=128, H=128, scale=0.025, precomp=True, cam_seed=102, gs_seed=4),
    "tiny_sh0_mod": dict(P=2000, W=96, H=64, scale=0.03, sh_degree=0, scale_modifier=0.7, cam_seed=103, gs_seed=5),
}


def build_case(c):
    cam = S.make_camera(c["W"], c["H"], seed=c.get("cam_seed", 100))
    gs = S.make_gaussians(c["P"], cam, seed=c.get("gs_seed", 0), scale=c["scale"])
    up = S.make_upstream(c["H"], c["W"], seed=c.get("up_seed", 1))
    kw = dict(sh_degree=c.get("sh_degree", 3), bg=c.get("bg", (0.0, 0.0, 0.0)), shift=c.get("shift", (0.0, 0.0, 0.0)),
              scale_modifier=c.get("scale_modifier", 1.0))
    if c.get("precomp"):
        g = torch.Generator().manual_seed(77)
        colors = torch.rand(c["P"], 3, generator=g)
        # cov3D = R S^2 R^T in fp32 on the host (what GaussianModel.get_covariance hands over)
        from oracle.raster_oracle import cov3d_from_scale_rot
        cov = cov3d_from_scale_rot(gs.scales.double(), 1.0, gs.rotations.double()).float()
        kw.update(use_colors_precomp=True, use_cov_precomp=True, colors_precomp=colors, cov3D_precomp=cov)
    return cam, gs, up, kw


def main():
    sys.path.insert(0, ROOT)
    dev = torch.device("cuda:0")
    ref = pc.load_ref()
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, c in CASES.items():
        cam, gs, up, kw = build_case(c)
        r = pc.run_view(ref, cam, gs, up, dev, **kw)
        d = {}
        for k in ("color", "radii", "depth", "weights", "mean2D"):
            d[k] = r[k].numpy()
        d["num_rendered"] = np.int64(r["num_rendered"])
        for k, v in r["state"].items():
            d["state_" + k] = v.numpy()
        for k, v in r["grads"].items():
            if v is not None:
                d["grad_" + k] = v.numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **d)
        print(name, "R =", r["num_rendered"], "visible =", int((r["radii"] > 0).sum()),
              "size", os.path.getsize(os.path.join(out_dir, name + ".npz")))


if __name__ == "__main__":
    main()
