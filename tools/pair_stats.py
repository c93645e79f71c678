#!/usr/bin/env python
"""Development tool: histogram of valid lanes per (warp, splat) iteration of blend_bwd, from a -DBAGS_STATS build
(tools/libbags_stats.so).  Not part of the product or the bench."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
import bags_capi as capi  # noqa: E402
capi.LIB_PATH = os.path.join(ROOT, "tools", "libbags_stats.so")
import torch  # noqa: E402
import bags_synth as S  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc  # noqa: E402

L = capi.lib()
L.bags_debug_stats.restype = ctypes.c_int
L.bags_debug_stats.argtypes = [ctypes.c_void_p, ctypes.c_int]
for scale in (0.004, 0.012):
    c = S.CONFIGS["C3"]
    cam = S.make_camera(c["W"], c["H"], seed=100)
    gs = S.make_gaussians(c["P"], cam, scale=scale)
    up = S.make_upstream(c["H"], c["W"])
    buf = (ctypes.c_ulonglong * 64)()
    L.bags_debug_stats(buf, 1)
    o = pc.run_view(pc.load_ours(), cam, gs, up, torch.device("cuda:0"))
    L.bags_debug_stats(buf, 1)
    h = [int(buf[i]) for i in range(33)]
    cells = [int(buf[40 + i]) for i in range(5)]
    tot = sum(h)
    pairs = sum(i * h[i] for i in range(33))
    print(json.dumps({"scale": scale, "R": o["num_rendered"], "iterations": tot, "valid_pairs": pairs,
                      "mean_valid_lanes": pairs / max(tot, 1), "zero_valid_iters": h[0],
                      "hist_valid_lanes": h, "cells_hit_hist(0..4)": cells}))
