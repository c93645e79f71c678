#!/usr/bin/env python
"""Optimizer step of the reference's six parameter groups at P Gaussians (59 floats each): BagsAdam (one launch,
28 B/element) against torch.optim.Adam as the reference constructs it (foreach) and torch's fused=True variant.

    python tools/bench_adam.py [--P 1000000] [--iters 50]

Development / reporting tool: prints one JSON line (CUDA-event times, algorithmic GB/s = 28 B x elements / time)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
import torch  # noqa: E402

import bags_optim  # noqa: E402

GROUPS = [("xyz", (3,), 1.6e-4), ("f_dc", (1, 3), 2.5e-3), ("f_rest", (15, 3), 2.5e-3 / 20.0), ("opacity", (1,), 0.05),
          ("scaling", (3,), 5e-3), ("rotation", (4,), 1e-3)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    out = {"P": a.P, "elements": a.P * 59, "algorithmic_bytes": a.P * 59 * 28}
    peak = None
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for name, make in [("bags", lambda l: bags_optim.BagsAdam(l, lr=0.0, eps=1e-15)),
                       ("torch_foreach", lambda l: torch.optim.Adam(l, lr=0.0, eps=1e-15)),
                       ("torch_fused", lambda l: torch.optim.Adam(l, lr=0.0, eps=1e-15, fused=True))]:
        params = [torch.nn.Parameter(torch.randn(a.P, *s, generator=g).to(dev)) for _, s, _ in GROUPS]
        for p in params:
            p.grad = torch.randn_like(p) * 1e-3
        opt = make([{"params": [p], "lr": lr, "name": n} for p, (n, _, lr) in zip(params, GROUPS)])
        for _ in range(5):
            opt.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        gbs = out["algorithmic_bytes"] / (ms * 1e-3) / 1e9
        out[name] = {"ms_per_step": round(ms, 4), "alg_GBps": round(gbs, 1),
                     "frac_of_measured_hbm_peak": round(gbs / peak, 4) if peak else None}
        del opt, params
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
