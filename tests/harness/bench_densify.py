#!/usr/bin/env python
"""densify_and_prune at bench size: the reference's GaussianModel method (staged under oracle/_ref/glue) against
bags_densify.densify_and_prune on identical 1 M-Gaussian models.  Wall-clock per call including the host
synchronisations each implementation needs (CUDA events around the call + a final synchronize).  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import test_densify_gpu as T  # noqa: E402
import bags_densify  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    ns = T.load_glue("ours")
    res = {"P": P}
    for name in ("reference", "ours"):
        times, rows = [], None
        for rep in range(4):
            m = T.build_model(ns, P, steps=1)
            torch.manual_seed(1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name == "reference":
                m.densify_and_prune(0.0002, 0.005, 4.0, 20)
            else:
                bags_densify.densify_and_prune(m, 0.0002, 0.005, 4.0, 20)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
            rows = m.get_xyz.shape[0]
            del m
            torch.cuda.empty_cache()
        res[name] = {"ms": [round(t, 3) for t in times], "best_ms": round(min(times[1:]), 3), "rows_after": rows}
    # where ours spends its time (each phase closed by a synchronisation; development aid)
    m = T.build_model(ns, P, steps=1)
    torch.manual_seed(1)
    phases = {}
    bags_densify.densify_and_prune(m, 0.0002, 0.005, 4.0, 20, _timing=phases)
    res["ours_phases_ms"] = phases
    del m
    res["speedup"] = round(res["reference"]["best_ms"] / res["ours"]["best_ms"], 2)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
