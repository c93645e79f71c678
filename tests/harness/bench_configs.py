#!/usr/bin/env python
"""Bench lines for the BASELINE.json configs other than the headline one, both arms, plus simple-knn on 6 M points.

    python tests/harness/bench_configs.py [--out gpurun_out/configs.jsonl] [--steps 10]

C1/C2/C4: `bench.py --config Cx` (ours, then --impl reference).  C5: `bench.py --config C3 --P n` for the
densification sweep n in {100k, 300k, 1M, 3M, 6M} and distCUDA2 on 6 M points U(-5,5)^3 (seed 2), ours and reference.
Development / reporting tool: writes one JSON object per line."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_bench(extra, steps):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", "3",
           "--no-cpu-baseline", "--no-e2e", "--no-drop-in", "--min-seconds", "1", *extra]
    r = subprocess.run(cmd, capture_output=True, text=True)
    for line in r.stdout.splitlines()[::-1]:
        if line.startswith("{"):
            return json.loads(line)
    return {"error": (r.stderr or r.stdout)[-400:], "cmd": " ".join(extra)}


def knn_lines():
    import torch
    import parity_common as pc
    from simple_knn._C import distCUDA2 as ours
    g = torch.Generator().manual_seed(2)
    pts = ((torch.rand(6_000_000, 3, generator=g) * 10) - 5).cuda()
    out = []
    impls = [("ours", ours)]
    ref_so = os.path.join(pc.REF_DIR, "simple_knn", "_C.so")
    if os.path.exists(ref_so):
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_simple_knn._C", ref_so)
        refC = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(refC)
        impls.append(("reference", refC.distCUDA2))
    for name, fn in impls:
        fn(pts)
        torch.cuda.synchronize()
        n = 5 if name == "ours" else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn(pts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out.append({"workload": "C5 knn: distCUDA2 on 6M points U(-5,5)^3", "impl": name, "ms": ms,
                    "points_per_s": 6e6 / (ms * 1e-3), "alg_GBps_at_136B_per_point": 136 * 6e6 / (ms * 1e-3) / 1e9})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.jsonl"))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--skip-ref", action="store_true")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    keep = ("metric", "value", "ms_per_step", "ms_per_step_regions", "config", "workload_stats", "stages", "preprocess_binning", "impl", "error", "clocks")
    with open(a.out, "w") as f:
        def emit(tag, line):
            line = {k: v for k, v in line.items() if k in keep}
            line["tag"] = tag
            f.write(json.dumps(line) + "\n")
            f.flush()
        for cfg in ("C1", "C2", "C4"):
            emit(cfg, run_bench(["--config", cfg], a.steps))
            if not a.skip_ref:
                emit(cfg, run_bench(["--config", cfg, "--impl", "reference"], max(a.steps // 2, 3)))
        for n in (100_000, 300_000, 1_000_000, 3_000_000, 6_000_000):
            emit(f"C5-{n}", run_bench(["--config", "C3", "--P", str(n)], a.steps))
            if not a.skip_ref and n in (100_000, 6_000_000):
                emit(f"C5-{n}", run_bench(["--config", "C3", "--P", str(n), "--impl", "reference"], 3))
        for line in knn_lines():
            line["tag"] = "C5-knn"
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
