#!/bin/bash
# Round 2, GPU visit 10: longest-first tile order in the blend kernels -- full GPU suite, bench batch / graph, per-camera times.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err; echo "bench batch rc=$?"; tail -3 gpurun_out/bench_batch.err
python bench.py --mode graph --no-cpu-baseline --no-drop-in --no-e2e > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench graph rc=$?"
python tools/tile_balance.py 0.004 4 > gpurun_out/tile_balance2.jsonl 2> gpurun_out/tile_balance2.err; cut -c1-200 gpurun_out/tile_balance2.jsonl
python - <<'PY'
import json
for f in ("bench_batch", "bench_graph"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None, "drop_in", round(d["drop_in"]["value"], 1) if d.get("drop_in") else None,
              "launches", d.get("gpu_launches"), {k: v["ms"] for k, v in d["stages"].items()}, d["preprocess_binning"])
    except Exception as e:
        print(f, "ERR", e)
PY
