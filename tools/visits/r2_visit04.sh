#!/bin/bash
# Round 2, GPU visit 4: batched launch set (N3) tests, bench batch vs graph mode, reference-glue test after the loader fix,
# densify phase timing.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_batch_gpu.py tests/test_reference_glue_gpu.py -q -x > gpurun_out/pytest_batch.log 2>&1; echo "pytest batch rc=$?"; tail -25 gpurun_out/pytest_batch.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python bench.py --mode batch --no-cpu-baseline > gpurun_out/bench_batch.json 2> gpurun_out/bench_batch.err; echo "bench batch rc=$?"; tail -3 gpurun_out/bench_batch.err
python bench.py --mode graph --no-cpu-baseline --no-drop-in > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench graph rc=$?"
python tests/harness/bench_densify.py 1000000 > gpurun_out/densify.jsonl 2> gpurun_out/densify.err; echo "densify rc=$?"; cat gpurun_out/densify.jsonl
python - <<'PY'
import json
for f in ("bench_batch", "bench_graph"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "launches", d.get("gpu_launches"),
              {k: v["ms"] for k, v in d["stages"].items()}, d["preprocess_binning"])
    except Exception as e:
        print(f, "ERR", e)
PY
