#!/bin/bash
# Round 2, GPU visit 13: the shared-memory variant of the blend backward's per-splat reduction (BLEND_BWD_REDUCE = 1):
# parity suite with the knob on, bench with / without, ncu instruction count.
set -u
mkdir -p gpurun_out
BAGS_TEST_TUNE="8=1" python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_reduce1.log 2>&1; echo "pytest (reduce=1) rc=$?"; tail -3 gpurun_out/pytest_gpu_reduce1.log
for t in 0 1; do
  python bench.py --no-cpu-baseline --no-drop-in --no-e2e --tune BLEND_BWD_REDUCE=$t > gpurun_out/bench_reduce$t.json 2> gpurun_out/bench_reduce$t.err; echo "bench reduce=$t rc=$?"
done
python - <<'PY'
import json
for t in (0, 1):
    d = json.load(open(f"gpurun_out/bench_reduce{t}.json"))
    print("reduce", t, round(d["value"], 1), {k: v["ms"] for k, v in d["stages"].items() if "blend" in k})
PY
python tools/stage_times.py --views 4 --ppt 2:2 --tune 8=0,8=1 2>/dev/null | cut -c1-400
