#!/bin/bash
# Round 2, GPU visit 6: full GPU suite after the blend_bwd instruction cuts (fast exp / rcp / RED), tile balance per bench
# camera, ncu of the blend kernels (same camera as visit 2 for an instruction-count comparison), densify timing.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python tools/tile_balance.py 0.004 4 > gpurun_out/tile_balance.jsonl 2> gpurun_out/tile_balance.err; echo "balance rc=$?"; cut -c1-420 gpurun_out/tile_balance.jsonl
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:blend -o gpurun_out/blend_v6 python tests/harness/ncu_target.py 1 0.004 ours > gpurun_out/ncu_blend_v6.out 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/blend_v6.ncu-rep --page raw --csv --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active 2>/dev/null | cut -c1-300 | tail -3
python tests/harness/bench_densify.py 1000000 > gpurun_out/densify.jsonl 2> gpurun_out/densify.err; echo "densify rc=$?"; tail -1 gpurun_out/densify.jsonl
