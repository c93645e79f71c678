#!/bin/bash
# GPU-box visit 9 (session 8): parity suite (sort variants), small-tile / match.any sweep of the sort passes, bench.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python tools/stage_times.py --ppt 2:2 --tune 4=0,4=1,5=1,5=0,4=0,4=1,5=1 > gpurun_out/sort_sweep.jsonl 2> gpurun_out/sort_sweep.err
python tools/stage_times.py --config C1 --ppt 2:2 --tune 4=0,4=1,5=1,5=0 > gpurun_out/sort_sweep_c1.jsonl 2>> gpurun_out/sort_sweep.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
ls -la gpurun_out
