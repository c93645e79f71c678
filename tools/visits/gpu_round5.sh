#!/bin/bash
# GPU-box visit 5 (session 8): validate HEAD (gradient sinks) -- parity suite, both bench arms, GPU timeline summary.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time python -m pytest tests -m gpu -q --durations=8 --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py --trace gpurun_out/trace_summary.json > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --no-grad-sinks --no-cpu-baseline > gpurun_out/bench_ours_nosink.json 2> gpurun_out/bench_ours_nosink.err
ls -la gpurun_out
