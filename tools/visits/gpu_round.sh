#!/bin/bash
# One GPU-box visit: parity suite, both bench arms, ncu launch list and full captures.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none -k regex:bags --launch-skip 20 --launch-count 20 -f -o gpurun_out/full_all \
    python tests/harness/ncu_target.py 2 > gpurun_out/full_all.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_ --launch-skip 2 --launch-count 2 -f \
    -o gpurun_out/full_blend python tests/harness/ncu_target.py 2 > gpurun_out/full_blend.log 2>&1
ls -la gpurun_out
