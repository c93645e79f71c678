#!/bin/bash
# GPU-box visit 4 (session 5): parity suite, both bench arms (+ GPU timeline of the device / e2e legs), ncu launch
# list, ncu --set full of every kernel of 2 bench views summarised ON THE BOX (the .ncu-rep is too large to travel).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time python -m pytest tests -m gpu -q --durations=8 --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py --trace gpurun_out/trace_summary.json > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'preprocess|onesweep|sort_hist|emit_|tile_ranges|blend_|pose_finalize' -f -o /tmp/full_all \
    python tests/harness/ncu_target.py 2 > gpurun_out/full_all.log 2>&1
python profiles/summarize_ncu.py traffic /tmp/full_all.ncu-rep > gpurun_out/ncu_traffic.json 2> gpurun_out/traffic.err
python profiles/summarize_ncu.py full /tmp/full_all.ncu-rep > gpurun_out/full_all_summary.txt 2>> gpurun_out/traffic.err
ls -la gpurun_out
