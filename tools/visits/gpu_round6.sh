#!/bin/bash
# GPU-box visit 6 (session 8): parity suite after the oracle/test fixes, look-back width sweep of the onesweep passes,
# bench, ncu --set full of every kernel of one bench view (summarised on the box).
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python tools/stage_times.py --ppt 2:2 --tune 3=8,3=16,3=32,3=64,3=8,3=32 > gpurun_out/lookback_sweep.jsonl 2> gpurun_out/lookback_sweep.err
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_run.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'preprocess|onesweep|sort_hist|emit_|tile_ranges|blend_|pose_finalize' -f -o /tmp/full_all \
    python tests/harness/ncu_target.py 1 > gpurun_out/full_all.log 2>&1
python profiles/summarize_ncu.py traffic /tmp/full_all.ncu-rep > gpurun_out/ncu_traffic.json 2> gpurun_out/traffic.err
python profiles/summarize_ncu.py full /tmp/full_all.ncu-rep > gpurun_out/full_all_summary.txt 2>> gpurun_out/traffic.err
sz=$(stat -c %s /tmp/full_all.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 0 ] && [ "$sz" -lt 45000000 ]; then cp /tmp/full_all.ncu-rep gpurun_out/full_all.ncu-rep; fi
ls -la gpurun_out
