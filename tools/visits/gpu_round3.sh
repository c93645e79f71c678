#!/bin/bash
# GPU-box visit 3: full parity suite (incl. the new post-step tests), ncu --set full of every kernel of the 4 bench
# views summarised ON THE BOX (the .ncu-rep is too large to travel), bench lines of the other configs.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
ncu --set full --clock-control none --profile-from-start off \
    -k regex:'preprocess|onesweep|sort_hist|emit_|tile_ranges|blend_|pose_finalize' -f -o /tmp/full_all \
    python tests/harness/ncu_target.py 4 > gpurun_out/full_all.log 2>&1
python profiles/summarize_ncu.py traffic /tmp/full_all.ncu-rep > gpurun_out/ncu_traffic.json 2> gpurun_out/traffic.err
python profiles/summarize_ncu.py full /tmp/full_all.ncu-rep > gpurun_out/full_all_summary.txt 2>> gpurun_out/traffic.err
python tests/harness/bench_configs.py --out gpurun_out/configs.jsonl > gpurun_out/configs.log 2>&1
ls -la gpurun_out
