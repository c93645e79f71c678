#!/bin/bash
# GPU-box visit 2: ncu --set full over every kernel of the 4 bench views (-> profiles/ncu_traffic.json), bench lines
# of the other configs, headline bench.
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --profile-from-start off \
    -k regex:'preprocess|onesweep|sort_hist|emit_|tile_ranges|blend_|pose_finalize' -f -o gpurun_out/full_all \
    python tests/harness/ncu_target.py 4 > gpurun_out/full_all.log 2>&1
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python tests/harness/bench_configs.py --out gpurun_out/configs.jsonl > gpurun_out/configs.log 2>&1
ls -la gpurun_out
