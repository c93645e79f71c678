#!/bin/bash
# Round 2, GPU visit 12: ncu --set full of every kernel of the batched launch set (4 bench cameras), DRAM traffic per stage.
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:preprocess|onesweep|sort_hist|emit_|tile_|blend_|pose_' -o gpurun_out/batch_all python tools/ncu_target_batch.py 4 0.004 > gpurun_out/ncu_batch_all.out 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_batch_all.out
python profiles/summarize_ncu.py full gpurun_out/batch_all.ncu-rep > gpurun_out/batch_all_summary.txt 2>&1; grep -c . gpurun_out/batch_all_summary.txt
python profiles/summarize_ncu.py traffic gpurun_out/batch_all.ncu-rep 4 > gpurun_out/ncu_traffic_batch.json 2> gpurun_out/ncu_traffic_batch.err; cat gpurun_out/ncu_traffic_batch.json | head -60
