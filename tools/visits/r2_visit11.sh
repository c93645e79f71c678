#!/bin/bash
# Round 2, GPU visit 11: ncu launch list of the default bench command (batch mode), ncu --set full of every kernel of the
# batched launch set, compute-sanitizer on the batch tests.
set -u
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file gpurun_out/launches_batch.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-drop-in --no-cpu-baseline --repeats 1 --min-seconds 0 > gpurun_out/launches_batch_run.log 2>&1; echo "launch list rc=$?"
python profiles/summarize_ncu.py launches gpurun_out/launches_batch.csv > gpurun_out/launches_batch.txt 2>&1; head -30 gpurun_out/launches_batch.txt
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bags -o gpurun_out/batch_all python tools/ncu_target_batch.py 4 0.004 > gpurun_out/ncu_batch_all.out 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_batch_all.out
python profiles/summarize_ncu.py full gpurun_out/batch_all.ncu-rep > gpurun_out/batch_all_summary.txt 2>&1; grep -c . gpurun_out/batch_all_summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_batch_gpu.py -q -x -k "bit_identical_per_view[3] or sum_of_single_views[4-3] or sinks or overflow" > gpurun_out/sanitizer_memcheck_batch.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_batch.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_batch_gpu.py -q -x -k "sum_of_single_views[3-1]" > gpurun_out/sanitizer_racecheck_batch.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_batch.log
