#!/bin/bash
# Round 2, GPU visit 1: the compiled extension + asynchronous forward + captured step: full GPU test suite, bench both
# arms, compute-sanitizer on the smoke, CUB head-to-head, launch list of the reference arm.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"; tail -3 gpurun_out/bench_ours.err
python bench.py --mode streams --no-drop-in --no-cpu-baseline > gpurun_out/bench_ours_streams.json 2> gpurun_out/bench_ours_streams.err; echo "streams rc=$?"
tools/_bin/sort_h2h > gpurun_out/sort_h2h.jsonl 2>&1; echo "h2h rc=$?"
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/sanitizer_$tool.log python __graft_entry__.py smoke > gpurun_out/sanitizer_$tool.out 2>&1; echo "$tool rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ref.csv python tests/harness/ncu_target.py 1 0.004 ref > gpurun_out/ncu_ref.out 2>&1; echo "ncu ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ours.csv python tests/harness/ncu_target.py 1 0.004 ours > gpurun_out/ncu_ours.out 2>&1; echo "ncu ours rc=$?"
cat gpurun_out/bench_ours.json | head -c 3000
ls -la gpurun_out
