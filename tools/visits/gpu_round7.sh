#!/bin/bash
# GPU-box visit 7 (session 8): parity suite after the lean onesweep pass / SSIM window fix, look-back 4 vs 8,
# bench with the fused and the torch L1 loss in the e2e leg (GPU timelines of both).
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python tools/stage_times.py --ppt 2:2 --tune 3=8,3=4,3=8,3=4 > gpurun_out/lookback_sweep.jsonl 2> gpurun_out/lookback_sweep.err
python bench.py --trace gpurun_out/trace_fused.json > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python bench.py --torch-loss --no-cpu-baseline --trace gpurun_out/trace_torch.json > gpurun_out/bench_ours_torchloss.json 2> gpurun_out/bench_ours_torchloss.err
ls -la gpurun_out
