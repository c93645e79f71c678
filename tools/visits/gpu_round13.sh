#!/bin/bash
# GPU-box visit 13 (session 8): new tests (l1_u8), preprocess_bwd occupancy knob, e2e with the one-kernel u8 loss.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 --tb=short -k "l1_u8 or adam or view_streams or sinks" ) > gpurun_out/pytest_gpu_sel.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_sel.log
python tools/stage_times.py --ppt 2:2 --tune 6=5,6=6,6=5,6=6 > gpurun_out/pbwd_sweep.jsonl 2> gpurun_out/pbwd_sweep.err
python bench.py --fused-loss --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
python bench.py --no-cpu-baseline > gpurun_out/bench_torch.json 2> gpurun_out/bench_torch.err
python bench.py --fused-loss --no-cpu-baseline --streams 3 > gpurun_out/bench_fused_s3.json 2> gpurun_out/bench_fused_s3.err
ls -la gpurun_out
