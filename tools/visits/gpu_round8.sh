#!/bin/bash
# GPU-box visit 8 (session 8): parity suite (new: view pipelining over two streams), bench matrix
# {1, 2 streams} x {fused, torch L1 loss in the e2e leg}, reference arm.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py --streams 2 > gpurun_out/bench_s2_fused.json 2> gpurun_out/bench_s2_fused.err
python bench.py --streams 1 --no-cpu-baseline > gpurun_out/bench_s1_fused.json 2> gpurun_out/bench_s1_fused.err
python bench.py --streams 2 --torch-loss --no-cpu-baseline > gpurun_out/bench_s2_torch.json 2> gpurun_out/bench_s2_torch.err
python bench.py --streams 1 --torch-loss --no-cpu-baseline > gpurun_out/bench_s1_torch.json 2> gpurun_out/bench_s1_torch.err
python bench.py --streams 3 --no-cpu-baseline > gpurun_out/bench_s3_fused.json 2> gpurun_out/bench_s3_fused.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ls -la gpurun_out
