#!/bin/bash
# GPU-box visit 12 (session 8): parity suite incl. the fused Adam, Adam step timing, bench lines of C1/C2/C4/C5 + knn.
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=5 --tb=short ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python tools/bench_adam.py > gpurun_out/adam.json 2> gpurun_out/adam.err
python tools/bench_adam.py --P 3000000 --iters 20 >> gpurun_out/adam.json 2>> gpurun_out/adam.err
timeout 600 python tests/harness/bench_configs.py --out gpurun_out/configs.jsonl --steps 10 > gpurun_out/configs.log 2>&1
ls -la gpurun_out
