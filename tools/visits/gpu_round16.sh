#!/bin/bash
# GPU-box visit 16 (8 GPUs): the N=8 launch of the bench contract, once.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus8.txt 2>&1
nproc >> gpurun_out/gpus8.txt
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 8 --steps 20 --warmup 5 --repeats 3 --no-cpu-baseline > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
echo "exit $?" >> gpurun_out/bench_8gpu.err
ls -la gpurun_out
