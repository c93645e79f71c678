#!/bin/bash
# GPU-box visit 10 (2 GPUs): the N>1 launch of the bench contract (torchrun, NCCL all-reduce of the flat gradients).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "exit $?" >> gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
echo "exit $?" >> gpurun_out/bench_2gpu_ref.err
ls -la gpurun_out
