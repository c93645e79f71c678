#!/bin/bash
# GPU-box visit 11 (2 GPUs): why do the device-leg regions slow down at N=2?  default / one stream / host sync per step.
set -u
mkdir -p gpurun_out
env | grep -i nccl > gpurun_out/nccl_env.txt 2>&1
nproc > gpurun_out/nproc.txt
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 4 --no-e2e --no-cpu-baseline "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; }
run 29521 d2_default
run 29522 d2_streams1 --streams 1
run 29523 d2_stepsync --step-sync
run 29524 d2_nosink --no-grad-sinks
ls -la gpurun_out
