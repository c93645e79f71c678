#!/bin/bash
# GPU-box visit 17 (2 GPUs): balanced view assignment at N=2 (and the round-robin one for comparison).
set -u
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --no-cpu-baseline "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; echo "exit $?" >> gpurun_out/$2.err; }
run 29541 d2_balanced
run 29542 d2_roundrobin --no-balance --no-e2e
ls -la gpurun_out
