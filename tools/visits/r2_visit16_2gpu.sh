#!/bin/bash
# Round 2, GPU visit 16 (2 GPUs): the peer-memory all-reduce kernel vs NCCL (correctness + time).
set -u
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tools/allreduce_p2p_probe.py > gpurun_out/p2p_probe_2gpu.jsonl 2> gpurun_out/p2p_probe_2gpu.err; echo "probe rc=$?"; cat gpurun_out/p2p_probe_2gpu.jsonl; grep -a "Error\|error" gpurun_out/p2p_probe_2gpu.err | head -8
