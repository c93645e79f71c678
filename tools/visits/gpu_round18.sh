#!/bin/bash
# GPU-box visit 18 (8 GPUs): the all-reduce of the flat gradient buffer alone, under a few NCCL settings.
set -u
mkdir -p gpurun_out
: > gpurun_out/allreduce_probe.jsonl
port=29550
probe() {
  port=$((port+1))
  ( env "$@" timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
      tools/allreduce_probe.py >> gpurun_out/allreduce_probe.jsonl 2>> gpurun_out/allreduce_probe.err ) || echo "{\"failed\": \"$*\"}" >> gpurun_out/allreduce_probe.jsonl
}
probe NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING,ENV
probe NCCL_DEBUG=WARN NCCL_ALGO=Ring
probe NCCL_DEBUG=WARN NCCL_ALGO=NVLS
probe NCCL_DEBUG=WARN NCCL_MIN_NCHANNELS=32
probe NCCL_DEBUG=WARN NCCL_ALGO=Tree
grep -i -E "nvls|channels|algo|Using|P2P|multicast" gpurun_out/allreduce_probe.err | head -60 > gpurun_out/allreduce_probe_info.txt
tail -c 3000 gpurun_out/allreduce_probe.err > gpurun_out/allreduce_probe_err_tail.txt
rm -f gpurun_out/allreduce_probe.err
ls -la gpurun_out
