#!/bin/bash
# Round 2, GPU visit 5 (2 GPUs): NCCL parity of the per-view and the batched + sliced-reduce paths against the reference,
# 2-GPU bench lines: batch (overlapped slices) / batch --no-overlap.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
echo "(parity worker: see the previous run)"
for m in "batch" "batch --no-overlap"; do
  tag=$(echo $m | tr -d ' -')
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --mode $m --no-e2e > gpurun_out/d2_$tag.json 2> gpurun_out/d2_$tag.err; echo "$tag rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/d2_$tag.json')); print('$tag', round(d['value'],1), d['ms_per_step'], d['implementation']['allreduce'])" || tail -5 gpurun_out/d2_$tag.err
done
