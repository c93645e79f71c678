#!/bin/bash
# Round 2, GPU visit 7 (2 GPUs): default bench (batch mode, step-end all-reduce) and --overlap, after the teardown fixes.
set -u
mkdir -p gpurun_out
for m in "" "--overlap"; do
  tag="default$(echo $m | tr -d ' -')"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 $m --no-e2e > gpurun_out/d2_$tag.json 2> gpurun_out/d2_$tag.err; echo "$tag rc=$?"
  grep -a "bench rank 0" gpurun_out/d2_$tag.err | tail -2
  python -c "
import json
d=json.load(open('gpurun_out/d2_$tag.json')); print('$tag', round(d['value'],1), d['ms_per_step'], d['implementation']['allreduce'], d['implementation']['view_assignment'])" || tail -5 gpurun_out/d2_$tag.err
done
