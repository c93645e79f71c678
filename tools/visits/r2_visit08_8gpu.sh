#!/bin/bash
# Round 2, GPU visit 8 (8 GPUs): default bench at N = 8 and N = 4, --overlap at N = 8.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
run() {  # n tag extra...
  n=$1; tag=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $n "$@" --no-e2e > gpurun_out/d${n}_$tag.json 2> gpurun_out/d${n}_$tag.err; echo "N=$n $tag rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/d${n}_$tag.json')); print('N=$n $tag', round(d['value'],1), d['ms_per_step'], d['implementation']['allreduce'], d['implementation']['view_assignment'])" || (grep -a "bench rank 0\|Error" gpurun_out/d${n}_$tag.err | tail -5)
}
run 8 default
run 8 overlap --overlap
run 4 default
