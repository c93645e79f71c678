#!/bin/bash
# Round 2, GPU visit 14 (8 GPUs): the driver's scaling run after all kernel work -- N = 8 default and --overlap, N = 4.
set -u
mkdir -p gpurun_out
run() {  # n tag extra...
  n=$1; tag=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $n "$@" > gpurun_out/e${n}_$tag.json 2> gpurun_out/e${n}_$tag.err; echo "N=$n $tag rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/e${n}_$tag.json')); print('N=$n $tag', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],1), d['implementation'].get('per_rank_compute_ms_per_step'), d['implementation']['view_assignment'][-60:])" || (grep -a "bench rank 0\|Error" gpurun_out/e${n}_$tag.err | tail -5)
}
run 8 default
run 8 overlap --overlap --no-e2e
run 4 default --no-e2e
