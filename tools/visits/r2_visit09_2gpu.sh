#!/bin/bash
# Round 2, GPU visit 9 (2 GPUs): the exact command line the driver uses at N = 2 (all legs), then the N = 1 line.
set -u
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/d2_full.json 2> gpurun_out/d2_full.err; echo "N=2 full rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/d2_full.json')); print('N=2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['implementation']['view_assignment'])" || (grep -a "bench rank 0\|Error" gpurun_out/d2_full.err | tail -5)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 --no-drop-in > gpurun_out/d2_ref.json 2> gpurun_out/d2_ref.err; echo "N=2 ref rc=$?"; cut -c1-200 gpurun_out/d2_ref.json
