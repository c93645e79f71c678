#!/bin/bash
# Round 2, GPU visit 18 (2 GPUs): the driver's exact N = 2 command line after the warm-up fix (auto -> peer-memory all-reduce).
set -u
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/g2_full.json 2> gpurun_out/g2_full.err; echo "N=2 full rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/g2_full.json')); print('N=2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['implementation']['allreduce'][:70], d['implementation'].get('per_rank_compute_ms_per_step'))" || (grep -a "bench rank 0\|Error" gpurun_out/g2_full.err | tail -5)
