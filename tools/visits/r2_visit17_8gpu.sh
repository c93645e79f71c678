#!/bin/bash
# Round 2, GPU visit 17 (8 GPUs): peer-memory all-reduce kernel vs NCCL (probe), NCCL parity worker at 4 ranks (now through
# the peer-memory reducer), bench at N = 8 with both exchanges.
set -u
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 tools/allreduce_p2p_probe.py > gpurun_out/p2p_probe_8gpu.jsonl 2> gpurun_out/p2p_probe_8gpu.err; echo "probe rc=$?"; cat gpurun_out/p2p_probe_8gpu.jsonl
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 tests/nccl_parity_worker.py > gpurun_out/nccl_parity_4gpu.log 2> gpurun_out/nccl_parity_4gpu.err; echo "parity rc=$?"; grep -a '^{' gpurun_out/nccl_parity_4gpu.log | tail -1 | cut -c1-300
run() {  # n tag extra...
  n=$1; tag=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $n "$@" > gpurun_out/f${n}_$tag.json 2> gpurun_out/f${n}_$tag.err; echo "N=$n $tag rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/f${n}_$tag.json')); print('N=$n $tag', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],1), d['implementation']['allreduce'][:60])" || (grep -a "bench rank 0\|Error" gpurun_out/f${n}_$tag.err | tail -5)
}
run 8 p2p
run 8 nccl --allreduce nccl --no-e2e
