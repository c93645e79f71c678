#!/usr/bin/env python
"""How fast is the one collective of the path?  all_reduce(SUM) of the flat gradient buffer (59 floats x 1 M Gaussians
= 236 MB fp32) over NCCL, as bags_view_parallel.allreduce_gradients issues it.  Launch under torchrun; rank 0 prints
one JSON line.  Development tool (NCCL_* environment variables are taken from the caller)."""
import json
import os
import sys

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    n = 59 * int(sys.argv[1]) if len(sys.argv) > 1 else 59_000_000
    buf = torch.ones(n, dtype=torch.float32, device=dev)
    for _ in range(5):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gb = n * 4 / 1e9
        print(json.dumps({"world": world, "bytes": n * 4, "ms": round(float(t), 4),
                          "algbw_GBps": round(gb / (float(t) * 1e-3), 1),
                          "busbw_GBps": round(gb / (float(t) * 1e-3) * 2 * (world - 1) / world, 1),
                          "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_") and "PACKAGE" not in k}}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
