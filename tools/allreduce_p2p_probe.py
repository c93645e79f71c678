#!/usr/bin/env python
"""The gradient exchange as one kernel over NVLink peer memory (bags_allreduce_p2p, P2PGradientReducer) against NCCL's
all_reduce on the same 59 x P float buffer: correctness (vs NCCL to fp32 summation order, bit-identical across ranks) and
time.  Launch under torchrun; rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bags_view_parallel as VP  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    g = torch.Generator().manual_seed(0)
    z = lambda *s: torch.zeros(*s)
    params = VP.FlatGaussianParams(z(P, 3), z(P, 16, 3), z(P, 1), z(P, 3), z(P, 4), device=dev)
    red = VP.P2PGradientReducer(params)
    out = {"world": world, "P": P, "bytes": params.flat_grad.numel() * 4, "p2p": red.describe()}

    def fill(seed):
        gen = torch.Generator(device=dev).manual_seed(1000 * seed + rank)
        params.flat_grad.copy_(torch.randn(params.flat_grad.shape, generator=gen, device=dev))

    # ---- correctness: three rounds with fresh data (exercises the epoch flags) ---------------------------------------------
    worst, identical = 0.0, True
    for rnd in range(3):
        fill(rnd)
        want = params.flat_grad.clone()
        dist.all_reduce(want)
        red.reduce()
        torch.cuda.synchronize()
        got = params.flat_grad
        worst = max(worst, float((got - want).norm() / want.norm()))
        ref0 = got.clone()
        dist.broadcast(ref0, 0)
        identical &= bool(torch.equal(ref0, got))
    flag = torch.tensor([1 if identical else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["rel_l2_vs_nccl"] = worst
    out["bit_identical_across_ranks"] = bool(int(flag))

    def timed(fn, iters=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    fill(9)
    out["nccl_ms"] = round(timed(lambda: dist.all_reduce(params.flat_grad)), 4)
    fill(9)
    out["p2p_ms"] = round(timed(red.reduce), 4)
    gb = out["bytes"] / 1e9
    for k in ("nccl", "p2p"):
        out[k + "_busbw_GBps"] = round(gb / (out[k + "_ms"] * 1e-3) * 2 * (world - 1) / world, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
