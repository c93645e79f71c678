#!/usr/bin/env python
"""How evenly is the blend work spread over the tiles, per bench camera?

For each camera: tile instances R, the distribution of the per-tile list lengths (forward work) and of the per-tile
maximum n_contrib (backward work), the measured blend kernel times, and a list-scheduling simulation of the CTAs over
148 SMs x `slots` resident CTAs in (a) launch order (blockIdx) and (b) longest-first order: makespan / ideal.
Development tool: one JSON line per camera."""
import heapq
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bags_capi as capi  # noqa: E402
import bags_synth as S  # noqa: E402
import parity_common as pc  # noqa: E402


def makespan(costs, slots):
    h = [0.0] * slots
    heapq.heapify(h)
    for c in costs:
        heapq.heappush(h, heapq.heappop(h) + c)
    return max(h)


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.004
    ncam = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    c = S.CONFIGS["C3"]
    W, H, P = c["W"], c["H"], c["P"]
    dev = torch.device("cuda:0")
    mod = pc.load_ours()
    cam0 = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam0, scale=scale)
    L = capi.lib()
    for v in range(ncam):
        cam = S.make_camera(W, H, seed=100 + v)
        up = S.make_upstream(H, W, seed=1 + v)
        pc.run_view(mod, cam, gs, up, dev)                       # warm
        L.bags_profile_enable(1)
        capi.profile_read()
        o = pc.run_view(mod, cam, gs, up, dev)
        torch.cuda.synchronize()
        ms, _ = capi.profile_read()
        L.bags_profile_enable(0)
        st = o["state"]
        rng = st["ranges"].to(torch.int64)
        length = (rng[:, 1] - rng[:, 0]).float()
        gx, gy = (W + 15) // 16, (H + 15) // 16
        nc = st["n_contrib"].to(torch.int64).view(H, W)
        pad = torch.zeros(gy * 16, gx * 16, dtype=torch.int64)
        pad[:H, :W] = nc
        bmax = pad.view(gy, 16, gx, 16).permute(0, 2, 1, 3).reshape(gy * gx, 256).max(1).values.float()
        out = {"camera": 100 + v, "R": int(o["num_rendered"]), "visible": int((o["radii"] > 0).sum()),
               "blend_fwd_ms": round(ms["blend_fwd"], 4), "blend_bwd_ms": round(ms["blend_bwd"], 4),
               "fwd_ns_per_kinst": round(ms["blend_fwd"] * 1e6 / max(1, o["num_rendered"]) * 1e3, 1),
               "bwd_ns_per_kinst": round(ms["blend_bwd"] * 1e6 / max(1, o["num_rendered"]) * 1e3, 1)}
        for name, cost in (("list_length", length), ("max_n_contrib", bmax)):
            cl = cost.tolist()
            tot = sum(cl)
            q = torch.quantile(cost, torch.tensor([0.5, 0.9, 0.99]))
            rec = {"mean": round(tot / len(cl), 1), "p50": float(q[0]), "p90": float(q[1]), "p99": float(q[2]),
                   "max": float(cost.max()), "empty_tiles": int((cost == 0).sum())}
            for slots in (148 * 4, 148 * 7):
                ideal = tot / slots
                rec[f"launch_order_{slots}"] = round(makespan([x + 20 for x in cl], slots) / (ideal + 20 * len(cl) / slots), 3)
                rec[f"longest_first_{slots}"] = round(makespan(sorted((x + 20 for x in cl), reverse=True), slots) /
                                                      (ideal + 20 * len(cl) / slots), 3)
            out[name] = rec
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
