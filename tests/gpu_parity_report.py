"""Stand-alone GPU report: ours vs the reference build on several seeded scenes (run under gpurun).
Writes gpurun_out/parity_report.json.  The pytest parity tests use the same helpers."""
import json
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_common as pc  # noqa: E402
S = pc.S


def sort_check(dev):
    import ctypes
    import bags_capi as capi
    L = capi.lib()
    res = {}
    for n, (b0, b1) in [(1, (0, 32)), (1000, (0, 32)), (8192, (0, 13)), (8193, (0, 32)), (300001, (0, 32)),
                        (2_000_000, (0, 16)), (5_000_000, (0, 13))]:
        g = torch.Generator().manual_seed(n)
        keys = torch.randint(0, 2**31 - 1, (n,), generator=g, dtype=torch.int64)
        if n > 1000:
            keys[: n // 2] = keys[: n // 2] & 0xFFF           # many ties -> stability matters
        keys = keys.to(torch.int32).to(dev)
        vals = torch.arange(n, dtype=torch.int32, device=dev)
        ko, vo = torch.empty_like(keys), torch.empty_like(vals)
        scratch = torch.empty(int(L.bags_sort_scratch_bytes(n)), dtype=torch.uint8, device=dev)
        kin, vin = keys.clone(), vals.clone()
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        rc = L.bags_sort_pairs_u32(n, p(kin), p(vin), p(ko), p(vo), b0, b1, p(scratch), None)
        torch.cuda.synchronize()
        mask = (1 << (b1 - b0)) - 1
        kk = (keys.to(torch.int64) >> b0) & mask
        ref_order = torch.sort(kk, stable=True).indices
        ok_v = bool((vo.to(torch.int64) == ref_order).all())
        ok_k = bool((ko == keys[ref_order]).all())
        res[f"n={n},bits=[{b0},{b1})"] = (rc, ok_k, ok_v)
        print("  sort", n, b0, b1, rc, ok_k, ok_v)
    return res


def main():
    dev = torch.device("cuda:0")
    report = {}
    print("torch", torch.__version__, torch.cuda.get_device_name(0))
    try:
        report["sort"] = sort_check(dev)
    except Exception:
        traceback.print_exc()
        report["sort"] = "EXC"
    ours = pc.load_ours()
    ref = pc.load_ref() if pc.ref_available() else None
    cases = [
        dict(name="tiny", P=3000, W=160, H=120, scale=0.02),
        dict(name="small_bg_shift", P=20000, W=400, H=304, scale=0.01, bg=(1.0, 1.0, 1.0), shift=(1e-3, -5e-4, 2e-4)),
        dict(name="C1", P=100000, W=800, H=800, scale=0.004),
        dict(name="C1_deg1", P=100000, W=800, H=800, scale=0.004, sh_degree=1),
        dict(name="ragged", P=50000, W=1297 // 2, H=840 // 2, scale=0.006),
    ]
    for c in cases:
        print("== case", c)
        try:
            cam = S.make_camera(c["W"], c["H"])
            gs = S.make_gaussians(c["P"], cam, scale=c["scale"])
            up = S.make_upstream(c["H"], c["W"])
            kw = dict(sh_degree=c.get("sh_degree", 3), bg=c.get("bg", (0, 0, 0)), shift=c.get("shift", (0, 0, 0)))
            t0 = time.time()
            o = pc.run_view(ours, cam, gs, up, dev, **kw)
            torch.cuda.synchronize()
            print("  ours done", time.time() - t0, "R =", o["num_rendered"], "visible =", int((o["radii"] > 0).sum()))
            if ref is not None:
                r = pc.run_view(ref, cam, gs, up, dev, **kw)
                torch.cuda.synchronize()
                m = pc.compare(o, r)
                gv_o, gv_r = o["grads"]["viewmatrix"], r["grads"]["viewmatrix"]
                print("    viewmatrix grad ours:", gv_o.flatten().tolist())
                print("    viewmatrix grad ref :", gv_r.flatten().tolist())
                report[c["name"]] = m
        except Exception:
            traceback.print_exc()
            report[c["name"]] = "EXC"
    os.makedirs(os.path.join(pc.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(pc.ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(report, f, indent=1, default=str)


if __name__ == "__main__":
    main()
