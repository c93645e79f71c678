#!/usr/bin/env python
"""Turns ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python profiles/summarize_ncu.py full     gpurun_out/prof.ncu-rep   > profiles/rNN_<kernel>_ncu.txt
"""
import csv
import os
import subprocess
import sys
from collections import Counter, OrderedDict

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
    "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_membar",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[ix["Metric Unit"]]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = r[ix["Kernel Name"]].split("(")[0][-70:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms of kernel time "
          "(ncu: cold-cache, serialised -- compare SHARES, not absolutes)")
    print(f"{'share':>7} {'total_us':>10} {'calls':>6} {'avg_us':>9}  kernel")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / tot * 100:6.2f}% {ns / 1e3:10.1f} {n:6d} {ns / n / 1e3:9.2f}  {name}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("=" * 100)
        print("kernel:", r[ix["Kernel Name"]][:150])
        for m in METRICS:
            if m in ix:
                print(f"  {m:75s} {r[ix[m]]:>18s} {units[ix[m]]}")
    sass = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(sass.splitlines()))
    kern, data = None, OrderedDict()
    for r in rows:
        if r and r[0] == "Kernel Name":
            kern = r[1][:60]
            data[kern] = []
        elif r and r[0] == "Address":
            continue
        elif kern and len(r) > 8:
            data[kern].append(r)
    for k, v in data.items():
        tot = sum(int(x[5]) for x in v)
        print("=" * 100)
        print(f"SASS regions of {k}: {tot} warp instructions; consecutive instructions with equal execution count")
        i = 0
        while i < len(v):
            n = int(v[i][5])
            j, ops = i, []
            while j < len(v) and int(v[j][5]) == n:
                t = v[j][1].strip().split()
                ops.append(t[1] if t[0].startswith("@") and len(t) > 1 else t[0])
                j += 1
            if n * (j - i) > tot * 0.003:
                print(f"  [{i:4d}-{j - 1:4d}] exec={n / 1e6:8.2f}M x {j - i:3d} instr = {n * (j - i) / tot * 100:5.1f}%  "
                      f"avg_threads={v[i][8]:>4}  {dict(Counter(ops).most_common(8))}")
            i = j


STAGE_OF = {"preprocess_fwd_kernel": "preprocess", "emit_count_kernel": "emit", "emit_scan_kernel": "emit",
            "emit_fill_kernel": "emit", "tile_ranges_kernel": "ranges", "blend_fwd_kernel": "blend_fwd",
            "blend_bwd_kernel": "blend_bwd", "preprocess_bwd_kernel": "preprocess_bwd",
            "preprocess_bwd_batch_kernel": "preprocess_bwd", "pose_finalize_kernel": "preprocess_bwd"}


def traffic(path, views_per_launch=1):
    """(`views_per_launch` > 1: a capture of the batched launch set, tools/ncu_target_batch.py -- every launch covers
    that many views.)
    Per-stage DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per view from an `ncu --set full` capture
    of whole views (tests/harness/ncu_target.py); the sort kernels are attributed to the depth sort until the emit kernels of
    the view have run and to the tile sort afterwards.  Prints the JSON that bench.py reads (profiles/ncu_traffic.json)."""
    import json
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def num(r, m):
        v = float(r[ix[m]].replace(",", ""))
        u = units[ix[m]]
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6,
                    "s": 1e3}.get(u, 1)

    stages, views, sort_stage, order_stage = OrderedDict(), 0, "depth_sort", "blend_fwd"
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
        if name == "preprocess_fwd_kernel":
            views += int(views_per_launch)
            sort_stage = "depth_sort"
            order_stage = "blend_fwd"
        if name in ("sort_hist_kernel", "onesweep_pass_kernel"):
            st = sort_stage
        elif name == "tile_order_kernel":      # first launch of a view: forward order, second: backward order
            st = order_stage
            order_stage = "blend_bwd"
        elif name in STAGE_OF:
            st = STAGE_OF[name]
        else:
            continue
        if name == "emit_fill_kernel":
            sort_stage = "tile_sort"
        a = stages.setdefault(st, {"dram_bytes": 0.0, "ms": 0.0, "launches": 0})
        a["dram_bytes"] += num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
        a["ms"] += num(r, "gpu__time_duration.sum")
        a["launches"] += 1
    views = max(views, 1)
    out = {"_source": os.path.basename(path), "_views": views,
           "_views_per_launch": int(views_per_launch),
           "_note": "per view: sums over the kernels of a stage divided by the number of captured views; ncu times are "
                    "serialised and cold-cache"}
    for st, a in stages.items():
        out[st] = {"dram_bytes_per_launch": round(a["dram_bytes"] / views), "ncu_ms_per_view": round(a["ms"] / views, 4),
                   "kernel_launches_per_view": a["launches"] / views}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
