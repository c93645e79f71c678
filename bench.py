#!/usr/bin/env python
"""bench.py -- fwd+bwd views/s of the B200-native rasterizer (and of the reference arm) on synthetic scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3] [--views-per-gpu 4]

Contract (see DESIGN.md "Measurement"): a step = one training iteration over `views_per_gpu` views per GPU
(weak scaling): every view is rendered forward and back-propagated through the public GaussianRasterizer API, the
Gaussian-parameter gradients accumulate in one flat buffer and (N>1) are summed with NCCL.
`value` = views/s with the upstream image gradients resident in HBM; `e2e` = the same through host buffers: each
view's 8-bit ground-truth image comes from pinned host memory (H2D on a copy stream, overlapping the previous step),
the L1 loss is computed on the device and read back (D2H) inside the timed region.

Our arm (default `--mode graph`): the step's enqueue is captured once in a CUDA graph (bags_view_parallel.CapturedStep
over the asynchronous forward, diff_gaussian_rasterization.AsyncForward) and replayed with one launch per step.
`--mode streams` issues the same work eagerly (blocking forward, two streams).  `drop_in` is the record for the
unchanged caller: ONE view per step through the reference's API exactly as train.py uses it -- blocking forward, one
stream, gradients accumulated by autograd -- for both arms.

Timing: >= W warm-up steps, then regions of exactly K steps (barrier + synchronize on both sides, CUDA events, max over
ranks), at least 5 regions / 100 steps / 2 s of them; the MEDIAN region is reported and all are listed.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200")
sys.path.insert(0, PKG)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bags_synth as S  # noqa: E402

SCALE = 0.004     # median Gaussian scale of the synthetic scenes (gives R/P ~ 5 at 1080p)


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(S.CONFIGS))
    ap.add_argument("--views-per-gpu", type=int, default=4)
    ap.add_argument("--P", type=int, default=0, help="override the Gaussian count of the config (C5 sweep)")
    ap.add_argument("--scale", type=float, default=SCALE,
                    help="median Gaussian scale; 0.012 is the blend-heavy scene (R/P ~ 50 at 1080p)")
    ap.add_argument("--repeats", type=int, default=5, help="minimum number of timed regions of --steps steps")
    ap.add_argument("--min-seconds", type=float, default=2.0, help="minimum total time of the timed regions of a leg")
    ap.add_argument("--mode", default="batch", choices=["batch", "graph", "streams"],
                    help="our arm: batch = the rank's views as ONE launch set (GaussianBatchRasterizer), the step "
                         "captured in a CUDA graph (default); graph = one asynchronous forward/backward per view, the "
                         "step captured in a CUDA graph; streams = eager launches, blocking forward, views pipelined "
                         "over --streams CUDA streams")
    ap.add_argument("--bwd-slices", type=int, default=4,
                    help="batch mode, N > 1: slices of the batched backward whose gradient all-reduce overlaps the "
                         "next slice (bags_view_parallel.SlicedGradientReducer)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-drop-in", action="store_true")
    ap.add_argument("--no-grad-sinks", action="store_true",
                    help="let autograd accumulate the Gaussian gradients (fresh tensors + adds) instead of the "
                         "kernel adding into the flat gradient buffer")
    ap.add_argument("--streams", type=int, default=2,
                    help="CUDA streams the views of a step are pipelined over (bags_view_parallel.ViewStreams): view "
                         "k's preprocess/binning/forward overlaps view k-1's backward; 1 = everything on one stream")
    ap.add_argument("--no-balance", action="store_true",
                    help="N > 1: keep the round-robin view assignment instead of balancing it by tile instances")
    ap.add_argument("--overlap", action="store_true",
                    help="N > 1, batch mode: exchange the gradients slice by slice inside the batched backward "
                         "(bags_view_parallel.SlicedGradientReducer, NCCL stream at high priority) instead of one "
                         "all-reduce after the step")
    ap.add_argument("--no-overlap", action="store_true", help=argparse.SUPPRESS)     # (the default; kept for old scripts)
    ap.add_argument("--allreduce", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N > 1: the step-end gradient sum -- p2p = one kernel over NVLink peer memory "
                         "(bags_allreduce_p2p, falls back to NCCL where the peers' buffers cannot be mapped); nccl = "
                         "torch.distributed.all_reduce; auto (default) = p2p on 2 GPUs (measured 0.38 vs 0.46 ms for the "
                         "236 MB buffer), nccl on more (8 GPUs: 0.71 vs 0.64 ms)")
    ap.add_argument("--torch-loss", action="store_true",
                    help="e2e leg of our arm: the torch operator sequence of utils/loss_utils.py:18-19 (what the "
                         "reference arm always runs) instead of bags_loss.l1_loss_u8")
    ap.add_argument("--tune", action="append", default=[], metavar="KNOB=VALUE",
                    help="development: bags_set_tuning knob (BLEND_FWD_PPT, BLEND_BWD_PPT, BLEND_CULL, SORT_LOOKBACK, "
                         "SORT_SMALL_TILES, SORT_HWMATCH, PBWD_MIN_BLOCKS, BLEND_STAGING) before anything runs")
    return ap.parse_args(argv)


TUNE_KNOBS = {"BLEND_FWD_PPT": 0, "BLEND_BWD_PPT": 1, "BLEND_CULL": 2, "SORT_LOOKBACK": 3, "SORT_SMALL_TILES": 4,
              "SORT_HWMATCH": 5, "PBWD_MIN_BLOCKS": 6, "BLEND_STAGING": 7, "BLEND_BWD_REDUCE": 8}


def load_impl(name):
    if name == "ours":
        import diff_gaussian_rasterization as m
        return m
    import importlib.util
    d = os.path.join(ROOT, "oracle", "_ref", "diff_gaussian_rasterization")
    if not os.path.exists(os.path.join(d, "_C.so")):
        return None
    spec = importlib.util.spec_from_file_location("ref_dgr", os.path.join(d, "__init__.py"),
                                                  submodule_search_locations=[d])
    m = importlib.util.module_from_spec(spec)
    sys.modules["ref_dgr"] = m
    spec.loader.exec_module(m)
    return m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 500 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "500", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_first_sample(self, timeout=15.0):
        """nvidia-smi takes a second or two to initialise NVML (and perturbs kernel launches while it does): do not
        start timing before it has printed its first line."""
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.f.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.05)

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(P, V, R, N, Tt, K=16):
    """SURVEY.md section 8d stage formulas (bytes per view)."""
    passes = 6
    return {
        "preprocess": 20 * P + (107 + 12 * K) * V,
        "scan": 8 * P,
        "duplicate": 4 * P + 20 * V + 12 * R,
        "sort": (8 + 24 * passes) * R,
        "ranges": 8 * R + 8 * Tt,
        "blend_fwd": 44 * R + 28 * N,
        "blend_bwd": 44 * R + 28 * N + 52 * R,
        "preprocess_bwd": 4 * P + (111 + 12 * K) * V + 296 * P,
    }


def median(xs):
    s = sorted(xs)
    n = len(s)
    return s[n // 2] if n % 2 else 0.5 * (s[n // 2 - 1] + s[n // 2])


def cpu_baseline_sample():
    """float64 host oracle (oracle/raster_oracle.py) fwd+bwd on a bounded sample, all host cores."""
    sys.path.insert(0, ROOT)
    from oracle import raster_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    P, W, H = 20000, 400, 400
    cam = S.make_camera(W, H)
    gs = S.make_gaussians(P, cam, scale=0.008)
    up = S.make_upstream(H, W)
    sc = O.Scene(gs.means3D, gs.opacities, cam.viewmatrix, cam.projmatrix, cam.campos, W, H, cam.tanfovx, cam.tanfovy,
                 (0, 0, 0), shs=gs.shs, scales=gs.scales, rotations=gs.rotations, shift_factors=torch.zeros(3),
                 sh_degree=3)
    t0 = time.time()
    n = 0
    while True:
        f = O.rasterize_forward(sc)
        O.rasterize_backward(sc, f, *up)
        n += 1
        if time.time() - t0 > 10.0 or n >= 5:
            break
    dt = time.time() - t0
    return {"value": n / dt, "unit": "views/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"float64 PyTorch host oracle, {P} Gaussians SH3 at {W}x{H}, R={f['num_rendered']}, "
                      f"{n} fwd+bwd views in {dt:.1f}s (NOT the bench workload: a bounded sample)"}


_REAL_STDOUT = None


def guard_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on
    stdout when the image sets NCCL_DEBUG=VERSION, whatever NCCL_DEBUG_FILE says): keep a private duplicate of the
    real stdout for the JSON line and point file descriptor 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    args = parse()
    guard_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ours = args.impl == "ours"
    if not ours and rank != 0:
        return 0
    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: bench.py measures the CUDA path and has no CPU fallback"})
        return 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    use_dist = world > 1 and ours
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its debug output (the image sets NCCL_DEBUG=VERSION: a version banner) to stdout, next to the
        # one JSON line; send it to stderr whatever level the caller chose
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if args.overlap:
            # NCCL's stream gets a high priority: the slice all-reduces that overlap the batched backward must be
            # scheduled as soon as SM slots free up, not after the compute kernel queued behind them has dispatched all
            # its CTAs.  (Only with --overlap: with the plain step-end all-reduce a high-priority NCCL stream was seen
            # to hang the 2-GPU run.)
            os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)

    def progress(what):
        if use_dist:
            print(f"[bench rank {rank}] {what}", file=sys.stderr, flush=True)

    mod = load_impl(args.impl)
    if mod is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref not built (run oracle/build_ref.py where "
                                                  "/root/reference exists)"})
        return 0

    cfg = S.CONFIGS[args.config]
    P, W, H = args.P or cfg["P"], cfg["W"], cfg["H"]
    N = W * H
    vpg = args.views_per_gpu
    nviews = vpg * (world if ours else 1)
    my_views = list(range(rank, nviews, world)) if ours else list(range(nviews))
    batch_mode = ours and args.mode == "batch"
    graph_mode = ours and args.mode in ("graph", "batch")

    # ---- scene: replicated Gaussians (seed 0, placed in front of camera 100), one camera per view (seed 100+v) -------
    cam0 = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam0, seed=0, scale=args.scale)
    import bags_view_parallel as VP
    params = VP.FlatGaussianParams(gs.means3D.to(dev), gs.shs.to(dev), gs.opacities.to(dev), gs.scales.to(dev),
                                   gs.rotations.to(dev), device=dev)
    grad_sinks = ours and not args.no_grad_sinks
    if grad_sinks:
        params.enable_direct_accumulation(mod)     # consulted only inside `with params.accumulating()`
    bg = torch.zeros(3, device=dev)
    shift = torch.zeros(3, device=dev, requires_grad=True)

    def make_view(v):
        cam = S.make_camera(W, H, seed=100 + v)
        g = torch.Generator().manual_seed(1000 + v)
        # ground truth as the dataset stores it: 8-bit RGB in pinned host memory, converted to fp32 on the device
        gt_host = torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8).pin_memory()
        up = [u.to(dev) for u in S.make_upstream(H, W, seed=1 + v)]
        return dict(
            id=v, view=cam.viewmatrix.to(dev).requires_grad_(True), proj=cam.projmatrix.to(dev).requires_grad_(True),
            campos=cam.campos.to(dev).requires_grad_(True), intrinsic=cam.intrinsic.to(dev), cam=cam,
            gt_host=gt_host, gt_dev=[torch.empty(3, H, W, device=dev, dtype=torch.uint8) for _ in range(2)], up=up)

    views = [make_view(v) for v in my_views]

    def render(vw):
        st = mod.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=vw["cam"].tanfovx, tanfovy=vw["cam"].tanfovy, bg=bg,
            scale_modifier=1.0, viewmatrix=vw["view"], projmatrix=vw["proj"], intrinsic=vw["intrinsic"], sh_degree=3,
            campos=vw["campos"], prefiltered=False, debug=False, debug_iter=0)
        rast = mod.GaussianRasterizer(raster_settings=st)
        m2d = torch.zeros_like(params["means3D"], requires_grad=True)
        m2dd = torch.zeros_like(params["means3D"], requires_grad=True)
        return rast(means3D=params["means3D"], means2D=m2d, means2D_densify=m2dd, shift_factors=shift,
                    opacities=params["opacities"], shs=params["shs"], scales=params["scales"],
                    rotations=params["rotations"])

    def clear_pose_grads():
        for vw in views:
            vw["view"].grad = vw["proj"].grad = vw["campos"].grad = None
        shift.grad = None

    # ---- view -> rank assignment balanced by measured work (N > 1) -----------------------------------------------
    # The step time of view-parallel training is the time of the slowest rank, and the tile-instance count of the
    # bench cameras differs by up to 3x.  Every rank renders its round-robin views once, the counts are all-gathered
    # and bags_view_parallel.balance_views deals the same 4N views out again, the same number per rank.
    assignment = "round-robin (rank r: views r, r+N, ...)"
    if use_dist and not args.no_balance:
        # cost of a view = its measured forward + backward time (CUDA events, second of two runs): the tile-instance
        # count is a poor proxy -- the bench cameras with the fewest instances are the slowest (larger footprints)
        costs = []
        for vw in views:
            best = None
            for rep_ in range(5):                               # first run untimed, then the fastest of four
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                c, _, d_, w_, _ = render(vw)
                torch.autograd.backward([c, d_, w_], vw["up"])
                e1.record()
                torch.cuda.synchronize()
                del c, d_, w_
                if rep_ > 0:
                    t = e0.elapsed_time(e1)
                    best = t if best is None else min(best, t)
            costs.append(int(1e3 * best))                       # microseconds
        params.zero_grad()
        clear_pose_grads()
        mine = torch.tensor(costs, dtype=torch.int64, device=dev)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        cost = [0] * nviews
        for r in range(world):
            for k, v in enumerate(range(r, nviews, world)):
                cost[v] = int(every[r][k])
        table = VP.balance_views(cost, world)
        loads = [sum(cost[v] for v in q) for q in table]
        rr = [sum(cost[v] for v in range(r, nviews, world)) for r in range(world)]
        assignment = (f"balanced by measured view time (bags_view_parallel.balance_views): max/mean work "
                      f"{max(loads) / (sum(loads) / world):.3f} (round-robin: {max(rr) / (sum(rr) / world):.3f})")
        if table[rank] != my_views:
            my_views = table[rank]
            views = [make_view(v) for v in my_views]
        torch.cuda.empty_cache()

    # ---- workload statistics: visible Gaussians V and tile instances R of every view of this rank (blocking forward) --
    Vs, Rs = [], []
    with torch.no_grad():
        for vw in views:
            c, radii, d, w_, _ = render(vw)
            Vs.append(int((radii > 0).sum()))
            del c, radii, d, w_
    for vw in views:
        c = render(vw)[0]
        Rs.append(int(c.grad_fn.num_rendered))
        del c
    V = sum(Vs) / len(Vs)
    R = sum(Rs) / len(Rs)
    capacity = int(max(Rs) * 1.25) + (64 << 10)

    pipe = VP.ViewStreams(dev, depth=args.streams if ours else 1)
    pipe_serial = VP.ViewStreams(dev, depth=1)

    l1_loss_u8 = None
    if ours and not args.torch_loss:
        import bags_loss          # l1_loss against the 8-bit ground truth as stored: loss + gradient in one kernel
        l1_loss_u8 = bags_loss.l1_loss_u8

    def l1_loss(network_output, gt):   # utils/loss_utils.py:18-19, the operator sequence the reference runs
        return torch.abs(network_output - gt).mean()

    def view_loss(color, gt_u8):
        if l1_loss_u8 is not None:
            return l1_loss_u8(color, gt_u8)
        return l1_loss(color, gt_u8.float().mul_(1.0 / 255.0))

    # ---- the step, as an enqueue ------------------------------------------------------------------------------------
    def body_device(p=None):
        """all views of the step: forward + backward with the upstream gradients resident in HBM"""
        p = p or pipe
        params.zero_grad()
        with params.accumulating():
            p.begin()
            for k, vw in enumerate(views):
                with p.view(k):
                    color, radii, depth, weights, _ = render(vw)
                    p.before_backward()
                    torch.autograd.backward([color, depth, weights], vw["up"])
                    p.after_backward()
            p.end()

    loss_dev = [torch.zeros(1, device=dev) for _ in range(2)]

    def body_e2e(buf):
        """the same with the L1 loss against the 8-bit ground truth in gt_dev[buf]; the step's loss goes to loss_dev[buf]"""
        params.zero_grad()
        losses = []
        with params.accumulating():
            pipe.begin()
            for k, vw in enumerate(views):
                with pipe.view(k):
                    color, radii, depth, weights, _ = render(vw)
                    loss = view_loss(color, vw["gt_dev"][buf])
                    pipe.before_backward()
                    loss.backward()
                    pipe.after_backward()
                    losses.append(loss.detach())
            pipe.end()
        loss_dev[buf].copy_(torch.stack(losses).sum().reshape(1))

    # ---- batch mode: the rank's views as one launch set ---------------------------------------------------------------
    batch = {}
    if batch_mode:
        nb = len(views)
        batch["up"] = [torch.stack([vw["up"][k] for vw in views]).contiguous() for k in range(3)]
        batch["gt"] = [torch.empty(nb, 3, H, W, device=dev, dtype=torch.uint8) for _ in range(2)]
        for k, vw in enumerate(views):            # the per-view ground-truth buffers become slices of the stacked ones
            vw["gt_dev"] = [batch["gt"][0][k], batch["gt"][1][k]]
        batch["m2d"] = torch.zeros(nb, P, 3, device=dev, requires_grad=True)
        batch["m2dd"] = torch.zeros(nb, P, 3, device=dev, requires_grad=True)
        batch["reducer"] = VP.SlicedGradientReducer(params) if (use_dist and args.overlap) else None
        settings = [mod.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=vw["cam"].tanfovx, tanfovy=vw["cam"].tanfovy, bg=bg,
            scale_modifier=1.0, viewmatrix=vw["view"], projmatrix=vw["proj"], intrinsic=vw["intrinsic"], sh_degree=3,
            campos=vw["campos"], prefiltered=False, debug=False, debug_iter=0) for vw in views]
        red = batch["reducer"]
        # the rank's views are ONE batch per step: its summed gradients replace the flat gradient buffer (sink_overwrite),
        # which therefore needs no zero fill at the start of the step
        batch["rast"] = mod.GaussianBatchRasterizer(settings, capacity,
                                                    backward_slices=args.bwd_slices if red is not None else 1,
                                                    on_slice=red.on_slice if red is not None else None,
                                                    sink_overwrite=grad_sinks)

    def render_batch():
        return batch["rast"](means3D=params["means3D"], means2D=batch["m2d"], means2D_densify=batch["m2dd"],
                             shift_factors=shift, opacities=params["opacities"], shs=params["shs"],
                             scales=params["scales"], rotations=params["rotations"])

    def clear_batch_grads():
        clear_pose_grads()
        if batch:
            batch["m2d"].grad = batch["m2dd"].grad = None

    def body_device_batch():
        if not grad_sinks:
            params.zero_grad()
        with params.accumulating():
            color, radii, depth, weights, _ = render_batch()
            torch.autograd.backward([color, depth, weights], batch["up"])
        if batch["reducer"] is not None:
            batch["reducer"].finish()

    def body_e2e_batch(buf):
        if not grad_sinks:
            params.zero_grad()
        with params.accumulating():
            color, radii, depth, weights, _ = render_batch()
            # sum over the views of each view's mean L1 = (number of views) x the mean over the stacked images
            loss = view_loss(color, batch["gt"][buf]) * float(len(views))
            loss.backward()
        if batch["reducer"] is not None:
            batch["reducer"].finish()
        loss_dev[buf].copy_(loss.detach().reshape(1))

    capi = None
    tuned = {}
    if ours:
        import bags_capi as capi
        for kv in args.tune:
            k, v = kv.split("=")
            capi.check(capi.lib().bags_set_tuning(TUNE_KNOBS[k], int(v)), "bags_set_tuning")
            tuned[k] = int(v)
    launches_per_step = None
    cap_dev = cap_e2e = None
    if batch_mode:
        clear_batch_grads()
        cap_dev = VP.CapturedStep(body_device_batch, dev, mod, None, forwards_per_step=1, reset=clear_batch_grads,
                                  check=batch["rast"].check)
        l0 = capi.lib().bags_launch_count()
        body_device_batch()
        launches_per_step = int(capi.lib().bags_launch_count() - l0)
        torch.cuda.synchronize()
        clear_batch_grads()
    elif graph_mode:
        clear_pose_grads()
        cap_dev = VP.CapturedStep(body_device, dev, mod, capacity, forwards_per_step=len(views),
                                  reset=clear_pose_grads)
        # kernels of ours inside one step (what one replay launches): counted on an eager run of the same enqueue
        with mod.AsyncForward(capacity, slots=len(views)):
            l0 = capi.lib().bags_launch_count()
            body_device()
            launches_per_step = int(capi.lib().bags_launch_count() - l0)
        torch.cuda.synchronize()
        clear_pose_grads()

    overlap = None
    if use_dist:
        use_p2p = args.allreduce == "p2p" or (args.allreduce == "auto" and world == 2)
        overlap = VP.P2PGradientReducer(params) if use_p2p else VP.GradientReducer(params)
    if batch_mode and batch["reducer"] is not None:
        overlap = batch["reducer"]

    def reduce_grads():
        if batch_mode and batch["reducer"] is not None:
            return          # exchanged slice by slice inside the (captured) backward
        if use_dist:
            if overlap is not None:
                overlap.reduce()
            else:
                VP.allreduce_gradients(params)

    def step_device():
        if cap_dev is not None:
            cap_dev.replay()
        else:
            body_device()
        reduce_grads()

    # ---- e2e: ground truth from pinned host memory, loss read back ------------------------------------------------------
    copy_stream = torch.cuda.Stream(device=dev)
    st8 = {"n": 0, "h2d": [None, None], "used": [None, None], "last_loss": None}
    loss_slots = [{"host": torch.zeros(1, dtype=torch.float32).pin_memory(), "ev": None} for _ in range(2)]

    def issue_h2d(buf):
        """host->device copies of every view's ground truth into buffer set `buf` on the copy stream, after the last
        step that read that set"""
        with torch.cuda.stream(copy_stream):
            if st8["used"][buf] is not None:
                copy_stream.wait_event(st8["used"][buf])
            for vw in views:
                vw["gt_dev"][buf].copy_(vw["gt_host"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        st8["h2d"][buf] = ev

    def step_e2e():
        """The step with host-resident ground truth.  Two device buffer sets alternate: the copies for step s+1 run on
        the copy stream underneath step s.  The step's loss is copied to pinned host memory every step and consumed
        one step later (how a training loop logs its loss without draining the GPU); finish_e2e() collects the last
        one inside the timed region."""
        buf = st8["n"] % 2
        if st8["h2d"][buf] is None:
            issue_h2d(buf)
        issue_h2d(1 - buf)                                  # next step's images
        cur = torch.cuda.current_stream()
        cur.wait_event(st8["h2d"][buf])
        if cap_e2e is not None:
            cap_e2e[buf].replay()
        else:
            body_e2e(buf)
        used = torch.cuda.Event()
        used.record(cur)
        st8["used"][buf] = used
        st8["h2d"][buf] = None
        reduce_grads()
        slot = loss_slots[st8["n"] % 2]
        if slot["ev"] is not None:
            slot["ev"].synchronize()
            st8["last_loss"] = float(slot["host"][0])
        slot["host"].copy_(loss_dev[buf], non_blocking=True)
        slot["ev"] = torch.cuda.Event()
        slot["ev"].record()
        st8["n"] += 1

    def finish_e2e():
        for slot in loss_slots:
            if slot["ev"] is not None:
                slot["ev"].synchronize()
                st8["last_loss"] = float(slot["host"][0])
                slot["ev"] = None

    # ---- drop-in record: ONE view per step through the API exactly as train.py uses it ----------------------------------
    dstate = {"n": 0, "ev_next": None, "last_loss": None}
    d_slots = [{"host": torch.zeros(1, dtype=torch.float32).pin_memory(), "ev": None} for _ in range(2)]

    def h2d_one(vw):
        with torch.cuda.stream(copy_stream):
            if vw.get("used_ev") is not None:
                copy_stream.wait_event(vw["used_ev"])
            vw["gt_dev"][0].copy_(vw["gt_host"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def step_dropin_device():
        vw = views[dstate["n"] % len(views)]
        params.zero_grad()
        color, radii, depth, weights, _ = render(vw)          # blocking forward, current stream, no sinks
        torch.autograd.backward([color, depth, weights], vw["up"])
        dstate["n"] += 1

    def step_dropin_e2e():
        k = dstate["n"] % len(views)
        vw = views[k]
        ev = dstate["ev_next"] or h2d_one(vw)
        dstate["ev_next"] = h2d_one(views[(k + 1) % len(views)])      # prefetch the next view's image
        params.zero_grad()
        cur = torch.cuda.current_stream()
        color, radii, depth, weights, _ = render(vw)
        cur.wait_event(ev)
        loss = l1_loss(color, vw["gt_dev"][0].float().mul_(1.0 / 255.0))   # the caller's own torch loss
        loss.backward()
        used = torch.cuda.Event()
        used.record(cur)
        vw["used_ev"] = used
        slot = d_slots[dstate["n"] % 2]
        if slot["ev"] is not None:
            slot["ev"].synchronize()
            dstate["last_loss"] = float(slot["host"][0])
        slot["host"].copy_(loss.detach().reshape(1), non_blocking=True)
        slot["ev"] = torch.cuda.Event()
        slot["ev"].record()
        dstate["n"] += 1

    def finish_dropin():
        for slot in d_slots:
            if slot["ev"] is not None:
                slot["ev"].synchronize()
                dstate["last_loss"] = float(slot["host"][0])
                slot["ev"] = None

    # ---- timing ---------------------------------------------------------------------------------------------------------
    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_once(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if use_dist:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(fn, steps, finish=None):
        """Regions of exactly `steps` steps each (barrier + synchronize on both sides, CUDA events, max over ranks): at
        least --repeats regions, at least 100 steps and at least --min-seconds in total (capped at 40 regions; every
        rank takes the same decisions because the times are already max-reduced).  The MEDIAN region is reported."""
        need = max(1, args.repeats, math.ceil(100 / max(1, steps)))
        all_ms = []
        while len(all_ms) < need or (sum(all_ms) < 1e3 * args.min_seconds and len(all_ms) < 40):
            all_ms.append(timed_once(fn, steps, finish))
        return median(all_ms), all_ms

    progress("scene, views and step built")
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # W untimed warm-up steps, and at least ~1.5 s of them (allocator growth, clocks, page-locked staging).
    # N > 1: a FIXED number of steps -- every step contains a collective, so every rank must run the same number of them;
    # a wall-clock criterion evaluated per rank let the ranks leave this loop a few steps apart and deadlocked the
    # exchange (seen as sporadic hangs of multi-GPU runs right after "scene, views and step built").
    t_warm, n_warm = time.time(), 0
    n_warm_fixed = max(args.warmup, 3, 160) if use_dist else None
    while (n_warm < n_warm_fixed) if use_dist else \
            (n_warm < max(args.warmup, 3) or (time.time() - t_warm < 1.5 and n_warm < 400)):
        step_device()
        n_warm += 1
        if n_warm % 8 == 0:
            torch.cuda.synchronize()
    barrier()
    if sampler is not None:
        sampler.wait_first_sample()

    launches = None
    if capi is not None and not graph_mode:
        launches0 = capi.lib().bags_launch_count()
    progress("warm-up done")
    ms, ms_all = timed(step_device, args.steps)
    progress("device leg timed")
    if capi is not None:
        if graph_mode:
            launches = launches_per_step * args.steps          # one replay = the captured launches of one step
        else:
            launches = int(capi.lib().bags_launch_count() - launches0) // len(ms_all)
    if cap_dev is not None:
        cap_dev.check()

    # N > 1 diagnostic (outside the timed regions): every rank's own compute time per step, without the exchange and
    # without waiting for anybody -- what the slowest rank costs the others shows up as max vs mean
    per_rank_ms = None
    if use_dist and cap_dev is not None and not (batch_mode and batch["reducer"] is not None):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            cap_dev.replay()
        e1.record()
        torch.cuda.synchronize()
        mine_ms = torch.tensor([e0.elapsed_time(e1) / 10.0], device=dev)
        every = [torch.zeros_like(mine_ms) for _ in range(world)]
        dist.all_gather(every, mine_ms)
        per_rank_ms = [round(float(t), 4) for t in every]

    e2e = None
    if not args.no_e2e:
        if batch_mode:
            clear_batch_grads()
            cap_e2e = [VP.CapturedStep(lambda b=b: body_e2e_batch(b), dev, mod, None, forwards_per_step=1,
                                      reset=clear_batch_grads, check=batch["rast"].check) for b in range(2)]
        elif graph_mode:
            clear_pose_grads()
            cap_e2e = [VP.CapturedStep(lambda b=b: body_e2e(b), dev, mod, capacity, forwards_per_step=len(views),
                                      reset=clear_pose_grads) for b in range(2)]
        for _ in range(max(args.warmup, 3)):
            step_e2e()
        finish_e2e()
        ms_e2e, ms_e2e_all = timed(step_e2e, args.steps, finish_e2e)
        if cap_e2e is not None:
            for c in cap_e2e:
                c.check()
        e2e = {"value": nviews * args.steps / (ms_e2e / 1e3), "unit": "views/s",
               "h2d_bytes_per_step": len(views) * 3 * N, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / args.steps,
               "ms_per_step_regions": [round(m / args.steps, 4) for m in ms_e2e_all],
               "loss_readback": "async copy to pinned memory every step, consumed one step later",
               "last_loss": st8["last_loss"]}

    # ---- drop-in record (N = 1): one view per step, blocking forward, one stream, autograd accumulation -----------------
    drop_in = None
    if world == 1 and not args.no_drop_in:
        nsteps = args.steps * len(views)                       # every camera equally often in a region
        for _ in range(2 * len(views)):
            step_dropin_device()
        d_ms, d_all = timed(step_dropin_device, nsteps)
        drop_in = {"value": nsteps / (d_ms / 1e3), "unit": "views/s", "ms_per_view": d_ms / nsteps,
                   "ms_per_view_regions": [round(m / nsteps, 4) for m in d_all],
                   "what": "ONE view per step through GaussianRasterizer exactly as train.py calls it: blocking "
                           "forward, current stream, gradients accumulated by autograd (no sinks, no graph, no "
                           "view pipelining); same cameras, upstream gradients resident in HBM"}
        if not args.no_e2e:
            for _ in range(2 * len(views)):
                step_dropin_e2e()
            finish_dropin()
            de_ms, de_all = timed(step_dropin_e2e, nsteps, finish_dropin)
            drop_in["e2e"] = {"value": nsteps / (de_ms / 1e3), "unit": "views/s", "ms_per_view": de_ms / nsteps,
                              "ms_per_view_regions": [round(m / nsteps, 4) for m in de_all],
                              "h2d_bytes_per_step": 3 * N, "d2h_bytes_per_step": 4,
                              "what": "+ 8-bit ground truth from pinned host memory (prefetched one view ahead), "
                                      "torch L1 loss (utils/loss_utils.py:18-19), loss read back"}
    clocks = sampler.stop() if sampler is not None else None
    progress("timed legs done")

    # per-stage CUDA-event times (events on the launching stream, recorded inside the library) in a separate short
    # pass, so that the event records do not sit in the timed region above
    stage_ms, stage_calls = {}, {}
    if capi is not None:
        capi.lib().bags_profile_enable(1)
        capi.profile_read()
        for _ in range(3):
            if batch_mode:
                clear_batch_grads()
                body_device_batch()
            else:
                body_device(pipe_serial)   # one stream: concurrent kernels would inflate each other's event times
        torch.cuda.synchronize()
        stage_ms, stage_calls = capi.profile_read()
        capi.lib().bags_profile_enable(0)

    def leave(code=0):
        """End of a multi-rank run: release the captured graphs (graph-captured collectives keep NCCL work alive inside
        them; destroy_process_group() was seen to hang behind that), meet at a barrier, tear the group down -- with a
        timer that ends the process if the teardown does not return."""
        nonlocal cap_dev, cap_e2e
        if use_dist:
            import gc
            import threading
            torch.cuda.synchronize()
            cap_dev = cap_e2e = None
            gc.collect()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            t = threading.Timer(15.0, lambda: os._exit(code))
            t.daemon = True
            t.start()
            dist.destroy_process_group()
            t.cancel()
        return code


    if rank != 0:
        return leave(0)

    value = nviews * args.steps / (ms / 1e3)
    line = {
        "metric": "fwd+bwd views/s (1M Gaussians, 1080p)" if args.config == "C3" else f"fwd+bwd views/s ({args.config})",
        "value": value, "unit": "views/s", "n_gpus": world if ours else 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "ms_per_step_regions": [round(m / args.steps, 4) for m in ms_all], "region_statistic": "median",
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config}: {P} random Gaussians SH3, {W}x{H}, {vpg} views/GPU/step "
                               f"(fwd+bwd incl. pose/intrinsic grads), median scale {args.scale}",
                   "l2": "per-view working set (236 B x P inputs + state) exceeds the 126 MB L2; no explicit flush",
                   "parallelism": f"view-dp{args.gpus}"},
        "workload_stats": {"views_per_step": nviews, "visible_mean": V, "tile_instances_mean": R},
        "clocks": clocks,
    }
    if not ours:
        line["impl"] = "reference"
        line["implementation"] = {"mode": "eager, blocking forward (the reference's cudaMemcpy), one stream, autograd "
                                          "accumulation"}
        line["cpu_baseline"] = {"value": value, "unit": "views/s", "cores": 0, "kind": "reference",
                                "sample": "the reference has no CPU rasterizer: this arm times its unmodified CUDA "
                                          "extension (oracle/_ref, built for sm_100) through its own Python API on "
                                          "the same workload, GPU-resident inputs"}
        line["e2e"] = {"value": e2e["value"] if e2e else value, "unit": "views/s",
                       "h2d_bytes_per_step": e2e["h2d_bytes_per_step"] if e2e else 0,
                       "d2h_bytes_per_step": e2e["d2h_bytes_per_step"] if e2e else 0,
                       "ms_per_step_regions": e2e["ms_per_step_regions"] if e2e else None}
        if drop_in is not None:
            line["drop_in"] = drop_in
        emit(line)
        return 0

    line["implementation"] = {
        "mode": (f"the rank's {len(views)} views as ONE launch set (GaussianBatchRasterizer: bags_rasterize_forward_batch"
                 f" / _backward_batch_*, binning capacity {capacity} tile instances per view), the step's enqueue "
                 "captured in a CUDA graph, one launch per step") if batch_mode
        else ("CUDA graph of the step's enqueue (bags_view_parallel.CapturedStep) over the asynchronous forward "
              f"(binning capacity {capacity} tile instances), one launch per step") if graph_mode
        else "eager launches, blocking forward",
        "streams": 1 if batch_mode else pipe.depth, "view_assignment": assignment,
        "grad_accumulation": ("the batched backward writes the views' summed gradients into the flat gradient buffer "
                              "(register_grad_sink, sink_overwrite)") if (grad_sinks and batch_mode)
        else "kernel adds into the flat gradient buffer (register_grad_sink)" if grad_sinks
        else "autograd (fresh gradient tensors + add)",
        "allreduce": (overlap.describe() if overlap is not None else "one all_reduce after the step") if use_dist
        else None}
    if tuned:
        line["implementation"]["tuning"] = tuned
    if per_rank_ms is not None:
        line["implementation"]["per_rank_compute_ms_per_step"] = per_rank_ms
    line["e2e"] = e2e
    if drop_in is not None:
        line["drop_in"] = drop_in
    line["gpu_launches"] = launches
    # ---- roofline of the dominant kernel + per-stage table -----------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    Tt = ((W + 15) // 16) * ((H + 15) // 16)
    ab = algorithmic_bytes(P, V, R, N, Tt)
    stage_bytes = {"preprocess": ab["preprocess"], "depth_sort": ab["scan"], "emit": ab["duplicate"],
                   "tile_sort": ab["sort"], "ranges": ab["ranges"], "blend_fwd": ab["blend_fwd"],
                   "blend_bwd": ab["blend_bwd"], "preprocess_bwd": ab["preprocess_bwd"], "bwd_clear": 0}
    stages = {}
    nviews_prof = 3 * len(views)                                # views rendered by the profiling pass above
    for k, tot in stage_ms.items():
        avg = tot / nviews_prof
        gbs = stage_bytes.get(k, 0) / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
        stages[k] = {"ms": round(avg, 4), "alg_GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}
    dom = max(stages, key=lambda k: stages[k]["ms"]) if stages else None
    traffic = None
    traffic_R = None
    try:  # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = tj.get(dom, {}).get("dram_bytes_per_launch")
        traffic_R = tj.get("_meta", {}).get("tile_instances")
    except Exception:
        pass
    if dom:
        line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["alg_GBps"], "peak": peak,
                            "unit": "GB/s", "frac": stages[dom]["frac"], "traffic": traffic,
                            "peak_source": peak_src, "R": R, "V": V,
                            "note": "blend kernels are FP32-issue bound, not HBM bound; see stages + profiles/.  "
                                    "achieved / traffic are per VIEW: in batch mode one launch covers the rank's "
                                    "views (launch / views).  traffic: DRAM bytes from the committed ncu --set full "
                                    f"capture of the same cameras as one batched launch set (mean R = {traffic_R} tile "
                                    "instances per view)"}
        pb = ["preprocess", "depth_sort", "emit", "tile_sort", "ranges"]
        t_pb = sum(stages[k]["ms"] for k in pb if k in stages)
        b_pb = ab["preprocess"] + ab["scan"] + ab["duplicate"] + ab["sort"] + ab["ranges"]
        line["stages"] = stages
        line["preprocess_binning"] = {"alg_GBps": round(b_pb / (t_pb * 1e-3) / 1e9, 1) if t_pb > 0 else None,
                                      "frac": round(b_pb / (t_pb * 1e-3) / 1e9 / peak, 4) if t_pb > 0 else None,
                                      "ms": round(t_pb, 4)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_sample()
        except Exception as ex:  # the baseline is a report, not the product
            line["cpu_baseline"] = {"error": repr(ex)}
    emit(line)
    return leave(0)


if __name__ == "__main__":
    sys.exit(main())
