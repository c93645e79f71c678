"""View-batched data parallelism for the rasterizer.

The reference trains on ONE view per iteration on one GPU and has no distributed code at all.  The natural sharding
axis is the view: every (view) forward+backward is independent given replicated Gaussian parameters, and the only
coupling is the sum of the Gaussian-parameter gradients.  This module provides exactly that and nothing more:

* ``shard_views(num_views, rank, world)``  -- rank r of G owns views r, r+G, ...
* ``balance_views(costs, world)``  -- the same number of views per rank, dealt out by measured work (tile instances).
* ``FlatGaussianParams``  -- the Gaussian parameters as views into ONE contiguous fp32 buffer (59 floats per Gaussian:
  means3D 3 | SH 48 | opacity 1 | scales 3 | rotations 4) with ``.grad`` views into ONE contiguous gradient buffer, so
  autograd accumulates every view's gradients in place and a single all-reduce moves them (no pack step).
  ``enable_direct_accumulation()`` additionally lets the backward kernel add into that buffer itself.
* ``allreduce_gradients(params, group)``  -- one ``all_reduce(SUM)`` over that buffer (NCCL over NVLink on GPUs,
  gloo in the CPU tests).  Per-camera pose / intrinsic gradients are NOT reduced: each camera is owned by one rank.

Parity definition: the reduced buffer equals the sum over all views of the single-view gradients.
"""
from typing import Callable, Iterable, List, Sequence

import torch
import torch.distributed as dist

FIELDS = (("means3D", (3,)), ("shs", (16, 3)), ("opacities", (1,)), ("scales", (3,)), ("rotations", (4,)))
FLOATS_PER_GAUSSIAN = 59


def shard_views(num_views: int, rank: int, world: int) -> List[int]:
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"invalid rank/world {rank}/{world}")
    return list(range(rank, num_views, world))


def balance_views(costs: Sequence[float], world: int) -> List[List[int]]:
    """Assignment of views to ranks that evens out the per-rank work: every rank gets the same number of views
    (+-1), and the views are dealt out longest-first to the rank with the least work so far (LPT with a cardinality
    limit).  `costs[v]` is any per-view work estimate -- the number of tile instances (`num_rendered`) of the view's
    previous render is what bench.py uses: the blend kernels, 3/4 of a view, are proportional to it, and it differs
    by up to 3x between cameras of one scene.  The step time of view-parallel training is the time of the slowest
    rank, so `r, r+G, ...` sharding of 32 views over 8 ranks (max/mean work 1.15 on the bench scene) costs that 15 %.
    Deterministic (ties broken by view index, then rank index): every rank computes the same table."""
    n = len(costs)
    if world <= 0:
        raise ValueError(f"invalid world size {world}")
    cap = -(-n // world)
    order = sorted(range(n), key=lambda v: (-float(costs[v]), v))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    full_ranks_allowed = n - (cap - 1) * world if cap > 0 else 0     # how many ranks may hold `cap` views
    for v in order:
        best = None
        for r in range(world):
            limit = cap if (len(out[r]) < cap - 1 or sum(1 for q in out if len(q) == cap) < full_ranks_allowed) else cap - 1
            if len(out[r]) >= limit:
                continue
            if best is None or load[r] < load[best]:
                best = r
        out[best].append(v)
        load[best] += float(costs[v])
    # refinement: swap one view of the most loaded rank with one of another rank while that lowers the larger of the
    # two loads (the greedy deal leaves the rank that drew the heaviest view a few percent above the rest)
    c = [float(x) for x in costs]
    for _ in range(4 * n + 8):
        a = max(range(world), key=lambda r: (load[r], -r))
        best_gain, best_swap = 1e-12 * (1.0 + load[a]), None
        for b in range(world):
            if b == a:
                continue
            for va in out[a]:
                for vb in out[b]:
                    d = c[va] - c[vb]
                    if d <= 0:
                        continue
                    new_max = max(load[a] - d, load[b] + d)
                    gain = load[a] - new_max
                    if gain > best_gain:
                        best_gain, best_swap = gain, (b, va, vb)
        if best_swap is None:
            break
        b, va, vb = best_swap
        out[a].remove(va); out[b].remove(vb)
        out[a].append(vb); out[b].append(va)
        d = c[va] - c[vb]
        load[a] -= d
        load[b] += d
    return [sorted(q) for q in out]


class FlatGaussianParams:
    """Gaussian parameters and their gradients in two flat buffers (field-major: all means, then all SH, ...)."""

    def __init__(self, means3D, shs, opacities, scales, rotations, device=None):
        src = dict(means3D=means3D, shs=shs, opacities=opacities, scales=scales, rotations=rotations)
        P = means3D.shape[0]
        device = device if device is not None else means3D.device
        self.P = P
        self.flat = torch.empty(P * FLOATS_PER_GAUSSIAN, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros_like(self.flat)
        self.params = {}
        self.grad_views = {}
        off = 0
        for name, shape in FIELDS:
            n = P
            for s in shape:
                n *= s
            view = self.flat[off:off + n].view(P, *shape)
            view.copy_(src[name].reshape(P, *shape))
            p = view.detach().requires_grad_(True)          # leaf sharing storage with the flat buffer
            self.grad_views[name] = self.flat_grad[off:off + n].view(P, *shape)
            p.grad = self.grad_views[name]
            self.params[name] = p
            off += n
        assert off == P * FLOATS_PER_GAUSSIAN

    def __getitem__(self, k):
        return self.params[k]

    def zero_grad(self):
        """Zeroes the flat gradient buffer and (re-)attaches every parameter's `.grad` to its slice of it.  Use this
        instead of `optimizer.zero_grad(set_to_none=True)` (the reference's train.py:421), which drops the `.grad`
        views: autograd would then allocate fresh gradient tensors outside the flat buffer, and with gradient sinks
        the kernel would keep adding onto stale sums."""
        self.flat_grad.zero_()
        for name, p in self.params.items():
            if p.grad is None or p.grad.data_ptr() != self.grad_views[name].data_ptr():
                p.grad = self.grad_views[name]

    def ensure_attached(self):
        """Repairs the `.grad` <-> flat-buffer link after foreign code touched it, keeping the values autograd means:
        a `.grad` of None (`zero_grad(set_to_none=True)`) stands for zero -> that field's slice is zeroed; a `.grad`
        that is some other tensor (allocated by autograd after set_to_none) holds the sum accumulated so far -> it is
        copied into the slice.  Called at the start of `run_views` and by `allreduce_gradients`."""
        for name, p in self.params.items():
            view = self.grad_views[name]
            if p.grad is None:
                if not getattr(self, "_sink_module", None):
                    view.zero_()
                # with direct accumulation the kernel has been adding into the slice while `.grad` was None: the
                # slice already holds the sum (the step must have started from zero_grad() / ensure_attached())
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def begin_step(self):
        """Start of an accumulation step for callers that do not use zero_grad(): `.grad is None` means zero."""
        for name, p in self.params.items():
            view = self.grad_views[name]
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def enable_direct_accumulation(self, rasterizer_module=None):
        """Registers every parameter's gradient view as a gradient sink of the rasterizer
        (diff_gaussian_rasterization.register_grad_sink): the backward kernel then adds each view's gradient straight
        into the flat buffer -- bit-identical to autograd's `grad += new`, without the fresh tensors and the add
        passes.  CUDA only (the sinks are a feature of the native library)."""
        if rasterizer_module is None:
            import diff_gaussian_rasterization as rasterizer_module
        for name, _ in FIELDS:
            rasterizer_module.register_grad_sink(self.params[name], self.grad_views[name])
        self._sink_module = rasterizer_module

    def disable_direct_accumulation(self):
        m = getattr(self, "_sink_module", None)
        if m is not None:
            for name, _ in FIELDS:
                m.unregister_grad_sink(self.params[name])
            self._sink_module = None

    def accumulating(self):
        """`with params.accumulating():` -- scope in which the registered sinks are active (the training forward +
        backward of a step).  Outside it -- evaluation renders, torch.autograd.grad probes, double backward -- the
        rasterizer behaves like the reference: fresh gradient tensors, hooks fire, `.grad` is only touched by
        autograd."""
        m = getattr(self, "_sink_module", None)
        if m is None:
            import contextlib
            return contextlib.nullcontext()
        return m.grad_sinks(True)


def allreduce_gradients(params: FlatGaussianParams, group=None, async_op: bool = False):
    """Sums the flat gradient buffer over all ranks (no-op without an initialised process group)."""
    params.ensure_attached()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(params.flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class GradientReducer:
    """The step-end sum of the flat gradient buffer over the ranks, plus the small per-step statistics that ride along
    (SURVEY.md section 8e): everything a view-parallel trainer has to exchange per iteration.

    * `reduce()`: all_reduce(SUM) of `params.flat_grad` (59 floats per Gaussian).
    * `reduce_stats(stats)`: the densification statistics (scene/gaussian_model.py:449-455: `xyz_gradient_accum`,
      `xyz_gradient_accum_abs`, `denom` -- per-view norms, so every rank accumulates its own views first and the sums
      are exchanged once per step), `max_radii2D` (train.py:377: a running maximum -> all_reduce(MAX)) and the gradient
      of the one global `shift_factors` 3-vector, packed so that the step costs two small collectives next to the big
      one: one SUM over [accum | accum_abs | denom | shift grad], one MAX over max_radii2D."""

    def __init__(self, params: "FlatGaussianParams", group=None):
        self.params = params
        self.group = group

    def describe(self):
        return "one all_reduce(SUM) of the flat gradient buffer after the step"

    def reduce(self):
        return allreduce_gradients(self.params, self.group)

    def reduce_stats(self, stats: "DensificationStats", shift_grad=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        if shift_grad is not None:
            stats.sum_buf[-3:].copy_(shift_grad.reshape(3))
        dist.all_reduce(stats.sum_buf, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(stats.max_radii2D, op=dist.ReduceOp.MAX, group=self.group)
        if shift_grad is not None:
            shift_grad.copy_(stats.sum_buf[-3:].view_as(shift_grad))


class SlicedGradientReducer(GradientReducer):
    """The gradient all-reduce overlapped with the backward that produces the gradients.

    The batched backward (`diff_gaussian_rasterization.GaussianBatchRasterizer(..., backward_slices=S, on_slice=...)`)
    finishes the Gaussian-parameter gradients of ALL of a rank's views slice by slice (consecutive Gaussians).  As soon
    as a slice has been enqueued its rows of the flat buffer are final, so their all-reduce is launched right there
    (`async_op=True`: NCCL's stream waits for the slice's kernel, the compute stream goes on with the next slice) and
    only the last slice's exchange is exposed -- instead of one 236 B/Gaussian all-reduce after the whole step.

    The flat buffer is field-major (means | SH | opacity | scales | rotations), so a slice of Gaussians is five
    contiguous ranges: they go as ONE coalesced all-reduce (one NCCL group launch) per slice.  `finish()` makes the
    current stream wait for everything."""

    def __init__(self, params: "FlatGaussianParams", group=None):
        super().__init__(params, group)
        self._works = []
        self._active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def describe(self):
        return ("one coalesced all_reduce(SUM) of each slice's gradient rows (5 ranges, 59 floats per Gaussian) as the "
                "batched backward finishes the slice, overlapped with the next slice")

    def slice_ranges(self, first: int, count: int):
        """The five contiguous pieces of the flat gradient buffer that hold Gaussians [first, first + count)."""
        P = self.params.P
        g = self.params.flat_grad
        out, off = [], 0
        for _, shape in FIELDS:
            w = 1
            for s in shape:
                w *= s
            out.append(g[off + w * first:off + w * (first + count)])
            off += w * P
        return out

    def on_slice(self, first: int, count: int):
        if not self._active or count <= 0:
            return
        pieces = self.slice_ranges(first, count)
        cm_factory = getattr(dist.distributed_c10d, "_coalescing_manager", None)
        if cm_factory is not None:
            with cm_factory(self.group, async_ops=True) as cm:
                for t in pieces:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self._works.append(cm)
        else:
            self._works.append(dist.all_reduce_coalesced(pieces, op=dist.ReduceOp.SUM, group=self.group,
                                                         async_op=True))

    def finish(self):
        """The current stream waits for every all-reduce launched since the last finish()."""
        for w in self._works:
            w.wait()
        self._works = []

    def reduce(self):
        # everything was exchanged during the backward
        self.finish()


class P2PGradientReducer(GradientReducer):
    """The step-end gradient sum as ONE kernel over NVLink peer memory instead of an NCCL ring (C ABI
    `bags_allreduce_p2p`, csrc/allreduce.cu): every rank maps the flat gradient buffers of its peers (CUDA IPC, one
    process per GPU of one NVSwitch node), adds the contributions to the slice it owns in rank order and stores the sums
    into all buffers; loads and stores stream concurrently, there are no ring steps.  All ranks end up with bit-identical
    buffers; the sums differ from NCCL's by fp32 summation order only.

    `P2PGradientReducer(params)` exchanges the IPC handles once (all_gather_object over the process group).  If the
    handles cannot be exchanged or mapped (no peer access, IPC not permitted, more than 8 ranks, a buffer NCCL reallocated)
    it says why in `self.fallback_reason` and `reduce()` uses NCCL's all_reduce."""

    def __init__(self, params: "FlatGaussianParams", group=None):
        super().__init__(params, group)
        self.fallback_reason = None
        self._peers = None
        self._keep = []
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            self.fallback_reason = "single process"
            return
        try:
            self._setup()
        except Exception as ex:          # every rank must take the same decision: agree below
            self.fallback_reason = f"{type(ex).__name__}: {ex}"
        ok = torch.tensor([0 if self.fallback_reason else 1], device=params.flat_grad.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok) == 0:
            if self.fallback_reason is None:
                self.fallback_reason = "another rank could not map the peer buffers"
            self._peers = None

    def _setup(self):
        import ctypes

        import bags_capi as capi
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if world > 8:
            raise RuntimeError("more than 8 ranks")
        g = self.params.flat_grad
        if g.numel() % 4 != 0 or g.data_ptr() % 16 != 0:
            raise RuntimeError("gradient buffer not a multiple of four floats / not 16-byte aligned")
        L = capi.lib()
        self._flags = torch.zeros(int(L.bags_allreduce_p2p_flag_bytes()) // 4, dtype=torch.int32, device=g.device)
        torch.cuda.synchronize(g.device)

        def export(t):
            # (device, 64-byte cudaIpcMemHandle_t of the allocation, its size, offset of the storage inside it, ...)
            sh = t.untyped_storage()._share_cuda_()
            h = bytes(sh[1])
            # newer caching allocators prefix the raw handle with a version byte and a type byte ('c': cudaMalloc block,
            # 'e': expandable segment -- a file descriptor, not a cudaIpcMemHandle_t)
            if len(h) == 66 and h[1:2] == b"c":
                h = h[2:]
            elif len(h) != 64:
                raise RuntimeError(f"unsupported CUDA IPC handle format ({len(h)} bytes, type {h[1:2]!r}); "
                                   "expandable segments are not supported")
            return h, int(sh[3]) + t.storage_offset() * t.element_size(), int(t.numel())

        every = [None] * world
        dist.all_gather_object(every, (export(g), export(self._flags)), group=self.group)
        bufs, flags = [], []
        with torch.cuda.device(g.device):
            for r in range(world):
                if r == rank:
                    bufs.append(g.data_ptr())
                    flags.append(self._flags.data_ptr())
                    continue
                ptrs = []
                for handle, offset, numel in every[r]:
                    base = ctypes.c_void_p()
                    capi.check(L.bags_ipc_open(handle, ctypes.byref(base)), "bags_ipc_open")
                    self._keep.append(base.value)
                    ptrs.append(base.value + offset)
                if every[r][0][2] != g.numel():
                    raise RuntimeError("ranks disagree on the size of the gradient buffer")
                bufs.append(ptrs[0])
                flags.append(ptrs[1])
        self._buf_ptrs = (ctypes.c_void_p * world)(*bufs)
        self._flag_ptrs = (ctypes.c_void_p * world)(*flags)
        self._own = (g.data_ptr(), g.numel())
        self._peers = (world, rank)
        # (no barrier here: a rank whose mapping failed never gets this far, and the agreement all-reduce in __init__ --
        # which every rank reaches, mapped or not -- is the point after which everyone has mapped everything)

    def close(self):
        """Unmaps the peers' buffers (collective in spirit: call it on every rank, after the last reduce())."""
        if self._keep:
            import bags_capi as capi
            for base in self._keep:
                capi.lib().bags_ipc_close(base)
            self._keep = []
        self._peers = None
        if self.fallback_reason is None:
            self.fallback_reason = "closed"

    def describe(self):
        if self._peers is None:
            return f"one all_reduce(SUM) of the flat gradient buffer after the step (NCCL; p2p unavailable: {self.fallback_reason})"
        return ("one kernel over NVLink peer memory after the step (bags_allreduce_p2p: every rank sums the slice it "
                "owns straight out of the peers' buffers and stores the result into all of them)")

    def reduce(self):
        if self._peers is None:
            return allreduce_gradients(self.params, self.group)
        import ctypes

        import bags_capi as capi
        self.params.ensure_attached()
        g = self.params.flat_grad
        if (g.data_ptr(), g.numel()) != self._own:
            # the peers hold mappings of the buffer that existed when the reducer was built
            raise RuntimeError("P2PGradientReducer: the flat gradient buffer was reallocated (e.g. after densification); "
                               "build a new reducer for the new FlatGaussianParams on every rank")
        world, rank = self._peers
        with torch.cuda.device(g.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
            rc = capi.lib().bags_allreduce_p2p(world, rank, self._buf_ptrs, self._flag_ptrs, g.numel(), stream)
        capi.check(rc, "bags_allreduce_p2p")
        return None


class DensificationStats:
    """Per-step densification statistics of the views a rank rendered, laid out for one SUM and one MAX collective.

    Mirrors `GaussianModel.add_densification_stats` (scene/gaussian_model.py:449-455) and the `max_radii2D` update of
    train.py:377: for every view, with vis = radii > 0,
        xyz_gradient_accum[vis] += ||viewspace_point_tensor.grad[vis, :2]||             (abs_grad False)
        xyz_gradient_accum[vis] += ||viewspace_point_tensor_densify.grad[vis, :2]||     (abs_grad True)
        denom[vis]              += 1
        max_radii2D[vis]         = max(max_radii2D[vis], radii[vis])
    Both norms are kept (accum / accum_abs) so that the caller picks at `apply_to`.  `sum_buf` = [accum (P) |
    accum_abs (P) | denom (P) | 3 spare words for the shift_factors gradient]; after `GradientReducer.reduce_stats`
    every rank holds the step's totals, which the caller adds to the model's running accumulators (`apply_to`)."""

    def __init__(self, P: int, device):
        self.P = P
        self.sum_buf = torch.zeros(3 * P + 3, dtype=torch.float32, device=device)
        self.max_radii2D = torch.zeros(P, dtype=torch.float32, device=device)

    @property
    def accum(self):
        return self.sum_buf[:self.P]

    @property
    def accum_abs(self):
        return self.sum_buf[self.P:2 * self.P]

    @property
    def denom(self):
        return self.sum_buf[2 * self.P:3 * self.P]

    def zero(self):
        self.sum_buf.zero_()
        self.max_radii2D.zero_()

    def add_view(self, viewspace_grad, viewspace_grad_densify, radii):
        vis = radii > 0
        self.accum.add_(torch.norm(viewspace_grad[:, :2], dim=-1) * vis)
        self.accum_abs.add_(torch.norm(viewspace_grad_densify[:, :2], dim=-1) * vis)
        self.denom.add_(vis.to(torch.float32))
        torch.maximum(self.max_radii2D, radii.to(torch.float32) * vis, out=self.max_radii2D)

    def apply_to(self, xyz_gradient_accum, denom, max_radii2D, abs_grad: bool):
        """Adds the (reduced) step totals to a GaussianModel's running accumulators (shapes (P,1) / (P,1) / (P,))."""
        src = self.accum_abs if abs_grad else self.accum
        xyz_gradient_accum.add_(src.view_as(xyz_gradient_accum))
        denom.add_(self.denom.view_as(denom))
        torch.maximum(max_radii2D, self.max_radii2D.view_as(max_radii2D).to(max_radii2D.dtype), out=max_radii2D)


def gather_view_costs(local_costs: Sequence[float], num_views: int, rank: int, world: int, group=None) -> List[float]:
    """Per-view work estimates of ALL views from the ones every rank measured for its round-robin share
    (`shard_views`): one all_gather_object of a few numbers.  Without a process group the local list is the result."""
    mine = [float(c) for c in local_costs]
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return mine
    every = [None] * world
    dist.all_gather_object(every, mine, group=group)
    costs = [0.0] * num_views
    for r in range(world):
        for k, v in enumerate(shard_views(num_views, r, world)):
            costs[v] = every[r][k]
    return costs


def run_views(params: FlatGaussianParams, views: Sequence, view_loss: Callable, num_views_total: int = None,
              rank: int = 0, world: int = 1, group=None, assignment: Sequence[Sequence[int]] = None):
    """Gradient accumulation over this rank's share of `views`, then one all-reduce.

    `view_loss(params, view) -> scalar loss tensor` renders one view and returns its loss; per-view pose gradients
    land on that view's own leaves.  `assignment` (e.g. from `balance_views`) overrides the round-robin share:
    `assignment[rank]` lists this rank's views.  Returns the list of local losses (python floats are NOT taken: no
    host sync).
    """
    n = len(views) if num_views_total is None else num_views_total
    losses = []
    mine = shard_views(n, rank, world) if assignment is None else list(assignment[rank])
    params.begin_step()
    with params.accumulating():
        for vi in mine:
            loss = view_loss(params, views[vi])
            loss.backward()
            losses.append(loss.detach())
    allreduce_gradients(params, group)
    return losses


class ViewStreams:
    """Software pipelining of a rank's views over a few CUDA streams (several views
    over the same Gaussians in flight at once).

    A view's preprocess + binning kernels are short and latency bound (they fill a fraction of the GPU), its blend
    kernels are long and issue bound.  Issuing consecutive views on alternating streams lets view k's preprocess /
    binning / forward blend run underneath view k-1's backward.  The backward passes themselves are chained with
    events (view k's backward starts after view k-1's has finished): they all add into the same flat gradient buffer,
    and two issue-bound backward kernels gain nothing from running side by side.

        vs = ViewStreams(device, depth=2)
        vs.begin()                         # after params.zero_grad(), on the current stream
        for k, view in enumerate(views):
            with vs.view(k):
                outputs = render(view)
                vs.before_backward()
                torch.autograd.backward(outputs, grads)
                vs.after_backward()
        vs.end()                           # the current stream waits for every view
    """

    def __init__(self, device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = max(1, int(depth))
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.depth)] if self.depth > 1 else []
        self._last_bwd = None
        self._cur = None
        self._ctx = None

    def begin(self):
        self._last_bwd = None
        if self.depth > 1:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            for s in self.streams:
                s.wait_event(ev)

    def view(self, k: int):
        vs = self

        class _Scope:
            def __enter__(self_inner):
                if vs.depth > 1:
                    vs._cur = vs.streams[k % vs.depth]
                    vs._ctx = torch.cuda.stream(vs._cur)
                    vs._ctx.__enter__()
                return vs

            def __exit__(self_inner, *exc):
                if vs.depth > 1:
                    vs._ctx.__exit__(*exc)
                    vs._cur = None
                return False
        return _Scope()

    def before_backward(self):
        if self.depth > 1 and self._last_bwd is not None:
            self._cur.wait_event(self._last_bwd)

    def after_backward(self):
        if self.depth > 1:
            ev = torch.cuda.Event()
            ev.record(self._cur)
            self._last_bwd = ev

    def end(self):
        if self.depth > 1:
            main = torch.cuda.current_stream(self.device)
            for s in self.streams:
                ev = torch.cuda.Event()
                ev.record(s)
                main.wait_event(ev)


class _NoAsyncScope:
    """Stand-in for AsyncForward when the captured body needs none."""

    def __init__(self, check=None):
        self._check = check
        self._k = 0
        self.counts = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def check(self):
        return self._check() if self._check is not None else 0


class CapturedStep:
    """A training step's whole enqueue -- any number of views: forward, loss, backward, gradient accumulation --
    captured once in a CUDA graph and replayed with ONE launch per step.

    This is what the asynchronous forward (`diff_gaussian_rasterization.AsyncForward`, C ABI
    `bags_rasterize_forward_async`) makes possible: the reference blocks the host on a 4-byte copy of `num_rendered` in
    every forward (rasterizer_impl.cu:291), which can neither be captured nor run ahead of.  With a caller-provided
    binning capacity a view is a pure enqueue; replayed as a graph, a step costs the host a few microseconds whatever
    else the machine's CPUs are doing.

        step = CapturedStep(body, device, rasterizer_module, capacity, forwards_per_step=len(views))
        step.replay()              # every iteration; inputs are updated IN PLACE between replays (static tensors)
        step.check()               # after a synchronisation point: raises if a view overflowed `capacity`

    `body()` must be capturable: static input tensors (camera matrices, ground truth, parameters are read from fixed
    addresses -- update them with copy_()), no host synchronisation, gradients of the Gaussian parameters accumulated
    through gradient sinks or into pre-attached `.grad` buffers.  Leaves whose gradient autograd accumulates itself
    (per-camera pose / intrinsic tensors, shift_factors) must have `.grad = None` when the capture starts, so that
    their gradient tensors come from the graph's private pool and are REWRITTEN by every replay -- with a `.grad`
    already in place autograd captures `grad += new` and the replays pile up.  `reset` (a callable that sets those
    `.grad`s to None) is invoked before every warm-up run and before the capture.
    The tile-instance counts of the step's forwards land in `step.counts` (pinned int64[forwards_per_step])."""

    def __init__(self, body: Callable, device, rasterizer_module, capacity: int, forwards_per_step: int,
                 warmup: int = 2, reset: Callable = None, check: Callable = None):
        """`capacity=None`: the body brings its own asynchronous forwards (a GaussianBatchRasterizer has its capacity
        and its counts); `check` is then what `CapturedStep.check()` calls."""
        self.device = torch.device(device)
        if capacity is None:
            self.af = _NoAsyncScope(check)
            self.capacity = None
        else:
            self.af = rasterizer_module.AsyncForward(capacity, slots=max(1, int(forwards_per_step)))
            self.capacity = int(capacity)
        main = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):      # eager runs: lazy initialisations happen outside the capture
                self.af._k = 0
                if reset is not None:
                    reset()
                with self.af:
                    body()
        main.wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.af.check()
        if reset is not None:
            reset()
        self.graph = torch.cuda.CUDAGraph()
        self.af._k = 0
        with torch.cuda.graph(self.graph):
            with self.af:
                self.outputs = body()
        self.counts = self.af.counts

    def replay(self):
        self.graph.replay()

    def check(self):
        """Call after the replays of interest have been synchronised: raises on a capacity overflow."""
        return self.af.check()
