"""Drop-in ``diff_gaussian_rasterization`` for Self-Cali-GS / Bundle-Adjusting-Gaussian-Splatting on B200.

Public surface and semantics follow 3dgs-pose/diff_gaussian_rasterization/__init__.py of the reference:

* ``GaussianRasterizationSettings`` -- the 14-field NamedTuple (reference lines 220-234);
* ``GaussianRasterizer(raster_settings)`` -- ``forward(means3D, means2D, means2D_densify, shift_factors, opacities,
  shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None)`` returning
  ``(color, radii, depth, weights, mean2D)`` and ``markVisible(positions)`` (lines 236-287);
* ``rasterize_gaussians(...)`` (lines 21-49) and ``compute_relocation(...)`` (lines 289-291).

The autograd function returns the 14 gradients in input order, with the view-matrix / camera-centre / projection /
pupil-shift gradients that ``--opt_cam`` and ``--opt_intrinsic`` consume (reference lines 119-218).
``gaussian_renderer/__init__.py:render()`` of the reference runs on top of this unchanged.

The native work happens in ``libbags_b200.so`` (hand-written sm_100a kernels, C ABI in include/bags_b200.h)
reached through ``_C``, a compiled PyTorch C++ extension with the reference's four pybind symbols
(3dgs-pose/ext.cpp:15-20; source: csrc/ext.cpp); there is no CPU path.

Extensions beyond the reference API (all off unless asked for): gradient sinks (``register_grad_sink`` +
``grad_sinks()``), the asynchronous forward (``AsyncForward``), several cameras over the same Gaussians as one launch
set (``GaussianBatchRasterizer``).
"""
import contextlib
from typing import NamedTuple

import torch
import torch.nn as nn

try:
    from . import _C
except ImportError as _e:       # no fallback: the extension is the product
    raise ImportError(
        "bags_b200: the compiled extension diff_gaussian_rasterization._C is missing or does not load (" + str(_e) +
        "); build it with `python bundle-adjusting-gaussian-splatting_b200/build.py` -- there is no CPU or PyTorch "
        "fallback for the rasterizer") from _e


def cpu_deep_copy_tuple(input_tuple):
    return tuple(item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple)


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    intrinsic: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    debug_iter: int


# ---- gradient sinks (extension; off unless a caller registers one AND activates them) ------------------------------
# The reference returns fresh gradient tensors and autograd then runs `param.grad += fresh` for each of the five
# Gaussian parameters: three more passes over 236 B/Gaussian per view.  A caller that accumulates over views (gradient
# accumulation, the view-parallel step of bags_view_parallel.py) can register the buffer a leaf's gradient is summed
# into; the backward kernel then adds into that buffer directly and this input's autograd slot returns None.
#
# Because a sink changes what `backward` does to a leaf (its `.grad` buffer is written by the kernel, tensor hooks and
# DDP hooks on the leaf do not fire, `torch.autograd.grad(...)` w.r.t. the leaf gets None), registered sinks are only
# consulted while they are ACTIVE: inside `with grad_sinks():` around the training forward+backward, or after
# `set_grad_sinks_active(True)`.  Evaluation renders, gradient probes and double-backward checks outside that scope see
# the reference's behaviour.
_GRAD_SINKS = {}        # (device index, data_ptr) -> (weakref to the leaf, buffer)
_SINKS_ACTIVE = False


def register_grad_sink(param, buffer):
    """Accumulate the rasterizer's gradient w.r.t. `param` (a leaf passed as means3D / shs / opacities / scales /
    rotations) directly into `buffer` (same shape, fp32, contiguous, same device) whenever sinks are active (see
    `grad_sinks`).  Typically buffer is param.grad."""
    import weakref
    if not (param.is_cuda and buffer.is_cuda and param.device == buffer.device):
        raise RuntimeError("bags_b200: gradient sinks need CUDA tensors on one device (there is no CPU path)")
    if tuple(param.shape) != tuple(buffer.shape) or buffer.dtype != torch.float32 or not buffer.is_contiguous():
        raise RuntimeError("bags_b200: a gradient sink must be a contiguous fp32 tensor of the parameter's shape")
    _GRAD_SINKS[(param.device.index, param.data_ptr())] = (weakref.ref(param), buffer)


def unregister_grad_sink(param):
    _GRAD_SINKS.pop((param.device.index, param.data_ptr()), None)


def set_grad_sinks_active(active):
    """Process-wide switch for the registered sinks; returns the previous state."""
    global _SINKS_ACTIVE
    prev, _SINKS_ACTIVE = _SINKS_ACTIVE, bool(active)
    return prev


@contextlib.contextmanager
def grad_sinks(active=True):
    """`with grad_sinks():` -- forwards issued inside use the registered gradient sinks in their backward."""
    prev = set_grad_sinks_active(active)
    try:
        yield
    finally:
        set_grad_sinks_active(prev)


def _sink_of(t):
    if t is None or not t.is_cuda or t.numel() == 0 or not t.requires_grad:
        return None
    ent = _GRAD_SINKS.get((t.device.index, t.data_ptr()))
    if ent is None:
        return None
    ref, buf = ent
    p = ref()
    if p is None or p.data_ptr() != t.data_ptr() or tuple(p.shape) != tuple(t.shape) or not t.is_leaf:
        return None
    return buf


# ---- asynchronous forward (extension) -----------------------------------------------------------------------------
class AsyncForward:
    """Forward passes issued inside `with AsyncForward(capacity) as af:` never wait for the GPU
    (bags_rasterize_forward_async): the binning buffer and the binning grids are sized by `capacity` tile instances
    instead of by a num_rendered the host has to read back (the reference blocks on that copy in every forward,
    rasterizer_impl.cu:291).  The whole view -- forward, loss, backward -- becomes a pure enqueue, so the host can run
    ahead of the GPU by any number of views and the sequence can be captured in a CUDA graph.

    The actual count of every forward lands, in stream order, in `af.counts` (pinned int64[slots], slot k % slots for
    the k-th forward of the scope, -1 = not written yet).  A count above the capacity means that view's tile lists
    were truncated and its image and gradients are wrong: call `af.check()` (after a synchronisation point of your
    choice -- e.g. where the loss is read back, one step late) to raise on that, and re-run with a larger capacity.
    `suggested_capacity()` gives 1.25 x the largest count seen + 64 Ki."""

    _current = None

    def __init__(self, capacity, slots=64, device=None):
        self.capacity = int(capacity)
        if self.capacity < 0 or self.capacity >= (1 << 30):
            raise ValueError("capacity must be in [0, 2^30)")
        self.slots = int(slots)
        self.counts = torch.full((self.slots,), -1, dtype=torch.int64).pin_memory()
        self._k = 0
        self._prev = None
        self.max_seen = 0

    def __enter__(self):
        self._prev, AsyncForward._current = AsyncForward._current, self
        return self

    def __exit__(self, *exc):
        AsyncForward._current = self._prev
        return False

    def _next_slot(self):
        s = self.counts[self._k % self.slots:self._k % self.slots + 1]
        self._k += 1
        return s

    def check(self):
        """Raises if any finished forward of this scope overflowed its capacity.  Only looks at host memory: call it
        after the stream (or an event behind the forwards of interest) has been synchronised."""
        m = int(self.counts.max())
        self.max_seen = max(self.max_seen, m)
        if m > self.capacity:
            raise RuntimeError(f"bags_b200: asynchronous forward overflowed its binning capacity ({m} tile instances > "
                               f"capacity {self.capacity}); the images / gradients of that view are truncated -- "
                               f"re-run with capacity >= {self.suggested_capacity()}")
        return m

    def suggested_capacity(self):
        m = max(self.max_seen, int(self.counts.max()))
        return int(m * 1.25) + (64 << 10)


def rasterize_gaussians(means3D, means2D, means2D_densify, shift_factors, sh, colors_precomp, opacities, scales,
                        rotations, cov3Ds_precomp, raster_settings):
    sinks = None
    if _SINKS_ACTIVE and _GRAD_SINKS and torch.is_grad_enabled():
        sinks = {k: b for k, b in (("means3D", _sink_of(means3D)), ("sh", _sink_of(sh)),
                                   ("opacity", _sink_of(opacities)), ("scales", _sink_of(scales)),
                                   ("rotations", _sink_of(rotations))) if b is not None} or None
    return _RasterizeGaussians.apply(
        means3D, means2D, means2D_densify, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
        raster_settings.viewmatrix, raster_settings.campos, raster_settings.projmatrix, shift_factors,
        raster_settings, sinks)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, means2D_densify, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, world_view, camera_center, full_proj, shift_factors, raster_settings, sinks=None):
        rs = raster_settings
        args = (rs.bg, means3D, shift_factors, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, world_view, full_proj, rs.intrinsic, rs.tanfovx, rs.tanfovy, rs.image_height,
                rs.image_width, sh, rs.sh_degree, camera_center, rs.prefiltered, rs.debug)
        af = AsyncForward._current
        if af is not None:
            out = _C.rasterize_gaussians_async(*args[:-1], af.capacity, af._next_slot())
        elif rs.debug:
            # (the reference's debug branch unpacks the wrong tuple arity and cannot run; this one works)
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _C.rasterize_gaussians(*args)
            except Exception:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            out = _C.rasterize_gaussians(*args)
        num_rendered, color, depth, weights, radii, geomBuffer, binningBuffer, imgBuffer, means2Dx, means2Dy = out

        mean2D = torch.cat((means2Dx, means2Dy), dim=-1)
        # undefined upstream gradients arrive as None instead of freshly zero-filled images (train.py's loss only
        # touches the colour image): the backward passes NULL for them and the kernel reads zeros
        ctx.set_materialize_grads(False)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.camera_center = camera_center
        ctx.full_proj = full_proj
        ctx.world_view = world_view
        ctx.grad_sinks = sinks
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer, weights)
        return color, radii, depth, weights, mean2D

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_depth, grad_out_weight, grad_mean2D):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer,
         weights) = ctx.saved_tensors
        if grad_out_color is None:
            if grad_out_depth is None and grad_out_weight is None:
                return (None,) * 15                   # nothing differentiable was used (mean2D / radii carry no gradient)
            grad_out_color = torch.zeros((3, rs.image_height, rs.image_width), dtype=torch.float32,
                                         device=means3D.device)
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
                ctx.world_view, ctx.full_proj, rs.intrinsic, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width,
                grad_out_color, grad_out_depth, grad_out_weight, sh, rs.sh_degree, ctx.camera_center, geomBuffer,
                ctx.num_rendered, binningBuffer, imgBuffer, weights, rs.debug)
        sk = ctx.grad_sinks or {}
        sinks = (sk.get("means3D"), sk.get("sh"), sk.get("opacity"), sk.get("scales"), sk.get("rotations"))
        if rs.debug:
            cpu_args = cpu_deep_copy_tuple(args)
            try:
                out = _C.rasterize_gaussians_backward(*args, *sinks)
            except Exception:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        else:
            out = _C.rasterize_gaussians_backward(*args, *sinks)
        (grad_means2D, grad_means2D_densify, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp,
         grad_sh, grad_scales, grad_rotations, grad_camera_center, grad_full_proj, grad_world_view,
         grad_shift_factors, covariance) = out

        grad_full_proj = grad_full_proj[:, None].reshape(4, 4)
        grad_world_view = grad_world_view[:, None].reshape(4, 4)

        return (grad_means3D, grad_means2D, grad_means2D_densify, grad_sh, grad_colors_precomp, grad_opacities,
                grad_scales, grad_rotations, grad_cov3Ds_precomp, grad_world_view, grad_camera_center,
                grad_full_proj, grad_shift_factors, None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, means2D_densify, shift_factors, opacities, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        if shs is None:
            shs = torch.Tensor([])
        if colors_precomp is None:
            colors_precomp = torch.Tensor([])
        if scales is None:
            scales = torch.Tensor([])
        if rotations is None:
            rotations = torch.Tensor([])
        if cov3D_precomp is None:
            cov3D_precomp = torch.Tensor([])
        return rasterize_gaussians(means3D, means2D, means2D_densify, shift_factors, shs, colors_precomp, opacities,
                                   scales, rotations, cov3D_precomp, rs)


# ---- B views of the same Gaussians as one launch set (extension) ----------------------
class GaussianBatchRasterizer(nn.Module):
    """Renders B cameras over the same Gaussians in ONE launch set (C ABI: bags_rasterize_forward_batch /
    bags_rasterize_backward_batch_*).  The reference renders its five cubemap faces with five render() calls
    (utils/cubemap_utils.py:229,265); this module takes the five `GaussianRasterizationSettings` at once:

        rasterizer = GaussianBatchRasterizer([settings_0, ..., settings_{B-1}], capacity)
        color, radii, depth, weights, mean2D = rasterizer(means3D, means2D, means2D_densify, shift_factors, opacities,
                                                          shs=..., scales=..., rotations=...)

    with `means2D` / `means2D_densify` of shape (B, P, 3) (the per-view gradient holders of render()), `shift_factors`
    (3,) shared or (B, 3), and outputs stacked along a leading view axis: color (B,3,H,W), radii (B,P), depth / weights
    (B,1,H,W), mean2D (B,P,2).  View v's slice is bit-identical to what `GaussianRasterizer(settings_v)` returns; the
    gradients w.r.t. the Gaussian parameters are the sums over the views, the per-camera gradients (view matrix, camera
    centre, projection, pupil shift) flow back to each settings' own tensors.

    The views must share image size, background, scale modifier, SH degree.  `capacity` = tile instances per view the
    binning buffers are sized for (the asynchronous protocol of `AsyncForward`: nothing waits on the host; counts land
    in `self.counts`, pinned int64[B]; `check()` raises on an overflow once the stream has been synchronised).
    `backward_slices` / `on_slice`: the per-Gaussian part of the backward runs in that many slices of consecutive
    Gaussians and `on_slice(first, count)` is called after each has been enqueued -- the rows [first, first+count) of the
    Gaussian-parameter gradients are final at that point of the stream (view-parallel trainers start the all-reduce of
    the slice there, see bags_view_parallel.SlicedGradientReducer).
    `sink_overwrite`: with active gradient sinks, this batch's summed gradients REPLACE the contents of the sink buffers
    (every row is written, zeros for the Gaussians no view sees) instead of being added onto them -- for the first, or
    only, batch of an accumulation step: the buffer needs no zeroing beforehand and is not read back."""

    def __init__(self, raster_settings_list, capacity, backward_slices=1, on_slice=None, sink_overwrite=False):
        super().__init__()
        rs = list(raster_settings_list)
        if not 1 <= len(rs) <= _C.MAX_BATCH:
            raise ValueError(f"a batch holds 1..{_C.MAX_BATCH} views, got {len(rs)}")
        r0 = rs[0]
        for r in rs[1:]:
            if (r.image_height, r.image_width, r.scale_modifier, r.sh_degree, r.prefiltered) != \
                    (r0.image_height, r0.image_width, r0.scale_modifier, r0.sh_degree, r0.prefiltered) or \
                    (r.bg is not r0.bg and not torch.equal(r.bg, r0.bg)):
                raise ValueError("the views of a batch must share image size, background, scale modifier, SH degree")
        self.raster_settings = rs
        self.capacity = int(capacity)
        if self.capacity < 0 or self.capacity >= (1 << 30):
            raise ValueError("capacity must be in [0, 2^30)")
        self.backward_slices = max(1, int(backward_slices))
        self.on_slice = on_slice
        self.sink_overwrite = bool(sink_overwrite)
        counts = torch.full((len(rs),), -1, dtype=torch.int64)
        self.counts = counts.pin_memory() if torch.cuda.is_available() else counts    # (no device: construction only)

    def check(self):
        m = int(self.counts.max())
        if m > self.capacity:
            raise RuntimeError(f"bags_b200: batched forward overflowed its binning capacity ({m} tile instances > "
                               f"capacity {self.capacity}); re-run with capacity >= {int(m * 1.25) + (64 << 10)}")
        return m

    def forward(self, means3D, means2D, means2D_densify, shift_factors, opacities, shs=None, colors_precomp=None,
                scales=None, rotations=None, cov3D_precomp=None):
        rs = self.raster_settings
        B = len(rs)
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        world_view = torch.stack([r.viewmatrix for r in rs])
        full_proj = torch.stack([r.projmatrix for r in rs])
        campos = torch.stack([r.campos.reshape(3) for r in rs])
        if shift_factors.dim() == 1:
            shift_factors = shift_factors.unsqueeze(0).expand(B, 3)
        sinks = None
        if _SINKS_ACTIVE and _GRAD_SINKS and torch.is_grad_enabled():
            sinks = {k: b for k, b in (("means3D", _sink_of(means3D)), ("sh", _sink_of(shs)),
                                       ("opacity", _sink_of(opacities)), ("scales", _sink_of(scales)),
                                       ("rotations", _sink_of(rotations))) if b is not None} or None
        return _RasterizeGaussiansBatch.apply(means3D, means2D, means2D_densify, shs, colors_precomp, opacities, scales,
                                              rotations, cov3D_precomp, world_view, campos, full_proj, shift_factors,
                                              self, sinks)


class _RasterizeGaussiansBatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, means2D_densify, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, world_view, camera_center, full_proj, shift_factors, module, sinks=None):
        rs = module.raster_settings
        r0 = rs[0]
        tanx = [float(r.tanfovx) for r in rs]
        tany = [float(r.tanfovy) for r in rs]
        out = _C.rasterize_gaussians_batch(r0.bg, means3D, shift_factors.contiguous(), colors_precomp, opacities, scales,
                                           rotations, r0.scale_modifier, cov3Ds_precomp, world_view, full_proj, tanx,
                                           tany, r0.image_height, r0.image_width, sh, r0.sh_degree, camera_center,
                                           r0.prefiltered, module.capacity, module.counts)
        color, depth, weights, radii, geomBuffer, binningBuffer, imgBuffer, means2Dx, means2Dy = out
        mean2D = torch.cat((means2Dx, means2Dy), dim=-1)
        ctx.set_materialize_grads(False)
        ctx.module = module
        ctx.tan = (tanx, tany)
        ctx.grad_sinks = sinks
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer, world_view, camera_center, full_proj)
        ctx.mark_non_differentiable(radii, mean2D)
        return color, radii, depth, weights, mean2D

    @staticmethod
    def backward(ctx, grad_out_color, _, grad_out_depth, grad_out_weight, grad_mean2D):
        module = ctx.module
        r0 = module.raster_settings[0]
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer,
         world_view, camera_center, full_proj) = ctx.saved_tensors
        B, P = radii.shape
        dev = means3D.device
        if grad_out_color is None:
            if grad_out_depth is None and grad_out_weight is None:
                return (None,) * 15
            grad_out_color = torch.zeros((B, 3, r0.image_height, r0.image_width), dtype=torch.float32, device=dev)
        scratch = _C.rasterize_gaussians_backward_batch_blend(r0.bg, P, module.capacity, grad_out_color, grad_out_depth,
                                                              grad_out_weight, geomBuffer, binningBuffer, imgBuffer)
        sk = ctx.grad_sinks or {}
        M = sh.shape[1] if sh.numel() else 0
        f32 = dict(dtype=torch.float32, device=dev)
        acc = 0

        def out(name, bit, shape):
            nonlocal acc
            b = sk.get(name)
            if b is None:
                return torch.empty(shape, **f32), False
            if not module.sink_overwrite:
                acc |= bit
            return b, True
        g_means3D, s_means = out("means3D", _C.ACC_MEANS3D, (P, 3))
        g_sh, s_sh = out("sh", _C.ACC_SH, (P, M, 3))
        g_op, s_op = out("opacity", _C.ACC_OPACITY, (P, 1))
        g_sc, s_sc = out("scales", _C.ACC_SCALE, (P, 3))
        g_rot, s_rot = out("rotations", _C.ACC_ROT, (P, 4))
        g_m2d = torch.empty((B, P, 3), **f32)
        g_m2d_dens = torch.empty((B, P, 3), **f32)
        pose = torch.empty((38 * B,), **f32)
        g_col = torch.empty((P, 3), **f32) if colors_precomp.numel() else None
        g_cov = torch.empty((P, 6), **f32) if cov3Ds_precomp.numel() else None
        n = module.backward_slices if P > 0 else 1
        bounds = [(P * k) // n for k in range(n + 1)]
        for k in range(n):
            first, count = bounds[k], bounds[k + 1] - bounds[k]
            if count == 0 and k < n - 1:
                continue
            _C.rasterize_gaussians_backward_batch_gaussians(
                means3D, radii, scales, rotations, r0.scale_modifier, cov3Ds_precomp, world_view, full_proj,
                ctx.tan[0], ctx.tan[1], r0.image_height, r0.image_width, sh, r0.sh_degree, camera_center, geomBuffer,
                scratch, g_m2d, g_m2d_dens, g_op, g_means3D, g_sh, g_sc, g_rot, pose, g_col, g_cov, None, acc,
                first, count)
            if module.on_slice is not None:
                module.on_slice(first, count)
        g_proj = pose[:16 * B].view(B, 4, 4)
        g_view = pose[16 * B:32 * B].view(B, 4, 4)
        g_campos = pose[32 * B:35 * B].view(B, 3)
        g_shift = pose[35 * B:38 * B].view(B, 3)
        return (None if s_means else g_means3D, g_m2d, g_m2d_dens, None if s_sh else (g_sh if M else None), g_col,
                None if s_op else g_op, None if s_sc else (g_sc if scales.numel() else None),
                None if s_rot else (g_rot if rotations.numel() else None), g_cov, g_view, g_campos, g_proj, g_shift,
                None, None)


def compute_relocation(opacity_old, scale_old, N, binoms, n_max):
    new_opacity, new_scale = _C.compute_relocation(opacity_old, scale_old, N.int(), binoms, n_max)
    return new_opacity, new_scale
