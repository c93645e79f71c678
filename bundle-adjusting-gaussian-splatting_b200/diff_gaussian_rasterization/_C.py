"""Stand-in for the reference's pybind module ``diff_gaussian_rasterization._C`` (3dgs-pose/ext.cpp:15-20).

Same four functions, same argument order and return tuples as RasterizeGaussiansCUDA,
RasterizeGaussiansBackwardCUDA, markVisible and ComputeRelocationCUDA (3dgs-pose/rasterize_points.cu:36-325), but
the work is done by libbags_b200.so through its C ABI: this file only allocates torch tensors and passes raw device
pointers + the current CUDA stream.
"""
import ctypes

import torch

import bags_capi as capi

NUM_CHANNELS = 3


def _ptr(t):
    """Device pointer of a tensor, or NULL for the 0-element placeholders the Python wrapper passes for absent
    optional inputs (mirrors data_ptr() of an empty tensor in the reference binding)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()          # a plain int: ctypes converts it for the c_void_p parameters of bags_capi.SIGNATURES


def _f32c(t, device):
    if t is None or t.numel() == 0:
        return t
    if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


class _Resizable:
    """Growable uint8 CUDA buffer handed to the library as an allocation callback (resizeFunctional,
    rasterize_points.cu:28-34)."""

    def __init__(self, device):
        self.t = torch.empty(0, dtype=torch.uint8, device=device)
        self.cb = capi.ALLOC_FN(self._alloc)

    def _alloc(self, nbytes, _user):
        self.t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.t.device)
        return self.t.data_ptr()


def rasterize_gaussians(background, means3D, shift_factors, colors, opacity, scales, rotations, scale_modifier,
                        cov3D_precomp, viewmatrix, projmatrix, intrinsic, tan_fovx, tan_fovy, image_height,
                        image_width, sh, degree, campos, prefiltered, debug):
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    if not means3D.is_cuda:
        raise RuntimeError("bags_b200: means3D must be a CUDA tensor (there is no CPU path)")
    dev = means3D.device
    L = capi.lib()
    P = means3D.size(0)
    H, W = int(image_height), int(image_width)
    with torch.cuda.device(dev):
        means3D = _f32c(means3D, dev)
        background = _f32c(background, dev)
        shift_factors = _f32c(shift_factors, dev)
        colors = _f32c(colors, dev)
        opacity = _f32c(opacity, dev)
        scales = _f32c(scales, dev)
        rotations = _f32c(rotations, dev)
        cov3D_precomp = _f32c(cov3D_precomp, dev)
        viewmatrix = _f32c(viewmatrix, dev)
        projmatrix = _f32c(projmatrix, dev)
        intrinsic = _f32c(intrinsic, dev)
        sh = _f32c(sh, dev)
        campos = _f32c(campos, dev)

        f32 = dict(dtype=torch.float32, device=dev)
        out_color = torch.empty((NUM_CHANNELS, H, W), **f32)
        out_depth = torch.empty((1, H, W), **f32)
        out_alpha = torch.empty((1, H, W), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        means2Dx = torch.empty((P, 1), **f32)
        means2Dy = torch.empty((P, 1), **f32)
        geom, binning, img = _Resizable(dev), _Resizable(dev), _Resizable(dev)

        M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rendered = L.bags_rasterize_forward(
            geom.cb, None, binning.cb, None, img.cb, None,
            P, int(degree), int(M),
            _ptr(background), W, H,
            _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(opacity),
            _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
            _ptr(viewmatrix), _ptr(projmatrix), _ptr(intrinsic), _ptr(campos), _ptr(shift_factors),
            float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
            _ptr(out_color), _ptr(out_depth), _ptr(out_alpha), _ptr(radii), _ptr(means2Dx), _ptr(means2Dy),
            int(bool(debug)), stream)
        capi.check(rendered, "bags_rasterize_forward")
    return (rendered, out_color, out_depth, out_alpha, radii, geom.t, binning.t, img.t, means2Dx, means2Dy)


ACC_MEANS3D, ACC_SH, ACC_OPACITY, ACC_SCALE, ACC_ROT = 1, 2, 4, 8, 16     # BAGS_ACC_* of include/bags_b200.h


def _sink_ok(buf, shape, dev):
    return (buf is not None and buf.is_cuda and buf.device == dev and buf.dtype == torch.float32
            and buf.is_contiguous() and tuple(buf.shape) == tuple(shape))


def rasterize_gaussians_backward(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                 cov3D_precomp, viewmatrix, projmatrix, intrinsic, tan_fovx, tan_fovy, image_height,
                                 image_width, dL_dout_color, dL_dout_depth, dL_dout_alpha, sh, degree, campos,
                                 geomBuffer, R, binningBuffer, imageBuffer, alphas, debug, sinks=None):
    """`sinks` (extension, default None = the reference's behaviour): dict name -> fp32 CUDA buffer for any of
    means3D / sh / opacity / scales / rotations.  The gradient of a listed input is ADDED into its buffer by the
    kernel (bags_rasterize_backward_accum) and None is returned in its slot of the tuple."""
    dev = means3D.device
    L = capi.lib()
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    with torch.cuda.device(dev):
        means3D = _f32c(means3D, dev)
        background = _f32c(background, dev)
        colors = _f32c(colors, dev)
        scales = _f32c(scales, dev)
        rotations = _f32c(rotations, dev)
        cov3D_precomp = _f32c(cov3D_precomp, dev)
        viewmatrix = _f32c(viewmatrix, dev)
        projmatrix = _f32c(projmatrix, dev)
        intrinsic = _f32c(intrinsic, dev)
        sh = _f32c(sh, dev)
        campos = _f32c(campos, dev)
        dL_dout_color = _f32c(dL_dout_color, dev)
        dL_dout_depth = _f32c(dL_dout_depth, dev)
        dL_dout_alpha = _f32c(dL_dout_alpha, dev)
        radii = radii.contiguous()

        M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
        f32 = dict(dtype=torch.float32, device=dev)
        sinks = sinks or {}
        acc = 0

        def out(name, bit, shape):
            nonlocal acc
            buf = sinks.get(name)
            if buf is None:
                return torch.empty(shape, **f32), False
            if not _sink_ok(buf, shape, dev):
                raise RuntimeError(f"bags_b200: gradient sink for {name} must be a contiguous fp32 CUDA tensor of "
                                   f"shape {tuple(shape)} on {dev}")
            acc |= bit
            return buf, True

        dL_dmeans3D, s_means = out("means3D", ACC_MEANS3D, (P, 3))
        dL_dmeans2D = torch.empty((P, 3), **f32)
        dL_dmeans2D_densify = torch.empty((P, 3), **f32)
        dL_dcolors = torch.empty((P, NUM_CHANNELS), **f32)
        covariance = torch.empty((P, 1), **f32)
        dL_dopacity, s_op = out("opacity", ACC_OPACITY, (P, 1))
        dL_dcov3D = torch.empty((P, 6), **f32)
        dL_dsh, s_sh = out("sh", ACC_SH, (P, M, 3))
        dL_dscales, s_sc = out("scales", ACC_SCALE, (P, 3))
        dL_drotations, s_rot = out("rotations", ACC_ROT, (P, 4))
        pose = torch.empty((38,), **f32)  # proj 16 | view 16 | campos 3 | shift 3
        scratch = torch.empty((int(L.bags_backward_scratch_bytes(P)),), dtype=torch.uint8, device=dev)

        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        pp = pose.data_ptr()
        rc = L.bags_rasterize_backward_accum(
            P, int(degree), int(M), int(R),
            _ptr(background), W, H,
            _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(alphas),
            _ptr(scales), float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp),
            _ptr(viewmatrix), _ptr(projmatrix), _ptr(intrinsic), _ptr(campos),
            float(tan_fovx), float(tan_fovy),
            _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
            _ptr(dL_dout_color), _ptr(dL_dout_depth), _ptr(dL_dout_alpha),
            _ptr(scratch),
            _ptr(dL_dmeans2D), _ptr(dL_dmeans2D_densify), _ptr(covariance), _ptr(dL_dopacity), _ptr(dL_dcolors),
            _ptr(dL_dmeans3D), _ptr(dL_dcov3D), _ptr(dL_dsh), _ptr(dL_dscales), _ptr(dL_drotations),
            ctypes.c_void_p(pp), ctypes.c_void_p(pp + 64), ctypes.c_void_p(pp + 128), ctypes.c_void_p(pp + 140),
            acc, int(bool(debug)), stream)
        capi.check(rc, "bags_rasterize_backward_accum")
    return (dL_dmeans2D, dL_dmeans2D_densify, dL_dcolors, None if s_op else dL_dopacity,
            None if s_means else dL_dmeans3D, dL_dcov3D, None if s_sh else dL_dsh, None if s_sc else dL_dscales,
            None if s_rot else dL_drotations, pose[32:35], pose[0:16], pose[16:32], pose[35:38], covariance)


def mark_visible(means3D, viewmatrix, projmatrix):
    dev = means3D.device
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=dev)
    if P != 0:
        with torch.cuda.device(dev):
            means3D = _f32c(means3D, dev)
            viewmatrix = _f32c(viewmatrix, dev)
            projmatrix = _f32c(projmatrix, dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = capi.lib().bags_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix), _ptr(present),
                                              stream)
            capi.check(rc, "bags_mark_visible")
    return present


def compute_relocation(opacity_old, scale_old, N, binoms, n_max):
    dev = opacity_old.device
    P = opacity_old.size(0)
    final_opacity = torch.zeros((P,), dtype=torch.float32, device=dev)
    final_scale = torch.zeros((3 * P,), dtype=torch.float32, device=dev)
    if P != 0:
        with torch.cuda.device(dev):
            opacity_old = _f32c(opacity_old, dev)
            scale_old = _f32c(scale_old, dev)
            binoms = _f32c(binoms, dev)
            N = N.to(device=dev, dtype=torch.int32).contiguous()
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = capi.lib().bags_compute_relocation(P, _ptr(opacity_old), _ptr(scale_old), _ptr(N), _ptr(binoms),
                                                    int(n_max), _ptr(final_opacity), _ptr(final_scale), stream)
            capi.check(rc, "bags_compute_relocation")
    return final_opacity, final_scale


def unpack_state(P, R, H, W, geomBuffer, binningBuffer, imageBuffer):
    """Test-only: private forward state as a dict of tensors (see bags_unpack_state in include/bags_b200.h)."""
    dev = geomBuffer.device
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    with torch.cuda.device(dev):
        out = {
            "tiles_touched": torch.empty((P,), dtype=torch.int32, device=dev),
            "depth_bits": torch.empty((P,), dtype=torch.int32, device=dev),
            "xy": torch.empty((P, 2), dtype=torch.float32, device=dev),
            "conic_opacity": torch.empty((P, 4), dtype=torch.float32, device=dev),
            "rgb": torch.empty((P, 3), dtype=torch.float32, device=dev),
            "point_list": torch.empty((R,), dtype=torch.int32, device=dev),
            "point_list_keys": torch.empty((R,), dtype=torch.int64, device=dev),
            "ranges": torch.empty((tiles, 2), dtype=torch.int32, device=dev),
            "n_contrib": torch.empty((H * W,), dtype=torch.int32, device=dev),
            "final_T": torch.empty((H * W,), dtype=torch.float32, device=dev),
        }
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = capi.lib().bags_unpack_state(
            P, int(R), W, H, _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
            _ptr(out["tiles_touched"]), _ptr(out["depth_bits"]), _ptr(out["xy"]), _ptr(out["conic_opacity"]),
            _ptr(out["rgb"]), _ptr(out["point_list"]), _ptr(out["point_list_keys"]), _ptr(out["ranges"]),
            _ptr(out["n_contrib"]), _ptr(out["final_T"]), stream)
        capi.check(rc, "bags_unpack_state")
    return out
