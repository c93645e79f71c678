"""ctypes binding of libbags_b200.so -- the C ABI declared in include/bags_b200.h.

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os
from ctypes import (CFUNCTYPE, POINTER, c_char_p, c_float, c_int, c_size_t, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbags_b200.so")

ALLOC_FN = CFUNCTYPE(c_void_p, c_size_t, c_void_p)

_lib = None

# every symbol include/bags_b200.h declares: (restype, argtypes)
_P = c_void_p
SIGNATURES = {
    "bags_last_error": (c_char_p, []),
    "bags_version": (c_int, []),
    "bags_geom_bytes": (c_size_t, [c_int]),
    "bags_image_bytes": (c_size_t, [c_int, c_int]),
    "bags_binning_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "bags_backward_scratch_bytes": (c_size_t, [c_int]),
    "bags_rasterize_forward": (c_int, [ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P,      # allocators
                                       c_int, c_int, c_int,                          # P, D, M
                                       _P, c_int, c_int,                             # background, width, height
                                       _P, _P, _P, _P,                               # means3D, shs, colors, opacities
                                       _P, c_float, _P, _P,                          # scales, mod, rotations, cov3D
                                       _P, _P, _P, _P, _P,                           # view, proj, intrinsic, campos, shift
                                       c_float, c_float, c_int,                      # tanfovx, tanfovy, prefiltered
                                       _P, _P, _P, _P, _P, _P,                       # color, depth, alpha, radii, m2Dx, m2Dy
                                       c_int, _P]),                                  # debug, stream
    "bags_rasterize_forward_async": (c_int, [ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P, c_int, c_int, c_int,
                                             _P, c_int, c_int, _P, _P, _P, _P, _P, c_float, _P, _P,
                                             _P, _P, _P, _P, _P, c_float, c_float, c_int,
                                             _P, _P, _P, _P, _P, _P,
                                             c_int, _P, _P]),                        # capacity, count out, stream
    "bags_rasterize_backward": (c_int, [c_int, c_int, c_int, c_int,                  # P, D, M, R
                                        _P, c_int, c_int,                            # background, width, height
                                        _P, _P, _P, _P,                              # means3D, shs, colors, alphas
                                        _P, c_float, _P, _P,                         # scales, mod, rotations, cov3D
                                        _P, _P, _P, _P,                              # view, proj, intrinsic, campos
                                        c_float, c_float,                            # tanfovx, tanfovy
                                        _P, _P, _P, _P,                              # radii, geom, binning, image
                                        _P, _P, _P,                                  # dL_dpix, dL_ddepth, dL_dalpha
                                        _P,                                          # scratch
                                        _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,      # dense outputs
                                        _P, _P, _P, _P,                              # proj, view, campos, shift
                                        c_int, _P]),                                 # debug, stream
    "bags_rasterize_backward_accum": (c_int, [c_int, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P, _P,
                                              _P, c_float, _P, _P, _P, _P, _P, _P, c_float, c_float,
                                              _P, _P, _P, _P, _P, _P, _P, _P,
                                              _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                              ctypes.c_uint, c_int, _P]),   # + accumulate_mask before debug
    "bags_rasterize_forward_batch": (c_int, [c_int, ALLOC_FN, _P, ALLOC_FN, _P, ALLOC_FN, _P, c_int, c_int, c_int,
                                             _P, c_int, c_int, _P, _P, _P, _P, _P, c_float, _P, _P,
                                             _P, _P, _P, _P,                         # view, proj, campos, shift: [B][..]
                                             _P, _P, c_int,                          # tan_fovx, tan_fovy (host [B]), prefiltered
                                             _P, _P, _P, _P, _P, _P,
                                             c_int, _P, _P]),                        # capacity, counts out [B], stream
    "bags_backward_batch_scratch_bytes": (c_size_t, [c_int, c_int]),
    "bags_rasterize_backward_batch_blend": (c_int, [c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P,
                                                    _P]),
    "bags_rasterize_backward_batch_gaussians": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int,
                                                        _P, _P, _P, c_float, _P, _P,     # means3D, shs, scales, mod, rot, cov3D
                                                        _P, _P, _P, _P, _P,              # view, proj, campos, tanx, tany
                                                        _P, _P, _P,                      # radii, geom, scratch
                                                        _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,   # dense outputs
                                                        _P, _P, _P, _P,                  # proj, view, campos, shift
                                                        ctypes.c_uint, c_int, c_int, _P]),
    "bags_allreduce_p2p_flag_bytes": (c_size_t, []),
    "bags_enable_peer_access": (c_int, [c_int]),
    "bags_ipc_open": (c_int, [c_char_p, POINTER(c_void_p)]),
    "bags_ipc_close": (c_int, [_P]),
    "bags_allreduce_p2p": (c_int, [c_int, c_int, _P, _P, ctypes.c_longlong, _P]),
    "bags_mark_visible": (c_int, [c_int, _P, _P, _P, _P, _P]),
    "bags_compute_relocation": (c_int, [c_int, _P, _P, _P, _P, c_int, _P, _P, _P]),
    "bags_knn_scratch_bytes": (c_size_t, [c_int]),
    "bags_knn_mean_dist2": (c_int, [c_int, _P, _P, _P, _P]),
    "bags_distort_forward": (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, _P, _P, _P, _P]),
    "bags_distort_backward": (c_int, [_P, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      _P, _P, _P, _P]),
    "bags_center_crop_forward": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "bags_center_crop_backward": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "bags_loss_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "bags_loss_window": (c_int, [c_void_p]),
    "bags_l1_u8_scratch_bytes": (c_size_t, [ctypes.c_longlong]),
    "bags_l1_u8_forward": (c_int, [_P, _P, ctypes.c_longlong, _P, _P, _P, _P]),
    "bags_adam_step": (c_int, [c_int, c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_double, _P]),
    "bags_densify_gather": (c_int, [ctypes.c_longlong, c_int, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_int,
                                    c_void_p, _P]),
    "bags_loss_forward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "bags_loss_backward": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "bags_unpack_state": (c_int, [c_int, c_int, c_int, c_int, _P, _P, _P,
                                  _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "bags_launch_count": (ctypes.c_ulonglong, []),
    "bags_profile_enable": (c_int, [c_int]),
    "bags_profile_read": (c_int, [_P, _P]),
    "bags_set_tuning": (c_int, [c_int, c_int]),
    "bags_sort_scratch_bytes": (c_size_t, [c_int]),
    "bags_sort_pairs_u32": (c_int, [c_int, _P, _P, _P, _P, c_int, c_int, _P, _P]),
}


class AdamTensor(ctypes.Structure):
    """struct bags_adam_tensor of include/bags_b200.h"""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("n", ctypes.c_longlong), ("step_size", ctypes.c_double), ("bias_correction2_sqrt", ctypes.c_double)]


class DensifyTensor(ctypes.Structure):
    """struct bags_densify_tensor of include/bags_b200.h"""
    _fields_ = [("src", c_void_p), ("dst", c_void_p), ("src_m", c_void_p), ("dst_m", c_void_p), ("src_v", c_void_p),
                ("dst_v", c_void_p), ("width", c_int), ("role", c_int)]


STAGES = ["preprocess", "depth_sort", "emit", "tile_sort", "ranges", "blend_fwd", "bwd_clear", "blend_bwd",
          "preprocess_bwd"]


def profile_read():
    """Returns ({stage: total_ms}, {stage: calls}) accumulated since the last read."""
    n = len(STAGES)
    ms = (ctypes.c_double * n)()
    calls = (ctypes.c_ulonglong * n)()
    check(lib().bags_profile_read(ms, calls), "bags_profile_read")
    return {STAGES[i]: ms[i] for i in range(n)}, {STAGES[i]: int(calls[i]) for i in range(n)}


class BagsError(RuntimeError):
    pass


def lib():
    """Loads the library on first use.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BagsError(
                f"{LIB_PATH} not found: build it with `python {os.path.join(_HERE, 'build.py')}` "
                "(there is no CPU or PyTorch fallback for the rasterizer)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc < 0:
        msg = lib().bags_last_error()
        raise BagsError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
    return rc
