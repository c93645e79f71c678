"""Stand-in for the reference's pybind module ``simple_knn._C`` (simple-knn/ext.cpp:15-17, spatial.cu:14-26)."""
import ctypes

import torch

import bags_capi as capi


def distCUDA2(points):
    if not points.is_cuda:
        raise RuntimeError("bags_b200: distCUDA2 needs a CUDA tensor (there is no CPU path)")
    dev = points.device
    P = points.size(0)
    means = torch.zeros((P,), dtype=torch.float32, device=dev)
    if P == 0:
        return means
    with torch.cuda.device(dev):
        pts = points.to(torch.float32).contiguous()
        L = capi.lib()
        scratch = torch.empty((int(L.bags_knn_scratch_bytes(P)),), dtype=torch.uint8, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = L.bags_knn_mean_dist2(P, ctypes.c_void_p(pts.data_ptr()), ctypes.c_void_p(means.data_ptr()),
                                   ctypes.c_void_p(scratch.data_ptr()), stream)
        capi.check(rc, "bags_knn_mean_dist2")
    return means
