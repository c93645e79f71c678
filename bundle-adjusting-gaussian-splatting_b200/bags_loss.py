"""Fused photometric loss: drop-in for ``l1_loss`` / ``ssim`` of the reference's ``utils/loss_utils.py`` (lines 18-19
and 49-76) plus the combination of train.py:325, SURVEY.md section 8f row N2.

``l1_ssim(image, gt)`` evaluates both terms in ONE forward kernel and ONE backward kernel of ``libbags_b200.so``
(``csrc/loss.cu``); ``l1_loss`` and ``ssim`` keep the reference's names and signatures on top of the same kernels.
Reductions are fixed-order (deterministic).  There is no CPU path: CPU tensors raise."""
import ctypes

import torch

import bags_capi as capi


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _L1SSIM(torch.autograd.Function):
    """(img1, img2) -> (sum |img1 - img2| per plane, sum ssim_map per plane) as float64 [planes] tensors."""

    @staticmethod
    def forward(ctx, img1, img2, want_ssim):
        if not (img1.is_cuda and img2.is_cuda):
            raise RuntimeError("bags_b200: the fused loss needs CUDA tensors (there is no CPU path)")
        if img1.shape != img2.shape or img1.dim() not in (3, 4):
            raise RuntimeError("loss: expected two (C,H,W) or (N,C,H,W) tensors of the same shape")
        dev = img1.device
        a = img1.detach().to(torch.float32).contiguous()
        b = img2.detach().to(torch.float32).contiguous()
        H, W = a.shape[-2], a.shape[-1]
        planes = a.numel() // (H * W)
        need_grad = ctx.needs_input_grad[0]
        maps = torch.empty((3, planes, H, W), dtype=torch.float32, device=dev) if (need_grad and want_ssim) else None
        if img2.dtype != img1.dtype and ctx.needs_input_grad[1]:
            raise RuntimeError("loss: img1 and img2 must share a dtype when both require gradients")
        L = capi.lib()
        scratch = torch.empty((int(L.bags_loss_scratch_bytes(planes, H, W)),), dtype=torch.uint8, device=dev)
        sums = torch.empty((planes, 2), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = L.bags_loss_forward(_ptr(a), _ptr(b), planes, H, W, int(bool(want_ssim)),
                                     _ptr(maps[0]) if maps is not None else ctypes.c_void_p(0),
                                     _ptr(maps[1]) if maps is not None else ctypes.c_void_p(0),
                                     _ptr(maps[2]) if maps is not None else ctypes.c_void_p(0),
                                     _ptr(scratch), _ptr(sums), stream)
        capi.check(rc, "bags_loss_forward")
        ctx.save_for_backward(a, b, maps if maps is not None else torch.empty(0, device=dev))
        ctx.meta = (planes, H, W, bool(want_ssim), img1.shape, img1.dtype)
        ctx.scratch_bytes = scratch.numel()
        return sums[:, 0].clone(), sums[:, 1].clone()

    @staticmethod
    def backward(ctx, g_l1, g_ssim):
        a, b, maps = ctx.saved_tensors
        planes, H, W, want_ssim, shape, dtype = ctx.meta
        dev = a.device
        coef = torch.stack([g_l1.to(torch.float32), g_ssim.to(torch.float32)], dim=1).contiguous()   # [planes, 2]
        out = torch.empty_like(a)
        has_maps = maps.numel() > 0
        rc = 0
        if ctx.needs_input_grad[0]:
            with torch.cuda.device(dev):
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                rc = capi.lib().bags_loss_backward(_ptr(a), _ptr(b), planes, H, W,
                                                   _ptr(maps[0]) if has_maps else ctypes.c_void_p(0),
                                                   _ptr(maps[1]) if has_maps else ctypes.c_void_p(0),
                                                   _ptr(maps[2]) if has_maps else ctypes.c_void_p(0),
                                                   _ptr(coef), _ptr(out), stream)
        capi.check(rc, "bags_loss_backward")
        g1 = out.view(shape).to(dtype) if ctx.needs_input_grad[0] else None
        g2 = None
        if ctx.needs_input_grad[1]:
            # The reference's second argument often carries gradients too (apply2gt: the ground truth is a
            # grid_sample through the lens network; after start_vignetting it is multiplied by a learnable mask).
            # Both terms are symmetric in their arguments, so the gradient w.r.t. img2 is the same pair of kernels with
            # the arguments swapped: the forward once more for the three derivative maps taken w.r.t. img2's moments,
            # then the backward.
            L = capi.lib()
            maps2 = torch.empty((3, planes, H, W), dtype=torch.float32, device=dev) if want_ssim else None
            out2 = torch.empty_like(b)
            with torch.cuda.device(dev):
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                if want_ssim:
                    scratch = torch.empty((int(ctx.scratch_bytes),), dtype=torch.uint8, device=dev)
                    sums = torch.empty((planes, 2), dtype=torch.float64, device=dev)
                    rc = L.bags_loss_forward(_ptr(b), _ptr(a), planes, H, W, 1, _ptr(maps2[0]), _ptr(maps2[1]),
                                             _ptr(maps2[2]), _ptr(scratch), _ptr(sums), stream)
                    capi.check(rc, "bags_loss_forward")
                rc = L.bags_loss_backward(_ptr(b), _ptr(a), planes, H, W,
                                          _ptr(maps2[0]) if want_ssim else ctypes.c_void_p(0),
                                          _ptr(maps2[1]) if want_ssim else ctypes.c_void_p(0),
                                          _ptr(maps2[2]) if want_ssim else ctypes.c_void_p(0),
                                          _ptr(coef), _ptr(out2), stream)
            capi.check(rc, "bags_loss_backward")
            g2 = out2.view(shape).to(dtype)
        return g1, g2, None


def _sums(img1, img2, want_ssim):
    s_l1, s_ssim = _L1SSIM.apply(img1, img2, want_ssim)
    H, W = img1.shape[-2], img1.shape[-1]
    return s_l1, s_ssim, H * W


def l1_loss(network_output, gt):
    """utils/loss_utils.py:18-19: mean |network_output - gt| (gradients flow to both arguments, like the reference's
    torch expression)."""
    s_l1, _, _ = _sums(network_output, gt, False)
    return (s_l1.sum() / network_output.numel()).to(torch.float32)


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:49-76.  Only the reference's window (11, sigma 1.5) is implemented."""
    if window_size != 11:
        raise RuntimeError("bags_b200: ssim supports window_size=11 only (the value the reference uses)")
    _, s_ssim, hw = _sums(img1, img2, True)
    if size_average:
        return (s_ssim.sum() / img1.numel()).to(torch.float32)
    if img1.dim() != 4:
        raise IndexError("ssim(size_average=False) needs a (N,C,H,W) input, like the reference's .mean(1).mean(1).mean(1)")
    n, c = img1.shape[0], img1.shape[1]
    return (s_ssim.view(n, c).sum(1) / (c * hw)).to(torch.float32)


def l1_ssim(image, gt_image):
    """Both terms from one forward / one backward kernel: returns (Ll1, ssim) like calling l1_loss and ssim."""
    s_l1, s_ssim, _ = _sums(image, gt_image, True)
    n = image.numel()
    return (s_l1.sum() / n).to(torch.float32), (s_ssim.sum() / n).to(torch.float32)


def training_loss(image, gt_image, lambda_dssim=0.2):
    """train.py:325: (1 - lambda) * Ll1 + lambda * (1 - ssim).  Returns (loss, Ll1, ssim)."""
    Ll1, s = l1_ssim(image, gt_image)
    return (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - s), Ll1, s


class _L1U8(torch.autograd.Function):
    """(render fp32, ground truth uint8) -> mean |render - gt / 255| with its gradient formed in the same pass."""

    @staticmethod
    def forward(ctx, img, gt_u8):
        if not (img.is_cuda and gt_u8.is_cuda):
            raise RuntimeError("bags_b200: the fused loss needs CUDA tensors (there is no CPU path)")
        if gt_u8.dtype != torch.uint8 or img.shape != gt_u8.shape:
            raise RuntimeError("l1_loss_u8: expected a uint8 ground truth of the render's shape")
        dev = img.device
        a = img.detach().to(torch.float32).contiguous()
        b = gt_u8.contiguous()
        n = a.numel()
        L = capi.lib()
        grad = torch.empty_like(a)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        scratch = torch.empty((int(L.bags_l1_u8_scratch_bytes(n)),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = L.bags_l1_u8_forward(_ptr(a), _ptr(b), n, _ptr(grad), _ptr(loss), _ptr(scratch), stream)
        capi.check(rc, "bags_l1_u8_forward")
        ctx.save_for_backward(grad)
        ctx.meta = (img.shape, img.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        shape, dtype = ctx.meta
        return (grad * g).view(shape).to(dtype), None


def l1_loss_u8(network_output, gt_u8):
    """l1_loss (utils/loss_utils.py:18-19) against the ground truth as the data set stores it: uint8, scaled by
    1/255 inside the kernel (what utils/general_utils.py:PILtoTorch does once at load time).  One kernel produces the
    loss and its gradient; the ground truth stays 8-bit in HBM (4x less memory and H2D traffic per view)."""
    return _L1U8.apply(network_output, gt_u8)
