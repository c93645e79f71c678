"""Fused lens-distortion post-step: drop-in for ``apply_distortion`` / ``center_crop`` of the reference's
``utils/util_distortion.py`` (lines 271-311 and 58-77), SURVEY.md section 8f row N1.

The reference runs, on every rendered canvas of an ``--outside_rasterizer`` iteration,
``F.interpolate`` (control-point flow -> canvas resolution), ``F.grid_sample`` (render -> distorted canvas),
``center_crop`` (a second ``F.grid_sample``) and a mask, each a full-canvas pass with its own backward.  Here the
chain is one gather kernel forward and one scatter kernel backward in ``libbags_b200.so``
(``csrc/distort.cu``, C ABI ``bags_distort_forward`` / ``bags_distort_backward``); the upsampled flow and the
intermediate canvas are never written to memory.  Same names, argument order, return values and error behaviour as
the reference functions.  There is no CPU path: CPU tensors raise.
"""
import ctypes

import torch

import bags_capi as capi


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _check_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("bags_b200: the distortion post-step needs CUDA tensors (there is no CPU path)")


class _DistortFunction(torch.autograd.Function):
    """(image [C,H,W], flow [Hs,Ws,2]) -> (out [C,Hc,Wc], mask [1,Hc,Wc], flow_full [Hf,Wf,2] or empty)."""

    @staticmethod
    def forward(ctx, image, flow, Hf, Wf, Hc, Wc, crop, mask_mode, want_flow):
        _check_cuda(image, flow)
        if image.dim() != 3 or flow.dim() != 3 or flow.size(2) != 2:
            raise RuntimeError("distort: expected image (C,H,W) and flow (Hs,Ws,2)")
        dev = image.device
        img = image.detach().to(torch.float32).contiguous()
        flw = flow.detach().to(torch.float32).contiguous()
        C, H, W = img.shape
        Hs, Ws = flw.shape[0], flw.shape[1]
        out = torch.empty((C, Hc, Wc), dtype=torch.float32, device=dev)
        mask = torch.empty((1, Hc, Wc), dtype=torch.float32, device=dev)
        flow_full = torch.empty((Hf, Wf, 2) if want_flow else (0,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = capi.lib().bags_distort_forward(_ptr(img), C, H, W, _ptr(flw), Hs, Ws, Hf, Wf, Hc, Wc, int(crop),
                                                 int(mask_mode), _ptr(out), _ptr(mask),
                                                 _ptr(flow_full) if want_flow else ctypes.c_void_p(0), stream)
        capi.check(rc, "bags_distort_forward")
        ctx.save_for_backward(img, flw)
        ctx.dims = (Hf, Wf, Hc, Wc, int(crop))
        ctx.mark_non_differentiable(mask, flow_full)
        return out, mask, flow_full

    @staticmethod
    def backward(ctx, grad_out, _gm, _gf):
        img, flw = ctx.saved_tensors
        Hf, Wf, Hc, Wc, crop = ctx.dims
        dev = img.device
        C, H, W = img.shape
        need_img, need_flow = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_img or need_flow):
            return (None,) * 9
        g = grad_out.to(torch.float32).contiguous()
        d_img = torch.empty_like(img) if need_img else None
        d_flow = torch.empty_like(flw) if need_flow else None
        with torch.cuda.device(dev):
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            rc = capi.lib().bags_distort_backward(_ptr(img), C, H, W, _ptr(flw), flw.shape[0], flw.shape[1], Hf, Wf, Hc,
                                                  Wc, crop, _ptr(g), _ptr(d_img), _ptr(d_flow), stream)
        capi.check(rc, "bags_distort_backward")
        return d_img, d_flow, None, None, None, None, None, None, None


def distort_image(image, flow, canvas_hw=None, crop_hw=None, mask_mode=0, return_flow=False):
    """Fused ``interpolate(flow) -> grid_sample(image, flow) -> center_crop -> mask``.

    image (C,H,W); flow (Hs,Ws,2) in normalised [-1,1] coordinates (x, y); canvas_hw = (Hf, Wf) size the flow is
    upsampled to (default: the flow's own size, i.e. no upsampling); crop_hw = (Hc, Wc) centre crop (default: none).
    Returns (out (C,Hc,Wc), mask (1,Hc,Wc)[, flow_full (Hf,Wf,2)])."""
    Hf, Wf = canvas_hw if canvas_hw is not None else (flow.shape[0], flow.shape[1])
    crop = crop_hw is not None
    Hc, Wc = crop_hw if crop else (Hf, Wf)
    out, mask, flow_full = _DistortFunction.apply(image, flow, int(Hf), int(Wf), int(Hc), int(Wc), crop, mask_mode,
                                                  return_flow)
    return (out, mask, flow_full) if return_flow else (out, mask)


def homogenize(X):
    """utils/util_distortion.py:16-21."""
    assert X.ndim == 2
    assert X.shape[1] in (2, 3)
    return torch.cat((X, torch.ones((X.shape[0], 1), dtype=X.dtype, device=X.device)), dim=1)


class _CropFunction(torch.autograd.Function):
    """center_crop of one (C,H,W) image: the crop grid's bilinear grid_sample alone."""

    @staticmethod
    def forward(ctx, image, th, tw):
        _check_cuda(image)
        img = image.detach().to(torch.float32).contiguous()
        C, H, W = img.shape
        out = torch.empty((C, th, tw), dtype=torch.float32, device=img.device)
        with torch.cuda.device(img.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)
            rc = capi.lib().bags_center_crop_forward(_ptr(img), C, H, W, th, tw, _ptr(out), stream)
        capi.check(rc, "bags_center_crop_forward")
        ctx.dims = (C, H, W, th, tw)
        return out

    @staticmethod
    def backward(ctx, g):
        C, H, W, th, tw = ctx.dims
        g = g.to(torch.float32).contiguous()
        gi = torch.empty((C, H, W), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
            rc = capi.lib().bags_center_crop_backward(_ptr(g), C, H, W, th, tw, _ptr(gi), stream)
        capi.check(rc, "bags_center_crop_backward")
        return gi, None, None


def center_crop(tensor, target_height, target_width):
    """utils/util_distortion.py:58-77: (N,C,H,W) -> (N,C,th,tw) through the crop grid's bilinear grid_sample."""
    if tensor.dim() != 4:
        raise RuntimeError("center_crop expects a (N,C,H,W) tensor")
    return torch.stack([_CropFunction.apply(tensor[i], int(target_height), int(target_width))
                        for i in range(tensor.size(0))], 0)


def apply_distortion(flow_apply2_gt_or_img, lens_net, P_view_insidelens_direction, P_sensor, viewpoint_cam, image,
                     apply2gt=False, flow_scale=None):
    """Drop-in for utils/util_distortion.py:271-311.

    Returns ``(image_or_gt, mask, flow)`` exactly like the reference: ``flow`` is the canvas-resolution flow that the
    training loop hands back as ``flow_apply2_gt_or_img`` on later iterations.  The lens network itself (an invertible
    ResNet, out of the hot path) is called as in the reference; everything after it is the fused kernel."""
    ctrl = None
    if flow_apply2_gt_or_img is None:
        P_out = lens_net.forward(P_view_insidelens_direction, sensor_to_frustum=apply2gt)
        camera_directions_w_lens = homogenize(P_out)
        control_points = camera_directions_w_lens.reshape((P_sensor.shape[0], P_sensor.shape[1], 3))[:, :, :2]
        projection_matrix = viewpoint_cam.flow4gt if apply2gt else viewpoint_cam.projection_matrix
        ctrl = control_points @ projection_matrix[:2, :2]

    if apply2gt:
        src = viewpoint_cam.fish_gt_image.cuda()
        if ctrl is not None:
            hw = (int(viewpoint_cam.image_height), int(viewpoint_cam.image_width))
            gt_image, mask, flow = distort_image(src, ctrl, canvas_hw=hw, mask_mode=1, return_flow=True)
        else:
            flow = flow_apply2_gt_or_img
            gt_image, mask = distort_image(src, flow, mask_mode=1)
        return gt_image, mask, flow

    res = viewpoint_cam.fish_gt_image_resolution
    crop_hw = (int(res[1]), int(res[2]))
    if ctrl is not None:
        hw = (int(res[1] * flow_scale[0]), int(res[2] * flow_scale[1]))
        out, mask, flow = distort_image(image, ctrl, canvas_hw=hw, crop_hw=crop_hw, mask_mode=0, return_flow=True)
    else:
        flow = flow_apply2_gt_or_img
        out, mask = distort_image(image, flow, crop_hw=crop_hw, mask_mode=0)
    return out, mask, flow
