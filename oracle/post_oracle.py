"""TEST INFRASTRUCTURE ONLY -- host restatement of the reference's post-rasterizer steps.

Nothing in the product (`bundle-adjusting-gaussian-splatting_b200/`) imports this file; the product path is the CUDA
kernels of csrc/distort.cu and csrc/loss.cu and has no CPU fallback.

The reference implements these steps as sequences of PyTorch operators, so the restatement IS that sequence, run on
host tensors (any float dtype; the tests use float64 for "ground truth" and float32 for the golden vectors):

  center_crop()          utils/util_distortion.py:58-77
  distort_render()       utils/util_distortion.py:271-283,296-311   (apply_distortion, apply2gt=False branch)
  distort_gt()           utils/util_distortion.py:271-295           (apply_distortion, apply2gt=True branch)
  l1_loss(), ssim()      utils/loss_utils.py:18-19, 36-76
  training_loss()        train.py:325  ((1 - lambda) * L1 + lambda * (1 - SSIM))

Parity pinning: the reference has no golden vectors for these functions.  oracle/make_golden_post.py imports the
UNMODIFIED reference modules (utils/util_distortion.py, utils/loss_utils.py) in the build container, runs them on
seeded inputs and stores inputs + outputs + gradients in tests/golden/post_*.npz;
tests/test_post_oracle_golden.py checks this file against them.  Status: pinned.
"""
import math

import torch
import torch.nn.functional as F


def center_crop(tensor, target_height, target_width):
    """utils/util_distortion.py:58-77 ((N,C,H,W) -> (N,C,th,tw)); the grid arithmetic is done in float32 like the
    reference (linspace, 2*g/(size-1)-1), then cast to the tensor's dtype."""
    _, _, height, width = tensor.size()
    start_y = (height - target_height) // 2
    start_x = (width - target_width) // 2
    grid_y, grid_x = torch.meshgrid(torch.linspace(start_y, start_y + target_height - 1, target_height),
                                    torch.linspace(start_x, start_x + target_width - 1, target_width), indexing="ij")
    grid = torch.stack((grid_x, grid_y), 2).unsqueeze(0)
    grid = 2.0 * grid / torch.tensor([width - 1, height - 1]) - 1.0
    grid = grid.expand(tensor.size(0), target_height, target_width, 2).to(device=tensor.device, dtype=tensor.dtype)
    return F.grid_sample(tensor, grid, align_corners=True)


def upsample_flow(flow_ctrl, size_hw):
    """utils/util_distortion.py:285,298: (Hs,Ws,2) -> (Hf,Wf,2), bilinear, align_corners=False."""
    return F.interpolate(flow_ctrl.permute(2, 0, 1).unsqueeze(0), size=(int(size_hw[0]), int(size_hw[1])),
                         mode='bilinear', align_corners=False).permute(0, 2, 3, 1).squeeze(0)


def distort_render(image, flow, canvas_hw, crop_hw):
    """apply2gt=False branch: returns (image (C,Hc,Wc), mask (1,Hc,Wc), flow (Hf,Wf,2)).  `flow` is the control-point
    flow (Hs,Ws,2) when canvas_hw is given, else the canvas-resolution flow itself."""
    if canvas_hw is not None:
        flow = upsample_flow(flow, canvas_hw)
    img = F.grid_sample(image.unsqueeze(0), flow.unsqueeze(0), mode="bilinear", padding_mode="zeros",
                        align_corners=True)
    img = center_crop(img, crop_hw[0], crop_hw[1]).squeeze(0)
    mask = (~((img[0] == 0.0000) & (img[1] == 0.0000)).unsqueeze(0)).to(image.dtype)
    return img, mask, flow


def distort_gt(gt, flow, canvas_hw):
    """apply2gt=True branch: returns (gt_image (C,Hf,Wf), mask, flow)."""
    if canvas_hw is not None:
        flow = upsample_flow(flow, canvas_hw)
    out = F.grid_sample(gt.unsqueeze(0), flow.unsqueeze(0), mode="bilinear", padding_mode="zeros",
                        align_corners=True).squeeze(0)
    mask = (~((out[0] < 0.00001) & (out[1] < 0.00001)).unsqueeze(0)).to(gt.dtype)
    return out, mask, flow


def l1_loss(network_output, gt):
    """utils/loss_utils.py:18-19."""
    return torch.abs((network_output - gt)).mean()


def _window(window_size, channel, dtype):
    """utils/loss_utils.py:36-46: 1-D Gaussian (sigma 1.5) built from python floats in float32, outer product."""
    gauss = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g1 = (gauss / gauss.sum()).unsqueeze(1)
    g2 = g1.mm(g1.t()).float().unsqueeze(0).unsqueeze(0)
    return g2.expand(channel, 1, window_size, window_size).contiguous().to(dtype)


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:49-76 for (C,H,W) or (N,C,H,W) inputs."""
    channel = img1.size(-3)
    window = _window(window_size, channel, img1.dtype).to(img1.device)
    squeeze = img1.dim() == 3
    if squeeze:
        img1, img2 = img1.unsqueeze(0), img2.unsqueeze(0)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    if size_average:
        return ssim_map.mean()
    return ssim_map.mean(1).mean(1).mean(1)


def training_loss(image, gt_image, lambda_dssim=0.2):
    """train.py:325: loss = (1 - lambda) * Ll1 + lambda * (1 - ssim)."""
    Ll1 = l1_loss(image, gt_image)
    s = ssim(image, gt_image)
    return (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - s), Ll1, s
