"""Plain-torch restatements of the reference's torch code either side of the rasterizer -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import this module; nothing
under online_lang_splatting_b200/ does.  Each function cites the reference lines it restates and is pinned by a golden
file produced from the REAL reference code (tests/golden/make_golden_*.py):

  reference_mapping_loss / reference_tracking_loss   utils/slam_utils.py:91-165, utils/slam_backend.py:576-592
                                                     pinned by tests/golden/losses_small.npz
  reference_ssim / reference_color_refinement_loss   gaussian_splatting/utils/loss_utils.py:41-101   (ssim_small.npz)
  reference_chain                                    language/autoencoder/model.py:15-62             (ae_*.npz)
  reference_update_stats / reference_densify_flags   gaussian_splatting/scene/gaussian_model.py:948-969
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

CLONE, SPLIT, PRUNE = 1, 2, 4   # bits of the densification flags (densification.py)

def reference_tracking_loss(image, depth, opacity, gt_image, gt_depth, grad_mask=None, *, alpha=0.95,
                            rgb_boundary_threshold=0.01, exposure_a=None, exposure_b=None):
    """Plain-torch restatement of utils/slam_utils.py:91-118 (test reference)."""
    if exposure_a is not None:
        image = torch.exp(torch.as_tensor(exposure_a, device=image.device)) * image + torch.as_tensor(exposure_b, device=image.device)
    rgb_pixel_mask = (gt_image.sum(dim=0) > rgb_boundary_threshold).view(*depth.shape)
    if grad_mask is not None:
        rgb_pixel_mask = rgb_pixel_mask * grad_mask
    l1 = opacity * torch.abs(image * rgb_pixel_mask - gt_image * rgb_pixel_mask)
    depth_mask = (gt_depth > 0.01).view(*depth.shape) * (opacity > 0.95).view(*depth.shape)
    l1_depth = torch.abs(depth * depth_mask - gt_depth * depth_mask)
    return alpha * l1.mean() + (1 - alpha) * l1_depth.mean()


def reference_mapping_loss(image, depth, gt_image, gt_depth, language=None, gt_lang_feat=None, *, alpha=0.95,
                           rgb_boundary_threshold=0.01, exposure_a=None, exposure_b=None, lambda_lang=1.0):
    """Plain-torch restatement of the reference lines cited above (used by the tests as the fp32 reference)."""
    if exposure_a is not None:
        image = torch.exp(torch.as_tensor(exposure_a, device=image.device)) * image + torch.as_tensor(exposure_b, device=image.device)
    rgb_pixel_mask = (gt_image.sum(dim=0) > rgb_boundary_threshold).view(*depth.shape)
    depth_pixel_mask = (gt_depth > 0.01).view(*depth.shape)
    l1_rgb = torch.abs(image * rgb_pixel_mask - gt_image * rgb_pixel_mask)
    l1_depth = torch.abs(depth * depth_pixel_mask - gt_depth * depth_pixel_mask)
    loss = alpha * l1_rgb.mean() + (1 - alpha) * l1_depth.mean()
    if language is not None and gt_lang_feat is not None:
        up = torch.nn.functional.interpolate(gt_lang_feat.unsqueeze(0), size=tuple(image.shape[1:]), mode="bilinear",
                                             align_corners=False).squeeze(0)
        loss = loss + lambda_lang * torch.abs(language - up).mean()
    return loss


def reference_ssim(img1, img2, window_size=11):
    """Plain-torch restatement of loss_utils.py:41-101 (test reference)."""
    import math
    g = torch.tensor([math.exp(-((x - window_size // 2) ** 2) / float(2 * 1.5 ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    channel = img1.size(-3)
    window = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()
    window = window.to(img1.device).type_as(img1)
    conv = lambda t: torch.nn.functional.conv2d(t, window, padding=window_size // 2, groups=channel)
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq, sigma2_sq, sigma12 = conv(img1 * img1) - mu1_sq, conv(img2 * img2) - mu2_sq, conv(img1 * img2) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def reference_color_refinement_loss(image, gt_image, lambda_dssim=0.2):
    return (1.0 - lambda_dssim) * torch.abs(image - gt_image).mean() + lambda_dssim * (1.0 - reference_ssim(image, gt_image))


def reference_chain(modules: Sequence[nn.Module], x: torch.Tensor) -> torch.Tensor:
    """The reference's arithmetic in plain torch (used by tests and the CPU baseline only)."""
    for m in modules:
        x = m(x)
    return x / x.norm(dim=-1, keepdim=True)


def reference_update_stats(radii, viewspace_grad, max_radii2D, xyz_gradient_accum=None, denom=None):
    vis = radii > 0
    max_radii2D[vis] = torch.max(max_radii2D[vis], radii[vis])
    if viewspace_grad is not None:
        xyz_gradient_accum[vis] += torch.norm(viewspace_grad[vis, :2], dim=-1, keepdim=True)
        denom[vis] += 1


def reference_densify_flags(xyz_gradient_accum, denom, scaling_raw, opacity_raw, max_radii2D, *, max_grad, min_opacity,
                            extent, max_screen_size, percent_dense=0.01):
    grads = xyz_gradient_accum / denom
    grads[grads.isnan()] = 0.0
    smax = torch.max(torch.exp(scaling_raw), dim=1).values
    hot = torch.norm(grads, dim=-1) >= max_grad
    clone = hot & (smax <= percent_dense * extent)
    split = (grads.squeeze(-1) >= max_grad) & (smax > percent_dense * extent)
    prune = (torch.sigmoid(opacity_raw) < min_opacity).squeeze(-1)
    if max_screen_size:
        prune = prune | (max_radii2D > max_screen_size) | (smax > 0.1 * extent)
    return clone.to(torch.uint8) * CLONE + split.to(torch.uint8) * SPLIT + prune.to(torch.uint8) * PRUNE
