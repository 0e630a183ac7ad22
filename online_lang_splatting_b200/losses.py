"""Mapping loss on the outputs of ``render()``, fused into two CUDA kernels (SURVEY 8f rows N2 + N3).

Mirrors what a mapping iteration of the reference computes per view right after ``render()``
(utils/slam_backend.py:576-592 with utils/slam_utils.py:121-165)::

    loss = get_loss_mapping(config, image, depth, viewpoint, opacity)          # alpha * l1_rgb + (1 - alpha) * l1_depth
         + lamda_lang * l1_loss(language, F.interpolate(gt_lang_feat[None], (H, W), mode="bilinear")[0])

but keeps the low-resolution language target on the device (the reference copies the up-sampled
15 x H x W map over PCIe every iteration) and produces the three image-space gradients directly.
There is no torch fallback: without the CUDA library the call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native as N


class _MappingLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, depth, language, gt_image, gt_depth, gt_lang, exposure_a, exposure_b, alpha, threshold,
                lambda_lang, opacity=None, grad_mask=None):
        N.require_cuda()
        if not image.is_cuda:
            raise RuntimeError("mapping_loss needs CUDA tensors: there is no CPU path")
        dev = image.device
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        image_, depth_, gt_image_, gt_depth_ = f32(image), f32(depth), f32(gt_image), f32(gt_depth)
        _, H, W = image_.shape
        has_lang = language is not None and gt_lang is not None
        lang_ = f32(language) if has_lang else None
        gt_lang_ = f32(gt_lang) if has_lang else None
        opacity_ = f32(opacity) if opacity is not None else None
        grad_mask_ = f32(grad_mask) if grad_mask is not None else None
        F = int(lang_.shape[0]) if has_lang else 0
        ea = float(exposure_a) if exposure_a is not None else 0.0
        eb = float(exposure_b) if exposure_b is not None else 0.0
        args = N.LossArgs(W=W, H=H, F=F, lang_w=int(gt_lang_.shape[2]) if has_lang else 0,
                          lang_h=int(gt_lang_.shape[1]) if has_lang else 0, alpha=float(alpha),
                          rgb_boundary_threshold=float(threshold), exposure_a=ea, exposure_b=eb,
                          lambda_lang=float(lambda_lang), d_image=image_.data_ptr(), d_depth=depth_.data_ptr(),
                          d_language=N.ptr(lang_), d_gt_image=gt_image_.data_ptr(), d_gt_depth=gt_depth_.data_ptr(),
                          d_gt_lang=N.ptr(gt_lang_), d_opacity=N.ptr(opacity_), d_grad_mask=N.ptr(grad_mask_))
        out = torch.empty(14, dtype=torch.float32, device=dev)  # [0:6] results, [6:14] reduction scratch
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_mapping_loss_forward(C.byref(args), out.data_ptr(), out[6:].data_ptr(), stream))
        ctx.args = args
        ctx.keep = (image_, depth_, lang_, gt_image_, gt_depth_, gt_lang_, opacity_, grad_mask_)
        ctx.terms = out
        ctx.exposure_tensors = (torch.is_tensor(exposure_a) and exposure_a.requires_grad,
                                torch.is_tensor(exposure_b) and exposure_b.requires_grad)
        return out[5]

    @staticmethod
    def backward(ctx, grad_loss):
        image_, depth_, lang_, _, _, _, opacity_, _ = ctx.keep
        dev = image_.device
        up = grad_loss.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        d_image, d_depth = torch.empty_like(image_), torch.empty_like(depth_)
        d_lang = torch.empty_like(lang_) if lang_ is not None else None
        d_op = torch.empty_like(opacity_) if opacity_ is not None else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_mapping_loss_backward(C.byref(ctx.args), up.data_ptr(), d_image.data_ptr(),
                                                      d_depth.data_ptr(), N.ptr(d_lang), N.ptr(d_op), stream))
        ga = ctx.terms[3] * up[0] if ctx.exposure_tensors[0] else None
        gb = ctx.terms[4] * up[0] if ctx.exposure_tensors[1] else None
        return d_image, d_depth, d_lang, None, None, None, ga, gb, None, None, None, d_op, None


def mapping_loss(image: torch.Tensor, depth: torch.Tensor, gt_image: torch.Tensor, gt_depth: torch.Tensor,
                 language: Optional[torch.Tensor] = None, gt_lang_feat: Optional[torch.Tensor] = None, *,
                 alpha: float = 0.95, rgb_boundary_threshold: float = 0.01, exposure_a=None, exposure_b=None,
                 lambda_lang: float = 1.0) -> torch.Tensor:
    """``alpha * l1_rgb + (1 - alpha) * l1_depth + lambda_lang * l1_lang`` of one rendered view.

    ``gt_lang_feat`` is the low-resolution ``[F, h, w]`` code map (``viewpoint.gt_lang_feat``); it is up-sampled
    bilinearly (``align_corners=False``) inside the kernel.  Pass ``exposure_a`` / ``exposure_b`` (tensors or floats)
    for the exposure-compensated form of ``get_loss_mapping``; leave them ``None`` for ``initialization=True``.
    """
    return _MappingLoss.apply(image, depth, language, gt_image, gt_depth, gt_lang_feat, exposure_a, exposure_b, alpha,
                              rgb_boundary_threshold, lambda_lang)


def tracking_loss(image: torch.Tensor, depth: torch.Tensor, opacity: torch.Tensor, gt_image: torch.Tensor,
                  gt_depth: torch.Tensor, grad_mask: Optional[torch.Tensor] = None, *, alpha: float = 0.95,
                  rgb_boundary_threshold: float = 0.01, exposure_a=None, exposure_b=None) -> torch.Tensor:
    """``get_loss_tracking`` / ``get_loss_tracking_rgbd`` (utils/slam_utils.py:91-118): the colour residual is weighted
    by the rendered opacity and masked by ``viewpoint.grad_mask``; depth counts where ``opacity > 0.95``.  The gradient
    w.r.t. ``opacity`` is produced, but -- as in the reference -- the rasterizer does not propagate it further."""
    return _MappingLoss.apply(image, depth, None, gt_image, gt_depth, None, exposure_a, exposure_b, alpha,
                              rgb_boundary_threshold, 0.0, opacity, grad_mask)


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
