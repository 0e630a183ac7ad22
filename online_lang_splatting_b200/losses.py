"""Mapping loss on the outputs of ``render()``, fused into two CUDA kernels (SURVEY 8f rows N2 + N3).

Mirrors what a mapping iteration of the reference computes per view right after ``render()``
(utils/slam_backend.py:576-592 with utils/slam_utils.py:121-165)::

    loss = get_loss_mapping(config, image, depth, viewpoint, opacity)          # alpha * l1_rgb + (1 - alpha) * l1_depth
         + lamda_lang * l1_loss(language, F.interpolate(gt_lang_feat[None], (H, W), mode="bilinear")[0])

(the colour-only sub-forms ``get_loss_tracking_rgb`` / ``get_loss_mapping_rgb`` of utils/slam_utils.py:96-105,128-137 are the
same kernels with ``alpha = 1.0``)
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
        # exposure parameters stay on the device (viewpoint.exposure_a / _b are shape-[1] nn.Parameters,
        # utils/camera_utils.py:59-64): no host read-back, so the call neither synchronises nor breaks graph capture
        ea_t = exposure_a.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:1].contiguous() if torch.is_tensor(exposure_a) else None
        eb_t = exposure_b.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:1].contiguous() if torch.is_tensor(exposure_b) else None
        ea = float(exposure_a) if (exposure_a is not None and ea_t is None) else 0.0
        eb = float(exposure_b) if (exposure_b is not None and eb_t is None) else 0.0
        args = N.LossArgs(W=W, H=H, F=F, lang_w=int(gt_lang_.shape[2]) if has_lang else 0,
                          lang_h=int(gt_lang_.shape[1]) if has_lang else 0, alpha=float(alpha),
                          rgb_boundary_threshold=float(threshold), exposure_a=ea, exposure_b=eb,
                          lambda_lang=float(lambda_lang), d_image=image_.data_ptr(), d_depth=depth_.data_ptr(),
                          d_language=N.ptr(lang_), d_gt_image=gt_image_.data_ptr(), d_gt_depth=gt_depth_.data_ptr(),
                          d_gt_lang=N.ptr(gt_lang_), d_opacity=N.ptr(opacity_), d_grad_mask=N.ptr(grad_mask_),
                          d_exposure_a=N.ptr(ea_t), d_exposure_b=N.ptr(eb_t))
        out = torch.empty(14, dtype=torch.float32, device=dev)  # [0:6] results, [6:14] reduction scratch
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_mapping_loss_forward(C.byref(args), out.data_ptr(), out[6:].data_ptr(), stream))
        ctx.args = args
        ctx.keep = (image_, depth_, lang_, gt_image_, gt_depth_, gt_lang_, opacity_, grad_mask_, ea_t, eb_t)
        ctx.exposure_shapes = (tuple(exposure_a.shape) if torch.is_tensor(exposure_a) else None,
                               tuple(exposure_b.shape) if torch.is_tensor(exposure_b) else None)
        ctx.terms = out
        ctx.exposure_tensors = (torch.is_tensor(exposure_a) and exposure_a.requires_grad,
                                torch.is_tensor(exposure_b) and exposure_b.requires_grad)
        return out[5]

    @staticmethod
    def backward(ctx, grad_loss):
        image_, depth_, lang_, _, _, _, opacity_, _, _, _ = ctx.keep
        dev = image_.device
        up = grad_loss.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        d_image, d_depth = torch.empty_like(image_), torch.empty_like(depth_)
        d_lang = torch.empty_like(lang_) if lang_ is not None else None
        d_op = torch.empty_like(opacity_) if opacity_ is not None else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_mapping_loss_backward(C.byref(ctx.args), up.data_ptr(), d_image.data_ptr(),
                                                      d_depth.data_ptr(), N.ptr(d_lang), N.ptr(d_op), stream))
        # gradients take the shape of the parameters they belong to ([1] for the reference's Camera.exposure_a / _b)
        ga = (ctx.terms[3] * up[0]).reshape(ctx.exposure_shapes[0]) if ctx.exposure_tensors[0] else None
        gb = (ctx.terms[4] * up[0]).reshape(ctx.exposure_shapes[1]) if ctx.exposure_tensors[1] else None
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


# ---- SSIM and the colour-refinement loss -----------------------------------------------------------
class _SsimLoss(torch.autograd.Function):
    """value = w_l1 * mean|image - gt| + w_ssim * ssim(image, gt); gradient w.r.t. image only (gt is data)."""

    @staticmethod
    def forward(ctx, image, gt, w_l1, w_ssim):
        N.require_cuda()
        if not image.is_cuda:
            raise RuntimeError("ssim needs CUDA tensors: there is no CPU path")
        dev = image.device
        x = image.detach().to(device=dev, dtype=torch.float32).contiguous()
        y = gt.detach().to(device=dev, dtype=torch.float32).contiguous()
        if x.shape != y.shape or x.dim() < 3:
            raise RuntimeError("ssim expects two [..., C, H, W] tensors of the same shape")
        Cc, H, W = int(x.numel() // (x.shape[-1] * x.shape[-2])), int(x.shape[-2]), int(x.shape[-1])
        args = N.SsimArgs(C=Cc, H=H, W=W, w_l1=float(w_l1), w_ssim=float(w_ssim), d_image=x.data_ptr(), d_gt=y.data_ptr())
        need_grad = image.requires_grad
        partial = torch.empty((3,) + tuple(x.shape), dtype=torch.float32, device=dev) if need_grad else None
        out = torch.empty(4, dtype=torch.float32, device=dev)
        scratch = torch.empty(2, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_ssim_loss_forward(C.byref(args), out.data_ptr(), N.ptr(partial), scratch.data_ptr(), stream))
        ctx.args, ctx.keep = args, (x, y, partial)
        ctx.terms = out
        return out[2].clone()

    @staticmethod
    def backward(ctx, grad_value):
        x, _, partial = ctx.keep
        if partial is None:
            return None, None, None, None
        dev = x.device
        up = grad_value.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        dx = torch.empty_like(x)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_ssim_loss_backward(C.byref(ctx.args), partial.data_ptr(), up.data_ptr(), dx.data_ptr(), stream))
        return dx, None, None, None


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, size_average: bool = True) -> torch.Tensor:
    """Drop-in for ``gaussian_splatting.utils.loss_utils.ssim`` (:61-101) as its two callers use it (window 11,
    ``size_average=True``; utils/slam_backend.py:801, utils/eval_utils.py:174).  Differentiable w.r.t. ``img1``."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("the fused kernel implements the reference's call: window_size=11, size_average=True")
    return _SsimLoss.apply(img1, img2, 0.0, 1.0)


def color_refinement_loss(image: torch.Tensor, gt_image: torch.Tensor, lambda_dssim: float = 0.2) -> torch.Tensor:
    """``(1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))`` (utils/slam_backend.py:797-801),
    one forward and one backward kernel."""
    return _SsimLoss.apply(image, gt_image, 1.0 - lambda_dssim, -lambda_dssim) + lambda_dssim
