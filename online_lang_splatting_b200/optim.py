"""Fused Adam over the flat Gaussian parameter buffer.

The reference optimises the Gaussian parameters with ``torch.optim.Adam(param_groups, lr=0.0, eps=1e-15)``, one
group per tensor (gaussian_splatting/scene/gaussian_model.py:393-437).  With the structure-of-arrays layout of
``sharding.FlatGradBuffer`` -- the buffer the rasterizer backward writes and NCCL all-reduces -- the whole step is
one streaming kernel over (param, grad, exp_avg, exp_avg_sq).  ``FlatAdam`` mirrors torch's interface where the
reference touches it: ``param_groups[k]["lr"]`` / ``["name"]`` (the schedulers write the xyz learning rate there),
``step()`` and ``zero_grad()``.  There is no torch fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Sequence, Tuple

import torch

from . import _native as N


class FlatAdam:
    def __init__(self, flat_param: torch.Tensor, flat_grad: torch.Tensor, groups: Sequence[Tuple[str, int, float]],
                 betas=(0.9, 0.999), eps: float = 1e-15, capturable: bool = False):
        """``groups``: (name, number of elements, lr[, activation[, period, head, lr_tail]]) in buffer order; they must
        tile the flat buffers.  ``activation`` (``_native.ACT_*``) says that ``flat_grad`` holds the gradient w.r.t. the
        ACTIVATED parameter -- what the rasterizer's backward writes -- while ``flat_param`` holds the raw one: the
        kernel applies the activation's Jacobian (exp / sigmoid / normalize) before the Adam update, which is what
        autograd does between the rasterizer and ``optimizer.step()`` in the reference.  ``period / head / lr_tail``
        give the tail of every ``period``-element row its own learning rate (f_rest inside the [P,M,3] SH block)."""
        N.require_cuda()
        if not (flat_param.is_cuda and flat_grad.is_cuda and flat_param.dtype == flat_grad.dtype == torch.float32):
            raise RuntimeError("FlatAdam needs fp32 CUDA buffers: there is no CPU path")
        if flat_param.numel() != flat_grad.numel() or sum(int(g[1]) for g in groups) != flat_param.numel():
            raise ValueError("groups must cover the flat buffers exactly")
        if len(groups) > 8:
            raise ValueError("at most 8 parameter groups")
        self.param, self.grad = flat_param, flat_grad
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self.betas, self.eps, self.steps = betas, eps, 0
        # capturable: the step count lives on the device (like torch.optim.Adam(capturable=True)), so a CUDA graph that
        # contains step() advances the bias correction on every replay instead of freezing the captured value
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=flat_param.device) if capturable else None
        self.param_groups: List[Dict] = []
        o = 0
        for grp in groups:
            name, count, lr = grp[:3]
            act = int(grp[3]) if len(grp) > 3 else N.ACT_NONE
            period, head, lr_tail = (int(grp[4]), int(grp[5]), float(grp[6])) if len(grp) > 6 else (0, 0, 0.0)
            if act == N.ACT_NORMALIZE4 and (o % 4 or int(count) % 4):
                raise ValueError("a normalize4 group must start at a multiple of 4 elements and hold whole rows")
            self.param_groups.append({"name": name, "lr": float(lr), "offset": o, "count": int(count), "activation": act,
                                      "period": period, "head": head, "lr_tail": lr_tail})
            o += int(count)

    def zero_grad(self):
        self.grad.zero_()

    def step(self):
        self.steps += 1
        arr = (N.AdamGroup * len(self.param_groups))(*[
            N.AdamGroup(offset=g["offset"], count=g["count"], lr=g["lr"], activation=g["activation"], period=g["period"],
                        head=g["head"], lr_tail=g["lr_tail"]) for g in self.param_groups])
        dev = self.param.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if self.step_dev is not None:
                N.check(N.lib().ols_adam_step_dev(self.param.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                                  self.exp_avg_sq.data_ptr(), self.param.numel(), arr, len(self.param_groups),
                                                  float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                                  self.step_dev.data_ptr(), stream))
            else:
                N.check(N.lib().ols_adam_step(self.param.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                              self.exp_avg_sq.data_ptr(), self.param.numel(), arr, len(self.param_groups),
                                              float(self.betas[0]), float(self.betas[1]), float(self.eps), self.steps, stream))


def activate_params(opacity_raw: torch.Tensor, scaling_raw: torch.Tensor, rotation_raw: torch.Tensor,
                    opacity: torch.Tensor, scaling: torch.Tensor, rotation: torch.Tensor) -> None:
    """``get_opacity`` / ``get_scaling`` / ``get_rotation`` (gaussian_model.py:93-130) into preallocated buffers, one kernel."""
    N.require_cuda()
    P = opacity_raw.numel()
    cols = scaling_raw.numel() // max(P, 1)
    dev = opacity_raw.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().ols_activate_params(P, cols, opacity_raw.data_ptr(), scaling_raw.data_ptr(), rotation_raw.data_ptr(),
                                            opacity.data_ptr(), scaling.data_ptr(), rotation.data_ptr(), stream))
