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
                 betas=(0.9, 0.999), eps: float = 1e-15):
        """``groups``: (name, number of elements, lr) in buffer order; they must tile the flat buffers."""
        N.require_cuda()
        if not (flat_param.is_cuda and flat_grad.is_cuda and flat_param.dtype == flat_grad.dtype == torch.float32):
            raise RuntimeError("FlatAdam needs fp32 CUDA buffers: there is no CPU path")
        if flat_param.numel() != flat_grad.numel() or sum(g[1] for g in groups) != flat_param.numel():
            raise ValueError("groups must cover the flat buffers exactly")
        if len(groups) > 8:
            raise ValueError("at most 8 parameter groups")
        self.param, self.grad = flat_param, flat_grad
        self.exp_avg = torch.zeros_like(flat_param)
        self.exp_avg_sq = torch.zeros_like(flat_param)
        self.betas, self.eps, self.steps = betas, eps, 0
        self.param_groups: List[Dict] = []
        o = 0
        for name, count, lr in groups:
            self.param_groups.append({"name": name, "lr": float(lr), "offset": o, "count": int(count)})
            o += int(count)

    def zero_grad(self):
        self.grad.zero_()

    def step(self):
        self.steps += 1
        arr = (N.AdamGroup * len(self.param_groups))(*[N.AdamGroup(offset=g["offset"], count=g["count"], lr=g["lr"])
                                                       for g in self.param_groups])
        dev = self.param.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_adam_step(self.param.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                          self.exp_avg_sq.data_ptr(), self.param.numel(), arr, len(self.param_groups),
                                          float(self.betas[0]), float(self.betas[1]), float(self.eps), self.steps, stream))
