"""Drop-in for the reference's HR module (language/supervisedNet.py):

* ``AttentionFusion``            (:6-43)
* ``HighResLanguageFeatureNet``  (:45-109)   fv [N,768,S,S] + res3 [N,384,.,.] + res2 [N,192,.,.] -> [N,768,8S,8S]
* ``LangSupervisedNet``          (:111-125)  the wrapper slam_backend.py:149-152 loads from a checkpoint

Same sub-module names and ``nn.Sequential`` indices, so the reference's ``state_dict`` (and the ``state_dict`` entry
of its Lightning checkpoint) loads unchanged.  The reference only ever runs this network in eval mode under
``torch.no_grad()`` (utils/slam_backend.py:370-386,547-552); that is the path implemented here: the 13 convolutions
run as implicit-GEMM tcgen05 kernels through the C ABI (``ols_hr_forward``), eval-mode BatchNorm folded into the
weights on the host side, ReLU / sigmoid gate / concatenation / bilinear resizing fused into the kernels.  Training
mode, autograd and CPU tensors are rejected -- there is no fallback path.

The result has the reference's shape ``[N,768,8S,8S]`` and values, stored channels-last: the caller's
``permute(0,2,3,1).view(-1,768)`` (slam_backend.py:392-394) is then a free view of a contiguous ``[M,768]`` matrix,
which is exactly what the fused autoencoder kernel reads by TMA.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _native as N


def _conv_bn_relu(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


def _up_bn_relu(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(cout),
                         nn.ReLU(inplace=True))


class AttentionFusion(nn.Module):
    """supervisedNet.py:6-43.  ``forward`` here is the torch restatement used for parameter-free shape checks only;
    the network's forward never calls it on the product path."""

    def __init__(self, in_channels_high_res: int, in_channels_low_res: int):
        super().__init__()
        self.low_res_align = (nn.Conv2d(in_channels_low_res, in_channels_high_res, kernel_size=1)
                              if in_channels_high_res != in_channels_low_res else nn.Identity())
        self.fusion = _conv_bn_relu(in_channels_high_res * 2, in_channels_high_res)
        self.attention = nn.Sequential(
            nn.Conv2d(in_channels_high_res, in_channels_high_res, kernel_size=3, padding=1),
            nn.BatchNorm2d(in_channels_high_res),
            nn.ReLU(inplace=True),
            nn.Conv2d(in_channels_high_res, in_channels_high_res, kernel_size=1),
            nn.Sigmoid(),
        )

    def forward(self, high_res_feat, low_res_feat):
        raise RuntimeError("AttentionFusion runs fused inside HighResLanguageFeatureNet.forward (ols_hr_forward)")


def _fold_bn(conv: nn.Module, bn: Optional[nn.BatchNorm2d]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(weight, bias) of conv followed by eval-mode BatchNorm, in the convolution's own torch layout."""
    w = conv.weight.detach().float()
    cout = w.shape[1] if isinstance(conv, nn.ConvTranspose2d) else w.shape[0]
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(cout, device=w.device)
    if bn is not None:
        g = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
        if bn.weight is not None:
            g = g * bn.weight.detach().float()
        beta = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(g)
        b = (b - bn.running_mean.detach().float()) * g + beta
        w = w * (g.view(1, -1, 1, 1) if isinstance(conv, nn.ConvTranspose2d) else g.view(-1, 1, 1, 1))
    return w.contiguous(), b.contiguous()


class HighResLanguageFeatureNet(nn.Module):
    """supervisedNet.py:45-109."""

    def __init__(self, desired_channels: int = 768):
        super().__init__()
        if desired_channels != 768:
            raise ValueError("the sm_100a HR path is built for the reference's 768-channel CLIP map")
        self.initial_conv = _conv_bn_relu(768, 512)
        self.upsample1 = _up_bn_relu(512, 512)
        self.attention_fusion1 = AttentionFusion(in_channels_high_res=512, in_channels_low_res=384)
        self.upsample2 = _up_bn_relu(512, 256)
        self.attention_fusion2 = AttentionFusion(in_channels_high_res=256, in_channels_low_res=192)
        self.upsample3 = _up_bn_relu(256, 128)
        self.final_conv = nn.Conv2d(128, desired_channels, kernel_size=1)
        self._plan = None
        self._key = None
        self._keep = None

    # ---- plan management --------------------------------------------------------------------------------------
    def _convs(self) -> List[Tuple[nn.Module, Optional[nn.BatchNorm2d]]]:
        """The 13 convolutions in the order include/ols_b200.h documents, each with the BatchNorm that follows it."""
        a1, a2 = self.attention_fusion1, self.attention_fusion2
        return [
            (self.initial_conv[0], self.initial_conv[1]),
            (self.upsample1[0], self.upsample1[1]),
            (a1.low_res_align, None),
            (a1.fusion[0], a1.fusion[1]),
            (a1.attention[0], a1.attention[1]),
            (a1.attention[3], None),
            (self.upsample2[0], self.upsample2[1]),
            (a2.low_res_align, None),
            (a2.fusion[0], a2.fusion[1]),
            (a2.attention[0], a2.attention[1]),
            (a2.attention[3], None),
            (self.upsample3[0], self.upsample3[1]),
            (self.final_conv, None),
        ]

    def _version(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _destroy(self):
        if self._plan is not None:
            N.lib().ols_hr_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def _ensure_plan(self, dev: torch.device, S_h: int, S_w: int, stream: int):
        key = (self._version(), dev.index, S_h, S_w)
        if self._plan is not None and self._key == key:
            return
        self._destroy()
        folded = [_fold_bn(c, bn) for c, bn in self._convs()]
        hw = N.HRWeights()
        for i, (w, b) in enumerate(folded):
            if w.device != dev:
                raise RuntimeError("HR module parameters and inputs must live on the same CUDA device")
            hw.d_weight[i] = w.data_ptr()
            hw.d_bias[i] = b.data_ptr()
        plan = C.c_void_p()
        N.check(N.lib().ols_hr_plan_create(C.byref(hw), S_h, S_w, C.byref(plan), stream))
        self._plan, self._key, self._keep = plan, key, folded

    # ---- forward ----------------------------------------------------------------------------------------------
    def forward(self, fv: torch.Tensor, f3: torch.Tensor, f2: torch.Tensor) -> torch.Tensor:
        N.require_cuda()
        if self.training:
            raise RuntimeError("HighResLanguageFeatureNet: only the eval-mode inference path exists (call .eval())")
        if torch.is_grad_enabled() and any(t.requires_grad for t in (fv, f3, f2)):
            raise RuntimeError("HighResLanguageFeatureNet: no autograd path (the reference runs it under no_grad)")
        if not (fv.is_cuda and f3.is_cuda and f2.is_cuda):
            raise RuntimeError("HighResLanguageFeatureNet inputs must be CUDA tensors: there is no CPU path")
        if fv.dim() != 4 or fv.shape[1] != 768 or f3.shape[1] != 384 or f2.shape[1] != 192:
            raise RuntimeError("expected fv [N,768,S,S], f3 [N,384,h,w], f2 [N,192,h,w]")
        if not (fv.shape[0] == f3.shape[0] == f2.shape[0]):
            raise RuntimeError("batch sizes differ")
        dev = fv.device
        n, _, S_h, S_w = fv.shape
        fv, f3, f2 = (t.detach().float().contiguous() for t in (fv, f3, f2))
        out = torch.empty((n, 8 * S_h, 8 * S_w, 768), dtype=torch.float32, device=dev)
        lib = N.lib()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self._ensure_plan(dev, S_h, S_w, stream)
            for i in range(n):
                N.check(lib.ols_hr_forward(self._plan, fv[i].data_ptr(), f3[i].data_ptr(), f3.shape[2], f3.shape[3],
                                           f2[i].data_ptr(), f2.shape[2], f2.shape[3], out[i].data_ptr(), stream))
        return out.permute(0, 3, 1, 2)  # [N,768,8S,8S], channels-last storage

    def features(self, fv: torch.Tensor, f3: torch.Tensor, f2: torch.Tensor) -> torch.Tensor:
        """Everything up to (not including) ``final_conv``: ``upsample3``'s output as bfloat16 ``[N, 8S*8S, 128]``
        (pixel-major).  Used by ``AutoencoderMLP.encode_hr``, which folds ``final_conv`` into the encoder."""
        N.require_cuda()
        if self.training:
            raise RuntimeError("HighResLanguageFeatureNet: only the eval-mode inference path exists (call .eval())")
        if not (fv.is_cuda and f3.is_cuda and f2.is_cuda):
            raise RuntimeError("HighResLanguageFeatureNet inputs must be CUDA tensors: there is no CPU path")
        dev = fv.device
        n, _, S_h, S_w = fv.shape
        fv, f3, f2 = (t.detach().float().contiguous() for t in (fv, f3, f2))
        feat = torch.empty((n, 64 * S_h * S_w, 128), dtype=torch.bfloat16, device=dev)
        lib = N.lib()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self._ensure_plan(dev, S_h, S_w, stream)
            for i in range(n):
                N.check(lib.ols_hr_forward_features(self._plan, fv[i].data_ptr(), f3[i].data_ptr(), f3.shape[2], f3.shape[3],
                                                    f2[i].data_ptr(), f2.shape[2], f2.shape[3], feat[i].data_ptr(), stream))
        return feat

    def read_activation(self, which: int) -> torch.Tensor:
        """Debug aid: output of convolution ``which`` (0..11) of the last forward as float32 [H,W,C]."""
        if self._plan is None:
            raise RuntimeError("no forward has run yet")
        S_h, S_w = self._key[2], self._key[3]
        level_out = [0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3][which]
        cout = [512, 512, 512, 512, 512, 512, 256, 256, 256, 256, 256, 128][which]
        dev = self._keep[0][0].device
        t = torch.empty((S_h << level_out, S_w << level_out, cout), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            N.check(N.lib().ols_hr_read_activation(self._plan, which, t.data_ptr(), t.numel(),
                                                   torch.cuda.current_stream(dev).cuda_stream))
        return t


class LangSupervisedNet(nn.Module):
    """supervisedNet.py:111-125 without the Lightning training loop (training the HR module is offline work in the
    reference and outside the hot path).  ``load_from_checkpoint`` reads the ``state_dict`` of a Lightning checkpoint."""

    def __init__(self, lambda_recon=1.0, lambda_edge=0.5, lambda_cosine=0.0, lambda_perceptual=0.0, lambda_tv=0.0):
        super().__init__()
        self.model = HighResLanguageFeatureNet()
        self.lambda_recon, self.lambda_edge, self.lambda_cosine = lambda_recon, lambda_edge, lambda_cosine
        self.lambda_tv, self.lambda_perpectual = lambda_tv, lambda_perceptual

    def forward(self, fv, f3, f2):
        return self.model(fv, f3, f2)

    @classmethod
    def load_from_checkpoint(cls, path: str, map_location="cpu", allow_pickle: bool = False, **kwargs) -> "LangSupervisedNet":
        """Reads the ``state_dict`` of a reference (Lightning) checkpoint.  Tensors-only loading is tried first; a
        checkpoint that needs arbitrary unpickling (it can execute code) is only read with ``allow_pickle=True``."""
        try:
            ckpt = torch.load(path, map_location=map_location, weights_only=True)
        except Exception:
            if not allow_pickle:
                raise
            ckpt = torch.load(path, map_location=map_location, weights_only=False)
        net = cls(**kwargs)
        net.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)
        return net
