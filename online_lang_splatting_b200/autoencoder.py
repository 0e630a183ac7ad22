"""Drop-in for the reference's language autoencoders (language/autoencoder/model.py):

* ``AutoencoderMLP``       (:15-62)   768 -> code -> 768 per-pixel MLP, BatchNorm in the encoder
* ``EncoderDecoderOnline`` (:314-354) the online 32 -> 24 -> 15 -> 24 -> 32 code compressor

Same constructor arguments, same sub-module names and indices (``encoder.0`` ... so reference
``state_dict``s load unchanged), same ``encode`` / ``decode`` / ``forward`` semantics.

Inference (``torch.no_grad()`` / eval mode, CUDA tensors) runs the whole layer chain as ONE fused
tcgen05 kernel through the C ABI (``ols_ae_forward``): eval-mode BatchNorm is folded into the
preceding Linear here, on the host side.  Calls that need autograd (the online autoencoder's Adam
step, utils/slam_backend.py:266-323) run the same module graph through torch so gradients exist;
CPU tensors are rejected -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _native as N


# "fast": tensor cores (tf32 first layer, bf16 inner layers, fp32 accumulate) -- cos >= 0.99999 against torch fp32;
# "fp32": parity mode, every layer in fp32 FMAs like the reference's nn.Linear (model.py:52-62), ~10x slower.
PRECISION = "fast"


class _FusedChain:
    """Owns an ``ols_ae_plan`` for a list of (weight, bias) pairs and rebuilds it when they change."""

    def __init__(self):
        self._plan: Optional[int] = None
        self._key = None
        self._keep = None

    def _destroy(self):
        if self._plan is not None:
            N.lib().ols_ae_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def run(self, layers: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]], version_key, normalize: bool,
            x: torch.Tensor, x_bf16: bool = False) -> torch.Tensor:
        N.require_cuda()
        if not x.is_cuda:
            raise RuntimeError("autoencoder input must be a CUDA tensor: the fused kernel has no CPU path")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        if x_bf16:
            if x2.dtype != torch.bfloat16:
                raise RuntimeError("expected a bfloat16 input matrix")
        elif x2.dtype != torch.float32:
            x2 = x2.float()
        x2 = x2.contiguous()
        dev = x2.device
        fp32 = PRECISION == "fp32" and not x_bf16
        key = (version_key, dev.index, normalize, x_bf16, fp32)
        lib = N.lib()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if self._plan is None or self._key != key:
                self._destroy()
                ws = [w.detach().to(dev, torch.float32).contiguous() for w, _ in layers]
                bs = [None if b is None else b.detach().to(dev, torch.float32).contiguous() for _, b in layers]
                chain = N.AEChain(n_layers=len(ws), normalize=int(normalize), input_bf16=int(x_bf16), precision=int(fp32))
                chain.dims[0] = ws[0].shape[1]
                for i, w in enumerate(ws):
                    chain.dims[i + 1] = w.shape[0]
                    chain.d_weight[i] = w.data_ptr()
                    chain.d_bias[i] = 0 if bs[i] is None else bs[i].data_ptr()
                plan = C.c_void_p()
                N.check(lib.ols_ae_plan_create(C.byref(chain), C.byref(plan), stream))
                self._plan, self._key, self._keep = plan, key, (ws, bs)
            if x2.shape[1] != self._keep[0][0].shape[1]:
                raise RuntimeError(f"expected input width {self._keep[0][0].shape[1]}, got {x2.shape[1]}")
            y = torch.empty((x2.shape[0], self._keep[0][-1].shape[0]), dtype=torch.float32, device=dev)
            fwd = lib.ols_ae_forward_bf16 if x_bf16 else lib.ols_ae_forward
            N.check(fwd(self._plan, x2.data_ptr(), y.data_ptr(), x2.shape[0], stream))
        return y.reshape(*lead, y.shape[-1])


def _fold(modules: Sequence[nn.Module]) -> List[Tuple[torch.Tensor, Optional[torch.Tensor]]]:
    """Linear [-> BatchNorm1d(eval)] [-> ReLU] ... -> list of (W, b) with the BatchNorm folded in."""
    out: List[Tuple[torch.Tensor, Optional[torch.Tensor]]] = []
    for m in modules:
        if isinstance(m, nn.Linear):
            out.append((m.weight.detach(), None if m.bias is None else m.bias.detach()))
        elif isinstance(m, nn.BatchNorm1d):
            W, b = out[-1]
            inv = torch.rsqrt(m.running_var.detach() + m.eps)
            g = inv if m.weight is None else m.weight.detach() * inv
            b0 = torch.zeros_like(m.running_mean) if b is None else b
            beta = torch.zeros_like(m.running_mean) if m.bias is None else m.bias.detach()
            out[-1] = (W * g[:, None], (b0 - m.running_mean.detach()) * g + beta)
        elif isinstance(m, nn.ReLU):
            continue
        else:
            raise RuntimeError(f"unsupported module in autoencoder chain: {type(m).__name__}")
    return out


def _version(modules: Sequence[nn.Module]):
    v = []
    for m in modules:
        for t in list(m.parameters(recurse=False)) + list(m.buffers(recurse=False)):
            v.append((t.data_ptr(), t._version))
        v.append(getattr(m, "_ols_updates", 0))   # in-place updates by fused kernels (invisible to the version counters)
    return tuple(v)


class _ChainMixin:
    def _run(self, which: str, modules, x: torch.Tensor) -> torch.Tensor:
        needs_graph = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for m in modules for p in m.parameters()))
        bn_training = any(isinstance(m, nn.BatchNorm1d) and m.training for m in modules)
        if needs_graph or bn_training:
            if not x.is_cuda:
                raise RuntimeError("autoencoder input must be a CUDA tensor")
            for m in modules:  # the reference's own graph (model.py:52-62), for autograd / batch statistics
                x = m(x)
            return x / x.norm(dim=-1, keepdim=True)
        fused = self.__dict__.setdefault("_fused_" + which, _FusedChain())
        # fold BatchNorm only when a parameter or buffer changed (folding launches ~7 small kernels per BatchNorm)
        key = _version(modules)
        cache = self.__dict__.get("_folded_" + which)
        if cache is None or cache[0] != key:
            cache = self.__dict__["_folded_" + which] = (key, _fold(modules))
        return fused.run(cache[1], key, True, x)


class AutoencoderMLP(nn.Module, _ChainMixin):
    def __init__(self, encoder_hidden_dims, decoder_hidden_dims, clip_dim=768):
        super().__init__()
        encoder_layers = []
        for i in range(len(encoder_hidden_dims)):
            if i == 0:
                encoder_layers.append(nn.Linear(clip_dim, encoder_hidden_dims[i]))
            else:
                encoder_layers.append(nn.BatchNorm1d(encoder_hidden_dims[i - 1]))
                encoder_layers.append(nn.ReLU())
                encoder_layers.append(nn.Linear(encoder_hidden_dims[i - 1], encoder_hidden_dims[i]))
        self.encoder = nn.ModuleList(encoder_layers)
        decoder_layers = []
        for i in range(len(decoder_hidden_dims)):
            if i == 0:
                decoder_layers.append(nn.Linear(encoder_hidden_dims[-1], decoder_hidden_dims[i]))
            else:
                decoder_layers.append(nn.ReLU())
                decoder_layers.append(nn.Linear(decoder_hidden_dims[i - 1], decoder_hidden_dims[i]))
        self.decoder = nn.ModuleList(decoder_layers)

    def forward(self, x):
        return self.decode(self.encode(x))

    def encode(self, x):
        return self._run("enc", list(self.encoder), x)

    def decode(self, x):
        return self._run("dec", list(self.decoder), x)

    def encode_hr(self, hr_model, fv: torch.Tensor, f3: torch.Tensor, f2: torch.Tensor) -> torch.Tensor:
        """``self.encode(hr_model(fv, f3, f2).permute(0, 2, 3, 1).view(-1, 768))`` (utils/slam_backend.py:381-395) without
        ever materialising the 768-channel map: the HR module's ``final_conv`` is a per-pixel Linear 128 -> 768 and the
        encoder starts with a Linear 768 -> h with nothing in between, so the two are folded into one Linear 128 -> h
        (``W' = W_enc0 W_final``, ``b' = W_enc0 b_final + b_enc0``, done here in float64) and the fused kernel reads the
        HR module's last bf16 activation directly.  Saves writing and re-reading 113 MB per frame plus 5/6 of the first
        layer's flops.  Same result up to rounding (``tests/test_hr.py::test_hr_fused_encode``).  Inference only."""
        net = getattr(hr_model, "model", hr_model)  # LangSupervisedNet or HighResLanguageFeatureNet
        mods = list(self.encoder)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and any(
                t.requires_grad for t in (fv, f3, f2)):
            raise RuntimeError("encode_hr is an inference path (the reference runs it under torch.no_grad())")
        if any(isinstance(m, nn.BatchNorm1d) and m.training for m in mods):
            raise RuntimeError("encode_hr needs eval mode (BatchNorm is folded)")
        feat = net.features(fv, f3, f2)  # [N, M, 128] bf16
        key = (_version(mods), _version([net.final_conv]))
        cache = self.__dict__.get("_hr_fold")
        if cache is None or cache[0] != key:
            layers = _fold(mods)
            W0, b0 = layers[0]
            Wf = net.final_conv.weight.detach().reshape(net.final_conv.out_channels, -1).double()
            bf = net.final_conv.bias.detach().double()
            b0d = torch.zeros(W0.shape[0], dtype=torch.float64, device=W0.device) if b0 is None else b0.double()
            layers[0] = ((W0.double() @ Wf).float(), (W0.double() @ bf + b0d).float())
            cache = self.__dict__["_hr_fold"] = (key, layers)
        fused = self.__dict__.setdefault("_fused_enc_hr", _FusedChain())
        return fused.run(cache[1], key, True, feat.reshape(-1, feat.shape[-1]), x_bf16=True)


class EncoderDecoderOnline(nn.Module, _ChainMixin):
    def __init__(self, method="mlp", input_dim=32, compressed_dim=15):
        super().__init__()
        if method != "mlp":
            # the reference's 'pca' branch (sklearn IncrementalPCA on the host) is documented as worse and unused
            raise NotImplementedError("only method='mlp' is provided (model.py:317)")
        self.method = method
        self.encoder = nn.Sequential(nn.Linear(input_dim, 24), nn.ReLU(), nn.Linear(24, 15))
        self.decoder = nn.Sequential(nn.Linear(15, 24), nn.ReLU(), nn.Linear(24, input_dim))

    def encode(self, x):
        return self._run("enc", list(self.encoder), x)

    def decode(self, x):
        return self._run("dec", list(self.decoder), x)

    def forward(self, x):
        return self.decode(self.encode(x))

    # ---- fused training step (utils/slam_backend.py:266-323) -----------------------------------------------------
    def _fused_state(self, dev):
        st = self.__dict__.get("_train_state")
        params = list(self.parameters())
        if st is not None and st["flat"].device == dev and all(
                p.data_ptr() == st["flat"].data_ptr() + 4 * o for p, o in zip(params, st["offsets"])):
            return st
        lib = N.lib()
        n = int(lib.ols_online_ae_param_count())
        if sum(p.numel() for p in params) != n or tuple(params[0].shape) != (24, 32) or tuple(params[2].shape) != (15, 24):
            raise RuntimeError("the fused step is built for the reference's EncoderDecoderOnline (32 -> 24 -> 15 -> 24 -> 32)")
        flat = torch.cat([p.detach().reshape(-1).float() for p in params]).to(dev).contiguous()
        offsets, o = [], 0
        for p in params:                      # the module's parameters become views of the flat vector the kernel updates
            p.data = flat[o:o + p.numel()].view(p.shape)
            offsets.append(o)
            o += p.numel()
        st = {"flat": flat, "offsets": offsets, "m": torch.zeros_like(flat), "v": torch.zeros_like(flat),
              "step": torch.zeros(1, dtype=torch.int64, device=dev),
              "scratch": torch.zeros(int(lib.ols_online_ae_scratch_bytes()), dtype=torch.uint8, device=dev),
              "loss": torch.zeros(1, dtype=torch.float32, device=dev)}
        self.__dict__["_train_state"] = st
        return st

    def fused_train_step(self, features: torch.Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        """``train_online_autoencoder`` (utils/slam_backend.py:266-323) -- encode, decode, ``l1_loss + 0.6 (1 - cos)``,
        backward and ``torch.optim.Adam(lr).step()`` -- as ONE kernel.  Returns ``(loss, comp_15)``: the loss as a
        0-dim device tensor (no host synchronisation; the reference calls ``.item()``) and the 15-dim codes computed
        with the parameters before the update, as the reference returns them.  The Adam moments and the step count live
        in this module; the parameters are updated in place (they are views of one flat vector from the first call on)."""
        N.require_cuda()
        if not features.is_cuda:
            raise RuntimeError("fused_train_step needs a CUDA tensor: there is no CPU path")
        x = features.detach().to(torch.float32).contiguous().view(-1, 32)
        dev = x.device
        st = self._fused_state(dev)
        code = torch.empty((x.shape[0], 15), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_online_ae_train_step(st["flat"].data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(),
                                                     st["step"].data_ptr(), x.data_ptr(), x.shape[0], float(lr), float(betas[0]),
                                                     float(betas[1]), float(eps), code.data_ptr(), st["loss"].data_ptr(),
                                                     st["scratch"].data_ptr(), st["scratch"].numel(), stream))
        for m in list(self.encoder) + list(self.decoder):
            m._ols_updates = getattr(m, "_ols_updates", 0) + 1
        return st["loss"][0], code
