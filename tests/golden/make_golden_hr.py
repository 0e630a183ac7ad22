#!/usr/bin/env python
"""Generates tests/golden/hr_small.npz by importing the REAL reference HR module (language/supervisedNet.py) in this
container (CPU torch); `pytorch_lightning` is stubbed (only LightningModule is referenced, as a base class).
The network's ~25 M parameters are not stored: they are re-created from a seed by oracle/hr_oracle.seeded_state_dict,
the same function the tests call.  Stored: the inputs' seed, and a strided sample of the reference's output and of
three intermediate activations.  Run here:  python tests/golden/make_golden_hr.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("OLS_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import hr_oracle  # noqa: E402

SEED_W, SEED_X, S_H, S_W = 1234, 99, 16, 24


def import_reference():
    m = types.ModuleType("pytorch_lightning")
    m.LightningModule = torch.nn.Module
    sys.modules["pytorch_lightning"] = m
    sys.path.insert(0, os.path.join(REF, "language"))
    import importlib
    return importlib.import_module("supervisedNet")


def inputs(seed=SEED_X, s_h=S_H, s_w=S_W):
    g = torch.Generator().manual_seed(seed)
    fv = torch.randn(1, 768, s_h, s_w, generator=g)
    f3 = torch.randn(1, 384, 4 * s_h, 4 * s_w, generator=g)       # ConvNeXt res3: stride 8 when fv is stride 32
    f2 = torch.randn(1, 192, 8 * s_h - 3, 8 * s_w + 5, generator=g)  # deliberately not an integer ratio
    return fv, f3, f2


def main():
    ref = import_reference()
    net = ref.HighResLanguageFeatureNet().eval()
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = hr_oracle.seeded_state_dict(shapes, SEED_W)
    net.load_state_dict(sd)
    fv, f3, f2 = inputs()
    with torch.no_grad():
        out = net(fv, f3, f2)
        ours, inter = hr_oracle.hr_forward(sd, fv, f3, f2, return_intermediates=True)
    print("reference out", tuple(out.shape), "rms", float(out.pow(2).mean().sqrt()),
          "oracle max |diff|", float((out - ours).abs().max()))
    np.savez_compressed(
        os.path.join(HERE, "hr_small.npz"),
        seed_w=SEED_W, seed_x=SEED_X, s_h=S_H, s_w=S_W,
        shape_keys=np.array([k for k, _ in shapes]), shape_dims=np.array([list(s) + [0] * (4 - len(s)) for _, s in shapes]),
        shape_rank=np.array([len(s) for _, s in shapes]),
        out_sample=out[0, ::16, ::8, ::8].numpy(),       # [48, 16, 24]
        out_rms=float(out.pow(2).mean().sqrt()), out_mean=float(out.mean()), out_abs_sum=float(out.abs().sum()),
    )


if __name__ == "__main__":
    main()
