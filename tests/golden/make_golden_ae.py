#!/usr/bin/env python
"""Generates tests/golden/ae_*.npz by importing the REAL reference autoencoder classes
(language/autoencoder/model.py) in this container (CPU torch) with import stubs for the packages the
file pulls in at module top but the two classes never use (lightning, matplotlib, open_clip,
torchvision, sklearn, eval.colormaps).  Run here:  python tests/golden/make_golden_ae.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OLS_REFERENCE_ROOT", "/root/reference")


def import_reference_model():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    pl = stub("lightning.pytorch", LightningModule=torch.nn.Module)
    stub("lightning", pytorch=pl)
    stub("matplotlib.pyplot")
    stub("matplotlib", pyplot=sys.modules["matplotlib.pyplot"])
    stub("open_clip")
    stub("torchvision.models")
    stub("torchvision", models=sys.modules["torchvision.models"])
    stub("sklearn.decomposition", IncrementalPCA=object)
    stub("sklearn", decomposition=sys.modules["sklearn.decomposition"])
    stub("eval.colormaps", apply_pca_colormap=None)
    stub("eval", colormaps=sys.modules["eval.colormaps"])
    sys.path.insert(0, os.path.join(REF, "language", "autoencoder"))
    import importlib
    return importlib.import_module("model")


def randomize_bn(model, g):
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(torch.empty(m.num_features).uniform_(0.5, 1.5, generator=g))
            m.weight.data.copy_(torch.empty(m.num_features).uniform_(0.8, 1.2, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))


def main():
    ref = import_reference_model()
    # weights are NOT stored (megabytes of noise): every case re-seeds torch, so the test rebuilds the
    # identical module with torch.manual_seed(SEED) and checks the per-tensor checksums stored here
    cases = {
        "ae_1stage": (lambda: ref.AutoencoderMLP([384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]), 768),
        "ae_2stage_general": (lambda: ref.AutoencoderMLP([512, 256, 128, 64, 32], [192, 256, 384, 512, 768]), 768),
        "ae_online": (lambda: ref.EncoderDecoderOnline(), 32),
    }
    for name, (make, din) in cases.items():
        torch.manual_seed(0)
        model = make()
        g = torch.Generator().manual_seed(1)
        model.eval()
        randomize_bn(model, g)
        x = torch.randn(64, din, generator=g)
        x = x / x.norm(dim=-1, keepdim=True)
        with torch.no_grad():
            code = model.encode(x)
            rec = model.decode(code)
        save = {"x": x.numpy(), "code": code.numpy(), "rec": rec.numpy()}
        for k, v in model.state_dict().items():
            if v.dtype.is_floating_point:
                save["cs." + k] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **save)
        print(name, "code", tuple(code.shape), "rec", tuple(rec.shape), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
