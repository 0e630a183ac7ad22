#!/usr/bin/env python
"""Generates tests/golden/ssim_small.npz with the REAL reference functions (gaussian_splatting/utils/loss_utils.py:
ssim, l1_loss) imported in this container on CPU torch: value and autograd gradient of the colour-refinement loss
(utils/slam_backend.py:797-801) on a seeded 3 x 37 x 45 image pair.  Run here: python tests/golden/make_golden_ssim.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OLS_REFERENCE_ROOT", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_loss_utils", os.path.join(REF, "gaussian_splatting", "utils", "loss_utils.py"))
    lu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lu)
    g = torch.Generator().manual_seed(21)
    gt = torch.rand(3, 37, 45, generator=g)
    img = (gt + 0.15 * torch.randn(3, 37, 45, generator=g)).clamp(0, 1).requires_grad_(True)
    lam = 0.2
    s = lu.ssim(img, gt)
    loss = (1.0 - lam) * lu.l1_loss(img, gt) + lam * (1.0 - s)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, "ssim_small.npz"), image=img.detach().numpy(), gt=gt.numpy(), lambda_dssim=lam,
                        ssim=float(s), loss=float(loss), grad=img.grad.numpy())
    print("ssim", float(s), "loss", float(loss), "|grad|", float(img.grad.abs().sum()))


if __name__ == "__main__":
    main()
