"""SSIM + colour-refinement loss (gaussian_splatting/utils/loss_utils.py:41-101, utils/slam_backend.py:797-801).

CPU: the torch restatement (oracle.torch_oracle.reference_ssim) reproduces value and autograd gradient of the REAL reference functions
(tests/golden/ssim_small.npz, written by make_golden_ssim.py).  GPU: the fused kernels against the golden and against
the restatement on larger, ragged images.  fp32; tolerance 2e-6 absolute on the value (the filter taps are summed in a
different order than cuDNN's), 1e-4 of the largest gradient entry per pixel.
"""
import os

import numpy as np
import pytest
import torch

from online_lang_splatting_b200 import losses as LS
from oracle import torch_oracle as TO  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssim_small.npz")


def test_restatement_matches_reference_golden():
    z = np.load(GOLD)
    img = torch.from_numpy(z["image"]).requires_grad_(True)
    gt = torch.from_numpy(z["gt"])
    lam = float(z["lambda_dssim"])
    s = TO.reference_ssim(img, gt)
    loss = TO.reference_color_refinement_loss(img, gt, lam)
    loss.backward()
    assert abs(float(s) - float(z["ssim"])) < 1e-6
    assert abs(float(loss) - float(z["loss"])) < 1e-6
    np.testing.assert_allclose(img.grad.numpy(), z["grad"], atol=1e-8, rtol=1e-4)


def test_ssim_rejects_cpu_and_other_windows():
    a, b = torch.rand(3, 20, 20), torch.rand(3, 20, 20)
    with pytest.raises(Exception):
        LS.ssim(a, b)  # no CPU path
    with pytest.raises(NotImplementedError):
        LS.ssim(a, b, window_size=7)


@pytest.mark.gpu
def test_fused_matches_reference_golden():
    dev = torch.device("cuda:0")
    z = np.load(GOLD)
    img = torch.from_numpy(z["image"]).to(dev).requires_grad_(True)
    gt = torch.from_numpy(z["gt"]).to(dev)
    lam = float(z["lambda_dssim"])
    s = LS.ssim(img.detach(), gt)
    loss = LS.color_refinement_loss(img, gt, lam)
    loss.backward()
    assert abs(s.item() - float(z["ssim"])) < 2e-6
    assert abs(loss.item() - float(z["loss"])) < 2e-6
    g, gr = img.grad.cpu().numpy(), z["grad"]
    assert np.abs(g - gr).max() < 1e-4 * np.abs(gr).max()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 540, 960), (3, 123, 77), (1, 3, 64, 48), (3, 11, 5), (1, 16, 16)])
def test_fused_matches_restatement(shape):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(sum(shape))
    gt = torch.rand(*shape, generator=g).to(dev)
    img = (gt + 0.2 * torch.randn(*shape, generator=g).to(dev)).clamp(0, 1).requires_grad_(True)
    # plain ssim, value and gradient
    ref = TO.reference_ssim(img, gt)
    (g_ref,) = torch.autograd.grad(ref * 1.3, img)
    ours = LS.ssim(img, gt)
    (g_ours,) = torch.autograd.grad(ours * 1.3, img)
    assert abs(ours.item() - ref.item()) < 2e-6
    assert (g_ours - g_ref).abs().max().item() < 1e-4 * g_ref.abs().max().item()
    # the colour-refinement loss
    ref = TO.reference_color_refinement_loss(img, gt, 0.2)
    (g_ref,) = torch.autograd.grad(ref, img)
    ours = LS.color_refinement_loss(img, gt, 0.2)
    (g_ours,) = torch.autograd.grad(ours, img)
    assert abs(ours.item() - ref.item()) < 2e-6
    bad = (g_ours - g_ref).abs() > 1e-4 * g_ref.abs().max()
    assert bad.float().mean().item() < 1e-4  # sign(0) ties of the L1 term only


@pytest.mark.gpu
def test_identical_images_give_ssim_one():
    dev = torch.device("cuda:0")
    a = torch.rand(3, 100, 130, device=dev)
    assert abs(LS.ssim(a, a.clone()).item() - 1.0) < 1e-6
