"""Fused mapping loss (SURVEY 8f N2+N3) against the plain-torch restatement of the reference's lines
(utils/slam_utils.py:121-165, utils/slam_backend.py:576-592).  fp32, tolerance 1e-5 relative on the loss and
2e-6 absolute on the per-pixel gradients (they are +-w/numel; a sign can only differ where |diff| ~ 1e-7)."""
import numpy as np
import pytest
import torch
from oracle import torch_oracle as TO  # noqa: E402

pytestmark = pytest.mark.gpu


def _case(H, W, F, lh, lw, seed, exposure):
    from online_lang_splatting_b200 import losses as LS
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    image, depth = r(3, H, W).to(dev).requires_grad_(True), (r(1, H, W) * 5).to(dev).requires_grad_(True)
    gt_image = r(3, H, W)
    gt_image[:, : H // 4] *= 0.001            # pixels below the rgb boundary threshold
    gt_depth = r(1, H, W) * 5
    gt_depth[:, :, : W // 5] = 0.0            # invalid depth
    gt_image, gt_depth = gt_image.to(dev), gt_depth.to(dev)
    lang = (torch.randn(F, H, W, generator=g) * 0.3).to(dev).requires_grad_(True) if F else None
    gt_lang = (torch.randn(F, lh, lw, generator=g) * 0.3).to(dev) if F else None
    ea = torch.tensor(0.07, device=dev, requires_grad=True) if exposure else None
    eb = torch.tensor(-0.02, device=dev, requires_grad=True) if exposure else None
    kw = dict(alpha=0.9, rgb_boundary_threshold=0.01, lambda_lang=0.7)
    ref = TO.reference_mapping_loss(image, depth, gt_image, gt_depth, lang, gt_lang, exposure_a=ea, exposure_b=eb, **kw)
    leaves = [t for t in (image, depth, lang, ea, eb) if t is not None]
    g_ref = torch.autograd.grad(ref * 1.7, leaves)
    ours = LS.mapping_loss(image, depth, gt_image, gt_depth, lang, gt_lang, exposure_a=ea, exposure_b=eb, **kw)
    g_ours = torch.autograd.grad(ours * 1.7, leaves)
    assert abs(ours.item() - ref.item()) <= 1e-5 * abs(ref.item())
    for a, b in zip(g_ours, g_ref):
        if a.dim() == 0:
            assert abs(a.item() - b.item()) <= 1e-4 * max(abs(b.item()), 1e-6)
        else:
            bad = (a - b).abs() > 2e-6 * b.abs().max().clamp_min(1e-12) + 1e-12
            assert bad.float().mean().item() < 1e-4


def test_mapping_loss_headline_shape():
    _case(540, 960, 15, 192, 192, seed=0, exposure=True)


def test_mapping_loss_ragged_and_upsample_edges():
    _case(75, 121, 3, 17, 23, seed=1, exposure=False)     # odd sizes, non-integer scale
    _case(64, 64, 15, 192, 192, seed=2, exposure=True)    # down-sampling direction of the same formula


def test_mapping_loss_without_language_term():
    _case(48, 80, 0, 0, 0, seed=3, exposure=True)


def test_mapping_loss_needs_cuda_tensors():
    from online_lang_splatting_b200 import losses as LS
    with pytest.raises(RuntimeError):
        LS.mapping_loss(torch.zeros(3, 4, 4), torch.zeros(1, 4, 4), torch.zeros(3, 4, 4), torch.zeros(1, 4, 4))


def test_tracking_loss_matches_reference_lines():
    from online_lang_splatting_b200 import losses as LS
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    H, W = 120, 161
    image = torch.rand(3, H, W, generator=g).to(dev).requires_grad_(True)
    depth = (torch.rand(1, H, W, generator=g) * 5).to(dev).requires_grad_(True)
    opacity = torch.rand(1, H, W, generator=g).pow(0.1).to(dev).requires_grad_(True)   # mostly > 0.95, some below
    gt_image, gt_depth = torch.rand(3, H, W, generator=g).to(dev), (torch.rand(1, H, W, generator=g) * 5).to(dev)
    gt_depth[:, :10] = 0
    grad_mask = (torch.rand(1, H, W, generator=g) > 0.4).float().to(dev)
    ea = torch.tensor(-0.05, device=dev, requires_grad=True)
    eb = torch.tensor(0.03, device=dev, requires_grad=True)
    kw = dict(alpha=0.9, rgb_boundary_threshold=0.01, exposure_a=ea, exposure_b=eb)
    ref = TO.reference_tracking_loss(image, depth, opacity, gt_image, gt_depth, grad_mask, **kw)
    g_ref = torch.autograd.grad(ref, [image, depth, opacity, ea, eb])
    ours = LS.tracking_loss(image, depth, opacity, gt_image, gt_depth, grad_mask, **kw)
    g_ours = torch.autograd.grad(ours, [image, depth, opacity, ea, eb])
    assert abs(ours.item() - ref.item()) <= 1e-5 * abs(ref.item())
    for a, b in zip(g_ours, g_ref):
        if a.dim() == 0:
            assert abs(a.item() - b.item()) <= 1e-4 * max(abs(b.item()), 1e-6)
        else:
            assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item() + 1e-12


# ---- pinned to the REAL reference functions (tests/golden/losses_small.npz, utils/slam_utils.py:91-165) -------------
def _golden():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_losses import golden_inputs
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses_small.npz"))
    d = golden_inputs()
    assert abs(float(sum(v.double().sum() for v in d.values())) - float(z["checksum"][0])) < 1e-6
    return d, z


def _run_golden_case(name, dev, fn_map, fn_track):
    d, z = _golden()
    t = {k: v.to(dev) for k, v in d.items()}
    image, depth, opacity = (t[k].clone().requires_grad_(True) for k in ("image", "depth", "opacity"))
    ea, eb = (torch.nn.Parameter(t[k].clone()) for k in ("exposure_a", "exposure_b"))   # shape [1], as in Camera
    gt_depth = t["gt_depth"][None]
    kw = dict(alpha=0.9, rgb_boundary_threshold=0.01)
    if name == "mapping":
        loss = fn_map(image, depth, t["gt_image"], gt_depth, exposure_a=ea, exposure_b=eb, **kw)
    elif name == "mapping_init":
        loss = fn_map(image, depth, t["gt_image"], gt_depth, **kw)
    else:
        loss = fn_track(image, depth, opacity, t["gt_image"], gt_depth, t["grad_mask"], exposure_a=ea, exposure_b=eb, **kw)
    (loss * 1.7).backward()
    assert abs(loss.item() - float(z[name + "_loss"][0])) <= 2e-6 * abs(float(z[name + "_loss"][0]))
    for key, leaf in (("dimage", image), ("ddepth", depth), ("dopacity", opacity), ("dexposure_a", ea), ("dexposure_b", eb)):
        if name + "_" + key not in z.files:
            continue
        ref = torch.from_numpy(z[name + "_" + key])
        got = leaf.grad.detach().cpu()
        assert got.shape == ref.shape, (name, key, got.shape, ref.shape)    # [1] for the exposure parameters
        if ref.numel() == 1:
            assert abs(got.item() - ref.item()) <= 1e-4 * max(abs(ref.item()), 1e-6), (name, key)
        else:
            assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item() + 1e-12, (name, key)


@pytest.mark.parametrize("name", ["mapping", "mapping_init", "tracking"])
def test_fused_losses_match_real_reference_golden(name):
    from online_lang_splatting_b200 import losses as LS
    _run_golden_case(name, torch.device("cuda:0"), LS.mapping_loss, LS.tracking_loss)


def test_exposure_parameters_of_shape_1_get_shape_1_gradients():
    """Camera.exposure_a / _b are nn.Parameter(torch.tensor([0.0])) (utils/camera_utils.py:59-64)."""
    from online_lang_splatting_b200 import losses as LS
    dev = torch.device("cuda:0")
    img = torch.rand(3, 32, 48, device=dev, requires_grad=True)
    dep = torch.rand(1, 32, 48, device=dev, requires_grad=True)
    ea = torch.nn.Parameter(torch.tensor([0.0], device=dev))
    eb = torch.nn.Parameter(torch.tensor([0.0], device=dev))
    loss = LS.mapping_loss(img, dep, torch.rand(3, 32, 48, device=dev), torch.rand(1, 32, 48, device=dev), exposure_a=ea, exposure_b=eb)
    loss.backward()
    assert ea.grad.shape == (1,) and eb.grad.shape == (1,) and torch.isfinite(ea.grad).all()
