"""Autoencoder tests.  CPU part: the module restatement reproduces the REAL reference classes
(golden vectors made by tests/golden/make_golden_ae.py from language/autoencoder/model.py).
GPU part (-m gpu): the fused tcgen05 kernel against the torch fp32 graph of the same module.

Tolerance (stated, not bit-exact -- SURVEY 8c: cuBLAS/torch accumulation order is unspecified and the
kernel computes layer 0 in TF32 and the inner layers in BF16 with fp32 accumulation): outputs are
unit vectors, so the metric is the cosine to the reference row (>= 0.9995) and the relative L2 error
of the whole tensor (<= 3e-2 through 6-7 layers, <= 3e-3 for a single TF32 layer)."""
import glob
import os

import numpy as np
import pytest
import torch

import _util as U
from online_lang_splatting_b200 import autoencoder as AE
from oracle import torch_oracle as TO  # noqa: E402

CASES = {
    "ae_1stage": (lambda: AE.AutoencoderMLP([384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]), 768),
    "ae_2stage_general": (lambda: AE.AutoencoderMLP([512, 256, 128, 64, 32], [192, 256, 384, 512, 768]), 768),
    "ae_online": (lambda: AE.EncoderDecoderOnline(), 32),
}


def randomize_bn(model, g):
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(torch.empty(m.num_features).uniform_(0.5, 1.5, generator=g))
            m.weight.data.copy_(torch.empty(m.num_features).uniform_(0.8, 1.2, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))


def build(name):
    make, din = CASES[name]
    torch.manual_seed(0)
    model = make()
    g = torch.Generator().manual_seed(1)
    model.eval()
    randomize_bn(model, g)
    return model, din


@pytest.mark.parametrize("name", sorted(CASES))
def test_module_restatement_matches_reference_golden(name):
    z = np.load(os.path.join(U.GOLDEN_DIR, name + ".npz"))
    model, din = build(name)
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point:
            cs = z["cs." + k]
            assert abs(v.double().sum().item() - cs[0]) < 1e-9 and abs(v.double().abs().sum().item() - cs[1]) < 1e-9, k
    x = torch.from_numpy(z["x"])
    with torch.no_grad():
        code = TO.reference_chain(list(model.encoder), x)
        rec = TO.reference_chain(list(model.decoder), code)
    assert np.abs(code.numpy() - z["code"]).max() < 1e-6
    assert np.abs(rec.numpy() - z["rec"]).max() < 1e-6


def test_bn_folding_is_exact():
    model, din = build("ae_1stage")
    x = torch.randn(50, din)
    folded = AE._fold(list(model.encoder))
    h = x
    for i, (W, b) in enumerate(folded):
        h = h @ W.t() + b
        if i < len(folded) - 1:
            h = torch.relu(h)
    h = h / h.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        ref = TO.reference_chain(list(model.encoder), x)
    assert (h - ref).abs().max() < 2e-5


def test_cpu_tensors_are_rejected():
    model, din = build("ae_online")
    with torch.no_grad(), pytest.raises(RuntimeError):
        model.encode(torch.randn(4, din))


def _metrics(y, ref):
    y, ref = y.double(), ref.double()
    cos = (y * ref).sum(-1) / (y.norm(dim=-1) * ref.norm(dim=-1))
    rel = float((y - ref).norm() / ref.norm())
    return float(cos.min()), rel


@pytest.mark.gpu
@pytest.mark.parametrize("dims,M", [([768, 384], 128), ([768, 384], 1000), ([768, 16], 300), ([768, 384, 192], 517),
                                    ([768, 512], 256), ([32, 24, 15], 700), ([15, 24, 32], 700), ([64, 256, 48], 129)])
def test_fused_chain_ladder(cuda, dims, M):
    """Single layers first (descriptor / swizzle / TMEM plumbing), then short chains; no normalisation."""
    g = torch.Generator().manual_seed(5)
    layers = [(torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5, 0.1 * torch.randn(dims[i + 1], generator=g))
              for i in range(len(dims) - 1)]
    x = torch.randn(M, dims[0], generator=g)
    ref = x.double()
    for i, (W, b) in enumerate(layers):
        ref = ref @ W.double().t() + b.double()
        if i < len(layers) - 1:
            ref = torch.relu(ref)
    chain = AE._FusedChain()
    y = chain.run([(W.to(cuda), b.to(cuda)) for W, b in layers], ("t", tuple(dims)), False, x.to(cuda)).cpu()
    assert torch.isfinite(y).all()
    cos, rel = _metrics(y, ref)
    assert rel < (3e-3 if len(dims) == 2 else 2e-2), (rel, cos)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_fused_autoencoder_matches_torch(cuda, name):
    model, din = build(name)
    model = model.to(cuda)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(36864 + 77, din, generator=g)
    x = (x / x.norm(dim=-1, keepdim=True)).to(cuda)
    with torch.no_grad():
        code = model.encode(x)
        rec = model.decode(code)
        ref_code = TO.reference_chain(list(model.encoder), x)
        ref_rec_same_code = TO.reference_chain(list(model.decoder), code)
    assert code.shape == ref_code.shape and rec.shape == (x.shape[0], din)
    assert torch.allclose(code.norm(dim=-1), torch.ones_like(code[:, 0]), atol=1e-4)
    # tolerance = 3x the measured error of the tensor-core mode (tools/ae_errors.py on B200: rel-L2 0.8e-3 .. 1.6e-3,
    # cos_min >= 0.999996 over the three reference architectures); the fp32 parity mode is tested below at 1e-5
    cos, rel = _metrics(code.cpu(), ref_code.cpu())
    assert cos > 0.99998 and rel < 5e-3, (cos, rel)
    cos, rel = _metrics(rec.cpu(), ref_rec_same_code.cpu())
    assert cos > 0.99998 and rel < 5e-3, (cos, rel)


@pytest.mark.gpu
def test_online_autoencoder_train_step_uses_autograd(cuda):
    """slam_backend.py:266-323: one Adam step on L1 + 0.6 (1 - cos); afterwards the fused path sees the new weights."""
    model, din = build("ae_online")
    model = model.to(cuda).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    x = torch.randn(4096, din, device=cuda)
    x = x / x.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        before = model.encode(x).clone()
    rec = model.decode(model.encode(x))
    loss = torch.nn.functional.l1_loss(rec, x) + 0.6 * (1 - torch.nn.functional.cosine_similarity(rec, x, dim=-1).mean())
    opt.zero_grad(); loss.backward(); opt.step()
    with torch.no_grad():
        after = model.encode(x)
        ref_after = TO.reference_chain(list(model.encoder), x)
    assert (after - before).abs().max() > 0
    cos, rel = _metrics(after.cpu(), ref_after.cpu())
    assert cos > 0.9995 and rel < 3e-2


@pytest.mark.gpu
def test_config4_batch32_full_size(cuda):
    """BASELINE config 4: batch 32 of 192x192x768 maps (1,179,648 rows) through encode -> decode.  The kernel is
    row-independent, so parity at full size is a random subset of rows against torch fp32 plus size-independent
    properties: unit norm of every code / reconstruction row, finite values, and equality with a separate run on
    the subset alone (the result of a row does not depend on which tile it lands in)."""
    model, din = build("ae_1stage")
    model = model.to(cuda)
    M = 32 * 192 * 192
    g = torch.Generator(device=cuda).manual_seed(0)
    x = torch.randn(M, din, device=cuda, generator=g)
    x /= x.norm(dim=-1, keepdim=True)
    with torch.no_grad():
        code = model.encode(x)
        rec = model.decode(code)
        idx = torch.randint(0, M, (4096,), device=cuda, generator=g)
        ref_code = TO.reference_chain(list(model.encoder), x[idx])
        sub_code = model.encode(x[idx].contiguous())
    assert code.shape == (M, 15) and rec.shape == (M, din)
    assert torch.isfinite(code).all() and torch.isfinite(rec).all()
    assert torch.allclose(code.norm(dim=-1), torch.ones(M, device=cuda), atol=1e-4)
    assert torch.allclose(rec.norm(dim=-1), torch.ones(M, device=cuda), atol=1e-4)
    cos = torch.nn.functional.cosine_similarity(code[idx], ref_code, dim=-1)
    assert cos.min().item() > 0.9995
    assert torch.equal(code[idx], sub_code)


@pytest.mark.gpu
def test_fused_online_train_step_matches_torch_adam():
    """train_online_autoencoder (utils/slam_backend.py:266-323) as one kernel.  Five steps against the
    reference's own sequence -- the module's torch graph, l1_loss + 0.6 (1 - cosine_similarity.mean()), backward,
    torch.optim.Adam(lr=1e-3) -- on [36864, 32] unit-norm codes: parameters within 1e-5, loss and returned codes too."""
    import copy
    from online_lang_splatting_b200 import autoencoder as AE
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    ref = AE.EncoderDecoderOnline().to(dev)
    ours = copy.deepcopy(ref)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(4)
    for it in range(5):
        x = torch.randn(36864 if it < 4 else 1000, 32, generator=g)      # last step: ragged row count (not a multiple of 128)
        x = (x / x.norm(dim=-1, keepdim=True)).to(dev)
        ref.train()
        opt.zero_grad()
        comp = ref.encode(x)
        recon = ref.decode(comp)
        loss = torch.nn.functional.l1_loss(recon, x) + 0.6 * (1 - torch.nn.functional.cosine_similarity(recon, x, dim=1).mean())
        loss.backward()
        opt.step()
        l2, c2 = ours.fused_train_step(x, lr=1e-3)
        assert abs(l2.item() - loss.item()) <= 2e-6 * abs(loss.item()) + 1e-7, (it, l2.item(), loss.item())
        assert (c2 - comp.detach()).abs().max().item() < 2e-6
    for p, q in zip(ours.parameters(), ref.parameters()):
        assert (p.detach() - q.detach()).abs().max().item() < 1e-5
    # the fused inference path sees the updated weights (its cache is keyed on the in-place update counter)
    ours.eval()
    with torch.no_grad():
        x = torch.randn(4096, 32, generator=g).to(dev)
        a, b = ours.encode(x), ref.eval().encode(x) if False else None
        want = torch.nn.functional.normalize(ref.encoder(x), dim=-1)
        assert torch.nn.functional.cosine_similarity(a, want, dim=-1).min().item() > 0.9995


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_fp32_parity_mode_matches_torch_fp32(cuda, name):
    """VERDICT r1 missing #4: an fp32-grade parity mode beside the tensor-core speed mode.  autoencoder.PRECISION = "fp32"
    runs every layer in fp32 FMAs (the arithmetic of the reference's nn.Linear, model.py:52-62); against torch fp32 on
    the same GPU the relative L2 error is bounded by summation order: <= 1e-5."""
    model, din = build(name)
    model = model.to(cuda)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4096 + 13, din, generator=g)
    x = (x / x.norm(dim=-1, keepdim=True)).to(cuda)
    saved = AE.PRECISION
    AE.PRECISION = "fp32"
    try:
        with torch.no_grad():
            code = model.encode(x)
            rec = model.decode(code)
            ref_code = TO.reference_chain(list(model.encoder), x)
            ref_rec = TO.reference_chain(list(model.decoder), code)
    finally:
        AE.PRECISION = saved
    cos, rel = _metrics(code.cpu(), ref_code.cpu())
    assert rel < 1e-5, (cos, rel)
    cos, rel = _metrics(rec.cpu(), ref_rec.cpu())
    assert rel < 1e-5, (cos, rel)
