"""Densification bookkeeping against the plain-torch restatement of the reference's lines
(gaussian_model.py:948-969, slam_backend.py:417-428).  max_radii2D / denom bit-exact; the accumulated gradient norm
within 1e-6 relative (sqrt(x*x + y*y) vs torch.norm); flags equal except where a value sits within 1e-6 relative of
its threshold (none in these seeded cases)."""
import pytest
import torch
from oracle import torch_oracle as TO  # noqa: E402

pytestmark = pytest.mark.gpu


def _state(P, cols, seed, dev):
    g = torch.Generator().manual_seed(seed)
    radii = torch.randint(-1, 40, (P,), generator=g, dtype=torch.int32)
    radii[radii < 8] = 0
    vgrad = torch.randn(P, 3, generator=g) * 1e-3
    max_r = torch.rand(P, generator=g) * 30
    accum = torch.rand(P, 1, generator=g) * 1e-2
    denom = torch.randint(0, 6, (P, 1), generator=g).float()
    accum[denom == 0] = 0.0  # never-seen Gaussians: 0 / 0 = nan -> 0
    scaling = torch.log(torch.rand(P, cols, generator=g) * 0.2 + 1e-3)
    opacity = torch.randn(P, 1, generator=g) * 3
    return [t.to(dev) for t in (radii, vgrad, max_r, accum, denom, scaling, opacity)]


@pytest.mark.parametrize("P,cols", [(100000, 3), (1237, 1), (1, 3), (0, 3)])
def test_update_stats(P, cols):
    from online_lang_splatting_b200 import densification as D
    dev = torch.device("cuda:0")
    radii, vgrad, max_r, accum, denom, _, _ = _state(P, cols, 3, dev)
    r_max, r_acc, r_den = max_r.clone(), accum.clone(), denom.clone()
    TO.reference_update_stats(radii, vgrad, r_max, r_acc, r_den)
    D.update_stats(radii, vgrad, max_r, accum, denom)
    assert torch.equal(max_r, r_max) and torch.equal(denom, r_den)
    assert torch.allclose(accum, r_acc, rtol=1e-6, atol=0)
    # colour refinement form: only max_radii2D
    m2, m2r = max_r.clone() * 0.5, max_r.clone() * 0.5
    D.update_stats(radii, None, m2)
    TO.reference_update_stats(radii, None, m2r)
    assert torch.equal(m2, m2r)


@pytest.mark.parametrize("P,cols,screen", [(100000, 3, 20.0), (5000, 1, None), (33, 3, 20.0), (0, 3, None)])
def test_densify_flags(P, cols, screen):
    from online_lang_splatting_b200 import densification as D
    dev = torch.device("cuda:0")
    _, _, max_r, accum, denom, scaling, opacity = _state(P, cols, 9, dev)
    kw = dict(max_grad=2e-3, min_opacity=0.3, extent=4.0, max_screen_size=screen, percent_dense=0.01)
    flags, counts = D.densify_flags(accum, denom, scaling, opacity, max_r, **kw)
    ref = TO.reference_densify_flags(accum.clone(), denom, scaling, opacity, max_r, **kw)
    assert torch.equal(flags, ref)
    c = counts.cpu().tolist()
    assert c == [int((ref & b).ne(0).sum()) for b in (D.CLONE, D.SPLIT, D.PRUNE)]
    if P >= 5000:
        assert c[0] > 0 and c[1] > 0 and c[2] > 0 and not bool(((ref & D.CLONE) != 0).logical_and((ref & D.SPLIT) != 0).any())
