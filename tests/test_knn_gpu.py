"""distCUDA2 drop-in against an exact brute-force 3-NN in torch and -- when oracle/_ref/ref_knn_C.so
travelled to the box -- the compiled reference submodules/simple-knn itself (same floats expected: both sides use the
reference's distance expression and (b0 + b1 + b2) / 3.0f)."""
import numpy as np
import pytest
import torch

import _util as U

pytestmark = pytest.mark.gpu


def _brute(pts):
    P = pts.shape[0]
    out = torch.empty(P, device=pts.device)
    for s in range(0, P, 4096):
        q = pts[s:s + 4096]
        d = q[:, None, :] - pts[None, :, :]
        d2 = (d * d).sum(-1)
        d2[torch.arange(q.shape[0], device=pts.device), torch.arange(s, s + q.shape[0], device=pts.device)] = float("inf")
        out[s:s + 4096] = d2.topk(3, dim=1, largest=False).values.sum(1) / 3.0
    return out


def _clouds():
    g = torch.Generator().manual_seed(0)
    yield "uniform", torch.rand(20000, 3, generator=g) * 4 - 2
    # a depth frame un-projected: a wavy surface seen through a pinhole camera
    v, u = torch.meshgrid(torch.arange(180.0), torch.arange(240.0), indexing="ij")
    z = 2.0 + 0.3 * torch.sin(u / 17) + 0.2 * torch.cos(v / 11) + 0.01 * torch.rand(180, 240, generator=g)
    yield "depth_frame", torch.stack([(u - 120) / 200 * z, (v - 90) / 200 * z, z], -1).reshape(-1, 3)
    c = torch.randn(6000, 3, generator=g) * 0.01
    c[:5] += torch.tensor([50.0, -30.0, 20.0])                 # a tight cluster and a few far outliers
    yield "cluster_and_outliers", c
    d = torch.rand(3000, 3, generator=g)
    d[1000:2000] = d[:1000]                                    # exact duplicates: zero distances count
    yield "duplicates", d
    yield "tiny", torch.rand(7, 3, generator=g)


@pytest.mark.parametrize("name,pts", list(_clouds()), ids=lambda x: x if isinstance(x, str) else "")
def test_distcuda2_matches_exact_knn(name, pts):
    from online_lang_splatting_b200.simple_knn._C import distCUDA2
    dev = torch.device("cuda:0")
    p = pts.to(dev).contiguous()
    ours = distCUDA2(p)
    ref = _brute(p)
    assert torch.isfinite(ours).all()
    assert (ours - ref).abs().max().item() <= 2e-6 * ref.abs().max().item() + 1e-12
    mod = U.ref_module("ref_knn_C")
    if mod is not None:
        theirs = mod.distCUDA2(p)
        assert torch.equal(ours, theirs), (ours - theirs).abs().max().item()


def test_distcuda2_edge_sizes():
    from online_lang_splatting_b200.simple_knn._C import distCUDA2
    dev = torch.device("cuda:0")
    assert distCUDA2(torch.zeros(0, 3, device=dev)).shape == (0,)
    mod = U.ref_module("ref_knn_C")
    for P in (1, 2, 3, 4):
        p = torch.rand(P, 3, generator=torch.Generator().manual_seed(P)).to(dev)
        ours = distCUDA2(p)
        if P >= 4:
            assert torch.isfinite(ours).all()
        else:  # fewer than 3 neighbours: FLT_MAX placeholders stay in the sum, as in the reference
            assert (ours > 1e37).all()
        if mod is not None:
            assert torch.equal(ours, mod.distCUDA2(p))
    with pytest.raises(RuntimeError):
        distCUDA2(torch.zeros(5, 3))
