"""Fused flat Adam against torch.optim.Adam with the reference's configuration
(one group per tensor, eps = 1e-15; gaussian_model.py:393-437).  fp32, tolerance 2e-6 relative after 5 steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_flat_adam_matches_torch_adam_groups():
    from online_lang_splatting_b200.optim import FlatAdam
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    P = 10007
    shapes = [("xyz", (P, 3), 1.6e-4), ("f_dc", (P, 1, 3), 2.5e-3), ("opacity", (P, 1), 5e-2), ("scaling", (P, 3), 1e-3),
              ("rotation", (P, 4), 1e-3), ("f_language", (P, 15), 2.5e-3)]
    tensors = [torch.randn(*s, generator=g).to(dev) for _, s, _ in shapes]
    n = sum(t.numel() for t in tensors)
    flat_p = torch.cat([t.reshape(-1) for t in tensors]).clone()
    flat_g = torch.zeros(n, device=dev)
    opt = FlatAdam(flat_p, flat_g, [(nm, t.numel(), lr) for (nm, _, lr), t in zip(shapes, tensors)])
    ref_params = [torch.nn.Parameter(t.clone()) for t in tensors]
    ref = torch.optim.Adam([{"params": [p], "lr": lr, "name": nm} for p, (nm, _, lr) in zip(ref_params, shapes)], lr=0.0, eps=1e-15)
    for it in range(5):
        grads = [torch.randn(*s, generator=g).to(dev) * (10.0 ** (it - 2)) for _, s, _ in shapes]
        grads[2][::7] = 0.0                                   # Gaussians without gradient (eps = 1e-15 matters there)
        if it == 3:                                           # the xyz learning-rate schedule writes into param_groups
            opt.param_groups[0]["lr"] = ref.param_groups[0]["lr"] = 8e-5
        flat_g.copy_(torch.cat([x.reshape(-1) for x in grads]))
        for p, x in zip(ref_params, grads):
            p.grad = x.clone()
        opt.step()
        ref.step()
    want = torch.cat([p.detach().reshape(-1) for p in ref_params])
    err = (flat_p - want).abs().max().item() / want.abs().max().item()
    assert err < 2e-6, err
    m_ref = torch.cat([ref.state[p]["exp_avg"].reshape(-1) for p in ref_params])
    v_ref = torch.cat([ref.state[p]["exp_avg_sq"].reshape(-1) for p in ref_params])
    # moments: elements that nearly cancel carry the rounding of the largest term, so compare against the tensor scale
    assert (opt.exp_avg - m_ref).abs().max().item() <= 2e-6 * m_ref.abs().max().item()
    assert (opt.exp_avg_sq - v_ref).abs().max().item() <= 2e-6 * v_ref.abs().max().item()


def test_flat_adam_rejects_bad_groups_and_cpu():
    from online_lang_splatting_b200.optim import FlatAdam
    dev = torch.device("cuda:0")
    p, g = torch.zeros(10, device=dev), torch.zeros(10, device=dev)
    with pytest.raises(ValueError):
        FlatAdam(p, g, [("a", 4, 1e-3)])
    with pytest.raises(RuntimeError):
        FlatAdam(torch.zeros(4), torch.zeros(4), [("a", 4, 1e-3)])
