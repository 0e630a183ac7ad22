"""Fused flat Adam against torch.optim.Adam with the reference's configuration
(one group per tensor, eps = 1e-15; gaussian_model.py:393-437).  fp32, tolerance 2e-6 relative after 5 steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("capturable", [False, True])
def test_flat_adam_matches_torch_adam_groups(capturable):
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
    opt = FlatAdam(flat_p, flat_g, [(nm, t.numel(), lr) for (nm, _, lr), t in zip(shapes, tensors)], capturable=capturable)
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
    assert err < (4e-6 if capturable else 2e-6), err     # capturable: bias corrections evaluated on the device
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


def test_flat_adam_applies_activation_jacobians_like_autograd():
    """the flat gradient buffer holds gradients w.r.t. the ACTIVATED rasterizer inputs.  With the
    groups of FlatGradBuffer.adam_groups the fused step equals torch.optim.Adam on the raw parameters with autograd
    applying exp / sigmoid / normalize in between (gaussian_model.py:93-130,393-437), incl. f_dc / f_rest learning rates."""
    from online_lang_splatting_b200.optim import FlatAdam
    from online_lang_splatting_b200.sharding import FlatParams, FlatGradBuffer
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    P, F, M = 4099, 15, 4
    raw = {"means3D": torch.randn(P, 3, generator=g), "sh": torch.randn(P, M, 3, generator=g), "opacity": torch.randn(P, 1, generator=g),
           "scales": torch.randn(P, 3, generator=g) * 0.5 - 3, "rotations": torch.randn(P, 4, generator=g) * 2, "language": torch.randn(P, F, generator=g)}
    fp = FlatParams({k: v.to(dev) for k, v in raw.items()}, F, M, device=dev)
    fg = FlatGradBuffer(P, F, M, device=dev)
    lr = {"xyz": 1.6e-4, "f_dc": 2.5e-3, "f_rest": 2.5e-3 / 20, "opacity": 0.05, "scaling": 1e-3, "rotation": 1e-3, "f_language": 2.5e-3}
    opt = FlatAdam(fp.flat, fg.flat, fg.adam_groups(lr))
    ref = {k: torch.nn.Parameter(v.to(dev).clone()) for k, v in raw.items()}
    f_dc = torch.nn.Parameter(raw["sh"][:, :1].to(dev).clone())
    f_rest = torch.nn.Parameter(raw["sh"][:, 1:].to(dev).clone())
    topt = torch.optim.Adam([{"params": [ref["means3D"]], "lr": lr["xyz"]}, {"params": [f_dc], "lr": lr["f_dc"]},
                             {"params": [f_rest], "lr": lr["f_rest"]}, {"params": [ref["opacity"]], "lr": lr["opacity"]},
                             {"params": [ref["scales"]], "lr": lr["scaling"]}, {"params": [ref["rotations"]], "lr": lr["rotation"]},
                             {"params": [ref["language"]], "lr": lr["f_language"]}], lr=0.0, eps=1e-15)
    for it in range(4):
        up = {k: torch.randn(v.shape, generator=g).to(dev) for k, v in raw.items()}      # dL/d(activated input)
        act = fp.activate()
        assert torch.allclose(act["scales"], torch.exp(ref["scales"]), rtol=2e-6)
        assert torch.allclose(act["rotations"], torch.nn.functional.normalize(ref["rotations"]), rtol=1e-5, atol=1e-7)
        assert torch.allclose(act["opacity"], torch.sigmoid(ref["opacity"]), rtol=2e-6)
        for k in up:
            fg.views[k].copy_(up[k])
        topt.zero_grad()
        sh = torch.cat((f_dc, f_rest), dim=1)
        loss = ((ref["means3D"] * up["means3D"]).sum() + (sh * up["sh"]).sum() + (torch.sigmoid(ref["opacity"]) * up["opacity"]).sum()
                + (torch.exp(ref["scales"]) * up["scales"]).sum() + (torch.nn.functional.normalize(ref["rotations"]) * up["rotations"]).sum()
                + (ref["language"] * up["language"]).sum())
        loss.backward()
        opt.step()
        topt.step()
    want = {"means3D": ref["means3D"], "sh": torch.cat((f_dc, f_rest), dim=1), "opacity": ref["opacity"], "scales": ref["scales"],
            "rotations": ref["rotations"], "language": ref["language"]}
    for k, v in want.items():
        err = (fp.views[k] - v.detach()).abs().max().item() / v.detach().abs().max().item()
        assert err < 5e-6, (k, err)
