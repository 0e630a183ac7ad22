"""Multi-view batch (ols_lang_forward_batch / ols_lang_backward_batch, render_batch): V views of the same Gaussians in
one set of launches must give exactly what V single-view calls give -- bit-identical images and lists per view, and
parameter gradients equal to the sum autograd forms over the V reference-style calls (up to the order of the float
atomics, which is not fixed in either path; the tolerance is stated)."""
import numpy as np
import pytest
import torch

import _util as U

pytestmark = pytest.mark.gpu


def _model(cuda, P=5000, W=128, H=80, seed=3):
    from online_lang_splatting_b200 import synthetic as S
    g = S.make_gaussians(P, 15, W, H, seed=seed, scale_px_sigma=0.05)
    return S.SyntheticGaussianModel(g, device=cuda, requires_grad=True), S


def _loss(out, w):
    return (out["render"] * w[0]).sum() + (out["language"] * w[1]).sum() + (out["depth"] * w[2]).sum()


@pytest.mark.parametrize("V", [1, 3, 5])
@pytest.mark.parametrize("mode", ["compat", "exact"])
def test_render_batch_equals_single_renders(cuda, V, mode):
    import online_lang_splatting_b200.gaussian_renderer as GR
    W, H = 128, 80
    pc, S = _model(cuda, W=W, H=H)
    pipe, bg = S.PipelineParams(), torch.tensor([0.1, 0.3, 0.2], device=cuda)
    weights = [tuple(t.to(cuda) for t in U.loss_weights(15, W, H, seed=10 + v)) for v in range(V)]
    saved = (GR.BACKWARD_MODE, GR.BITEXACT_BLEND)
    GR.BACKWARD_MODE, GR.BITEXACT_BLEND = mode, True
    try:
        cams = [S.make_camera(W, H, view=v, seed=3, device="cuda") for v in range(V)]
        singles, total = [], 0.0
        for v in range(V):
            o = GR.render(cams[v], pc, pipe, bg)
            singles.append(o)
            total = total + _loss(o, weights[v])
        total.backward()
        ref_grads = [p.grad.detach().clone() for p in pc.parameters() if p.numel()]
        ref_vs = [o["viewspace_points"].grad.detach().clone() for o in singles]
        ref_pose = [(c.cam_rot_delta.grad.detach().clone(), c.cam_trans_delta.grad.detach().clone()) for c in cams]
        for p in pc.parameters():
            p.grad = None
        cams = [S.make_camera(W, H, view=v, seed=3, device="cuda") for v in range(V)]
        outs = GR.render_batch(cams, pc, pipe, bg)
        total = 0.0
        for v in range(V):
            for k in ("render", "language", "depth", "opacity", "radii", "n_touched"):
                assert torch.equal(outs[v][k], singles[v][k]), (v, k)
            total = total + _loss(outs[v], weights[v])
        total.backward()
    finally:
        GR.BACKWARD_MODE, GR.BITEXACT_BLEND = saved
    got = [p.grad.detach() for p in pc.parameters() if p.numel()]
    assert len(got) == len(ref_grads)
    for a, b in zip(got, ref_grads):
        err = (a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)
        assert err.item() < 2e-5, err.item()     # float-atomic ordering only
    for v in range(V):
        a, b = outs[v]["viewspace_points"].grad, ref_vs[v]
        assert ((a - b).norm() / b.norm().clamp_min(1e-30)).item() < 2e-5
        for a, b in zip((cams[v].cam_rot_delta.grad, cams[v].cam_trans_delta.grad), ref_pose[v]):
            assert (a - b).abs().max().item() <= 2e-4 * b.abs().max().item() + 1e-9


def test_batch_validation(cuda):
    from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr
    sc = U.make_scene(P=500, F=15, W=64, H=48, seed=1)
    d = lambda k: sc[k].to(cuda)
    rs = U.settings(sc, cuda)
    e = torch.Tensor([])
    with pytest.raises(RuntimeError, match="share image size"):
        dgr._forward_native_batch(d("means3D"), d("shs"), e, d("language"), d("opacities"), d("scales"), d("rotations"), e,
                                  [rs, rs._replace(image_width=80)])
    with pytest.raises(RuntimeError, match="1..16"):
        dgr._forward_native_batch(d("means3D"), d("shs"), e, d("language"), d("opacities"), d("scales"), d("rotations"), e,
                                  [rs] * 17)


def test_batch_headline_shape_lists_match_single(cuda):
    """8 views at 200k Gaussians / 960x540: every view's sorted list, ranges and images equal the single-view call."""
    from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr
    from online_lang_splatting_b200.debug import workspace_arrays
    P, W, H = 200000, 960, 540
    scs = [U.make_scene(P=P, F=15, W=W, H=H, seed=0, view=v, scale=0.01) for v in range(8)]
    d = lambda k: scs[0][k].to(cuda)
    e = torch.Tensor([])
    rs_list = [U.settings(sc, cuda, bitexact=False)._replace(debug=False) for sc in scs]
    params = (d("means3D"), d("shs"), e, d("language"), d("opacities"), d("scales"), d("rotations"), e)
    outs, st = dgr._forward_native_batch(*params, rs_list)
    torch.cuda.synchronize()
    for v in (0, 3, 7):
        R, color, language, radii, depth, opacity, n_touched, st1 = dgr._forward_native(*params, rs_list[v])
        assert torch.equal(outs[v][0], color) and torch.equal(outs[v][1], language) and torch.equal(outs[v][2], radii)
        assert torch.equal(outs[v][5], n_touched)


def test_fused_densification_stats_equal_per_view_updates(cuda):
    """out["stats"] of the batched backward == densification.update_stats applied view by view to the same gradients
    (utils/slam_backend.py:719-728 + gaussian_model.py:965-969)."""
    from online_lang_splatting_b200 import densification as DN
    from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr
    P, W, H, V = 6000, 128, 80, 4
    scs = [U.make_scene(P=P, F=15, W=W, H=H, seed=2, view=v, scale=0.05) for v in range(V)]
    d = lambda k: scs[0][k].to(cuda)
    e = torch.Tensor([])
    rs_list = [U.settings(sc, cuda, bitexact=True)._replace(debug=False) for sc in scs]
    params = (d("means3D"), d("shs"), e, d("language"), d("opacities"), d("scales"), d("rotations"), e)
    w = [t.to(cuda) for t in U.loss_weights(15, W, H, seed=4)]
    outs, st = dgr._forward_native_batch(*params, rs_list)
    radii = [o[2] for o in outs]
    mx = torch.rand(P, device=cuda) * 3
    acc, den = torch.rand(P, device=cuda), torch.rand(P, device=cuda).round()
    mx0, acc0, den0 = mx.clone(), acc.clone(), den.clone()
    g = dgr._backward_native_batch(st, radii, [w[0]] * V, [w[1]] * V, [w[2]] * V, out={"stats": (mx, acc, den)})
    for v in range(V):
        DN.update_stats(radii[v], g["means2D"][v].contiguous(), mx0, acc0.view(P, 1), den0.view(P, 1))
    assert torch.equal(mx, mx0) and torch.equal(den, den0)
    assert torch.allclose(acc, acc0, rtol=1e-5, atol=1e-7)
