"""The optional-input paths of the reference API: precomputed colours, precomputed 3D
covariances, scale modifier, the non-language rasterizer, and render()'s pipe flags / mask -- each against the oracle
or against the default path."""
import numpy as np
import pytest
import torch

import _util as U

pytestmark = pytest.mark.gpu


def _l2rel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _oracle(sc, tile, grads, **kw):
    from oracle.oracle import OracleRasterizer
    base = dict(means3D=sc["means3D"], opacities=sc["opacities"], language=sc["language"], W=sc["W"], H=sc["H"],
                tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], viewmatrix=sc["viewmatrix"], projmatrix=sc["projmatrix"],
                projmatrix_raw=sc["projmatrix_raw"], campos=sc["campos"], bg=sc["bg"], sh_degree=sc["sh_degree"], tile=tile)
    base.update(kw)
    o = OracleRasterizer(**base)
    out = o.forward()
    out["grads"] = o.backward(grads[0], grads[1], grads[2], compat=False)
    return out


def _ours(sc, dev, tile, grads, shs=None, colors=None, scales=None, rots=None, cov=None, scale_modifier=1.0):
    from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr
    rs = U.settings(sc, dev, tile=tile, backward_mode="exact", bitexact=True)._replace(scale_modifier=scale_modifier)
    leaf = lambda t: None if t is None else t.to(dev).clone().requires_grad_(True)
    t = {"means3D": leaf(sc["means3D"]), "language": leaf(sc["language"]), "opacities": leaf(sc["opacities"]),
         "shs": leaf(shs), "colors": leaf(colors), "scales": leaf(scales), "rots": leaf(rots), "cov": leaf(cov)}
    rast = dgr.LanguageGaussianRasterizer(rs)
    color, language, radii, depth, opacity, n_touched = rast(
        means3D=t["means3D"], means2D=torch.zeros_like(t["means3D"], requires_grad=True), opacities=t["opacities"],
        shs=t["shs"], colors_precomp=t["colors"], language_precomp=t["language"], scales=t["scales"], rotations=t["rots"],
        cov3D_precomp=t["cov"])
    wc, wl, wd = (g.to(dev) for g in grads)
    ((color * wc).sum() + (language * wl).sum() + (depth * wd).sum()).backward()
    return {"color": color.detach().cpu().numpy(), "language": language.detach().cpu().numpy(),
            "radii": radii.cpu().numpy(), "t": t}


def test_precomputed_colours():
    dev = torch.device("cuda:0")
    sc = U.make_scene(P=3000, F=15, W=120, H=80, seed=4, scale=0.07, bg=(0.2, 0.1, 0.4))
    grads = U.loss_weights(15, 120, 80, seed=3)
    colors = torch.rand(3000, 3, generator=torch.Generator().manual_seed(1))
    ours = _ours(sc, dev, 15, grads, colors=colors, scales=sc["scales"], rots=sc["rotations"])
    ora = _oracle(sc, 15, grads, colors_precomp=colors, scales=sc["scales"], rotations=sc["rotations"])
    assert np.array_equal(ours["radii"], ora["radii"])
    assert U.rel_err(ours["color"], ora["color"]) < 1e-4
    assert _l2rel(ours["t"]["colors"].grad.cpu().numpy(), ora["grads"]["dL_dcolors"]) < 2e-3
    assert ours["t"]["shs"] is None


def test_precomputed_covariance_and_scale_modifier():
    dev = torch.device("cuda:0")
    from online_lang_splatting_b200 import synthetic as S
    sc = U.make_scene(P=3000, F=3, W=120, H=80, seed=5, scale=0.07)
    grads = U.loss_weights(3, 120, 80, seed=3)
    pc = S.SyntheticGaussianModel({k: sc[k] for k in ("means3D", "scales", "rotations", "opacities", "shs", "language")},
                                  device="cpu")
    cov = pc.get_covariance(1.0).detach()
    ours = _ours(sc, dev, 16, grads, shs=sc["shs"], cov=cov)
    ora = _oracle(sc, 16, grads, shs=sc["shs"], cov3D_precomp=cov)
    assert np.array_equal(ours["radii"], ora["radii"])
    assert U.rel_err(ours["language"], ora["language"]) < 1e-4
    assert _l2rel(ours["t"]["cov"].grad.cpu().numpy(), ora["grads"]["dL_dcov3D"]) < 2e-3
    # scale modifier: kernels (scales * 0.5) == oracle with scale_modifier = 0.5
    ours = _ours(sc, dev, 16, grads, shs=sc["shs"], scales=sc["scales"], rots=sc["rotations"], scale_modifier=0.5)
    ora = _oracle(sc, 16, grads, shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"], scale_modifier=0.5)
    assert np.array_equal(ours["radii"], ora["radii"])
    assert _l2rel(ours["t"]["scales"].grad.cpu().numpy(), ora["grads"]["dL_dscales"]) < 2e-3


def test_plain_rasterizer_and_render_flags():
    from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr, synthetic as S
    from online_lang_splatting_b200.gaussian_renderer import render
    dev = torch.device("cuda:0")
    W, H = 128, 80
    g = S.make_gaussians(2500, 15, W, H, seed=2, scale_px_sigma=0.06)
    cam = S.make_camera(W, H, view=1, seed=2, device="cuda:0")
    pipe = S.PipelineParams()
    bg = torch.tensor([0.1, 0.3, 0.2], device=dev)
    pc = S.SyntheticGaussianModel(g, device=dev)
    ref = render(cam, pc, pipe, bg)
    # non-language model -> GaussianRasterizer (reference :421-480): same colour / depth, no "language" key
    pc_plain = S.SyntheticGaussianModel(g, device=dev, is_language=False)
    plain = render(cam, pc_plain, pipe, bg)
    assert "language" not in plain
    assert torch.equal(plain["render"], ref["render"]) and torch.equal(plain["depth"], ref["depth"])
    assert torch.equal(plain["radii"], ref["radii"])
    # pipe.compute_cov3D_python / convert_SHs_python: the Python-side precomputation gives the same picture
    pipe2 = S.PipelineParams()
    pipe2.compute_cov3D_python = True
    pipe2.convert_SHs_python = True
    alt = render(cam, pc, pipe2, bg)
    assert (alt["radii"] != ref["radii"]).float().mean().item() < 1e-3
    assert (alt["render"] - ref["render"]).abs().max().item() < 2e-3
    # scaling_modifier argument reaches the kernels
    small = render(cam, pc, pipe, bg, 0.5)
    assert small["radii"].float().mean().item() < ref["radii"].float().mean().item()
    # mask (reference Q4: dead code upstream, implemented as intended here): only the selected Gaussians are drawn
    mask = torch.zeros(2500, dtype=torch.bool, device=dev)
    mask[::2] = True
    part = render(cam, pc, pipe, bg, mask=mask)
    assert part["n_touched"] is None and part["radii"].shape[0] == int(mask.sum())
    assert part["opacity"].sum().item() < ref["opacity"].sum().item()
