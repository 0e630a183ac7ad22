"""Shared helpers of the test-suite: seeded scenes, and the three implementations side by side
(ours through the C ABI, the CPU oracle, the compiled reference when it is present on a GPU box)."""
from __future__ import annotations

import importlib.util
import math
import os
import sys
from typing import Dict, Optional

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from online_lang_splatting_b200 import synthetic as S  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def make_scene(P=2000, F=15, W=96, H=64, seed=0, view=0, sh_degree=0, scale=0.05, bg=(0.0, 0.0, 0.0)) -> Dict:
    g = S.make_gaussians(P, F, W, H, seed=seed, sh_degree=sh_degree, scale_px_sigma=scale)
    cam = S.make_camera(W, H, view=view, seed=seed)
    sc = dict(g)
    sc.update(P=P, F=F, W=W, H=H, sh_degree=sh_degree, tanfovx=math.tan(cam.FoVx * 0.5),
              tanfovy=math.tan(cam.FoVy * 0.5), viewmatrix=cam.world_view_transform.clone(),
              projmatrix=cam.full_proj_transform.clone(), projmatrix_raw=cam.projection_matrix.clone(),
              campos=cam.camera_center.clone(), bg=torch.tensor(bg, dtype=torch.float32))
    return sc


def add_lang_footprint(sc: Dict, seed=7, scale_sigma=0.35) -> Dict:
    """The disentangled variant's second footprint: independent opacity / scale / rotation for the language pass."""
    P = sc["P"]
    g = torch.Generator().manual_seed(seed)
    sc = dict(sc)
    sc["opacities_lang"] = torch.sigmoid(torch.randn(P, 1, generator=g) * 1.5)
    sc["scales_lang"] = sc["scales"] * torch.exp(scale_sigma * torch.randn(P, 3, generator=g))
    q = torch.randn(P, 4, generator=g)
    sc["rotations_lang"] = q / q.norm(dim=1, keepdim=True)
    return sc


def loss_weights(F, W, H, seed=1):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(3, H, W, generator=g), torch.randn(F, H, W, generator=g), torch.randn(1, H, W, generator=g))


def scene_from_npz(z) -> Dict:
    sc = {k: torch.from_numpy(np.asarray(z[k])) for k in ("means3D", "scales", "rotations", "opacities", "shs",
                                                         "language", "viewmatrix", "projmatrix", "projmatrix_raw",
                                                         "campos", "bg")}
    for k in ("opacities_lang", "scales_lang", "rotations_lang"):
        if k in z:
            sc[k] = torch.from_numpy(np.asarray(z[k]))
    for k in ("P", "F", "W", "H", "sh_degree"):
        sc[k] = int(z[k])
    sc["tanfovx"], sc["tanfovy"] = float(z["tanfovx"]), float(z["tanfovy"])
    return sc


def run_oracle(sc: Dict, tile=15, grads=None, compat=True):
    from oracle.oracle import OracleRasterizer
    o = OracleRasterizer(means3D=sc["means3D"], opacities=sc["opacities"], language=sc["language"], W=sc["W"],
                         H=sc["H"], tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], viewmatrix=sc["viewmatrix"],
                         projmatrix=sc["projmatrix"], projmatrix_raw=sc["projmatrix_raw"], campos=sc["campos"],
                         bg=sc["bg"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"],
                         sh_degree=sc["sh_degree"], tile=tile)
    out = o.forward()
    if grads is not None:
        out["grads"] = o.backward(grads[0], grads[1], grads[2], compat=compat)
    return out


def settings(sc: Dict, device, tile=15, backward_mode="compat", bitexact=True):
    from online_lang_splatting_b200.diff_gaussian_rasterization import GaussianRasterizationSettings
    d = lambda t: t.to(device)
    return GaussianRasterizationSettings(
        image_height=sc["H"], image_width=sc["W"], tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], bg=d(sc["bg"]),
        scale_modifier=1.0, viewmatrix=d(sc["viewmatrix"]), projmatrix=d(sc["projmatrix"]),
        projmatrix_raw=d(sc["projmatrix_raw"]), sh_degree=sc["sh_degree"], campos=d(sc["campos"]), prefiltered=False,
        debug=True, tile_size=tile, backward_mode=backward_mode, bitexact_blend=bitexact)


def run_ours(sc: Dict, device, tile=15, grads=None, backward_mode="compat", bitexact=True):
    """Forward (and optionally backward) through the public module surface; returns numpy arrays
    plus the decoded internal state."""
    from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr
    from online_lang_splatting_b200.debug import workspace_arrays
    rs = settings(sc, device, tile, backward_mode, bitexact)
    req = grads is not None
    leaf = lambda t: t.to(device).clone().requires_grad_(req)
    means3D, shs, lang = leaf(sc["means3D"]), leaf(sc["shs"]), leaf(sc["language"])
    opac, scales, rots = leaf(sc["opacities"]), leaf(sc["scales"]), leaf(sc["rotations"])
    means2D = torch.zeros_like(means3D, requires_grad=req)
    theta = torch.zeros(3, device=device, requires_grad=req)
    rho = torch.zeros(3, device=device, requires_grad=req)
    fn = dgr._RasterizeLanguageGaussians
    color, language, radii, depth, opacity, n_touched = fn.apply(
        means3D, means2D, shs, torch.Tensor([]), lang, opac, scales, rots, torch.Tensor([]), theta, rho, rs)
    st = color.grad_fn.state if req else None
    if st is None:
        # no autograd graph: run the native forward again only to get at the state
        R, color, language, radii, depth, opacity, n_touched, st = dgr._forward_native(
            means3D, shs, torch.Tensor([]), lang, opac, scales, rots, torch.Tensor([]), rs)
    ws = {k: v.detach().cpu().numpy() for k, v in workspace_arrays(st).items()}
    out = {"color": color, "language": language, "radii": radii, "depth": depth, "opacity": opacity,
           "n_touched": n_touched}
    out = {k: v.detach().cpu().numpy() for k, v in out.items()}
    out["R"] = st.R
    out["ws"] = ws
    if req:
        wc, wl, wd = (g.to(device) for g in grads)
        loss = (color * wc).sum() + (language * wl).sum() + (depth * wd).sum()
        loss.backward()
        out["grads"] = {"means3D": means3D.grad, "means2D": means2D.grad, "shs": shs.grad, "language": lang.grad,
                        "opacities": opac.grad, "scales": scales.grad, "rotations": rots.grad, "theta": theta.grad,
                        "rho": rho.grad}
        out["grads"] = {k: v.detach().cpu().numpy() for k, v in out["grads"].items()}
    return out


def ref_module(name="ref_P_C"):
    """The reference's own CUDA extension compiled by oracle/build_ref.py (None if absent)."""
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(path) or not torch.cuda.is_available():
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _a128(x):
    return (x + 127) // 128 * 128


def run_ref(mod, sc: Dict, device, grads=None, F=15, tile=15):
    """Run the compiled reference (P/ variant) and decode its opaque buffers."""
    d = lambda t: t.to(device).contiguous()
    P, W, H = sc["P"], sc["W"], sc["H"]
    e = torch.Tensor([])
    args = (d(sc["bg"]), d(sc["means3D"]), e, d(sc["language"]), d(sc["opacities"]), d(sc["scales"]),
            d(sc["rotations"]), 1.0, e, d(sc["viewmatrix"]), d(sc["projmatrix"]), d(sc["projmatrix_raw"]),
            sc["tanfovx"], sc["tanfovy"], H, W, d(sc["shs"]), sc["sh_degree"], d(sc["campos"]), False, False)
    R, color, language, radii, geom, binning, img, depth, opacity, n_touched = mod.rasterize_language_gaussians(*args)
    torch.cuda.synchronize()
    out = {"color": color, "language": language, "radii": radii, "depth": depth, "opacity": opacity,
           "n_touched": n_touched}
    out = {k: v.cpu().numpy() for k, v in out.items()}
    out["R"] = int(R)
    gb, bb, ib = geom.cpu().numpy(), binning.cpu().numpy(), img.cpu().numpy()
    HW = H * W
    gx, gy = (W + tile - 1) // tile, (H + tile - 1) // tile
    o = 0
    out["depths"] = gb[o:o + 4 * P].view(np.float32).copy(); o = _a128(o + 4 * P)
    out["clamped"] = gb[o:o + 3 * P].reshape(P, 3).copy(); o = _a128(o + 3 * P)
    o = _a128(o + 4 * P)  # internal_radii
    out["means2D"] = gb[o:o + 8 * P].view(np.float32).reshape(P, 2).copy(); o = _a128(o + 8 * P)
    out["cov3D"] = gb[o:o + 24 * P].view(np.float32).reshape(P, 6).copy(); o = _a128(o + 24 * P)
    out["conic_opacity"] = gb[o:o + 16 * P].view(np.float32).reshape(P, 4).copy(); o = _a128(o + 16 * P)
    out["rgb"] = gb[o:o + 12 * P].view(np.float32).reshape(P, 3).copy(); o = _a128(o + 12 * P)
    o = _a128(o + 4 * F * P)  # language slab (never written)
    out["tiles_touched"] = gb[o:o + 4 * P].view(np.uint32).copy()
    end = len(gb) - 128
    out["point_offsets"] = gb[end - 4 * P:end].view(np.uint32).copy()
    o = 0
    out["point_list"] = bb[o:o + 4 * R].view(np.uint32).copy(); o = _a128(o + 4 * R)
    o = _a128(o + 4 * R)
    out["keys_sorted"] = bb[o:o + 8 * R].view(np.uint64).copy()
    o = 0
    out["final_T"] = ib[o:o + 4 * HW].view(np.float32).reshape(H, W).copy(); o = _a128(o + 4 * HW)
    out["n_contrib"] = ib[o:o + 4 * HW].view(np.uint32).reshape(H, W).copy(); o = _a128(o + 4 * HW)
    out["ranges"] = ib[o:o + 8 * gx * gy].view(np.uint32).reshape(gx * gy, 2).copy()
    if grads is not None:
        wc, wl, wd = (d(g) for g in grads)
        bargs = (d(sc["bg"]), d(sc["means3D"]), radii, e, d(sc["language"]), d(sc["scales"]), d(sc["rotations"]), 1.0,
                 e, d(sc["viewmatrix"]), d(sc["projmatrix"]), d(sc["projmatrix_raw"]), sc["tanfovx"], sc["tanfovy"],
                 wc, wl, wd, d(sc["shs"]), sc["sh_degree"], d(sc["campos"]), geom, R, binning, img, False)
        res = mod.rasterize_language_gaussians_backward(*bargs)
        torch.cuda.synchronize()
        names = ("means2D", "colors", "language", "opacities", "means3D", "cov3D", "shs", "scales", "rotations", "tau")
        g = {n: t.cpu().numpy() for n, t in zip(names, res)}
        tau = g["tau"].reshape(-1, 6).sum(0)
        g["rho"], g["theta"] = tau[:3], tau[3:]
        out["grads"] = g
    return out


def keys_from_ours(ws: Dict, gx: int) -> np.ndarray:
    """Rebuild the reference's sorted 64-bit keys (tile<<32 | depth bits) from our per-tile buckets."""
    ranges = ws["ranges"].astype(np.int64)
    keys = ws["keys"].view(np.uint64)
    out = np.zeros(len(keys), np.uint64)
    for t in range(ranges.shape[0]):
        a, b = ranges[t]
        if b > a:
            out[a:b] = (np.uint64(t) << np.uint64(32)) | (keys[a:b] >> np.uint64(32))
    return out


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)


# ---------------------------------------------------------------------------------------------------
# disentangled variant (D/)
# ---------------------------------------------------------------------------------------------------
def run_oracle_dis(sc: Dict, tile=16, grads=None, compat=True):
    from oracle.oracle import OracleDisRasterizer
    o = OracleDisRasterizer(means3D=sc["means3D"], opacities=sc["opacities"], language=sc["language"], W=sc["W"],
                            H=sc["H"], tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], viewmatrix=sc["viewmatrix"],
                            projmatrix=sc["projmatrix"], projmatrix_raw=sc["projmatrix_raw"], campos=sc["campos"],
                            bg=sc["bg"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"],
                            sh_degree=sc["sh_degree"], tile=tile, opacities_lang=sc["opacities_lang"],
                            scales_lang=sc["scales_lang"], rotations_lang=sc["rotations_lang"])
    out = o.forward()
    if grads is not None:
        out["grads"] = o.backward(grads[0], grads[1], grads[2], compat=compat)
    return out


def settings_dis(sc: Dict, device, tile=16, backward_mode="compat", bitexact=True):
    from online_lang_splatting_b200.diff_gaussian_rasterization_disentangle import GaussianRasterizationSettings
    d = lambda t: t.to(device)
    return GaussianRasterizationSettings(
        image_height=sc["H"], image_width=sc["W"], tanfovx=sc["tanfovx"], tanfovy=sc["tanfovy"], bg=d(sc["bg"]),
        scale_modifier=1.0, viewmatrix=d(sc["viewmatrix"]), projmatrix=d(sc["projmatrix"]),
        projmatrix_raw=d(sc["projmatrix_raw"]), sh_degree=sc["sh_degree"], campos=d(sc["campos"]), prefiltered=False,
        debug=True, tile_size=tile, backward_mode=backward_mode, bitexact_blend=bitexact)


DIS_GRAD_NAMES = ("means3D", "means2D", "shs", "language", "opacities", "opacities_lang", "scales", "scales_lang",
                  "rotations", "rotations_lang", "theta", "rho")


def run_ours_dis(sc: Dict, device, tile=16, grads=None, backward_mode="compat", bitexact=True):
    """Forward (+ backward) through the public disentangled module; numpy outputs + decoded state of both lists."""
    from online_lang_splatting_b200 import diff_gaussian_rasterization_disentangle as dd
    from online_lang_splatting_b200.debug import workspace_arrays_dis
    rs = settings_dis(sc, device, tile, backward_mode, bitexact)
    req = grads is not None
    leaf = lambda t: t.to(device).clone().requires_grad_(req)
    t = {k: leaf(sc[k]) for k in ("means3D", "shs", "language", "opacities", "opacities_lang", "scales", "scales_lang",
                                  "rotations", "rotations_lang")}
    t["means2D"] = torch.zeros_like(t["means3D"], requires_grad=req)
    t["theta"] = torch.zeros(3, device=device, requires_grad=req)
    t["rho"] = torch.zeros(3, device=device, requires_grad=req)
    rast = dd.LanguageGaussianRasterizer(rs)
    color, language, radii, radii_lang, depth, opacity, opacity_lang, n_touched, n_touched_lang = rast(
        means3D=t["means3D"], means2D=t["means2D"], opacities=t["opacities"], opacities_lang=t["opacities_lang"],
        shs=t["shs"], language_precomp=t["language"], scales=t["scales"], scales_lang=t["scales_lang"],
        rotations=t["rotations"], rotations_lang=t["rotations_lang"], theta=t["theta"], rho=t["rho"])
    if req:
        st = color.grad_fn.state
    else:
        e = torch.Tensor([])
        st = dd._forward_native(t["means3D"], t["shs"], e, t["language"], t["opacities"], t["opacities_lang"], t["scales"],
                                t["scales_lang"], t["rotations"], t["rotations_lang"], e, e, rs)[-1]
    wc, wl = workspace_arrays_dis(st)
    out = {"color": color, "language": language, "radii": radii, "radii_lang": radii_lang, "depth": depth,
           "opacity": opacity, "opacity_lang": opacity_lang, "n_touched": n_touched, "n_touched_lang": n_touched_lang}
    out = {k: v.detach().cpu().numpy() for k, v in out.items()}
    out["R"], out["R_lang"] = st.R, st.R_lang
    out["ws"] = {k: v.detach().cpu().numpy() for k, v in wc.items()}
    out["ws_lang"] = {k: v.detach().cpu().numpy() for k, v in wl.items()}
    if req:
        gc, gl, gd = (g.to(device) for g in grads)
        loss = (color * gc).sum() + (language * gl).sum() + (depth * gd).sum()
        loss.backward()
        out["grads"] = {k: t[k].grad.detach().cpu().numpy() for k in DIS_GRAD_NAMES}
    return out


def run_ref_dis(mod, sc: Dict, device, grads=None, tile=16):
    """Run the compiled reference D/ (F=3, 16x16) and decode its opaque buffers (D/rasterizer_impl.cu:173-241)."""
    d = lambda t: t.to(device).contiguous()
    P, W, H, F = sc["P"], sc["W"], sc["H"], sc["F"]
    e = torch.Tensor([])
    args = (d(sc["bg"]), d(sc["means3D"]), e, d(sc["language"]), d(sc["opacities"]), d(sc["opacities_lang"]),
            d(sc["scales"]), d(sc["scales_lang"]), d(sc["rotations"]), d(sc["rotations_lang"]), 1.0, e, e,
            d(sc["viewmatrix"]), d(sc["projmatrix"]), d(sc["projmatrix_raw"]), sc["tanfovx"], sc["tanfovy"], H, W,
            d(sc["shs"]), sc["sh_degree"], d(sc["campos"]), False, False)
    (R, Rl, color, language, radii, radii_lang, geom, binning, binning_lang, img, depth, opacity, opacity_lang, n_touched,
     n_touched_lang) = mod.rasterize_language_gaussians(*args)
    torch.cuda.synchronize()
    out = {"color": color, "language": language, "radii": radii, "radii_lang": radii_lang, "depth": depth,
           "opacity": opacity, "opacity_lang": opacity_lang, "n_touched": n_touched, "n_touched_lang": n_touched_lang}
    out = {k: v.cpu().numpy() for k, v in out.items()}
    out["R"], out["R_lang"] = int(R), int(Rl)
    gb, ib = geom.cpu().numpy(), img.cpu().numpy()
    HW = H * W
    gx, gy = (W + tile - 1) // tile, (H + tile - 1) // tile
    o = 0

    def take(nbytes, dtype, shape):
        nonlocal o
        a = gb[o:o + nbytes].view(dtype).reshape(shape).copy()
        o = _a128(o + nbytes)
        return a
    out["depths"] = take(4 * P, np.float32, (P,))
    out["clamped"] = take(3 * P, np.uint8, (P, 3))
    take(4 * P, np.int32, (P,)); take(4 * P, np.int32, (P,))  # internal radii (unused: radii tensors are passed in)
    out["means2D"] = take(8 * P, np.float32, (P, 2))
    out["cov3D"] = take(24 * P, np.float32, (P, 6))
    out["cov3D_lang"] = take(24 * P, np.float32, (P, 6))
    out["conic_opacity"] = take(16 * P, np.float32, (P, 4))
    out["conic_opacity_lang"] = take(16 * P, np.float32, (P, 4))
    out["rgb"] = take(12 * P, np.float32, (P, 3))
    take(4 * F * P, np.float32, (P, F))  # language slab (never written)
    out["tiles_touched"] = take(4 * P, np.uint32, (P,))
    out["tiles_touched_lang"] = take(4 * P, np.uint32, (P,))
    for key, bufname, n in (("", binning, int(R)), ("_lang", binning_lang, int(Rl))):
        bb = bufname.cpu().numpy()
        q = 0
        out["point_list" + key] = bb[q:q + 4 * n].view(np.uint32).copy(); q = _a128(q + 4 * n)
        q = _a128(q + 4 * n)
        out["keys_sorted" + key] = bb[q:q + 8 * n].view(np.uint64).copy()
    q = 0
    out["final_T"] = ib[q:q + 4 * HW].view(np.float32).reshape(H, W).copy(); q = _a128(q + 4 * HW)
    out["n_contrib"] = ib[q:q + 4 * HW].view(np.uint32).reshape(H, W).copy(); q = _a128(q + 4 * HW)
    out["final_T_lang"] = ib[q:q + 4 * HW].view(np.float32).reshape(H, W).copy(); q = _a128(q + 4 * HW)
    out["n_contrib_lang"] = ib[q:q + 4 * HW].view(np.uint32).reshape(H, W).copy(); q = _a128(q + 4 * HW)
    out["ranges"] = ib[q:q + 8 * gx * gy].view(np.uint32).reshape(gx * gy, 2).copy(); q = _a128(q + 8 * HW)
    out["ranges_lang"] = ib[q:q + 8 * gx * gy].view(np.uint32).reshape(gx * gy, 2).copy()
    if grads is not None:
        wc, wl, wd = (d(g) for g in grads)
        bargs = (d(sc["bg"]), d(sc["means3D"]), radii, radii_lang, e, d(sc["language"]), d(sc["scales"]), d(sc["scales_lang"]),
                 d(sc["rotations"]), d(sc["rotations_lang"]), 1.0, e, e, d(sc["viewmatrix"]), d(sc["projmatrix"]),
                 d(sc["projmatrix_raw"]), sc["tanfovx"], sc["tanfovy"], wc, wl, wd, d(sc["shs"]), sc["sh_degree"],
                 d(sc["campos"]), geom, R, Rl, binning, binning_lang, img, False)
        res = mod.rasterize_language_gaussians_backward(*bargs)
        torch.cuda.synchronize()
        names = ("means2D", "colors", "language", "opacities", "opacities_lang", "means3D", "cov3D", "cov3D_lang", "shs",
                 "scales", "scales_lang", "rotations", "rotations_lang", "tau")
        g = {n: t.cpu().numpy() for n, t in zip(names, res)}
        tau = g["tau"].reshape(-1, 6).sum(0)
        g["rho"], g["theta"] = tau[:3], tau[3:]
        out["grads"] = g
    return out
