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


def loss_weights(F, W, H, seed=1):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(3, H, W, generator=g), torch.randn(F, H, W, generator=g), torch.randn(1, H, W, generator=g))


def scene_from_npz(z) -> Dict:
    sc = {k: torch.from_numpy(np.asarray(z[k])) for k in ("means3D", "scales", "rotations", "opacities", "shs",
                                                         "language", "viewmatrix", "projmatrix", "projmatrix_raw",
                                                         "campos", "bg")}
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
