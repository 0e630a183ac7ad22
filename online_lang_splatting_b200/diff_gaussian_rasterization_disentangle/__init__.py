"""Drop-in for the *disentangled* rasterizer package of the reference
(submodules/diff-gaussian-rasterization-disentangle-optim/diff_gaussian_rasterization/__init__.py, "D/").

Every Gaussian carries a second footprint for the language pass -- ``opacities_lang``, ``scales_lang``,
``rotations_lang`` (or ``cov3D_precomp_lang``) -- so colour + depth and the language features are binned,
sorted and blended independently.  Same public names, argument order, return arity (9 forward values,
16 gradient slots) and error behaviour as the reference:

* ``GaussianRasterizationSettings``   (D/ :467-481; three optional trailing fields added)
* ``LanguageGaussianRasterizer``      (D/ :543-666)  -> 9 returns
* ``rasterize_language_gaussians``    (D/ :170-211)

The reference builds D/ with 16x16 tiles (D/cuda_rasterizer/config.h:15-18), which is the default here.
Everything goes through the ``ols_dis_*`` entry points of ``include/ols_b200.h``; there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, NamedTuple, Tuple

import torch
import torch.nn as nn

from .. import _native as N
from ..diff_gaussian_rasterization import _RasterizerBase, _f32c, _flags, _empty, _SLACK


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    # --- extensions (defaults reproduce the reference build: D/config.h BLOCK_X = BLOCK_Y = 16) ---
    tile_size: int = 16
    backward_mode: str = "compat"  # "compat": reference gradients incl. its quirks; "exact": true gradients
    bitexact_blend: bool = False


_R_HINT: Dict[Tuple, Tuple[int, int]] = {}
CHECK_OVERFLOW = True


class _Ctx:
    __slots__ = ("args", "keep", "R", "R_lang")


def _forward_native(means3D, sh, colors_precomp, language_precomp, opacities, opacities_lang, scales, scales_lang,
                    rotations, rotations_lang, cov3Ds_precomp, cov3Ds_precomp_lang, rs: GaussianRasterizationSettings):
    N.require_cuda()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # D/rasterize_points.cu
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor: the rasterizer has no CPU path")
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)
    tile = int(getattr(rs, "tile_size", 16))
    if language_precomp is None or language_precomp.dim() != 2 or (P > 0 and language_precomp.numel() == 0):
        raise RuntimeError("language_precomp is required by the language rasterizer")
    F = int(language_precomp.shape[1])
    if P == 0:  # nothing to rasterize
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        e = lambda: torch.empty((0,), dtype=torch.int32, device=dev)
        return 0, 0, z(3, H, W), z(F, H, W), e(), e(), z(1, H, W), z(1, H, W), z(1, H, W), e(), e(), None
    keep = {
        "means3D": _f32c(means3D), "language": _f32c(language_precomp), "opacities": _f32c(opacities),
        "opacities_lang": _f32c(opacities_lang),
        "bg": _f32c(rs.bg).to(dev), "viewmatrix": _f32c(rs.viewmatrix).to(dev),
        "projmatrix": _f32c(rs.projmatrix).to(dev), "projmatrix_raw": _f32c(rs.projmatrix_raw).to(dev),
        "campos": _f32c(rs.campos).to(dev),
    }
    for name, t in (("shs", sh), ("colors_precomp", colors_precomp), ("scales", scales), ("rotations", rotations),
                    ("cov3D_precomp", cov3Ds_precomp), ("scales_lang", scales_lang), ("rotations_lang", rotations_lang),
                    ("cov3D_precomp_lang", cov3Ds_precomp_lang)):
        keep[name] = None if (t is None or t.numel() == 0) else _f32c(t).to(dev)
    if keep["shs"] is None and keep["colors_precomp"] is None:
        raise RuntimeError("For non-RGB, provide precomputed Gaussian colors!")  # D/rasterizer_impl.cu:448-451
    M = 0 if keep["shs"] is None else int(keep["shs"].shape[1])

    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    color, language = torch.empty((3, H, W), **f32), torch.empty((F, H, W), **f32)
    depth, opacity, opacity_lang = (torch.empty((1, H, W), **f32) for _ in range(3))
    radii, radii_lang, n_touched, n_touched_lang = (torch.empty((P,), **i32) for _ in range(4))
    lib = N.lib()
    key = (dev.index, P, W, H, tile)
    stream = torch.cuda.current_stream(dev).cuda_stream
    hint = _R_HINT.get(key)
    cap, cap_l = (8 * P + _SLACK, 8 * P + _SLACK) if hint is None else (int(hint[0] * 1.25) + _SLACK, int(hint[1] * 1.25) + _SLACK)
    with torch.cuda.device(dev):
        for attempt in range(3):
            nbytes = lib.ols_dis_workspace_size(P, F, W, H, tile, cap, cap_l)
            if nbytes == 0:
                raise RuntimeError("invalid rasterizer configuration")
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            base = N.RasterArgs(
                P=P, F=F, sh_degree=int(rs.sh_degree), M=M, W=W, H=H, tile=tile, flags=_flags(rs),
                tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy), scale_modifier=float(rs.scale_modifier),
                d_bg=N.ptr(keep["bg"]), d_means3D=N.ptr(keep["means3D"]), d_shs=N.ptr(keep["shs"]),
                d_colors_precomp=N.ptr(keep["colors_precomp"]), d_language=N.ptr(keep["language"]),
                d_opacities=N.ptr(keep["opacities"]), d_scales=N.ptr(keep["scales"]),
                d_rotations=N.ptr(keep["rotations"]), d_cov3D_precomp=N.ptr(keep["cov3D_precomp"]),
                d_viewmatrix=N.ptr(keep["viewmatrix"]), d_projmatrix=N.ptr(keep["projmatrix"]),
                d_projmatrix_raw=N.ptr(keep["projmatrix_raw"]), d_campos=N.ptr(keep["campos"]),
                d_workspace=ws.data_ptr(), workspace_bytes=nbytes, R_cap=cap)
            args = N.DisArgs(base=base, d_opacities_lang=N.ptr(keep["opacities_lang"]),
                             d_scales_lang=N.ptr(keep["scales_lang"]), d_rotations_lang=N.ptr(keep["rotations_lang"]),
                             d_cov3D_precomp_lang=N.ptr(keep["cov3D_precomp_lang"]), R_cap_lang=cap_l)
            out = N.DisFwdOut(d_color=color.data_ptr(), d_language=language.data_ptr(), d_depth=depth.data_ptr(),
                              d_opacity=opacity.data_ptr(), d_opacity_lang=opacity_lang.data_ptr(),
                              d_radii=radii.data_ptr(), d_radii_lang=radii_lang.data_ptr(),
                              d_n_touched=n_touched.data_ptr(), d_n_touched_lang=n_touched_lang.data_ptr())
            N.check(lib.ols_dis_forward(C.byref(args), C.byref(out), stream))
            R = R_l = -1
            if CHECK_OVERFLOW:
                ic, il = N.FwdInfo(), N.FwdInfo()
                N.check(lib.ols_dis_read_info(C.byref(args), C.byref(ic), C.byref(il), stream))
                if ic.overflow or il.overflow:
                    cap = max(cap, int(ic.R) + _SLACK)
                    cap_l = max(cap_l, int(il.R) + _SLACK)
                    continue
                R, R_l = int(ic.R), int(il.R)
                _R_HINT[key] = (max(R, 1), max(R_l, 1))
            break
        else:
            raise N.OlsError(N.OLS_ERR_OVERFLOW, "instance capacity overflow after 3 attempts")
    keep["workspace"] = ws
    st = _Ctx()
    st.args, st.keep, st.R, st.R_lang = args, keep, R, R_l
    return R, R_l, color, language, radii, radii_lang, depth, opacity, opacity_lang, n_touched, n_touched_lang, st


GRAD_SHAPES = lambda P, F, M: {
    "means2D": (P, 3), "colors": (P, 3), "language": (P, F), "opacity": (P, 1), "opacity_lang": (P, 1), "means3D": (P, 3),
    "cov3D": (P, 6), "cov3D_lang": (P, 6), "sh": (P, M, 3), "scales": (P, 3), "scales_lang": (P, 3), "rotations": (P, 4),
    "rotations_lang": (P, 4), "tau": (P, 6)}


def _backward_native(st: _Ctx, radii, radii_lang, grad_color, grad_language, grad_depth):
    k, d = st.keep, st.args
    a = d.base
    dev = k["means3D"].device
    g = {n: torch.empty(s, dtype=torch.float32, device=dev) for n, s in GRAD_SHAPES(a.P, a.F, a.M).items()}
    gc, gl, gd = _f32c(grad_color), _f32c(grad_language), _f32c(grad_depth)
    b = N.DisBwdArgs(
        d_dL_dout_color=gc.data_ptr(), d_dL_dout_language=gl.data_ptr(), d_dL_dout_depth=gd.data_ptr(),
        d_radii=radii.data_ptr(), d_radii_lang=radii_lang.data_ptr(), d_dL_dmeans2D=g["means2D"].data_ptr(),
        d_dL_dcolors=g["colors"].data_ptr(), d_dL_dlanguage=g["language"].data_ptr(),
        d_dL_dopacity=g["opacity"].data_ptr(), d_dL_dopacity_lang=g["opacity_lang"].data_ptr(),
        d_dL_dmeans3D=g["means3D"].data_ptr(), d_dL_dcov3D=g["cov3D"].data_ptr(),
        d_dL_dcov3D_lang=g["cov3D_lang"].data_ptr(), d_dL_dsh=N.ptr(g["sh"]), d_dL_dscales=g["scales"].data_ptr(),
        d_dL_dscales_lang=g["scales_lang"].data_ptr(), d_dL_drotations=g["rotations"].data_ptr(),
        d_dL_drotations_lang=g["rotations_lang"].data_ptr(), d_dL_dtau=g["tau"].data_ptr())
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().ols_dis_backward(C.byref(d), C.byref(b), stream))
    return g


class _RasterizeLanguageGaussians(torch.autograd.Function):
    """Reference: _RasterizeLanguageGaussians of D/ (:213-465)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, language_precomp, opacities, opacities_lang, scales,
                scales_lang, rotations, rotations_lang, cov3Ds_precomp, cov3Ds_precomp_lang, theta, rho, raster_settings):
        (R, R_l, color, language, radii, radii_lang, depth, opacity, opacity_lang, n_touched, n_touched_lang,
         st) = _forward_native(means3D, sh, colors_precomp, language_precomp, opacities, opacities_lang, scales,
                               scales_lang, rotations, rotations_lang, cov3Ds_precomp, cov3Ds_precomp_lang,
                               raster_settings)
        ctx.raster_settings = raster_settings
        ctx.num_rendered, ctx.num_rendered_lang = R, R_l
        ctx.state = st
        ctx.save_for_backward(radii, radii_lang)
        ctx.mark_non_differentiable(radii, radii_lang, n_touched, n_touched_lang)
        return color, language, radii, radii_lang, depth, opacity, opacity_lang, n_touched, n_touched_lang

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_language, grad_out_radii, grad_out_radii_lang, grad_out_depth,
                 grad_out_opacity, grad_out_opacity_lang, grad_n_touched, grad_n_touched_lang):
        # the two opacity-map gradients are ignored exactly like the reference (D/ :351-384: not passed to C++)
        radii, radii_lang = ctx.saved_tensors
        st = ctx.state
        if st is None:
            raise RuntimeError("backward called on an empty render")
        a = st.args.base
        dev = radii.device
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, a.H, a.W), device=dev)
        if grad_out_language is None:
            grad_out_language = torch.zeros((a.F, a.H, a.W), device=dev)
        if grad_out_depth is None:
            grad_out_depth = torch.zeros((1, a.H, a.W), device=dev)
        g = _backward_native(st, radii, radii_lang, grad_out_color, grad_out_language, grad_out_depth)
        grad_tau = torch.sum(g["tau"].view(-1, 6), dim=0)  # D/ :440-442
        grad_rho = grad_tau[:3].view(1, -1)
        grad_theta = grad_tau[3:].view(1, -1)
        k = st.keep
        opt = lambda name, key: g[name] if k[key] is not None else None
        return (
            g["means3D"], g["means2D"], opt("sh", "shs"), opt("colors", "colors_precomp"), g["language"],
            g["opacity"], g["opacity_lang"], opt("scales", "scales"), opt("scales_lang", "scales_lang"),
            opt("rotations", "rotations"), opt("rotations_lang", "rotations_lang"), opt("cov3D", "cov3D_precomp"),
            opt("cov3D_lang", "cov3D_precomp_lang"), grad_theta, grad_rho, None,
        )


def rasterize_language_gaussians(means3D, means2D, sh, colors_precomp, language_precomp, opacities, opacities_lang,
                                 scales, scales_lang, rotations, rotations_lang, cov3Ds_precomp, cov3Ds_precomp_lang,
                                 theta, rho, raster_settings):
    return _RasterizeLanguageGaussians.apply(means3D, means2D, sh, colors_precomp, language_precomp, opacities,
                                             opacities_lang, scales, scales_lang, rotations, rotations_lang,
                                             cov3Ds_precomp, cov3Ds_precomp_lang, theta, rho, raster_settings)


class LanguageGaussianRasterizer(_RasterizerBase):
    """Reference: LanguageGaussianRasterizer of D/ (:543-666).  ``markVisible`` is inherited."""

    def forward(self, means3D, means2D, opacities, opacities_lang, shs=None, colors_precomp=None, language_precomp=None,
                scales=None, scales_lang=None, rotations=None, rotations_lang=None, cov3D_precomp=None,
                cov3D_precomp_lang=None, theta=None, rho=None):
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if ((scales_lang is None or rotations_lang is None) and cov3D_precomp_lang is None) or (
                (scales_lang is not None or rotations_lang is not None) and cov3D_precomp_lang is not None):
            raise Exception(
                "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance for language!")
        e = lambda t: _empty() if t is None else t
        return rasterize_language_gaussians(
            means3D, means2D, e(shs), e(colors_precomp), e(language_precomp), opacities, opacities_lang, e(scales),
            e(scales_lang), e(rotations), e(rotations_lang), e(cov3D_precomp), e(cov3D_precomp_lang), e(theta), e(rho),
            self.raster_settings)
