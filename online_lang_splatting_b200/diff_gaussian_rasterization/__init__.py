"""Drop-in for the reference's ``diff_gaussian_rasterization`` package (the "P/" variant that
``gaussian_renderer.render`` calls: submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py).

Same public names, argument names, return arity and error behaviour:

* ``GaussianRasterizationSettings``   (reference :405-419; three optional trailing fields added)
* ``LanguageGaussianRasterizer``      (reference :482-576)  -> 6 returns
* ``GaussianRasterizer``              (reference :421-480)  -> 5 returns
* ``rasterize_language_gaussians`` / ``rasterize_gaussians``

Underneath, everything goes through the C ABI of ``include/ols_b200.h`` (hand-written sm_100a
kernels); torch only owns the memory and the stream.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, NamedTuple, Optional, Tuple

import torch
import torch.nn as nn

from .. import _native as N


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
    # --- extensions (defaults reproduce the reference build: config.h BLOCK_X = BLOCK_Y = 15) ---
    tile_size: int = 15          # 15 = reference P/ geometry, 16 = D/ geometry / performance mode
    backward_mode: str = "compat"  # "compat": reference gradients incl. quirks Q1-Q3; "exact": true gradients
    bitexact_blend: bool = False   # accumulate c*alpha*T in the reference's exact operation order


# Capacity policy for Gaussian/tile instances.  The reference reads R back from the device in the
# middle of every forward (rasterizer_impl.cu:455); we size the workspace from a running estimate
# and only read the 32-byte info header after everything has been queued.
_R_HINT: Dict[Tuple, int] = {}
CHECK_OVERFLOW = True  # set False for fully asynchronous forwards (caller guarantees capacity)
_SLACK = 65536


def _capacity(key, P: int) -> int:
    hint = _R_HINT.get(key)
    if hint is None:
        return 8 * P + _SLACK
    return int(hint * 1.25) + _SLACK


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _Ctx:
    """What forward leaves for backward (the reference keeps geomBuffer/binningBuffer/imgBuffer)."""
    __slots__ = ("args", "keep", "R", "info")


def _flags(rs: GaussianRasterizationSettings) -> int:
    f = 0
    if rs.prefiltered:
        f |= N.FLAG_PREFILTERED
    if rs.debug:
        f |= N.FLAG_DEBUG
    if getattr(rs, "bitexact_blend", False):
        f |= N.FLAG_BITEXACT_BLEND
    mode = getattr(rs, "backward_mode", "compat")
    if mode == "exact":
        f |= N.FLAG_BWD_EXACT
    elif mode != "compat":
        raise Exception(f"backward_mode must be 'compat' or 'exact', got {mode!r}")
    return f


def _forward_native(means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    rs: GaussianRasterizationSettings):
    N.require_cuda()
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:159-161
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor: the rasterizer has no CPU path")
    dev = means3D.device
    P = means3D.shape[0]
    H, W = int(rs.image_height), int(rs.image_width)
    tile = int(getattr(rs, "tile_size", 15))
    if language_precomp is None or language_precomp.dim() != 2 or (P > 0 and language_precomp.numel() == 0):
        raise RuntimeError("language_precomp is required by the language rasterizer")
    F = int(language_precomp.shape[1])
    if P == 0:  # nothing to rasterize (render() returns None before getting here, reference :76,:210)
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        e = torch.empty((0,), dtype=torch.int32, device=dev)
        return 0, z(3, H, W), z(F, H, W), e, z(1, H, W), z(1, H, W), e.clone(), None
    keep = {
        "means3D": _f32c(means3D), "language": _f32c(language_precomp), "opacities": _f32c(opacities),
        "bg": _f32c(rs.bg).to(dev), "viewmatrix": _f32c(rs.viewmatrix).to(dev),
        "projmatrix": _f32c(rs.projmatrix).to(dev), "projmatrix_raw": _f32c(rs.projmatrix_raw).to(dev),
        "campos": _f32c(rs.campos).to(dev),
    }
    for name, t in (("shs", sh), ("colors_precomp", colors_precomp), ("scales", scales), ("rotations", rotations),
                    ("cov3D_precomp", cov3Ds_precomp)):
        keep[name] = None if (t is None or t.numel() == 0) else _f32c(t).to(dev)
    if keep["shs"] is None and keep["colors_precomp"] is None:
        # rasterizer_impl.cu:414-417
        raise RuntimeError("For non-RGB, provide precomputed Gaussian colors!")
    M = 0 if keep["shs"] is None else int(keep["shs"].shape[1])

    color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
    language = torch.empty((F, H, W), dtype=torch.float32, device=dev)
    depth = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    opacity = torch.empty((1, H, W), dtype=torch.float32, device=dev)
    radii = torch.empty((P,), dtype=torch.int32, device=dev)
    n_touched = torch.empty((P,), dtype=torch.int32, device=dev)
    lib = N.lib()
    key = (dev.index, P, W, H, tile)
    stream = torch.cuda.current_stream(dev).cuda_stream
    cap = _capacity(key, P)
    with torch.cuda.device(dev):
        for attempt in range(3):
            nbytes = lib.ols_lang_workspace_size(P, F, W, H, tile, cap)
            if nbytes == 0:
                raise RuntimeError("invalid rasterizer configuration")
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            args = N.RasterArgs(
                P=P, F=F, sh_degree=int(rs.sh_degree), M=M, W=W, H=H, tile=tile, flags=_flags(rs),
                tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy), scale_modifier=float(rs.scale_modifier),
                d_bg=N.ptr(keep["bg"]), d_means3D=N.ptr(keep["means3D"]), d_shs=N.ptr(keep["shs"]),
                d_colors_precomp=N.ptr(keep["colors_precomp"]), d_language=N.ptr(keep["language"]),
                d_opacities=N.ptr(keep["opacities"]), d_scales=N.ptr(keep["scales"]),
                d_rotations=N.ptr(keep["rotations"]), d_cov3D_precomp=N.ptr(keep["cov3D_precomp"]),
                d_viewmatrix=N.ptr(keep["viewmatrix"]), d_projmatrix=N.ptr(keep["projmatrix"]),
                d_projmatrix_raw=N.ptr(keep["projmatrix_raw"]), d_campos=N.ptr(keep["campos"]),
                d_workspace=ws.data_ptr(), workspace_bytes=nbytes, R_cap=cap)
            out = N.FwdOut(d_color=color.data_ptr(), d_language=language.data_ptr(), d_depth=depth.data_ptr(),
                           d_opacity=opacity.data_ptr(), d_radii=radii.data_ptr(), d_n_touched=n_touched.data_ptr())
            N.check(lib.ols_lang_forward(C.byref(args), C.byref(out), stream))
            info = None
            if CHECK_OVERFLOW:
                info = N.FwdInfo()
                N.check(lib.ols_lang_read_info(ws.data_ptr(), C.byref(info), stream))
                if info.overflow:
                    cap = int(info.R) + _SLACK
                    continue
                _R_HINT[key] = max(int(info.R), 1)
            break
        else:
            raise N.OlsError(N.OLS_ERR_OVERFLOW, "instance capacity overflow after 3 attempts")
    keep["workspace"] = ws
    st = _Ctx()
    st.args, st.keep, st.info = args, keep, info
    st.R = int(info.R) if info is not None else -1
    return st.R, color, language, radii, depth, opacity, n_touched, st


GRAD_SHAPES = lambda P, F, M: {"means2D": (P, 3), "colors": (P, 3), "language": (P, F), "opacity": (P, 1),
                                "means3D": (P, 3), "cov3D": (P, 6), "sh": (P, M, 3), "scales": (P, 3),
                                "rotations": (P, 4), "tau": (P, 6)}


def _backward_native(st: _Ctx, radii, grad_color, grad_language, grad_depth, out=None, accumulate=False):
    """Calls ols_lang_backward.  ``out`` may hold preallocated (contiguous fp32) gradient tensors keyed like
    GRAD_SHAPES -- e.g. views into one flat buffer that is all-reduced across ranks; with
    ``accumulate=True`` the parameter gradients are added to what ``out`` already holds."""
    k = st.keep
    a = st.args
    dev = k["means3D"].device
    P, F, M = a.P, a.F, a.M
    g = {}
    for name, shape in GRAD_SHAPES(P, F, M).items():
        t = None if out is None else out.get(name)
        if t is None:
            t = (torch.zeros if accumulate else torch.empty)(shape, dtype=torch.float32, device=dev)
        g[name] = t
    gc, gl, gd = _f32c(grad_color), _f32c(grad_language), _f32c(grad_depth)
    b = N.BwdArgs(d_dL_dout_color=gc.data_ptr(), d_dL_dout_language=gl.data_ptr(), d_dL_dout_depth=gd.data_ptr(),
                  d_radii=radii.data_ptr(), d_dL_dmeans2D=g["means2D"].data_ptr(), d_dL_dcolors=g["colors"].data_ptr(),
                  d_dL_dlanguage=g["language"].data_ptr(), d_dL_dopacity=g["opacity"].data_ptr(),
                  d_dL_dmeans3D=g["means3D"].data_ptr(), d_dL_dcov3D=g["cov3D"].data_ptr(),
                  d_dL_dsh=N.ptr(g["sh"]), d_dL_dscales=g["scales"].data_ptr(),
                  d_dL_drotations=g["rotations"].data_ptr(), d_dL_dtau=g["tau"].data_ptr())
    flags = a.flags
    if accumulate:
        a.flags = flags | N.FLAG_BWD_ACCUMULATE
    try:
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_lang_backward(C.byref(a), C.byref(b), stream))
    finally:
        a.flags = flags
    return g


class _RasterizeLanguageGaussians(torch.autograd.Function):
    """Reference: _RasterizeLanguageGaussians (diff_gaussian_rasterization/__init__.py:205-403)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, language_precomp, opacities, scales, rotations,
                cov3Ds_precomp, theta, rho, raster_settings):
        R, color, language, radii, depth, opacity, n_touched, st = _forward_native(
            means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations, cov3Ds_precomp,
            raster_settings)
        ctx.raster_settings = raster_settings
        ctx.num_rendered = R
        ctx.state = st
        ctx.save_for_backward(radii)
        ctx.mark_non_differentiable(radii, n_touched)
        return color, language, radii, depth, opacity, n_touched

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_language, grad_out_radii, grad_out_depth, grad_out_opacity,
                 grad_n_touched):
        # grad_out_opacity is ignored exactly like the reference (:296, not passed to C++)
        (radii,) = ctx.saved_tensors
        st = ctx.state
        if st is None:
            raise RuntimeError("backward called on an empty render")
        dev = radii.device
        P, F = st.args.P, st.args.F
        H, W = st.args.H, st.args.W
        if grad_out_color is None:
            grad_out_color = torch.zeros((3, H, W), device=dev)
        if grad_out_language is None:
            grad_out_language = torch.zeros((F, H, W), device=dev)
        if grad_out_depth is None:
            grad_out_depth = torch.zeros((1, H, W), device=dev)
        g = _backward_native(st, radii, grad_out_color, grad_out_language, grad_out_depth)
        grad_tau = torch.sum(g["tau"].view(-1, 6), dim=0)  # reference :383-385
        grad_rho = grad_tau[:3].view(1, -1)
        grad_theta = grad_tau[3:].view(1, -1)
        k = st.keep
        return (
            g["means3D"],
            g["means2D"],
            g["sh"] if k["shs"] is not None else None,
            g["colors"] if k["colors_precomp"] is not None else None,
            g["language"],
            g["opacity"],
            g["scales"] if k["scales"] is not None else None,
            g["rotations"] if k["rotations"] is not None else None,
            g["cov3D"] if k["cov3D_precomp"] is not None else None,
            grad_theta,
            grad_rho,
            None,
        )


def rasterize_language_gaussians(means3D, means2D, sh, colors_precomp, language_precomp, opacities, scales, rotations,
                                 cov3Ds_precomp, theta, rho, raster_settings):
    return _RasterizeLanguageGaussians.apply(means3D, means2D, sh, colors_precomp, language_precomp, opacities,
                                             scales, rotations, cov3Ds_precomp, theta, rho, raster_settings)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
                        raster_settings):
    """Non-language rasterizer (reference _RasterizeGaussians, :79-203): same kernels with a zero
    3-channel language input, returning the reference's 5-tuple."""
    lang = torch.zeros((means3D.shape[0], 3), dtype=torch.float32, device=means3D.device)
    color, _, radii, depth, opacity, n_touched = rasterize_language_gaussians(
        means3D, means2D, sh, colors_precomp, lang, opacities, scales, rotations, cov3Ds_precomp, theta, rho,
        raster_settings)
    return color, radii, depth, opacity, n_touched


def _empty():
    return torch.Tensor([])


def _validate(shs, colors_precomp, scales, rotations, cov3D_precomp):
    if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or (
            (scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")


class _RasterizerBase(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """Reference: markVisible (:487-496) -> checkFrustum (rasterizer_impl.cu:54-66)."""
        with torch.no_grad():
            N.require_cuda()
            rs = self.raster_settings
            pos = _f32c(positions)
            present = torch.empty((pos.shape[0],), dtype=torch.bool, device=pos.device)
            vm, pm = _f32c(rs.viewmatrix).to(pos.device), _f32c(rs.projmatrix).to(pos.device)
            with torch.cuda.device(pos.device):
                stream = torch.cuda.current_stream(pos.device).cuda_stream
                N.check(N.lib().ols_mark_visible(pos.shape[0], N.ptr(pos), vm.data_ptr(), pm.data_ptr(),
                                                 N.ptr(present), stream))
        return present


class GaussianRasterizer(_RasterizerBase):
    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None):
        _validate(shs, colors_precomp, scales, rotations, cov3D_precomp)
        shs = _empty() if shs is None else shs
        colors_precomp = _empty() if colors_precomp is None else colors_precomp
        scales = _empty() if scales is None else scales
        rotations = _empty() if rotations is None else rotations
        cov3D_precomp = _empty() if cov3D_precomp is None else cov3D_precomp
        theta = _empty() if theta is None else theta
        rho = _empty() if rho is None else rho
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                   theta, rho, self.raster_settings)


class LanguageGaussianRasterizer(_RasterizerBase):
    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, language_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, theta=None, rho=None):
        _validate(shs, colors_precomp, scales, rotations, cov3D_precomp)
        shs = _empty() if shs is None else shs
        colors_precomp = _empty() if colors_precomp is None else colors_precomp
        language_precomp = _empty() if language_precomp is None else language_precomp
        scales = _empty() if scales is None else scales
        rotations = _empty() if rotations is None else rotations
        cov3D_precomp = _empty() if cov3D_precomp is None else cov3D_precomp
        theta = _empty() if theta is None else theta
        rho = _empty() if rho is None else rho
        colors, language, radii, depth, opacity, n_touched = rasterize_language_gaussians(
            means3D, means2D, shs, colors_precomp, language_precomp, opacities, scales, rotations, cov3D_precomp,
            theta, rho, self.raster_settings)
        return colors, language, radii, depth, opacity, n_touched
