"""Drop-in for the reference's ``diff_gaussian_rasterization`` package (the "P/" variant that
``gaussian_renderer.render`` calls: submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py).

Same public names, argument names, return arity and error behaviour:

* ``GaussianRasterizationSettings``   (reference :405-419; three optional trailing fields added)
* ``LanguageGaussianRasterizer``      (reference :482-576)  -> 6 returns
* ``GaussianRasterizer``              (reference :421-480)  -> 5 returns
* ``rasterize_language_gaussians`` / ``rasterize_gaussians``
* ``rasterize_language_gaussians_batch`` (extension): V views of the same Gaussians in one set of launches

Underneath, everything goes through the C ABI of ``include/ols_b200.h`` (hand-written sm_100a
kernels); torch only owns the memory and the stream.  There is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

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


# ---------------------------------------------------------------------------------------------------
# Capacity policy for Gaussian/tile instances, and how the overflow flag reaches the host.
#
# The reference reads R back from the device in the middle of every forward (rasterizer_impl.cu:455) and
# sizes binningBuffer exactly.  Here the workspace is sized from a running estimate of R / P per
# (device, image size, tile) and the kernels raise a device-side flag when it does not fit.  How that flag
# is read is selected by ``CHECK_OVERFLOW``:
#
#   "deferred" (default)  The first forward of a configuration (no estimate yet) is checked synchronously and
#              retried with the exact size.  Every later forward is fully asynchronous: the 32-byte info header
#              is copied into pinned host memory behind the kernels, followed by an event; the header is looked
#              at -- without blocking -- by the next forward, and by this render's own backward (by then the
#              forward has long finished).  A forward that did overflow hands back NaN images (never a
#              plausible-looking empty frame), raises the estimate so the next render fits, and its backward raises.
#   True / "sync"         every forward waits for its header and retries on overflow (the reference's behaviour).
#   False                 never looked at (the caller guarantees capacity, e.g. inside a captured CUDA graph).
# ---------------------------------------------------------------------------------------------------
CHECK_OVERFLOW = "deferred"
_R_RATIO: Dict[Tuple, float] = {}   # (device index, W, H, tile) -> largest R / P seen
_SLACK = 65536
_GROWTH = 1.25
_PENDING: List["_Pending"] = []
_PIN_SLOTS = 64
_pin_pool: Dict[str, object] = {}


def _capacity(key, P: int) -> int:
    ratio = _R_RATIO.get(key)
    if ratio is None:
        return 8 * P + _SLACK
    return int(ratio * P * _GROWTH) + _SLACK


def _note_R(key, P: int, R: int) -> None:
    _R_RATIO[key] = max(_R_RATIO.get(key, 0.0), max(int(R), 1) / max(P, 1))


class _Pending:
    """Info headers of one asynchronous forward on their way to pinned host memory."""
    __slots__ = ("key", "P", "V", "slot", "event", "done", "overflow", "Rs")

    def poll(self, block: bool = False) -> bool:
        if self.done:
            return True
        if block:
            self.event.synchronize()
        elif not self.event.query():
            return False
        buf = _pin_pool["buf"]
        self.Rs, self.overflow = [], False
        for v in range(self.V):
            row = buf[self.slot, v]
            self.Rs.append((int(row[0].item()) & 0xffffffff) | ((int(row[1].item()) & 0xffffffff) << 32))
            self.overflow = self.overflow or bool(row[2].item())
        _note_R(self.key, self.P, max(self.Rs))
        _pin_pool["free"].append(self.slot)
        self.done = True
        return True


def _pinned_slot():
    if "buf" not in _pin_pool:
        _pin_pool["buf"] = torch.zeros((_PIN_SLOTS, N.MAX_BATCH_VIEWS, 8), dtype=torch.int32).pin_memory()
        _pin_pool["free"] = list(range(_PIN_SLOTS))
    if not _pin_pool["free"]:   # every slot is in flight: retire the oldest (blocks until its forward is done)
        _PENDING.pop(0).poll(block=True)
    return _pin_pool["free"].pop()


def _poll_pending() -> None:
    """Non-blocking look at the headers of earlier asynchronous forwards (updates the capacity estimate)."""
    while _PENDING and _PENDING[0].poll():
        _PENDING.pop(0)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _Ctx:
    """What forward leaves for backward (the reference keeps geomBuffer/binningBuffer/imgBuffer)."""
    __slots__ = ("args", "args_arr", "keep", "_R", "Rs", "info", "V", "pending", "workspaces")

    @property
    def R(self) -> int:
        """num_rendered of the first view (reference: the first return of rasterize_language_gaussians).  In the deferred
        mode it is not known when forward returns; reading it waits for the header (tests / debugging only)."""
        if self._R < 0 and self.pending is not None:
            self.pending.poll(block=True)
            self.Rs = list(self.pending.Rs)
            self._R = self.Rs[0]
        return self._R

    def resolve(self) -> None:
        """Called by backward: the forward's overflow flag is known by now (waits for it if it is not)."""
        pend = self.pending
        if pend is None:
            return
        pend.poll(block=True)
        if pend in _PENDING:
            _PENDING.remove(pend)
        self.Rs = list(pend.Rs)
        self._R = self.Rs[0]
        self.pending = None
        if pend.overflow:
            raise N.OlsError(N.OLS_ERR_OVERFLOW,
                             "the forward of this render exceeded its instance capacity (its images were NaN); the "
                             "capacity estimate has been raised -- render again")


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


def _forward_native_batch(means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations, cov3Ds_precomp,
                          rs_list: Sequence[GaussianRasterizationSettings]):
    """V views of the same Gaussians through ols_lang_forward_batch (one set of launches, grid.y = view).
    Returns ([(color, language, radii, depth, opacity, n_touched)] * V, state)."""
    N.require_cuda()
    V = len(rs_list)
    if not (1 <= V <= N.MAX_BATCH_VIEWS):
        raise RuntimeError(f"a batch holds 1..{N.MAX_BATCH_VIEWS} views, got {V}")
    if means3D.dim() != 2 or means3D.shape[1] != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # rasterize_points.cu:159-161
    if not means3D.is_cuda:
        raise RuntimeError("means3D must be a CUDA tensor: the rasterizer has no CPU path")
    dev = means3D.device
    P = means3D.shape[0]
    rs0 = rs_list[0]
    H, W = int(rs0.image_height), int(rs0.image_width)
    tile = int(getattr(rs0, "tile_size", 15))
    flags = _flags(rs0)
    for rs in rs_list[1:]:
        if (int(rs.image_height), int(rs.image_width), int(getattr(rs, "tile_size", 15)), _flags(rs), int(rs.sh_degree),
                float(rs.scale_modifier)) != (H, W, tile, flags, int(rs0.sh_degree), float(rs0.scale_modifier)):
            raise RuntimeError("the views of a batch must share image size, tile size, SH degree, scale modifier and flags")
    if language_precomp is None or language_precomp.dim() != 2 or (P > 0 and language_precomp.numel() == 0):
        raise RuntimeError("language_precomp is required by the language rasterizer")
    F = int(language_precomp.shape[1])
    if P == 0:  # nothing to rasterize (render() returns None before getting here, reference :76,:210)
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        e = torch.empty((0,), dtype=torch.int32, device=dev)
        st = None
        return [(z(3, H, W), z(F, H, W), e.clone(), z(1, H, W), z(1, H, W), e.clone()) for _ in range(V)], st
    keep = {"means3D": _f32c(means3D), "language": _f32c(language_precomp), "opacities": _f32c(opacities)}
    for name, t in (("shs", sh), ("colors_precomp", colors_precomp), ("scales", scales), ("rotations", rotations),
                    ("cov3D_precomp", cov3Ds_precomp)):
        keep[name] = None if (t is None or t.numel() == 0) else _f32c(t).to(dev)
    if keep["shs"] is None and keep["colors_precomp"] is None:
        # rasterizer_impl.cu:414-417
        raise RuntimeError("For non-RGB, provide precomputed Gaussian colors!")
    M = 0 if keep["shs"] is None else int(keep["shs"].shape[1])
    cams = []
    for rs in rs_list:
        cams.append({"bg": _f32c(rs.bg).to(dev), "viewmatrix": _f32c(rs.viewmatrix).to(dev),
                     "projmatrix": _f32c(rs.projmatrix).to(dev), "projmatrix_raw": _f32c(rs.projmatrix_raw).to(dev),
                     "campos": _f32c(rs.campos).to(dev)})
    keep["cams"] = cams
    keep.update(cams[0])

    outs = []
    for _ in range(V):
        outs.append((torch.empty((3, H, W), dtype=torch.float32, device=dev),
                     torch.empty((F, H, W), dtype=torch.float32, device=dev),
                     torch.empty((P,), dtype=torch.int32, device=dev),
                     torch.empty((1, H, W), dtype=torch.float32, device=dev),
                     torch.empty((1, H, W), dtype=torch.float32, device=dev),
                     torch.empty((P,), dtype=torch.int32, device=dev)))
    lib = N.lib()
    key = (dev.index, W, H, tile)
    capturing = torch.cuda.is_current_stream_capturing()
    mode = CHECK_OVERFLOW
    if capturing:
        mode = False
    elif mode == "deferred":
        _poll_pending()
        if key not in _R_RATIO:
            mode = "sync"       # no estimate for this configuration yet: check this one on the spot
    stream = torch.cuda.current_stream(dev).cuda_stream
    cap = _capacity(key, P)
    pend = None
    infos = None
    with torch.cuda.device(dev):
        for attempt in range(3):
            nbytes = lib.ols_lang_workspace_size(P, F, W, H, tile, cap)
            if nbytes == 0:
                raise RuntimeError("invalid rasterizer configuration")
            stride = (nbytes + 255) // 256 * 256
            ws_all = torch.empty((stride * V,), dtype=torch.uint8, device=dev)
            workspaces = [ws_all[v * stride:v * stride + nbytes] for v in range(V)]
            args = (N.RasterArgs * V)()
            fo = (N.FwdOut * V)()
            for v, rs in enumerate(rs_list):
                c = cams[v]
                args[v] = N.RasterArgs(
                    P=P, F=F, sh_degree=int(rs.sh_degree), M=M, W=W, H=H, tile=tile, flags=flags,
                    tanfovx=float(rs.tanfovx), tanfovy=float(rs.tanfovy), scale_modifier=float(rs.scale_modifier),
                    d_bg=N.ptr(c["bg"]), d_means3D=N.ptr(keep["means3D"]), d_shs=N.ptr(keep["shs"]),
                    d_colors_precomp=N.ptr(keep["colors_precomp"]), d_language=N.ptr(keep["language"]),
                    d_opacities=N.ptr(keep["opacities"]), d_scales=N.ptr(keep["scales"]),
                    d_rotations=N.ptr(keep["rotations"]), d_cov3D_precomp=N.ptr(keep["cov3D_precomp"]),
                    d_viewmatrix=N.ptr(c["viewmatrix"]), d_projmatrix=N.ptr(c["projmatrix"]),
                    d_projmatrix_raw=N.ptr(c["projmatrix_raw"]), d_campos=N.ptr(c["campos"]),
                    d_workspace=workspaces[v].data_ptr(), workspace_bytes=nbytes, R_cap=cap)
                o = outs[v]
                fo[v] = N.FwdOut(d_color=o[0].data_ptr(), d_language=o[1].data_ptr(), d_depth=o[3].data_ptr(),
                                 d_opacity=o[4].data_ptr(), d_radii=o[2].data_ptr(), d_n_touched=o[5].data_ptr())
            N.check(lib.ols_lang_forward_batch(args, fo, V, stream))
            if mode is True or mode == "sync":
                infos = []
                over = False
                for v in range(V):
                    info = N.FwdInfo()
                    N.check(lib.ols_lang_read_info(workspaces[v].data_ptr(), C.byref(info), stream))
                    infos.append(info)
                    over = over or bool(info.overflow)
                Rmax = max(int(i.R) for i in infos)
                if over:
                    cap = Rmax + _SLACK
                    continue
                _note_R(key, P, Rmax)
            elif mode == "deferred":
                pend = _Pending()
                pend.key, pend.P, pend.V, pend.done, pend.overflow, pend.Rs = key, P, V, False, False, []
                pend.slot = _pinned_slot()
                N.check(lib.ols_lang_read_info_async(args, V, _pin_pool["buf"][pend.slot].data_ptr(), stream))
                pend.event = torch.cuda.Event()
                pend.event.record(torch.cuda.current_stream(dev))
                _PENDING.append(pend)
            break
        else:
            raise N.OlsError(N.OLS_ERR_OVERFLOW, "instance capacity overflow after 3 attempts")
    keep["workspace"] = workspaces[0]
    keep["workspace_all"] = ws_all
    st = _Ctx()
    st.args_arr, st.args, st.keep, st.V, st.pending, st.workspaces = args, args[0], keep, V, pend, workspaces
    st.info = infos[0] if infos else None
    st.Rs = [int(i.R) for i in infos] if infos else [-1] * V
    st._R = st.Rs[0]
    return outs, st


def _forward_native(means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations, cov3Ds_precomp,
                    rs: GaussianRasterizationSettings):
    """One view (the reference's call): R, color, language, radii, depth, opacity, n_touched, state."""
    outs, st = _forward_native_batch(means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, [rs])
    color, language, radii, depth, opacity, n_touched = outs[0]
    return (st._R if st is not None else 0), color, language, radii, depth, opacity, n_touched, st


GRAD_SHAPES = lambda P, F, M: {"colors": (P, 3), "language": (P, F), "opacity": (P, 1),
                                "means3D": (P, 3), "cov3D": (P, 6), "sh": (P, M, 3), "scales": (P, 3),
                                "rotations": (P, 4)}


def _backward_native_batch(st: _Ctx, radii, grad_color, grad_language, grad_depth, out=None, accumulate=False,
                           per_gaussian_tau=False):
    """Calls ols_lang_backward_batch.  ``radii`` / ``grad_*`` are per-view lists.  Returns a dict with the
    parameter gradients summed over the views (keys of GRAD_SHAPES), "means2D" [V,P,3] (each view's screen-space
    gradient), "tau_sum" [V,6] (each view's pose gradient summed over the Gaussians) and, on request, "tau"
    [V,P,6] (the reference's per-Gaussian form).  ``out`` may hold preallocated contiguous fp32 tensors under
    the same keys -- e.g. views into one flat buffer that is all-reduced across ranks; with ``accumulate=True``
    the parameter gradients are added to what ``out`` already holds.  ``out["stats"] = (max_radii2D, xyz_gradient_accum,
    denom)`` (fp32 [P] each) folds the densification statistics of all the views into the same launch."""
    st.resolve()
    k = st.keep
    a = st.args
    V = st.V
    dev = k["means3D"].device
    P, F, M = a.P, a.F, a.M
    g = {}
    for name, shape in GRAD_SHAPES(P, F, M).items():
        t = None if out is None else out.get(name)
        if t is None:
            t = (torch.zeros if accumulate else torch.empty)(shape, dtype=torch.float32, device=dev)
        g[name] = t
    g["means2D"] = (out or {}).get("means2D")
    if g["means2D"] is None or g["means2D"].numel() != V * P * 3:
        g["means2D"] = torch.empty((V, P, 3), dtype=torch.float32, device=dev)
    g["tau_sum"] = (out or {}).get("tau_sum")
    if g["tau_sum"] is None or g["tau_sum"].numel() != V * 6:
        g["tau_sum"] = torch.empty((V, 6), dtype=torch.float32, device=dev)
    if per_gaussian_tau:
        g["tau"] = (out or {}).get("tau")
        if g["tau"] is None or g["tau"].numel() != V * P * 6:
            g["tau"] = torch.empty((V, P, 6), dtype=torch.float32, device=dev)
    m2, ts = g["means2D"].view(V, P, 3), g["tau_sum"].view(V, 6)
    tg = g["tau"].view(V, P, 6) if per_gaussian_tau else None
    b = (N.BwdArgs * V)()
    live = []
    for v in range(V):
        gc, gl, gd = _f32c(grad_color[v]), _f32c(grad_language[v]), _f32c(grad_depth[v])
        live.append((gc, gl, gd))
        b[v] = N.BwdArgs(d_dL_dout_color=gc.data_ptr(), d_dL_dout_language=gl.data_ptr(), d_dL_dout_depth=gd.data_ptr(),
                         d_radii=radii[v].data_ptr(), d_dL_dmeans2D=m2[v].data_ptr(), d_dL_dcolors=g["colors"].data_ptr(),
                         d_dL_dlanguage=g["language"].data_ptr(), d_dL_dopacity=g["opacity"].data_ptr(),
                         d_dL_dmeans3D=g["means3D"].data_ptr(), d_dL_dcov3D=g["cov3D"].data_ptr(),
                         d_dL_dsh=N.ptr(g["sh"]), d_dL_dscales=g["scales"].data_ptr(),
                         d_dL_drotations=g["rotations"].data_ptr(),
                         d_dL_dtau=tg[v].data_ptr() if tg is not None else None, d_dL_dtau_sum=ts[v].data_ptr())
    stats = (out or {}).get("stats")
    if stats is not None:
        b[0].d_stat_max_radii2D, b[0].d_stat_xyz_gradient_accum, b[0].d_stat_denom = (t.data_ptr() for t in stats)
    arr = st.args_arr
    flags = arr[0].flags
    try:
        if accumulate:
            for v in range(V):
                arr[v].flags = flags | N.FLAG_BWD_ACCUMULATE
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            N.check(N.lib().ols_lang_backward_batch(arr, b, V, stream))
    finally:
        for v in range(V):
            arr[v].flags = flags
    return g


def _backward_native(st: _Ctx, radii, grad_color, grad_language, grad_depth, out=None, accumulate=False):
    """One view: the dict of _backward_native_batch with "means2D" [P,3], "tau_sum" [6] and the reference's
    per-Gaussian "tau" [P,6]."""
    g = _backward_native_batch(st, [radii], [grad_color], [grad_language], [grad_depth], out=out, accumulate=accumulate,
                               per_gaussian_tau=True)
    g["means2D"] = g["means2D"].view(-1, 3)
    g["tau"] = g["tau"].view(-1, 6)
    g["tau_sum"] = g["tau_sum"].view(6)
    return g


def _param_grads(g, k):
    return (g["means3D"],
            g["sh"] if k["shs"] is not None else None,
            g["colors"] if k["colors_precomp"] is not None else None,
            g["language"],
            g["opacity"],
            g["scales"] if k["scales"] is not None else None,
            g["rotations"] if k["rotations"] is not None else None,
            g["cov3D"] if k["cov3D_precomp"] is not None else None)


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
        # the parameter tensors are saved so that autograd's version counters catch an in-place update between
        # forward and backward (the kernels read them again through the pointers kept in the state)
        ctx.save_for_backward(radii, means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations,
                              cov3Ds_precomp)
        ctx.mark_non_differentiable(radii, n_touched)
        return color, language, radii, depth, opacity, n_touched

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_language, grad_out_radii, grad_out_depth, grad_out_opacity,
                 grad_n_touched):
        # grad_out_opacity is ignored exactly like the reference (:296, not passed to C++)
        radii = ctx.saved_tensors[0]
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
        g = _backward_native_batch(st, [radii], [grad_out_color], [grad_out_language], [grad_out_depth])
        grad_tau = g["tau_sum"].view(6)                       # reference :383-385 (summed inside the kernel)
        grad_rho = grad_tau[:3].view(1, -1)
        grad_theta = grad_tau[3:].view(1, -1)
        pg = _param_grads(g, st.keep)
        return (pg[0], g["means2D"].view(P, 3)) + pg[1:] + (grad_theta, grad_rho, None)


class _RasterizeLanguageGaussiansBatch(torch.autograd.Function):
    """V views of the same Gaussians in one forward and one backward (utils/slam_backend.py:510-662 renders the
    window keyframes one by one, sums their losses and calls backward once: autograd then runs the reference's
    backward V times and adds the results -- here the V backward passes are one batched launch whose kernels sum
    over the views).  Inputs: the shared parameters, then (means2D, theta, rho) per view; outputs: the reference's
    6-tuple per view, flattened."""

    @staticmethod
    def forward(ctx, rs_list, means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations,
                cov3Ds_precomp, *per_view):
        outs, st = _forward_native_batch(means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations,
                                         cov3Ds_precomp, rs_list)
        ctx.state = st
        ctx.V = len(rs_list)
        ctx.save_for_backward(means3D, sh, colors_precomp, language_precomp, opacities, scales, rotations,
                              cov3Ds_precomp, *[o[2] for o in outs])
        flat = []
        for o in outs:
            ctx.mark_non_differentiable(o[2], o[5])
            flat.extend(o)
        return tuple(flat)

    @staticmethod
    def backward(ctx, *grads):
        st, V = ctx.state, ctx.V
        if st is None:
            raise RuntimeError("backward called on an empty render")
        radii = list(ctx.saved_tensors[8:])
        dev = radii[0].device
        P, F, H, W = st.args.P, st.args.F, st.args.H, st.args.W
        gc, gl, gd = [], [], []
        zc = zl = zd = None
        for v in range(V):
            c, l, d = grads[6 * v], grads[6 * v + 1], grads[6 * v + 3]
            if c is None:
                zc = torch.zeros((3, H, W), device=dev) if zc is None else zc
                c = zc
            if l is None:
                zl = torch.zeros((F, H, W), device=dev) if zl is None else zl
                l = zl
            if d is None:
                zd = torch.zeros((1, H, W), device=dev) if zd is None else zd
                d = zd
            gc.append(c); gl.append(l); gd.append(d)
        g = _backward_native_batch(st, radii, gc, gl, gd)
        pg = _param_grads(g, st.keep)
        per_view = []
        for v in range(V):
            tau = g["tau_sum"][v]
            per_view.extend((g["means2D"][v], tau[3:].view(1, -1), tau[:3].view(1, -1)))
        return (None,) + pg + tuple(per_view)


def rasterize_language_gaussians_batch(means3D, means2D_list, sh, colors_precomp, language_precomp, opacities, scales,
                                       rotations, cov3Ds_precomp, theta_list, rho_list, raster_settings_list):
    """-> [(color, language, radii, depth, opacity, n_touched)] per view."""
    per_view = []
    for m2, th, rh in zip(means2D_list, theta_list, rho_list):
        per_view.extend((m2, th, rh))
    flat = _RasterizeLanguageGaussiansBatch.apply(tuple(raster_settings_list), means3D, sh, colors_precomp,
                                                  language_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                                  *per_view)
    return [tuple(flat[6 * v:6 * v + 6]) for v in range(len(raster_settings_list))]


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
