"""Drop-in for ``gaussian_splatting.gaussian_renderer`` of the reference
(gaussian_splatting/gaussian_renderer/__init__.py:25-347): the same ``render(...)`` signature,
the same attributes read from the camera / Gaussian model / pipeline objects and the same
result dictionary, so ``utils/slam_frontend.py:218``, ``utils/slam_backend.py:352,515,610,787``,
``utils/eval_utils.py:152,168`` and ``gui/slam_gui.py:588,598`` call it unchanged.

Divergences from the reference, all documented in DESIGN.md:

* ``mask=...``: the reference raises ``ValueError`` on this path (it unpacks 5 of
  the 6 values the rasterizer returns).  Here the masked subset is rendered and ``n_touched`` is
  ``None`` as the reference intended.
* ``override_color`` is honoured in language mode (the reference tests a local it has just set to
  ``None`` and so silently ignores it, :271-288).
* ``render_batch(viewpoint_cameras, pc, pipe, bg_color)`` (extension): the window keyframes of a mapping iteration
  (utils/slam_backend.py:510-662) in one set of launches; returns the list of dictionaries ``render()`` would.
* Two module-level knobs select the rasterizer build the reference fixes at compile time:
  ``TILE_SIZE`` (15 = reference config.h) and ``BACKWARD_MODE`` ("compat" | "exact").
"""
from __future__ import annotations

import math

import torch

from ..diff_gaussian_rasterization import (
    GaussianRasterizationSettings,
    GaussianRasterizer,
    LanguageGaussianRasterizer,
    rasterize_language_gaussians_batch,
)

TILE_SIZE = 15
BACKWARD_MODE = "compat"
BITEXACT_BLEND = False

_SH_C0 = 0.28209479177387814
_SH_C1 = 0.4886025119029199
_SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
_SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """Real spherical harmonics at unit directions (restates gaussian_splatting/utils/sh_utils.py
    eval_sh, used only by the ``pipe.convert_SHs_python`` branch).  sh: [..., C, (deg+1)^2]."""
    result = _SH_C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - _SH_C1 * y * sh[..., 1] + _SH_C1 * z * sh[..., 2] - _SH_C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz = x * x, y * y, z * z
            xy, yz, xz = x * y, y * z, x * z
            result = (result + _SH_C2[0] * xy * sh[..., 4] + _SH_C2[1] * yz * sh[..., 5] +
                      _SH_C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] + _SH_C2[3] * xz * sh[..., 7] +
                      _SH_C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + _SH_C3[0] * y * (3 * xx - yy) * sh[..., 9] + _SH_C3[1] * xy * z * sh[..., 10] +
                          _SH_C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] +
                          _SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
                          _SH_C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + _SH_C3[5] * z * (xx - yy) * sh[..., 14] +
                          _SH_C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None, mask=None):
    """Render the scene.  Background tensor (bg_color) must be on GPU!  (reference :25-58)"""
    if pc.get_xyz.shape[0] == 0:  # reference :76, :210
        return None
    language_mode = bool(getattr(pc, "is_language", False))

    screenspace_points = torch.zeros_like(pc.get_xyz, dtype=pc.get_xyz.dtype, requires_grad=True, device="cuda")
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass

    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    raster_settings = GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=tanfovx,
        tanfovy=tanfovy,
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        projmatrix_raw=viewpoint_camera.projection_matrix,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=False,
        tile_size=TILE_SIZE,
        backward_mode=BACKWARD_MODE,
        bitexact_blend=BITEXACT_BLEND,
    )
    rasterizer = (LanguageGaussianRasterizer if language_mode else GaussianRasterizer)(raster_settings=raster_settings)

    means3D = pc.get_xyz
    means2D = screenspace_points
    opacity = pc.get_opacity

    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        if pc.get_scaling.shape[-1] == 1:  # isotropic Gaussians (reference :263-266)
            scales = pc.get_scaling.repeat(1, 3)
        else:
            scales = pc.get_scaling
        rotations = pc.get_rotation

    shs = colors_precomp = None
    if override_color is None:
        if pipe.convert_SHs_python:
            shs_view = pc.get_features.transpose(1, 2).view(-1, 3, (pc.max_sh_degree + 1) ** 2)
            dir_pp = pc.get_xyz - viewpoint_camera.camera_center.repeat(pc.get_features.shape[0], 1)
            dir_pp_normalized = dir_pp / dir_pp.norm(dim=1, keepdim=True)
            sh2rgb = eval_sh(pc.active_sh_degree, shs_view, dir_pp_normalized)
            colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)
        else:
            shs = pc.get_features
    else:
        colors_precomp = override_color

    sel = (lambda t: t) if mask is None else (lambda t: None if t is None else t[mask])
    kwargs = dict(means3D=sel(means3D), means2D=sel(means2D), shs=sel(shs), colors_precomp=sel(colors_precomp),
                  opacities=sel(opacity), scales=sel(scales), rotations=sel(rotations),
                  cov3D_precomp=sel(cov3D_precomp), theta=viewpoint_camera.cam_rot_delta,
                  rho=viewpoint_camera.cam_trans_delta)
    language = None
    if language_mode:
        kwargs["language_precomp"] = sel(pc.get_language_features)
        rendered_image, language, radii, depth, opacity_map, n_touched = rasterizer(**kwargs)
    else:
        rendered_image, radii, depth, opacity_map, n_touched = rasterizer(**kwargs)
    if mask is not None:
        n_touched = None

    out = {
        "render": rendered_image,
        "viewspace_points": screenspace_points,
        "visibility_filter": radii > 0,
        "radii": radii,
        "depth": depth,
        "opacity": opacity_map,
        "n_touched": n_touched,
    }
    if language_mode:
        out["language"] = language
    return out


def _settings_of(viewpoint_camera, pc, bg_color, scaling_modifier):
    return GaussianRasterizationSettings(
        image_height=int(viewpoint_camera.image_height),
        image_width=int(viewpoint_camera.image_width),
        tanfovx=math.tan(viewpoint_camera.FoVx * 0.5),
        tanfovy=math.tan(viewpoint_camera.FoVy * 0.5),
        bg=bg_color,
        scale_modifier=scaling_modifier,
        viewmatrix=viewpoint_camera.world_view_transform,
        projmatrix=viewpoint_camera.full_proj_transform,
        projmatrix_raw=viewpoint_camera.projection_matrix,
        sh_degree=pc.active_sh_degree,
        campos=viewpoint_camera.camera_center,
        prefiltered=False,
        debug=False,
        tile_size=TILE_SIZE,
        backward_mode=BACKWARD_MODE,
        bitexact_blend=BITEXACT_BLEND,
    )


def render_batch(viewpoint_cameras, pc, pipe, bg_color: torch.Tensor, scaling_modifier=1.0, override_color=None):
    """``[render(cam, pc, pipe, bg_color, ...) for cam in viewpoint_cameras]`` as ONE batched rasterizer call.

    The reference's mapping iteration renders the current window (8-12 keyframes) and two random older keyframes one
    after the other over the same Gaussians, sums their losses and back-propagates once
    (utils/slam_backend.py:510-662).  Here each Gaussian is read and its 3D covariance computed once for all the views,
    every kernel covers all views (grid.y = view), and the single backward sums the views' parameter gradients inside
    the kernels.  All cameras must have the same image size.  Each returned dictionary has its own
    ``viewspace_points`` leaf, so ``viewspace_points.grad`` is that view's screen-space gradient as in the reference.
    """
    if pc.get_xyz.shape[0] == 0:
        return [None for _ in viewpoint_cameras]
    if not bool(getattr(pc, "is_language", False)):
        return [render(cam, pc, pipe, bg_color, scaling_modifier, override_color) for cam in viewpoint_cameras]
    V = len(viewpoint_cameras)
    means3D = pc.get_xyz
    screenspace = []
    for _ in range(V):
        sp = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True, device="cuda")
        try:
            sp.retain_grad()
        except Exception:
            pass
        screenspace.append(sp)
    scales = rotations = cov3D_precomp = None
    if pipe.compute_cov3D_python:
        cov3D_precomp = pc.get_covariance(scaling_modifier)
    else:
        scales = pc.get_scaling.repeat(1, 3) if pc.get_scaling.shape[-1] == 1 else pc.get_scaling
        rotations = pc.get_rotation
    shs = colors_precomp = None
    if override_color is not None:
        colors_precomp = override_color
    elif pipe.convert_SHs_python:
        raise NotImplementedError("render_batch: convert_SHs_python produces per-view colours; use render() per view")
    else:
        shs = pc.get_features
    e = torch.Tensor([])
    nz = lambda t: e if t is None else t
    rs_list = [_settings_of(cam, pc, bg_color, scaling_modifier) for cam in viewpoint_cameras]
    res = rasterize_language_gaussians_batch(
        means3D, screenspace, nz(shs), nz(colors_precomp), pc.get_language_features, pc.get_opacity, nz(scales),
        nz(rotations), nz(cov3D_precomp), [nz(cam.cam_rot_delta) for cam in viewpoint_cameras],
        [nz(cam.cam_trans_delta) for cam in viewpoint_cameras], rs_list)
    outs = []
    for v, (rendered_image, language, radii, depth, opacity_map, n_touched) in enumerate(res):
        outs.append({"render": rendered_image, "viewspace_points": screenspace[v], "visibility_filter": radii > 0,
                     "radii": radii, "depth": depth, "opacity": opacity_map, "n_touched": n_touched,
                     "language": language})
    return outs
