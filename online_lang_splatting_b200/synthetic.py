"""Deterministic synthetic scenes / cameras / CLIP maps for tests and bench.

There is no network for datasets or checkpoints, so every configuration of BASELINE.json is
driven by the generators below.  Everything is produced on the CPU with a seeded
``torch.Generator`` and moved to the GPU by the caller, so the oracle (CPU) and the CUDA
path (GPU) see bit-identical inputs.

Camera conventions follow the reference (``utils/camera_utils.py:103-117``,
``gaussian_splatting/utils/graphics_utils.py:33-93``): matrices are handed to the rasterizer
*transposed* (row-vector convention).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch


def projection_matrix2(znear, zfar, cx, cy, fx, fy, W, H) -> torch.Tensor:
    """OpenCV-intrinsics projection (restates graphics_utils.getProjectionMatrix2, :64-93)."""
    left = ((2 * cx - W) / W - 1.0) * W / 2.0
    right = ((2 * cx - W) / W + 1.0) * W / 2.0
    top = ((2 * cy - H) / H + 1.0) * H / 2.0
    bottom = ((2 * cy - H) / H - 1.0) * H / 2.0
    left, right = znear / fx * left, znear / fx * right
    top, bottom = znear / fy * top, znear / fy * bottom
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def so3_exp(w: torch.Tensor) -> torch.Tensor:
    th = float(w.norm())
    K = torch.tensor([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=torch.float64)
    if th < 1e-8:
        return torch.eye(3, dtype=torch.float64) + K
    return torch.eye(3, dtype=torch.float64) + math.sin(th) / th * K + (1 - math.cos(th)) / th ** 2 * (K @ K)


@dataclass
class SyntheticCamera:
    """Carries exactly the attributes ``render()`` reads from a viewpoint."""
    image_width: int
    image_height: int
    fx: float
    fy: float
    cx: float
    cy: float
    R: torch.Tensor  # world->camera rotation [3,3]
    T: torch.Tensor  # world->camera translation [3]
    device: str = "cpu"
    uid: int = 0
    cam_rot_delta: Optional[torch.Tensor] = None
    cam_trans_delta: Optional[torch.Tensor] = None
    gt_lang_feat: Optional[torch.Tensor] = None
    _cache: Dict[str, torch.Tensor] = field(default_factory=dict, repr=False)

    def __post_init__(self):
        self.FoVx = 2 * math.atan(self.image_width / (2 * self.fx))
        self.FoVy = 2 * math.atan(self.image_height / (2 * self.fy))
        proj = projection_matrix2(0.01, 100.0, self.cx, self.cy, self.fx, self.fy,
                                  self.image_width, self.image_height).transpose(0, 1)
        self.projection_matrix = proj.to(self.device)
        if self.cam_rot_delta is None:
            self.cam_rot_delta = torch.zeros(3, device=self.device, requires_grad=True)
        if self.cam_trans_delta is None:
            self.cam_trans_delta = torch.zeros(3, device=self.device, requires_grad=True)

    @property
    def world_view_transform(self) -> torch.Tensor:
        if "wv" not in self._cache:
            Rt = torch.eye(4, dtype=torch.float32)
            Rt[:3, :3] = self.R.float()
            Rt[:3, 3] = self.T.float()
            self._cache["wv"] = Rt.transpose(0, 1).contiguous().to(self.device)
        return self._cache["wv"]

    @property
    def full_proj_transform(self) -> torch.Tensor:
        if "fp" not in self._cache:
            self._cache["fp"] = (self.world_view_transform.unsqueeze(0).bmm(
                self.projection_matrix.unsqueeze(0))).squeeze(0).contiguous()
        return self._cache["fp"]

    @property
    def camera_center(self) -> torch.Tensor:
        if "cc" not in self._cache:
            self._cache["cc"] = self.world_view_transform.cpu().inverse()[3, :3].contiguous().to(self.device)
        return self._cache["cc"]

    def to(self, device) -> "SyntheticCamera":
        return SyntheticCamera(self.image_width, self.image_height, self.fx, self.fy, self.cx, self.cy,
                               self.R, self.T, device=str(device), uid=self.uid, gt_lang_feat=self.gt_lang_feat)


def make_camera(W: int, H: int, view: int = 0, seed: int = 0, fx: Optional[float] = None,
                fy: Optional[float] = None, device: str = "cpu") -> SyntheticCamera:
    """View 0 is the identity pose; view k>0 is a small seeded SE(3) perturbation
    (sigma_rot 0.05 rad, sigma_trans 0.1 m)."""
    fx = W / 2.0 if fx is None else fx
    fy = W / 2.0 if fy is None else fy
    R = torch.eye(3, dtype=torch.float64)
    T = torch.zeros(3, dtype=torch.float64)
    if view > 0:
        g = torch.Generator().manual_seed(1000 * seed + view)
        R = so3_exp(0.05 * torch.randn(3, generator=g, dtype=torch.float64))
        T = 0.1 * torch.randn(3, generator=g, dtype=torch.float64)
    return SyntheticCamera(W, H, float(fx), float(fy), (W - 1) / 2.0, (H - 1) / 2.0, R, T, device=device, uid=view)


def make_gaussians(P: int, F: int, W: int, H: int, seed: int = 0, sh_degree: int = 0,
                   scale_px_sigma: float = 0.01) -> Dict[str, torch.Tensor]:
    """Activated Gaussian parameters in camera-0 space (CPU float32).  See SURVEY 8d."""
    g = torch.Generator().manual_seed(seed)
    tanx, tany = 1.0, H / W  # fx = fy = W/2
    n_near = max(P // 100, 0)
    z = torch.empty(P).uniform_(0.5, 6.0, generator=g)
    if n_near:
        z[:n_near] = torch.empty(n_near).uniform_(-1.0, 0.2, generator=g)
        z = z[torch.randperm(P, generator=g)]
    x = z * tanx * torch.empty(P).uniform_(-1.1, 1.1, generator=g)
    y = z * tany * torch.empty(P).uniform_(-1.1, 1.1, generator=g)
    means = torch.stack([x, y, z], 1).contiguous()
    log_s = math.log(scale_px_sigma) + 0.5 * torch.randn(P, 1, generator=g) + 0.2 * torch.randn(P, 3, generator=g)
    scales = torch.exp(log_s).contiguous()
    q = torch.randn(P, 4, generator=g)
    rot = (q / q.norm(dim=1, keepdim=True)).contiguous()
    opac = torch.sigmoid(1.5 * torch.randn(P, 1, generator=g)).contiguous()
    M = (sh_degree + 1) ** 2
    shs = torch.randn(P, M, 3, generator=g).contiguous()
    if M > 1:
        shs[:, 1:] *= 0.2
    lang = torch.randn(P, F, generator=g)
    lang = (lang / lang.norm(dim=1, keepdim=True)).contiguous()
    return {"means3D": means, "scales": scales, "rotations": rot, "opacities": opac, "shs": shs, "language": lang}


def make_clip_maps(B: int, seed: int = 0, C: int = 768, HW: int = 192) -> torch.Tensor:
    """Random unit-norm CLIP-like maps, already flattened to [B*HW*HW, C] as the back-end does
    (utils/slam_backend.py:392-395)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B * HW * HW, C, generator=g)
    return x / x.norm(dim=1, keepdim=True)


class PipelineParams:
    """The two flags ``render()`` reads from ``pipe`` (configs/rgbd/replicav2/base_config.yaml:97-99)."""

    def __init__(self, compute_cov3D_python: bool = False, convert_SHs_python: bool = False):
        self.compute_cov3D_python = compute_cov3D_python
        self.convert_SHs_python = convert_SHs_python


class SyntheticGaussianModel:
    """Stand-in for ``gaussian_splatting/scene/gaussian_model.py:GaussianModel`` exposing exactly the
    ``get_*`` properties ``render()`` reads.  Parameters are stored *pre-activation*
    like the reference (:51-57) and activated on access (:67-72, :93-130): exp for scales,
    sigmoid for opacity, L2-normalise for rotations."""

    def __init__(self, g: Dict[str, torch.Tensor], device="cuda", sh_degree: int = 0, is_language: bool = True,
                 requires_grad: bool = False):
        dev = torch.device(device)
        mk = lambda t: t.to(dev).clone().requires_grad_(requires_grad)
        self._xyz = mk(g["means3D"])
        self._scaling = mk(torch.log(g["scales"]))
        self._rotation = mk(g["rotations"])
        op = g["opacities"].clamp(1e-6, 1 - 1e-6)
        self._opacity = mk(torch.log(op / (1 - op)))
        self._features_dc = mk(g["shs"][:, :1, :])
        self._features_rest = mk(g["shs"][:, 1:, :])
        self._language_feature = mk(g["language"])
        self.active_sh_degree = sh_degree
        self.max_sh_degree = sh_degree
        self.is_language = is_language

    def parameters(self):
        return [self._xyz, self._scaling, self._rotation, self._opacity, self._features_dc, self._features_rest,
                self._language_feature]

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_scaling(self):
        return torch.exp(self._scaling)

    @property
    def get_rotation(self):
        return torch.nn.functional.normalize(self._rotation)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_language_features(self):
        return self._language_feature

    def get_covariance(self, scaling_modifier=1.0):
        """build_covariance_from_scaling_rotation (gaussian_model.py:59-65): L = R S, Sigma = L L^T, upper 6."""
        s = self.get_scaling * scaling_modifier
        q = self.get_rotation
        r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                         2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                         2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)
        L = R * s[:, None, :]
        S = L @ L.transpose(1, 2)
        return torch.stack([S[:, 0, 0], S[:, 0, 1], S[:, 0, 2], S[:, 1, 1], S[:, 1, 2], S[:, 2, 2]], -1)
