"""Device-resident camera + fused pose step for the tracking loop (the callers either side of
``render()`` in utils/slam_frontend.py:163-277).

``DeviceCamera`` carries exactly the attributes ``render()`` / ``render_batch()`` read from a viewpoint
(utils/camera_utils.py:14-135) but keeps R, T and the three derived tensors (``world_view_transform``,
``full_proj_transform``, ``camera_center``) in fixed device buffers, so a pose update is an in-place rewrite the
rasterizer sees through the same pointers -- no host round trip, CUDA-graph friendly.  ``PoseOptimizer.step()`` is
``pose_optimizer.step()`` + ``update_pose(viewpoint)`` (utils/pose_utils.py:60-95) as one kernel.  No CPU path."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _native as N
from .synthetic import projection_matrix2


class DeviceCamera:
    def __init__(self, width: int, height: int, fx: float, fy: float, cx: float, cy: float, R: torch.Tensor, T: torch.Tensor,
                 device="cuda", uid: int = 0):
        dev = torch.device(device)
        self.uid = uid
        self.device = dev
        self.image_width, self.image_height = int(width), int(height)
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy
        self.FoVx = 2 * math.atan(width / (2 * fx))
        self.FoVy = 2 * math.atan(height / (2 * fy))
        self.projection_matrix = projection_matrix2(0.01, 100.0, cx, cy, fx, fy, width, height).transpose(0, 1).contiguous().to(dev)
        self.R = R.to(dev, torch.float32).contiguous().clone()
        self.T = T.to(dev, torch.float32).contiguous().clone()
        self.world_view_transform = torch.empty(4, 4, device=dev)
        self.full_proj_transform = torch.empty(4, 4, device=dev)
        self.camera_center = torch.empty(3, device=dev)
        # pose / exposure parameters as the reference's Camera holds them; their .grad tensors are fixed views of two
        # small buffers so that the fused pose step reads (and clears) them without any glue kernels
        self._grad_tau = torch.zeros(6, device=dev)         # (rho | theta)
        self._grad_exposure = torch.zeros(2, device=dev)
        self.cam_trans_delta = torch.zeros(3, device=dev, requires_grad=True)
        self.cam_rot_delta = torch.zeros(3, device=dev, requires_grad=True)
        self._exposure = torch.zeros(2, device=dev)
        self.exposure_a = self._exposure[0:1].detach().requires_grad_(True)      # shape [1] like nn.Parameter(tensor([0.0]))
        self.exposure_b = self._exposure[1:2].detach().requires_grad_(True)
        self.attach_grads()
        self.original_image: Optional[torch.Tensor] = None
        self.depth: Optional[torch.Tensor] = None
        self.grad_mask: Optional[torch.Tensor] = None
        self.gt_lang_feat: Optional[torch.Tensor] = None
        self.coco_lang_feat: Optional[torch.Tensor] = None
        self.refresh()

    def attach_grads(self):
        self.cam_trans_delta.grad = self._grad_tau[0:3]
        self.cam_rot_delta.grad = self._grad_tau[3:6]
        self.exposure_a.grad = self._grad_exposure[0:1]
        self.exposure_b.grad = self._grad_exposure[1:2]

    def refresh(self):
        """Rebuilds the derived tensors from R, T on the device (torch ops; the fused pose step does this itself)."""
        with torch.no_grad():
            Rt = torch.eye(4, device=self.device)
            Rt[:3, :3] = self.R
            Rt[:3, 3] = self.T
            self.world_view_transform.copy_(Rt.t())
            self.full_proj_transform.copy_(self.world_view_transform @ self.projection_matrix)
            self.camera_center.copy_(-(self.R.t() @ self.T))

    def update_RT(self, R: torch.Tensor, T: torch.Tensor):
        with torch.no_grad():
            self.R.copy_(R.to(self.device, torch.float32))
            self.T.copy_(T.to(self.device, torch.float32))
        self.refresh()


class PoseOptimizer:
    """``Adam([cam_rot_delta (lr_rot), cam_trans_delta (lr_trans), exposure_a, exposure_b (0.01)])`` +
    ``update_pose`` of one camera (utils/slam_frontend.py:183-262), fused into one kernel per iteration."""

    def __init__(self, camera: DeviceCamera, lr_rot: float = 0.003, lr_trans: float = 0.001, lr_exposure: float = 0.01,
                 betas=(0.9, 0.999), eps: float = 1e-8, converged_threshold: float = 1e-4, optimize_exposure: bool = True):
        N.require_cuda()
        self.cam = camera
        dev = camera.device
        self.m = torch.zeros(8, device=dev)
        self.v = torch.zeros(8, device=dev)
        self.steps = torch.zeros(1, dtype=torch.int64, device=dev)
        self.converged = torch.zeros(1, dtype=torch.int32, device=dev)
        c = camera
        self.args = N.PoseStep(
            d_grad_tau=c._grad_tau.data_ptr(), d_grad_exposure=c._grad_exposure.data_ptr() if optimize_exposure else None,
            d_exposure=c._exposure.data_ptr() if optimize_exposure else None, d_exp_avg=self.m.data_ptr(),
            d_exp_avg_sq=self.v.data_ptr(), d_step=self.steps.data_ptr(), d_R=c.R.data_ptr(), d_T=c.T.data_ptr(),
            d_projection=c.projection_matrix.data_ptr(), d_viewmatrix=c.world_view_transform.data_ptr(),
            d_projmatrix=c.full_proj_transform.data_ptr(), d_campos=c.camera_center.data_ptr(),
            d_converged=self.converged.data_ptr(), lr_rot=lr_rot, lr_trans=lr_trans, lr_exposure=lr_exposure,
            beta1=betas[0], beta2=betas[1], eps=eps, converged_threshold=converged_threshold, zero_grads=1)

    def step(self):
        """pose_optimizer.step(); update_pose(viewpoint); pose_optimizer.zero_grad() -- asynchronous."""
        dev = self.cam.device
        with torch.cuda.device(dev):
            N.check(N.lib().ols_pose_adam_step(C.byref(self.args), torch.cuda.current_stream(dev).cuda_stream))

    def has_converged(self) -> bool:
        """Reads the device flag (synchronises): call it every few iterations, not every iteration."""
        return bool(self.converged.item())
