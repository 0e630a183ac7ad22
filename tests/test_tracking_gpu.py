"""Fused pose step (ols_pose_adam_step) against the REAL reference sequence -- torch.optim.Adam with the front-end's four
groups + utils/pose_utils.py:update_pose -- captured in tests/golden/pose_small.npz; and a short tracking run through
render() + tracking_loss() that must pull a perturbed pose back towards the true one."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_pose_step_matches_real_reference_golden(cuda):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_pose import gradient_sequence
    from online_lang_splatting_b200.tracking import DeviceCamera, PoseOptimizer
    z = np.load(os.path.join(HERE, "golden", "pose_small.npz"))
    cam = DeviceCamera(64, 48, 40.0, 40.0, 31.5, 23.5, torch.from_numpy(z["R0"]), torch.from_numpy(z["T0"]), device=cuda)
    opt = PoseOptimizer(cam)
    for i, (g_rot, g_trans, g_a, g_b) in enumerate(gradient_sequence()):
        cam._grad_tau[0:3].copy_(g_trans); cam._grad_tau[3:6].copy_(g_rot)
        cam._grad_exposure[0:1].copy_(g_a); cam._grad_exposure[1:2].copy_(g_b)
        opt.step()
        assert np.abs(cam.R.cpu().numpy() - z[f"R{i + 1}"]).max() < 2e-6, i
        assert np.abs(cam.T.cpu().numpy() - z[f"T{i + 1}"]).max() < 2e-6, i
        assert float(cam._grad_tau.abs().sum()) == 0.0          # zero_grad() folded into the step
        Rt = torch.eye(4)
        Rt[:3, :3], Rt[:3, 3] = cam.R.cpu(), cam.T.cpu()
        assert torch.allclose(cam.world_view_transform.cpu(), Rt.t(), atol=1e-6)
        assert torch.allclose(cam.full_proj_transform.cpu(), Rt.t() @ cam.projection_matrix.cpu(), atol=1e-5)
        assert torch.allclose(cam.camera_center.cpu(), torch.linalg.inv(Rt.t())[3, :3], atol=1e-5)
    assert np.abs(cam._exposure.cpu().numpy() - z["exposure"]).max() < 2e-6
    assert bool(opt.has_converged()) == bool(z["converged"].any())


def test_tracking_loop_recovers_pose(cuda):
    from online_lang_splatting_b200 import synthetic as S
    from online_lang_splatting_b200.gaussian_renderer import render
    from online_lang_splatting_b200.losses import tracking_loss
    from online_lang_splatting_b200.tracking import DeviceCamera, PoseOptimizer
    W, H = 160, 120
    g = S.make_gaussians(20000, 15, W, H, seed=5, scale_px_sigma=0.03)
    pc = S.SyntheticGaussianModel(g, device=cuda, requires_grad=False)
    pipe, bg = S.PipelineParams(), torch.zeros(3, device=cuda)
    true = S.make_camera(W, H, view=3, seed=5)
    cam = DeviceCamera(W, H, W / 2.0, W / 2.0, (W - 1) / 2.0, (H - 1) / 2.0, true.R, true.T, device=cuda)
    with torch.no_grad():
        gt = render(cam, pc, pipe, bg)
        cam.original_image, cam.depth = gt["render"].clone(), gt["depth"].clone()
    # perturb: start from a nearby pose
    T0 = true.T.clone().float()
    T0[0] += 0.03
    cam.update_RT(true.R.float(), T0)
    err0 = float((cam.T.cpu() - true.T.float()).norm())
    opt = PoseOptimizer(cam)
    mask = torch.ones(1, H, W, device=cuda)
    for it in range(60):
        out = render(cam, pc, pipe, bg)
        loss = tracking_loss(out["render"], out["depth"], out["opacity"], cam.original_image, cam.depth, mask,
                             exposure_a=cam.exposure_a, exposure_b=cam.exposure_b)
        loss.backward()
        opt.step()
    err1 = float((cam.T.cpu() - true.T.float()).norm())
    assert err1 < 0.5 * err0, (err0, err1)
