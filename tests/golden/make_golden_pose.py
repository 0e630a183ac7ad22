"""Golden for the fused pose step: the REAL utils/pose_utils.py (update_pose, SE3_exp) + torch.optim.Adam with the
reference's four parameter groups (utils/slam_frontend.py:183-213), driven by a seeded sequence of gradients on the CPU.
    python tests/golden/make_golden_pose.py     # needs /root/reference; writes tests/golden/pose_small.npz"""
import importlib.util
import os
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OLS_REFERENCE_ROOT", "/root/reference")


def gradient_sequence(n=6, seed=11):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(3, generator=g) * 10 ** (i % 3 - 1), torch.randn(3, generator=g) * 10 ** (i % 3 - 1),
             torch.randn(1, generator=g), torch.randn(1, generator=g)) for i in range(n)]


def main():
    spec = importlib.util.spec_from_file_location("ref_pose_utils", os.path.join(REF, "utils", "pose_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = torch.Generator().manual_seed(7)
    q = torch.randn(4, generator=g)
    q = q / q.norm()
    r, x, y, z = q.tolist()
    R0 = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
                       [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
                       [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]])
    T0 = torch.randn(3, generator=g)
    cam = types.SimpleNamespace(R=R0.clone(), T=T0.clone(), device="cpu",
                                cam_rot_delta=torch.nn.Parameter(torch.zeros(3)), cam_trans_delta=torch.nn.Parameter(torch.zeros(3)),
                                exposure_a=torch.nn.Parameter(torch.tensor([0.0])), exposure_b=torch.nn.Parameter(torch.tensor([0.0])))

    def update_RT(R, t):
        cam.R, cam.T = R, t
    cam.update_RT = update_RT
    opt = torch.optim.Adam([{"params": [cam.cam_rot_delta], "lr": 0.003}, {"params": [cam.cam_trans_delta], "lr": 0.001},
                            {"params": [cam.exposure_a], "lr": 0.01}, {"params": [cam.exposure_b], "lr": 0.01}])
    out = {"R0": R0.numpy(), "T0": T0.numpy()}
    conv = []
    for i, (g_rot, g_trans, g_a, g_b) in enumerate(gradient_sequence()):
        opt.zero_grad()
        cam.cam_rot_delta.grad, cam.cam_trans_delta.grad = g_rot.clone(), g_trans.clone()
        cam.exposure_a.grad, cam.exposure_b.grad = g_a.clone(), g_b.clone()
        with torch.no_grad():
            opt.step()
            conv.append(bool(mod.update_pose(cam)))
        out[f"R{i + 1}"] = cam.R.detach().numpy().copy()
        out[f"T{i + 1}"] = cam.T.detach().numpy().copy()
    out["exposure"] = np.array([cam.exposure_a.item(), cam.exposure_b.item()])
    out["converged"] = np.array(conv)
    np.savez_compressed(os.path.join(HERE, "pose_small.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
