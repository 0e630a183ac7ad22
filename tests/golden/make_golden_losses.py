"""Golden values of the REAL reference loss functions (utils/slam_utils.py:91-165: get_loss_mapping,
get_loss_mapping_rgbd, get_loss_tracking, get_loss_tracking_rgbd, get_loss_tracking_rgb), run in this container on
torch-CPU (the module imports only torch; its `.cuda()` calls are redirected to the CPU for the run).

    python tests/golden/make_golden_losses.py      # needs /root/reference; writes tests/golden/losses_small.npz

Inputs are re-created from the seed by tests/test_losses.py::_golden_inputs (checksums stored); the file holds the
loss values and the autograd gradients w.r.t. image / depth / opacity / exposure_a / exposure_b.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OLS_REFERENCE_ROOT", "/root/reference")


def golden_inputs(seed=5, H=40, W=56):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    d = {"image": r(3, H, W), "depth": r(1, H, W) * 4.0, "opacity": r(1, H, W) * 0.2 + 0.85,
         "gt_image": r(3, H, W), "gt_depth": r(H, W) * 4.0, "grad_mask": (r(1, H, W) > 0.3).float()}
    d["gt_image"][:, : H // 4] *= 0.001         # below the rgb boundary threshold
    d["gt_depth"][:, : W // 5] = 0.0            # invalid depth
    d["exposure_a"] = torch.tensor([0.07])      # shape [1] like Camera.exposure_a (utils/camera_utils.py:59-64)
    d["exposure_b"] = torch.tensor([-0.02])
    return d


def main():
    spec = importlib.util.spec_from_file_location("ref_slam_utils", os.path.join(REF, "utils", "slam_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.Tensor.cuda = lambda self, *a, **k: self          # the reference hard-codes .cuda(); run it on the CPU
    d = golden_inputs()
    config = {"Training": {"alpha": 0.9, "rgb_boundary_threshold": 0.01}}
    out = {"checksum": np.array([float(sum(v.double().sum() for v in d.values()))])}
    for name in ("mapping", "mapping_init", "tracking"):
        image = d["image"].clone().requires_grad_(True)
        depth = d["depth"].clone().requires_grad_(True)
        opacity = d["opacity"].clone().requires_grad_(True)
        vp = types.SimpleNamespace(original_image=d["gt_image"], depth=d["gt_depth"].numpy(), grad_mask=d["grad_mask"],
                                   exposure_a=torch.nn.Parameter(d["exposure_a"].clone()),
                                   exposure_b=torch.nn.Parameter(d["exposure_b"].clone()))
        if name == "mapping":
            loss = mod.get_loss_mapping(config, image, depth, vp, opacity)
        elif name == "mapping_init":
            loss = mod.get_loss_mapping(config, image, depth, vp, opacity, initialization=True)
        else:
            loss = mod.get_loss_tracking(config, image, depth, opacity, vp)
        (loss * 1.7).backward()
        out[name + "_loss"] = np.array([loss.item()])
        out[name + "_dimage"] = image.grad.numpy()
        out[name + "_ddepth"] = depth.grad.numpy()
        if opacity.grad is not None:
            out[name + "_dopacity"] = opacity.grad.numpy()
        if vp.exposure_a.grad is not None:
            out[name + "_dexposure_a"] = vp.exposure_a.grad.numpy()
            out[name + "_dexposure_b"] = vp.exposure_b.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "losses_small.npz"), **out)
    print({k: (v.shape, float(np.abs(v).sum())) for k, v in out.items()})


if __name__ == "__main__":
    main()
