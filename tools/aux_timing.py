"""CUDA-event times of the kernels either side of render() at the benchmark's sizes, each next to the
torch code of the reference it replaces, on the same GPU:  python tools/aux_timing.py  -> one JSON object on stdout."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from online_lang_splatting_b200 import losses as LS, densification as DN
from online_lang_splatting_b200.optim import FlatAdam
from online_lang_splatting_b200.simple_knn._C import distCUDA2
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_oracle as TO  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
P, H, W, F = 1_000_000, 540, 960, 15
PEAK = 6454.6

def entry(name, ms, ref_ms, nbytes, note):
    out[name] = {"ms": round(ms, 4), "torch_reference_ms": round(ref_ms, 4), "speedup": round(ref_ms / ms, 2),
                 "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / PEAK, 3),
                 "note": note}

# ---- flat Adam over the 1M-Gaussian parameter buffer (59 floats per Gaussian at SH degree 3 features + 15 language) ----
shapes = [("xyz", 3, 1.6e-4), ("f_dc", 3, 2.5e-3), ("f_rest", 45, 1.25e-4), ("opacity", 1, 5e-2), ("scaling", 3, 1e-3),
          ("rotation", 4, 1e-3), ("f_language", 15, 2.5e-3)]
n = sum(c for _, c, _ in shapes) * P
flat_p, flat_g = torch.randn(n, device=dev), torch.randn(n, device=dev)
opt = FlatAdam(flat_p, flat_g, [(nm, c * P, lr) for nm, c, lr in shapes])
ref_params = [torch.nn.Parameter(torch.randn(P, c, device=dev)) for _, c, _ in shapes]
for p in ref_params:
    p.grad = torch.randn_like(p)
ref = torch.optim.Adam([{"params": [p], "lr": lr} for p, (_, _, lr) in zip(ref_params, shapes)], lr=0.0, eps=1e-15)
entry("adam_step", timeit(opt.step), timeit(ref.step), 28 * n, "1M Gaussians x 74 floats; reads p,g,m,v, writes p,m,v")
del ref, ref_params, opt, flat_p, flat_g
torch.cuda.empty_cache()

# ---- mapping loss forward + backward ----
img = torch.rand(3, H, W, device=dev, requires_grad=True); dep = torch.rand(1, H, W, device=dev, requires_grad=True)
lang = torch.randn(F, H, W, device=dev, requires_grad=True)
gti, gtd, gtl = torch.rand(3, H, W, device=dev), torch.rand(1, H, W, device=dev), torch.randn(F, 192, 192, device=dev)
def ours_map():
    LS.mapping_loss(img, dep, gti, gtd, lang, gtl).backward()
def ref_map():
    TO.reference_mapping_loss(img, dep, gti, gtd, lang, gtl).backward()
entry("mapping_loss_fwd_bwd", timeit(ours_map), timeit(ref_map), (19 + 4) * 4 * H * W + 19 * 4 * H * W * 2,
      "960x540, 15 language channels, low-resolution target 192x192 resident")

# ---- SSIM colour-refinement loss forward + backward ----
im2 = torch.rand(3, H, W, device=dev, requires_grad=True)
def ours_ssim():
    LS.color_refinement_loss(im2, gti).backward()
def ref_ssim():
    TO.reference_color_refinement_loss(im2, gti).backward()
entry("color_refinement_loss_fwd_bwd", timeit(ours_ssim), timeit(ref_ssim), (2 + 3 + 3 + 2 + 1) * 3 * 4 * H * W,
      "3x540x960; forward reads 2 images, writes 3 maps; backward reads 3 maps + 2 images, writes 1")

# ---- densification statistics + flags ----
radii = torch.randint(0, 30, (P,), device=dev, dtype=torch.int32); radii[radii < 12] = 0
vg = torch.randn(P, 3, device=dev) * 1e-3
mr, acc, den = torch.zeros(P, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev)
mr2, acc2, den2 = mr.clone(), acc.clone(), den.clone()
entry("densify_stats", timeit(lambda: DN.update_stats(radii, vg, mr, acc, den)),
      timeit(lambda: TO.reference_update_stats(radii, vg, mr2, acc2, den2)), P * (4 + 8 + 3 * 8 * 0.6), "1M Gaussians, 60 % visible")
sc, op = torch.randn(P, 3, device=dev) - 3, torch.randn(P, 1, device=dev)
kw = dict(max_grad=2e-4, min_opacity=0.3, extent=4.0, max_screen_size=20.0)
entry("densify_flags", timeit(lambda: DN.densify_flags(acc, den, sc, op, mr, **kw)),
      timeit(lambda: TO.reference_densify_flags(acc.clone(), den, sc, op, mr, **kw)), P * (4 + 4 + 12 + 4 + 4 + 1), "1M Gaussians")

# ---- online autoencoder training step (a17): fused kernel vs the reference's torch sequence (slam_backend.py:266-323) ----
import copy
from online_lang_splatting_b200 import autoencoder as AE
torch.manual_seed(0)
on_ref = AE.EncoderDecoderOnline().to(dev)
on_fused = copy.deepcopy(on_ref)
feats = torch.randn(192 * 192, 32, device=dev)
feats = feats / feats.norm(dim=-1, keepdim=True)
o_opt = torch.optim.Adam(on_ref.parameters(), lr=1e-3)
def torch_online_step():
    on_ref.train()
    o_opt.zero_grad()
    comp = on_ref.encode(feats)
    recon = on_ref.decode(comp)
    loss = torch.nn.functional.l1_loss(recon, feats) + 0.6 * (1 - torch.nn.functional.cosine_similarity(recon, feats, dim=1).mean())
    loss.backward()
    o_opt.step()
entry("online_ae_train_step", timeit(lambda: on_fused.fused_train_step(feats, lr=1e-3)), timeit(torch_online_step),
      192 * 192 * (32 + 15) * 4, "36,864 rows x 32; forward + L1 + 0.6 (1 - cos) + backward + Adam; reference = the module's torch graph + torch.optim.Adam")

# ---- tracking pose step: fused kernel vs torch Adam + the reference's update_pose arithmetic on the device ----
from online_lang_splatting_b200.tracking import DeviceCamera, PoseOptimizer
cam = DeviceCamera(W, H, W / 2.0, W / 2.0, (W - 1) / 2.0, (H - 1) / 2.0, torch.eye(3), torch.zeros(3), device=dev)
popt = PoseOptimizer(cam)
def fused_pose():
    cam._grad_tau.fill_(1e-3)
    popt.step()
rot, trans = torch.nn.Parameter(torch.zeros(3, device=dev)), torch.nn.Parameter(torch.zeros(3, device=dev))
ea, eb = torch.nn.Parameter(torch.zeros(1, device=dev)), torch.nn.Parameter(torch.zeros(1, device=dev))
t_opt = torch.optim.Adam([{"params": [rot], "lr": 0.003}, {"params": [trans], "lr": 0.001}, {"params": [ea], "lr": 0.01}, {"params": [eb], "lr": 0.01}])
Rm, Tm = torch.eye(3, device=dev), torch.zeros(3, device=dev)
def skew(x):
    m = torch.zeros(3, 3, device=dev)
    m[0, 1], m[0, 2], m[1, 0], m[1, 2], m[2, 0], m[2, 1] = -x[2], x[1], x[2], -x[0], -x[1], x[0]
    return m
def torch_pose():
    global Rm, Tm
    for p_ in (rot, trans, ea, eb):
        p_.grad = torch.full_like(p_, 1e-3)
    t_opt.step()
    with torch.no_grad():      # utils/pose_utils.py:24-95 (SO3_exp, V, SE3_exp, update_pose) incl. its host-side branches
        tau = torch.cat([trans, rot])
        Wm = skew(tau[3:]); W2 = Wm @ Wm
        angle = torch.norm(tau[3:])
        I = torch.eye(3, device=dev)
        if angle < 1e-5:
            Rd, Vm = I + Wm + 0.5 * W2, I + 0.5 * Wm + W2 / 6.0
        else:
            Rd = I + (torch.sin(angle) / angle) * Wm + ((1 - torch.cos(angle)) / angle ** 2) * W2
            Vm = I + Wm * ((1.0 - torch.cos(angle)) / angle ** 2) + W2 * ((angle - torch.sin(angle)) / angle ** 3)
        T4 = torch.eye(4, device=dev); T4[:3, :3] = Rd; T4[:3, 3] = Vm @ tau[:3]
        Tw = torch.eye(4, device=dev); Tw[:3, :3] = Rm; Tw[:3, 3] = Tm
        new = T4 @ Tw
        Rm, Tm = new[:3, :3], new[:3, 3]
        converged = bool(tau.norm() < 1e-4)
        rot.data.fill_(0); trans.data.fill_(0)
        wv = new.t(); fp_ = wv @ cam.projection_matrix; cc = wv.inverse()[3, :3]
entry("tracking_pose_step", timeit(fused_pose), timeit(torch_pose), 0, "Adam over rot / trans / exposure + SE3_exp + camera matrices; reference = torch.optim.Adam + pose_utils.update_pose + Camera properties")

# ---- distCUDA2 ----
pts = torch.rand(P, 3, device=dev) * 4
ms_knn = timeit(lambda: distCUDA2(pts), n=5, warm=2)
out["distCUDA2"] = {"ms": round(ms_knn, 3), "note": "1M uniformly random points; exact 3-NN mean squared distance (the reference's simple-knn is not importable on the GPU box)"}
print(json.dumps(out))
