"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): joint + disentangled rasterizer forward and
backward in both gradient modes, the fused mapping loss and the autoencoder."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _util as U
dev = torch.device("cuda:0")
for tile, F in ((15, 15), (16, 3)):
    sc = U.make_scene(P=1500, F=F, W=100, H=70, seed=2, scale=0.1, bg=(0.1, 0.2, 0.3))
    grads = U.loss_weights(F, 100, 70, seed=1)
    for mode in ("compat", "exact"):
        o = U.run_ours(sc, dev, tile=tile, grads=grads, backward_mode=mode, bitexact=(mode == "exact"))
    scd = U.add_lang_footprint(sc, seed=3)
    for mode in ("compat", "exact"):
        o = U.run_ours_dis(scd, dev, tile=tile, grads=grads, backward_mode=mode)
from online_lang_splatting_b200 import losses as LS, autoencoder as AE
img = torch.rand(3, 70, 100, device=dev, requires_grad=True); dep = torch.rand(1, 70, 100, device=dev, requires_grad=True)
lang = torch.randn(15, 70, 100, device=dev, requires_grad=True)
l = LS.mapping_loss(img, dep, torch.rand(3, 70, 100, device=dev), torch.rand(1, 70, 100, device=dev), lang,
                    torch.randn(15, 19, 23, device=dev), exposure_a=0.1, exposure_b=0.0)
l.backward()
torch.manual_seed(0)
ae = AE.AutoencoderMLP([384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]).eval().to(dev)
x = torch.randn(300, 768, device=dev)
with torch.no_grad():
    y = ae.decode(ae.encode(x))
from online_lang_splatting_b200.optim import FlatAdam
fp, fg = torch.randn(1003, device=dev), torch.randn(1003, device=dev)
opt = FlatAdam(fp, fg, [("a", 500, 1e-3), ("b", 503, 1e-2)])
opt.step(); opt.step()
t = LS.tracking_loss(img, dep, torch.rand(1, 70, 100, device=dev, requires_grad=True), torch.rand(3, 70, 100, device=dev),
                     torch.rand(1, 70, 100, device=dev), (torch.rand(1, 70, 100, device=dev) > 0.5).float())
t.backward()
# sort paths: exact depth ties + a clump (radix kernel) next to spread depths (bucket kernel)
sc = U.make_scene(P=3000, F=3, W=60, H=45, seed=4, scale=0.6)
sc["means3D"][:1200, 2] = 2.0
o = U.run_ours(sc, dev, tile=15)
# HR up-sampler (TMA loads / stores, clusters with DSMEM split-K, PDL, side stream), fused HR -> AE encode, SSIM, densification
from online_lang_splatting_b200 import supervised_net as SN, densification as DN
torch.manual_seed(1)
hr = SN.HighResLanguageFeatureNet().eval().to(dev)
fv, f3, f2 = torch.randn(1, 768, 16, 24, device=dev), torch.randn(1, 384, 37, 50, device=dev), torch.randn(1, 192, 70, 111, device=dev)
with torch.no_grad():
    hm = hr(fv, f3, f2)
    code = ae.encode_hr(hr, fv, f3, f2)
im2 = torch.rand(3, 45, 61, device=dev, requires_grad=True)
LS.color_refinement_loss(im2, torch.rand(3, 45, 61, device=dev)).backward()
P = 1237
radii = torch.randint(0, 30, (P,), device=dev, dtype=torch.int32)
mr, acc, den = torch.zeros(P, device=dev), torch.zeros(P, 1, device=dev), torch.zeros(P, 1, device=dev)
DN.update_stats(radii, torch.randn(P, 3, device=dev), mr, acc, den)
fl, cnt = DN.densify_flags(acc, den, torch.randn(P, 3, device=dev), torch.randn(P, 1, device=dev), mr, max_grad=0.5,
                           min_opacity=0.3, extent=4.0, max_screen_size=20.0)
# round 2: multi-view batch (forward + backward, fused statistics), flat Adam with activation Jacobians + device step count,
# parameter activation, fused online-AE training step, fused pose step, fp32 parity mode of the autoencoder
from online_lang_splatting_b200 import synthetic as S, diff_gaussian_rasterization as dgr
import online_lang_splatting_b200.gaussian_renderer as GR
g_ = S.make_gaussians(1700, 15, 100, 70, seed=6, scale_px_sigma=0.08)
pc = S.SyntheticGaussianModel(g_, device=dev, requires_grad=True)
for mode in ("compat", "exact"):
    GR.BACKWARD_MODE = mode
    cams = [S.make_camera(100, 70, view=v, seed=6, device="cuda") for v in range(3)]
    outs = GR.render_batch(cams, pc, S.PipelineParams(), torch.zeros(3, device=dev))
    sum((o["render"].sum() + o["language"].sum() + o["depth"].sum()) for o in outs).backward()
GR.BACKWARD_MODE = "compat"
scs = [U.make_scene(P=900, F=15, W=64, H=48, seed=2, view=v, scale=0.1) for v in range(2)]
d_ = lambda k: scs[0][k].to(dev)
e_ = torch.Tensor([])
rsl = [U.settings(sc_, dev, bitexact=False)._replace(debug=False) for sc_ in scs]
bo, bst = dgr._forward_native_batch(d_("means3D"), d_("shs"), e_, d_("language"), d_("opacities"), d_("scales"), d_("rotations"), e_, rsl)
w_ = [t.to(dev) for t in U.loss_weights(15, 64, 48, seed=4)]
stt = (torch.zeros(900, device=dev), torch.zeros(900, device=dev), torch.zeros(900, device=dev))
dgr._backward_native_batch(bst, [o[2] for o in bo], [w_[0]] * 2, [w_[1]] * 2, [w_[2]] * 2, out={"stats": stt})
from online_lang_splatting_b200.sharding import FlatParams, FlatGradBuffer
raw = {"means3D": torch.randn(333, 3), "sh": torch.randn(333, 4, 3), "opacity": torch.randn(333, 1), "scales": torch.randn(333, 3),
       "rotations": torch.randn(333, 4), "language": torch.randn(333, 15)}
fpar = FlatParams({k: v.to(dev) for k, v in raw.items()}, 15, 4, device=dev)
fgr = FlatGradBuffer(333, 15, 4, device=dev, extra=666)
fgr.flat.normal_()
lr = {"xyz": 1e-4, "f_dc": 2.5e-3, "opacity": 0.05, "scaling": 1e-3, "rotation": 1e-3, "f_language": 2.5e-3}
fa = FlatAdam(fpar.flat, fgr.grads, fgr.adam_groups(lr), capturable=True)
fa.step(); fpar.activate(); fa.step()
online = AE.EncoderDecoderOnline().to(dev)
feats = torch.randn(1000, 32, device=dev)
online.fused_train_step(feats / feats.norm(dim=-1, keepdim=True)); online.fused_train_step(feats / feats.norm(dim=-1, keepdim=True))
from online_lang_splatting_b200.tracking import DeviceCamera, PoseOptimizer
cam = DeviceCamera(100, 70, 50.0, 50.0, 49.5, 34.5, torch.eye(3), torch.zeros(3), device=dev)
po = PoseOptimizer(cam)
cam._grad_tau.fill_(0.01); po.step(); po.step()
AE.PRECISION = "fp32"
with torch.no_grad():
    y32 = ae.decode(ae.encode(x))
AE.PRECISION = "fast"
torch.cuda.synchronize()
print("sanitize pass done", float(l), tuple(y.shape), tuple(hm.shape), tuple(code.shape), cnt.tolist())
