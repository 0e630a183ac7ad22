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
torch.cuda.synchronize()
print("sanitize pass done", float(l), tuple(y.shape), tuple(hm.shape), tuple(code.shape), cnt.tolist())
