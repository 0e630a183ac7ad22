"""Forward+backward timing at the headline shape (development aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _util as U
from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["compat", "exact"]
dev = torch.device("cuda:0")
sc = U.make_scene(P=P, F=15, W=960, H=540, seed=0, scale=0.01, view=int(os.environ.get("OLS_VIEW", "0")))
grads = [g.to(dev) for g in U.loss_weights(15, 960, 540, seed=1)]
for tile in ((15,) if os.environ.get("OLS_TILE15") else (15, 16)):
    for mode in modes:
        rs = U.settings(sc, dev, tile=tile, bitexact=False, backward_mode=mode)._replace(debug=False)
        d = lambda k: sc[k].to(dev)
        args = (d("means3D"), d("shs"), torch.Tensor([]), d("language"), d("opacities"), d("scales"), d("rotations"),
                torch.Tensor([]), rs)
        def step():
            R, color, language, radii, depth, opacity, n_touched, st = dgr._forward_native(*args)
            return dgr._backward_native(st, radii, grads[0], grads[1], grads[2])
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        tf = tb = 0.0
        from online_lang_splatting_b200 import _native as N
        N.timing_begin(iters * 16 + 64)
        for _ in range(iters):
            e0.record()
            R, color, language, radii, depth, opacity, n_touched, st = dgr._forward_native(*args)
            e1.record()
            dgr._backward_native(st, radii, grads[0], grads[1], grads[2])
            e2.record(); torch.cuda.synchronize()
            tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
        marks = N.timing_end()
        print(f"P={P} tile={tile} mode={mode} R={R} fwd_ms={tf/iters:.3f} bwd_ms={tb/iters:.3f}  " +
              " ".join(f"{k}={v[0]/max(v[1],1):.3f}" for k, v in marks.items() if v[1]), flush=True)
if os.environ.get("OLS_SKIP_REF"):
    sys.exit(0)
mod = U.ref_module("ref_P_C")
if mod is not None:
    for it in range(3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        d = lambda t: t.to(dev).contiguous()
        e = torch.Tensor([])
        a = (d(sc["bg"]), d(sc["means3D"]), e, d(sc["language"]), d(sc["opacities"]), d(sc["scales"]),
             d(sc["rotations"]), 1.0, e, d(sc["viewmatrix"]), d(sc["projmatrix"]), d(sc["projmatrix_raw"]),
             sc["tanfovx"], sc["tanfovy"], 540, 960, d(sc["shs"]), 0, d(sc["campos"]), False, False)
        torch.cuda.synchronize()
        e0.record()
        R, color, language, radii, geom, binning, img, depth, opacity, n_touched = mod.rasterize_language_gaussians(*a)
        e1.record()
        b = (a[0], a[1], radii, e, a[3], a[5], a[6], 1.0, e, a[9], a[10], a[11], a[12], a[13], grads[0], grads[1], grads[2],
             a[16], 0, a[18], geom, R, binning, img, False)
        mod.rasterize_language_gaussians_backward(*b)
        e2.record(); torch.cuda.synchronize()
        print(f"reference P/ CUDA fwd_ms={e0.elapsed_time(e1):.3f} bwd_ms={e1.elapsed_time(e2):.3f} R={R}", flush=True)
