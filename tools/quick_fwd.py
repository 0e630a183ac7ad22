"""Times the forward at the headline shape (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import _util as U
from online_lang_splatting_b200 import diff_gaussian_rasterization as dgr

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
dev = torch.device("cuda:0")
sc = U.make_scene(P=P, F=15, W=960, H=540, seed=0, scale=0.01)
for tile in (15, 16):
    for bitexact in (True, False):
        rs = U.settings(sc, dev, tile=tile, bitexact=bitexact)._replace(debug=False)
        d = lambda k: sc[k].to(dev)
        args = (d("means3D"), d("shs"), torch.Tensor([]), d("language"), d("opacities"), d("scales"), d("rotations"),
                torch.Tensor([]), rs)
        for _ in range(3):
            R = dgr._forward_native(*args)[0]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 20
        for _ in range(n):
            dgr._forward_native(*args)
        e1.record(); torch.cuda.synchronize()
        print(f"P={P} tile={tile} bitexact={bitexact} R={R} fwd_ms={e0.elapsed_time(e1)/n:.3f}", flush=True)
mod = U.ref_module("ref_P_C")
if mod is not None:
    for _ in range(3):
        r = U.run_ref(mod, sc, dev)
    d = lambda t: t.to(dev).contiguous()
    e = torch.Tensor([])
    a = (d(sc["bg"]), d(sc["means3D"]), e, d(sc["language"]), d(sc["opacities"]), d(sc["scales"]),
         d(sc["rotations"]), 1.0, e, d(sc["viewmatrix"]), d(sc["projmatrix"]), d(sc["projmatrix_raw"]),
         sc["tanfovx"], sc["tanfovy"], 540, 960, d(sc["shs"]), 0, d(sc["campos"]), False, False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        mod.rasterize_language_gaussians(*a)
    e1.record(); torch.cuda.synchronize()
    print(f"reference P/ CUDA fwd_ms={e0.elapsed_time(e1)/10:.3f} R={r['R']}")
