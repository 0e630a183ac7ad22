"""Per-layer error of the HR path against the oracle + timing (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import hr_oracle
import test_hr as T
dev = torch.device("cuda:0")
z, _ = T.golden()
sd, net = T._build(int(z["seed_w"]), dev)
fv, f3, f2 = T.make_inputs(int(z["seed_x"]), int(z["s_h"]), int(z["s_w"]))
ref, inter = hr_oracle.hr_forward(sd, fv, f3, f2, return_intermediates=True)
with torch.no_grad():
    out = net(fv.to(dev), f3.to(dev), f2.to(dev))
torch.cuda.synchronize()
for which, t in inter.items():
    a = net.read_activation(which).cpu().permute(2, 0, 1)[None]
    print("act", which, tuple(a.shape), "rel rms %.3e max/rms %.3e" % T._errors(a, t), flush=True)
print("out rel rms %.3e max/rms %.3e" % T._errors(out.cpu(), ref), flush=True)
# timing at the reference size
sd, net = T._build(77, dev)
g = torch.Generator().manual_seed(3)
fv = torch.randn(1, 768, 24, 24, generator=g).to(dev); f3 = torch.randn(1, 384, 96, 96, generator=g).to(dev)
f2 = torch.randn(1, 192, 192, 192, generator=g).to(dev)
with torch.no_grad():
    for _ in range(3): net(fv, f3, f2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): net(fv, f3, f2)
    e1.record(); torch.cuda.synchronize()
print("HR forward 24x24 -> 192x192x768: %.3f ms" % (e0.elapsed_time(e1) / 20))
if os.environ.get("OLS_HR_TRACE"):
    with torch.no_grad():
        net(fv, f3, f2)
# fused HR -> AE encode timing
from online_lang_splatting_b200 import autoencoder as AE
torch.manual_seed(0)
ae = AE.AutoencoderMLP([384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]).eval().to(dev)
def timeit(fn, n=20):
    with torch.no_grad():
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("HR + AE encode, unfused: %.3f ms   fused (final_conv folded into the encoder): %.3f ms" % (
    timeit(lambda: ae.encode(net(fv, f3, f2).permute(0, 2, 3, 1).view(-1, 768))), timeit(lambda: ae.encode_hr(net, fv, f3, f2))))
