"""Two HR forwards at the reference size (fv 24x24 -> 192x192x768) for ncu:  ncu -k regex:k_hr_conv -s 13 -c 13 ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_hr as T
dev = torch.device("cuda:0")
sd, net = T._build(77, dev)
g = torch.Generator().manual_seed(3)
fv = torch.randn(1, 768, 24, 24, generator=g).to(dev); f3 = torch.randn(1, 384, 96, 96, generator=g).to(dev)
f2 = torch.randn(1, 192, 192, 192, generator=g).to(dev)
with torch.no_grad():
    for _ in range(2):
        out = net(fv, f3, f2)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
