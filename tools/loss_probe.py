"""Timing of the fused mapping loss at two image sizes (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from online_lang_splatting_b200 import losses as LS
dev = torch.device("cuda:0")
for (W, H) in ((960, 540), (1200, 680), (1216, 680)):
    img = torch.rand(3, H, W, device=dev, requires_grad=True); dep = torch.rand(1, H, W, device=dev, requires_grad=True)
    lang = torch.randn(15, H, W, device=dev, requires_grad=True)
    gti, gtd, gtl = torch.rand(3, H, W, device=dev), torch.rand(1, H, W, device=dev), torch.randn(15, 192, 192, device=dev)
    def f():
        return LS.mapping_loss(img, dep, gti, gtd, lang, gtl)
    for _ in range(3):
        f().backward()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(20):
        e[0].record(); l = f(); e[1].record(); l.backward(); e[2].record(); torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print(W, H, "fwd ms", tf / 20, "bwd ms", tb / 20, flush=True)
