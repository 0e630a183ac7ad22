import os, sys
sys.path.insert(0, "/root/repo")
import torch
from online_lang_splatting_b200 import autoencoder as AE
dev = torch.device("cuda:0")
torch.manual_seed(0)
ae = AE.AutoencoderMLP([384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]).eval().to(dev)
xs = [torch.randn(36864, 768, device=dev) for _ in range(8)]
with torch.no_grad():
    for _ in range(3):
        for x in xs: y = ae.encode(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        for x in xs: y = ae.encode(x)
    e1.record(); torch.cuda.synchronize()
print("AE encode ms/launch", e0.elapsed_time(e1) / 160, "checksum", float(y.double().sum()))
