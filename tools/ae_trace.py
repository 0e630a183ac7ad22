"""Phase timeline of the fused autoencoder kernel (development aid): OLS_AE_TRACE=1 python tools/ae_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from online_lang_splatting_b200 import autoencoder as AE
dev = torch.device("cuda:0")
torch.manual_seed(0)
ae = AE.AutoencoderMLP([384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]).eval().to(dev)
x = torch.randn(36864, 768, device=dev); x /= x.norm(dim=-1, keepdim=True)
with torch.no_grad():
    for _ in range(3):
        ae.encode(x)
torch.cuda.synchronize()
