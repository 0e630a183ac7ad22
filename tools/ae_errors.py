"""Measured errors of the fused autoencoder against torch fp32 (both precision modes): python tools/ae_errors.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from online_lang_splatting_b200 import autoencoder as AE
from oracle import torch_oracle as TO
import test_ae as T
dev = torch.device("cuda:0")
for name in sorted(T.CASES):
    model, din = T.build(name)
    model = model.to(dev)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(36864, din, generator=g)
    x = (x / x.norm(dim=-1, keepdim=True)).to(dev)
    for mode in ("fast", "fp32"):
        AE.PRECISION = mode
        with torch.no_grad():
            code = model.encode(x); rec = model.decode(code)
            rc = TO.reference_chain(list(model.encoder), x); rr = TO.reference_chain(list(model.decoder), code)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                model.encode(x)
            e1.record(); torch.cuda.synchronize()
        print(name, mode, "encode (cos_min, rel_l2)", T._metrics(code.cpu(), rc.cpu()), "decode", T._metrics(rec.cpu(), rr.cpu()),
              "encode ms", e0.elapsed_time(e1) / 5, flush=True)
