"""bench.py --config 4 / --config 5: the two BASELINE.json configurations that are not the headline line.

config 4 (configs[3]): autoencoder only -- batch 32 of 192x192x768 random unit-norm CLIP maps -> 15-dim encode ->
  768-dim decode on one GPU (2.20 TFLOP and 7.25 GB of mandatory I/O), 1-stage chain; the 2-stage chain
  (general 768->32 + online 32->15 and back) is timed beside it.
config 5 (configs[4]): Replica-room0-shaped loop (1200x680, fx = fy = 600, window 10 + 2, 2-stage AE) in the
  reference's call pattern -- see config5().
Both print ONE JSON line in bench.py's format.
"""
from __future__ import annotations

import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

ENC1, DEC1 = [384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]           # 1-stage (language/autoencoder defaults)
ENC2, DEC2 = [512, 256, 128, 64, 32], [192, 256, 384, 512, 768]                      # 2-stage "general" AE (768 -> 32)


def _flops_per_row(dims_in, dims):
    f, k = 0, dims_in
    for d in dims:
        f += 2 * k * d
        k = d
    return f


def _bf16_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 2250.0, 6650.0, "fallback (B200_PROFILING.md)"


def config4(args, emit, peaks, ClockSampler):
    import torch
    from online_lang_splatting_b200 import _native as N
    from online_lang_splatting_b200 import autoencoder as AE
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path")
    N.require_cuda()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    B = 32
    M = B * 192 * 192
    torch.manual_seed(0)
    ae1 = AE.AutoencoderMLP(ENC1, DEC1).eval().to(dev)
    ae2 = AE.AutoencoderMLP(ENC2, DEC2).eval().to(dev)
    online = AE.EncoderDecoderOnline().eval().to(dev)
    for m in (ae1, ae2, online):
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):       # non-trivial eval-mode statistics
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
                mod.weight.data.uniform_(0.8, 1.2); mod.bias.data.normal_(0, 0.1)
        for p in m.parameters():
            p.requires_grad_(False)
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(M, 768, device=dev, generator=gen)
    x = x / x.norm(dim=-1, keepdim=True)                      # 3.6 GB: far larger than the 126 MB L2

    def chain1():
        with torch.no_grad():
            return ae1.decode(ae1.encode(x))

    def chain2():
        with torch.no_grad():
            return ae2.decode(online.decode(online.encode(ae2.encode(x))))

    def enc1():
        with torch.no_grad():
            return ae1.encode(x)

    code = enc1()

    def dec1():
        with torch.no_grad():
            return ae1.decode(code)

    def timed(fn, steps, warm):
        for _ in range(max(warm, 3)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(chain1, args.steps, args.warmup)
    clocks = sampler.stop()
    ms_enc, ms_dec = timed(enc1, args.steps, 3), timed(dec1, args.steps, 3)
    ms2 = timed(chain2, max(args.steps // 2, 2), 3)
    # parity against torch fp32 on a sample (the checker, outside every timed region)
    sys.path.insert(0, ROOT)
    from oracle import torch_oracle as TO
    idx = torch.arange(0, M, M // 4096, device=dev)[:4096]
    with torch.no_grad():
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        ref_code = TO.reference_chain(list(ae1.encoder), x[idx])
        ref_rec = TO.reference_chain(list(ae1.decoder), ref_code)
        torch.backends.cuda.matmul.allow_tf32 = prev
        rec = chain1()
        cos_code = torch.nn.functional.cosine_similarity(code[idx], ref_code, dim=-1).min().item()
        cos_rec = torch.nn.functional.cosine_similarity(rec[idx], ref_rec, dim=-1).min().item()
    del rec
    # e2e: maps from pinned host memory, codes + reconstruction quality back
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(M, 768).pin_memory()
        xh.copy_(x)
        code_h = torch.empty(M, 15).pin_memory()
        xd = torch.empty_like(x)

        def step_e2e():
            xd.copy_(xh, non_blocking=True)
            with torch.no_grad():
                c = ae1.encode(xd)
                r = ae1.decode(c)
                q = torch.nn.functional.cosine_similarity(r[::64], xd[::64], dim=-1).mean()
            code_h.copy_(c, non_blocking=True)
            return float(q.item())

        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = max(args.steps // 2, 2)
        for _ in range(n):
            step_e2e()
        torch.cuda.synchronize()
        ms_e = (time.perf_counter() - t0) / n * 1e3
        e2e = {"value": B / (ms_e * 1e-3), "unit": "maps/s", "h2d_bytes_per_step": M * 768 * 4, "d2h_bytes_per_step": M * 15 * 4 + 4,
               "ms_per_step": ms_e, "api": "AutoencoderMLP.encode() + .decode() on maps copied from pinned host memory"}
    f_enc, f_dec = _flops_per_row(768, ENC1), _flops_per_row(15, DEC1)
    tflop = M * (f_enc + f_dec) / 1e12
    gbytes = M * (768 * 4 + 15 * 4 + 15 * 4 + 768 * 4) / 1e9
    peak_tf, peak_bw, src = _bf16_peak()
    line = {"metric": "AE-only encode+decode maps/s @ batch 32 of 192x192x768 -> 15 -> 768 (BASELINE configs[3])",
            "value": B / (ms * 1e-3), "unit": "maps/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 first layer + bf16 inner layers, fp32 accumulate, fp32 I/O", "data": "synthetic",
            "config": {"workload": "AE only: 32 x 192x192x768 random unit-norm CLIP maps (1,179,648 rows) -> 15-dim encode -> 768-dim decode, "
                                   "1-stage AutoencoderMLP 768-384-192-96-48-24-15 / 15-24-48-96-192-384-384-768, BatchNorm folded, eval mode",
                       "rows": M, "l2_policy": "3.6 GB input and 3.6 GB output per step exceed the 126 MB L2"},
            "roofline": {"kernel": "k_ae_chain (encode) + k_ae_chain (decode)", "bound": "tensor", "achieved": tflop / (ms * 1e-3),
                         "peak": peak_tf, "unit": "TFLOP/s", "frac": tflop / (ms * 1e-3) / peak_tf, "traffic": None, "peak_source": src,
                         "tflop_per_step": tflop, "mandatory_io_GB": gbytes, "io_GBps": gbytes / (ms * 1e-3),
                         "io_frac_of_hbm_peak": gbytes / (ms * 1e-3) / peak_bw,
                         "encode_ms": ms_enc, "decode_ms": ms_dec,
                         "encode_TFLOPs": M * f_enc / 1e12 / (ms_enc * 1e-3), "decode_TFLOPs": M * f_dec / 1e12 / (ms_dec * 1e-3),
                         "encode_GBps": M * (768 * 4 + 60) / 1e9 / (ms_enc * 1e-3), "decode_GBps": M * (768 * 4 + 60) / 1e9 / (ms_dec * 1e-3)},
            "two_stage": {"ms_per_step": ms2, "maps_per_s": B / (ms2 * 1e-3),
                          "chain": "general 768-512-256-128-64-32 -> online 32-24-15 -> online 15-24-32 -> general 32-192-256-384-512-768"},
            "parity": {"cos_min_code_vs_torch_fp32": cos_code, "cos_min_reconstruction_vs_torch_fp32": cos_rec, "rows_checked": int(idx.numel())},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": args.steps * 2, "clocks": clocks}
    emit(line)


def config5(args, emit, peaks, ClockSampler):
    raise SystemExit("config 5 is implemented in a later commit of this round")
