"""HR module: oracle pinned to the real reference class, then the tcgen05 path against the oracle.

Tolerance: the CUDA path keeps activations in bf16 between the 13 convolutions (fp32 accumulation); measured against
the fp32 oracle the error of the final map is ~0.65 % of its RMS (0.3 % after the first convolution, growing with
depth) and the worst single element of the ~19 M outputs is off by 4-7 % of the RMS.  The tests allow a relative RMS
error of 1.5e-2 and a maximum absolute error of 12 % of the RMS; a wrong tap, class or channel gives errors of order 100 %.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hr_oracle  # noqa: E402
from online_lang_splatting_b200 import supervised_net as SN  # noqa: E402
from oracle import torch_oracle as TO  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "hr_small.npz")
REL_RMS_TOL = 1.5e-2
MAX_ABS_TOL = 12e-2  # of the output RMS


def golden():
    z = np.load(GOLD)
    shapes = [(str(k), tuple(int(d) for d in dims[:r])) for k, dims, r in zip(z["shape_keys"], z["shape_dims"], z["shape_rank"])]
    return z, shapes


def make_inputs(seed, s_h, s_w):
    g = torch.Generator().manual_seed(seed)
    fv = torch.randn(1, 768, s_h, s_w, generator=g)
    f3 = torch.randn(1, 384, 4 * s_h, 4 * s_w, generator=g)
    f2 = torch.randn(1, 192, 8 * s_h - 3, 8 * s_w + 5, generator=g)
    return fv, f3, f2


def test_oracle_matches_reference_golden():
    z, shapes = golden()
    sd = hr_oracle.seeded_state_dict(shapes, int(z["seed_w"]))
    fv, f3, f2 = make_inputs(int(z["seed_x"]), int(z["s_h"]), int(z["s_w"]))
    out = hr_oracle.hr_forward(sd, fv, f3, f2)
    np.testing.assert_allclose(out[0, ::16, ::8, ::8].numpy(), z["out_sample"], rtol=1e-5, atol=1e-5)
    assert abs(float(out.pow(2).mean().sqrt()) - float(z["out_rms"])) < 1e-5
    assert abs(float(out.abs().sum()) / float(z["out_abs_sum"]) - 1.0) < 1e-6


def test_state_dict_keys_match_reference():
    z, shapes = golden()
    net = SN.HighResLanguageFeatureNet()
    ours = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert ours == shapes
    # and the Lightning wrapper prefixes them with "model."
    assert [k for k in SN.LangSupervisedNet().state_dict()] == ["model." + k for k, _ in shapes]


def test_transposed_conv_parity_classes():
    """The decomposition the kernel uses: ConvTranspose2d(4,2,1) = four 2x2 convolutions, one per output parity."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 6, 5, 7, generator=g)
    w = torch.randn(6, 4, 4, 4, generator=g)
    ref = torch.nn.functional.conv_transpose2d(x, w, stride=2, padding=1)
    d = [[0, -1], [0, 1]]
    k = [[1, 3], [2, 0]]
    out = torch.zeros_like(ref)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    for py in range(2):
        for px in range(2):
            acc = 0
            for a in range(2):
                for b in range(2):
                    dy, dx = d[py][a], d[px][b]
                    shifted = xp[:, :, 1 + dy:1 + dy + 5, 1 + dx:1 + dx + 7]
                    acc = acc + torch.einsum("nchw,co->nohw", shifted, w[:, :, k[py][a], k[px][b]])
            out[:, :, py::2, px::2] = acc
    assert torch.allclose(out, ref, atol=1e-5)


def test_batchnorm_folding_matches_torch():
    """Host-side folding (supervised_net._fold_bn) of eval-mode BatchNorm into Conv2d / ConvTranspose2d weights."""
    import torch.nn as nn
    torch.manual_seed(4)
    for conv in (nn.Conv2d(6, 10, 3, padding=1), nn.ConvTranspose2d(6, 10, 4, stride=2, padding=1)):
        bn = nn.BatchNorm2d(10).eval()
        with torch.no_grad():
            bn.running_mean.normal_(0, 0.3); bn.running_var.uniform_(0.5, 1.5); bn.weight.uniform_(0.7, 1.3); bn.bias.normal_(0, 0.2)
        x = torch.randn(2, 6, 9, 7)
        w, b = SN._fold_bn(conv, bn)
        with torch.no_grad():
            ref = bn(conv(x))
            got = (torch.nn.functional.conv_transpose2d(x, w, b, stride=2, padding=1) if isinstance(conv, nn.ConvTranspose2d)
                   else torch.nn.functional.conv2d(x, w, b, padding=1))
        assert torch.allclose(got, ref, atol=1e-5)


def test_rejects_training_and_cpu():
    net = SN.HighResLanguageFeatureNet()
    fv, f3, f2 = make_inputs(0, 16, 16)
    with pytest.raises(RuntimeError):
        net.train()(fv, f3, f2)
    with pytest.raises(RuntimeError):
        net.eval()(fv, f3, f2)  # CPU tensors: no CPU path


# ------------------------------------------------------------------------------------------------------------------
def _build(seed_w, dev):
    _, shapes = golden()
    sd = hr_oracle.seeded_state_dict(shapes, seed_w)
    net = SN.HighResLanguageFeatureNet()
    net.load_state_dict(sd)
    return sd, net.eval().to(dev)


def _errors(ours, ref):
    rms = float(ref.pow(2).mean().sqrt())
    diff = ours - ref
    return float(diff.pow(2).mean().sqrt()) / rms, float(diff.abs().max()) / rms


@pytest.mark.gpu
def test_hr_matches_oracle_small():
    dev = torch.device("cuda:0")
    z, _ = golden()
    sd, net = _build(int(z["seed_w"]), dev)
    fv, f3, f2 = make_inputs(int(z["seed_x"]), int(z["s_h"]), int(z["s_w"]))
    ref, inter = hr_oracle.hr_forward(sd, fv, f3, f2, return_intermediates=True)
    with torch.no_grad():
        out = net(fv.to(dev), f3.to(dev), f2.to(dev))
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    assert out.permute(0, 2, 3, 1).is_contiguous()  # the AE's [M,768] view is free
    for which, t in inter.items():
        a = net.read_activation(which).cpu().permute(2, 0, 1)[None]
        rel, mx = _errors(a, t)
        assert rel < REL_RMS_TOL and mx < MAX_ABS_TOL, f"activation {which}: rel rms {rel:.3e}, max {mx:.3e}"
    rel, mx = _errors(out.cpu(), ref)
    assert rel < REL_RMS_TOL and mx < MAX_ABS_TOL, f"output: rel rms {rel:.3e}, max {mx:.3e}"
    # the stored sample of the REAL reference's output
    np.testing.assert_allclose(out.cpu()[0, ::16, ::8, ::8].numpy(), z["out_sample"], atol=MAX_ABS_TOL * float(z["out_rms"]))


@pytest.mark.gpu
def test_hr_full_size_feeds_autoencoder():
    """fv 24x24 -> 192x192x768 (the configuration the reference runs), then straight into the fused autoencoder."""
    from online_lang_splatting_b200 import autoencoder as AE
    dev = torch.device("cuda:0")
    sd, net = _build(77, dev)
    g = torch.Generator().manual_seed(3)
    fv = torch.randn(1, 768, 24, 24, generator=g)
    f3 = torch.randn(1, 384, 96, 96, generator=g)
    f2 = torch.randn(1, 192, 192, 192, generator=g)
    ref = hr_oracle.hr_forward(sd, fv, f3, f2)
    with torch.no_grad():
        out = net(fv.to(dev), f3.to(dev), f2.to(dev))
        out2 = net(fv.to(dev), f3.to(dev), f2.to(dev))
    assert torch.equal(out, out2)  # deterministic
    rel, mx = _errors(out.cpu(), ref)
    assert rel < REL_RMS_TOL and mx < MAX_ABS_TOL, f"rel rms {rel:.3e}, max {mx:.3e}"
    torch.manual_seed(0)
    ae = AE.AutoencoderMLP([384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]).eval().to(dev)
    with torch.no_grad():
        x = out.permute(0, 2, 3, 1).view(-1, 768)  # slam_backend.py:392-394
        code = ae.encode(x)
        code_ref = ae.encode(ref.to(dev).permute(0, 2, 3, 1).reshape(-1, 768))
    assert code.shape == (192 * 192, 15)
    assert float((code - code_ref).abs().max()) < 5e-2


@pytest.mark.gpu
def test_hr_weight_update_rebuilds_plan():
    dev = torch.device("cuda:0")
    sd, net = _build(5, dev)
    fv, f3, f2 = (t.to(dev) for t in make_inputs(1, 16, 16))
    with torch.no_grad():
        a = net(fv, f3, f2).clone()
        net.final_conv.bias.add_(1.0)
        b = net(fv, f3, f2)
    assert torch.allclose(b, a + 1.0, atol=1e-5)


@pytest.mark.gpu
def test_hr_fused_encode():
    """AutoencoderMLP.encode_hr (final_conv folded into the encoder's first Linear, bf16 features fed straight to the
    autoencoder kernel) against the unfused product path and against the fp32 oracle chain."""
    from online_lang_splatting_b200 import autoencoder as AE
    dev = torch.device("cuda:0")
    sd, net = _build(77, dev)
    g = torch.Generator().manual_seed(3)
    fv = torch.randn(1, 768, 24, 24, generator=g)
    f3 = torch.randn(1, 384, 96, 96, generator=g)
    f2 = torch.randn(1, 192, 192, 192, generator=g)
    torch.manual_seed(0)
    ae = AE.AutoencoderMLP([384, 192, 96, 48, 24, 15], [24, 48, 96, 192, 384, 384, 768]).eval().to(dev)
    with torch.no_grad():
        fused = ae.encode_hr(net, fv.to(dev), f3.to(dev), f2.to(dev))
        unfused = ae.encode(net(fv.to(dev), f3.to(dev), f2.to(dev)).permute(0, 2, 3, 1).view(-1, 768))
        ref = TO.reference_chain(list(ae.encoder), hr_oracle.hr_forward(sd, fv, f3, f2).to(dev).permute(0, 2, 3, 1).reshape(-1, 768))
    assert fused.shape == (192 * 192, 15)
    assert torch.allclose(fused.norm(dim=-1), torch.ones_like(fused[:, 0]), atol=1e-4)
    cos_u = (fused * unfused).sum(-1)
    cos_r = (fused * ref).sum(-1)
    assert cos_u.min().item() > 0.999 and cos_r.min().item() > 0.995, (cos_u.min().item(), cos_r.min().item())
    # the wrapper class the reference loads from a checkpoint works the same way
    wrap = SN.LangSupervisedNet()
    wrap.model = net
    with torch.no_grad():
        again = ae.encode_hr(wrap, fv.to(dev), f3.to(dev), f2.to(dev))
    assert torch.equal(again, fused)


@pytest.mark.gpu
def test_hr_cuda_graph_capture():
    """The whole HR forward (16 launches, side stream fork/join, programmatic dependent launch, cluster launches) can be
    captured into a CUDA graph by the caller and replays to the same bits."""
    dev = torch.device("cuda:0")
    sd, net = _build(5, dev)
    fv, f3, f2 = (t.to(dev) for t in make_inputs(1, 16, 24))
    with torch.no_grad():
        eager = net(fv, f3, f2).clone()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(fv, f3, f2)  # warm-up on the capture stream
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = net(fv, f3, f2)
        out.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager)
        fv.mul_(0.5)  # new input, same buffers
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, net(fv, f3, f2))
