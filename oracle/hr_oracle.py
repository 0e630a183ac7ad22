"""CPU restatement of the reference's HR module forward (TEST INFRASTRUCTURE ONLY -- the product never imports this).

Follows language/supervisedNet.py: AttentionFusion.forward (:30-43) and HighResLanguageFeatureNet.forward (:83-109),
eval-mode BatchNorm (running statistics), fp32 on the CPU through torch.nn.functional.  Parameters come as a flat
``state_dict`` with the reference's key names (``initial_conv.0.weight`` ...).  Pinned against outputs of the real
reference class by tests/golden/make_golden_hr.py -> tests/golden/hr_small.npz (tests/test_hr.py).
"""
import torch
import torch.nn.functional as F


def _bn(sd, prefix, x, eps=1e-5):
    # nn.BatchNorm2d in eval mode: (x - running_mean) / sqrt(running_var + eps) * weight + bias
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=eps)


def _conv(sd, prefix, x, padding):
    return F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], padding=padding)


def _up(sd, prefix, x):
    # nn.ConvTranspose2d(kernel_size=4, stride=2, padding=1) + BatchNorm2d + ReLU (supervisedNet.py:58-62,67-71,76-80)
    x = F.conv_transpose2d(x, sd[prefix + ".0.weight"], sd[prefix + ".0.bias"], stride=2, padding=1)
    return F.relu(_bn(sd, prefix + ".1", x))


def attention_fusion(sd, prefix, high, low):
    # supervisedNet.py:30-43
    low = _conv(sd, prefix + ".low_res_align", low, 0)
    fused = torch.cat([high, low], dim=1)
    fused = F.relu(_bn(sd, prefix + ".fusion.1", _conv(sd, prefix + ".fusion.0", fused, 1)))
    att = F.relu(_bn(sd, prefix + ".attention.1", _conv(sd, prefix + ".attention.0", fused, 1)))
    att = torch.sigmoid(_conv(sd, prefix + ".attention.3", att, 0))
    return fused * att + fused


def hr_forward(sd, fv, f3, f2, return_intermediates=False):
    """supervisedNet.py:83-109.  fv [N,768,S,S], f3 [N,384,.,.], f2 [N,192,.,.] -> [N,768,8S,8S]."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    fv, f3, f2 = fv.float().cpu(), f3.float().cpu(), f2.float().cpu()
    inter = {}
    x = F.relu(_bn(sd, "initial_conv.1", _conv(sd, "initial_conv.0", fv, 1)))
    inter[0] = x
    x = _up(sd, "upsample1", x)
    inter[1] = x
    f3r = F.interpolate(f3, size=(x.size(2), x.size(3)), mode="bilinear", align_corners=False)
    x = attention_fusion(sd, "attention_fusion1", x, f3r)
    inter[5] = x
    x = _up(sd, "upsample2", x)
    inter[6] = x
    f2r = F.interpolate(f2, size=(x.size(2), x.size(3)), mode="bilinear", align_corners=False)
    x = attention_fusion(sd, "attention_fusion2", x, f2r)
    inter[10] = x
    x = _up(sd, "upsample3", x)
    inter[11] = x
    x = _conv(sd, "final_conv", x, 0)
    return (x, inter) if return_intermediates else x


def seeded_state_dict(shapes, seed):
    """Deterministic parameters for a list of (key, shape): conv weights ~ N(0, 1/fan_in-ish) so activations stay O(1),
    BatchNorm statistics away from the identity.  Used by the golden generator and by the tests (same values)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in shapes:
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros(shape, dtype=torch.long)
        elif key.endswith("running_var"):
            sd[key] = torch.empty(shape).uniform_(0.5, 1.5, generator=g)
        elif key.endswith("running_mean"):
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 4:
            # Conv2d [Cout,Cin,kh,kw] / ConvTranspose2d [Cin,Cout,4,4] (stride 2: 4 of the 16 taps reach one output)
            transposed = ".0.weight" in key and key.startswith("upsample")
            fan_in = shape[0] * 4 if transposed else shape[1] * shape[2] * shape[3]
            sd[key] = torch.randn(shape, generator=g) * (1.4 / fan_in ** 0.5)
        elif key.endswith(".1.weight"):  # BatchNorm gamma
            sd[key] = torch.empty(shape).uniform_(0.8, 1.2, generator=g)
        else:  # biases, BatchNorm beta
            sd[key] = 0.1 * torch.randn(shape, generator=g)
    return sd
