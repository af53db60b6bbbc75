"""CPU oracle for the full-resolution conv stages of ADNM-UNet (TEST INFRASTRUCTURE - never on the product path).

Plain-PyTorch restatement, differentiable by autograd (run it in float64), of
  * `Conv2dLayer` (models/model_untils.py:71-93) as a dense 3x3 / 1x1 convolution over token-major (B, L, C) activations,
  * `WTConvLayer` (:96-116): WTConv2d, `scale * InstanceNorm2d(x) + shift`, optional GELU,
  * `Mlp` (:52-68): fc1, GELU, fc2 (act2 is declared, never applied),
  * `WTLayer.forward` (:402-426) incl. its quirk: with `residual` the concat of `features` is built and thrown away (:407-408),
    so gama3 / gama4 receive no gradient on that path,
  * `PatchEmbed.forward` (:296-314) and `OutProj.forward` (:866-892).
The restatement is token-major end to end (the layout the CUDA path keeps); the reference permutes to NCHW and back.
Parity pinning: the reference ships no golden vectors; this file is pinned against the unmodified reference
modules run in the build container (tests/golden/make_golden.py -> tests/golden/{wtlayer,patchembed,outproj}_*.npz,
tests/test_oracle_vs_golden.py).  Parameter names are the reference modules' state_dict keys.
"""
import math

import torch
import torch.nn.functional as F

from oracle import wtconv_oracle as WO

IN_EPS = 1e-5      # nn.InstanceNorm2d default eps (models/model_untils.py:284,371,814 construct it with defaults)


def to_planes(x, H, W):
    B, L, C = x.shape
    return x.reshape(B, H, W, C).permute(0, 3, 1, 2)


def to_tokens(x):
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, H * W, C)


def conv_tokens(x, H, W, weight, bias=None, gamma=None):
    """nn.Conv2d(k = 3, padding 1 | k = 1) on token-major x, with the optional `x.mul(gamma)` that precedes it (:420-421,:882-883)."""
    if gamma is not None:
        x = x * gamma
    pad = weight.shape[-1] // 2
    return to_tokens(F.conv2d(to_planes(x, H, W), weight, bias, padding=pad))


def instance_norm(y):
    """nn.InstanceNorm2d(C): per (sample, channel) plane, biased variance, no affine."""
    m = y.mean(dim=(2, 3), keepdim=True)
    v = y.var(dim=(2, 3), keepdim=True, unbiased=False)
    return (y - m) * torch.rsqrt(v + IN_EPS)


def sub(p, prefix):
    return {k[len(prefix):]: v for k, v in p.items() if k.startswith(prefix)}


def wtconvlayer(p, prefix, xs, levels, norm, act):
    """WTConvLayer.forward (:108-116) on planes xs (B, C, H, W)."""
    y = WO.wtconv_forward(sub(p, prefix + "conv."), xs, levels)
    if norm:
        y = p[prefix + "scale"] * instance_norm(y) + p[prefix + "shift"]
    if act:
        y = F.gelu(y)
    return y


def plane_mix(y, xs, alpha, beta, scale=None, shift=None, gamma=None, norm=False, act=False):
    """tokens = gamma * (alpha * act(scale * IN(y) + shift) + beta * xs): the op the CUDA path fuses (adn_plane_mix_*)."""
    u = y
    if norm:
        u = scale * instance_norm(u) + shift
    if act:
        u = F.gelu(u)
    t = alpha * u + beta * xs
    if gamma is not None:
        t = t * gamma.view(1, -1, 1, 1)
    return to_tokens(t)


def mlp(p, prefix, x):
    h = F.gelu(x @ p[prefix + "fc1.weight"].t() + p[prefix + "fc1.bias"])
    return h @ p[prefix + "fc2.weight"].t() + p[prefix + "fc2.bias"]


def wtlayer_forward(p, x, levels, residual=None, features=None):
    """models/model_untils.py:402-426; x / residual / features token-major (B, L, C)."""
    if residual is not None:
        x = torch.cat((p["gama1"] * x, p["gama2"] * residual), dim=-1)
    elif features is not None:
        x = x + p["gama3"] * features
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    xs = to_planes(x, H, W)
    t = p["alpha"] * wtconvlayer(p, "wtconv.", xs, levels, norm=True, act=False) + p["beta"] * xs
    t = mlp(p, "mlp.", to_tokens(t))
    return F.gelu(conv_tokens(t, H, W, p["conv.conv.weight"], p["conv.conv.bias"], p.get("gamma")))


def patchembed_forward(p, x, levels):
    """models/model_untils.py:296-314; returns (tokens (B, L, embed_dim), res (B, H, W) = the last input frame)."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    xs = to_planes(x, H, W)
    res = xs[:, -1]
    t = p["alpha1"] * wtconvlayer(p, "conv1.0.", xs, levels, norm=False, act=True) + p["beta1"] * xs
    s = F.gelu(F.conv2d(t, p["conv2.0.conv.weight"], None, padding=1))
    t = p["alpha2"] * wtconvlayer(p, "conv3.0.", s, levels, norm=True, act=False) + p["beta2"] * s
    if "gamma" in p:
        t = t * p["gamma"].view(1, -1, 1, 1)
    return to_tokens(t), res


def outproj_forward(p, x, residual, H, W):
    """models/model_untils.py:866-892; x (B, L, C) tokens, residual (B, H, W) or None -> (B, frames, H, W).  wt_levels is 3 (:811)."""
    xs = to_planes(x, H, W)
    t = p["alpha"] * wtconvlayer(p, "wtconv.", xs, 3, norm=True, act=True) + p["beta"] * xs
    if "gamma" in p:
        t = t * p["gamma"].view(1, -1, 1, 1)
    t = F.gelu(F.conv2d(t, p["conv.0.conv.weight"], None, padding=1))
    t = F.gelu(F.conv2d(t, p["conv.1.conv.weight"], None))
    if residual is not None:
        t = p["alpha1"] * t + p["alpha2"] * residual.unsqueeze(1)
    t = F.conv2d(t, p["conv2.conv.weight"], None, padding=1)
    return t * torch.sigmoid(p["conv2.act.beta"] * t)                  # Swish (:162-169)


def perturb_params(module_state, seed, scale=0.1):
    """float32-representable perturbed copies of a module's state_dict (the frozen Haar filters set to their exact db1 values
    +-0.5: a float32 1/sqrt(2) squared, as pywt-derived filters are built, is 0.49999997), so that scalar gates,
    layer scales and biases are away from their 1 / 0 initial values and every term of the backward is exercised:
    0-dim / 1-dim tensors get N(0, scale^2) added, weight matrices / kernels N(0, (0.3 mean|w|)^2)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in module_state.items():
        if k.endswith("wt_filter") or k.endswith("iwt_filter"):
            out[k] = WO.haar_filters(v.shape[0] // 4, torch.float32)[0]
            continue
        sd = scale if v.dim() < 2 else 0.3 * float(v.double().abs().mean().clamp_min(1e-3))
        out[k] = (v.double() + sd * torch.randn(v.shape, generator=g, dtype=torch.float64)).float()
    return out


def gconv4_tokens(x, H, W, weight, bias=None):
    """nn.Conv2d(C, C, (kh, kw), padding (kh // 2, kw // 2), groups = C / 4) of the EncoderToDecoder bridges
    (models/model_untils.py:621-675) on token-major x (B, L, C)."""
    kh, kw = weight.shape[2], weight.shape[3]
    return to_tokens(F.conv2d(to_planes(x, H, W), weight, bias, padding=(kh // 2, kw // 2), groups=weight.shape[0] // 4))
