"""CPU oracle for the ADNM-UNet `Block` around the mixer (TEST INFRASTRUCTURE - never on the product path).

Plain-PyTorch restatement, differentiable by autograd (run it in float64), of
  * the standalone RMSNorm of the reference README (README.md:22-30) with the Block's scalar affine
    `scale * norm(x) + shift` (models/ADNMUNet.py:149,155),
  * `FeedForward` (models/model_untils.py:172-197): 1x1 conv D -> 4D, depthwise 3x3 on 4D channels (zero padding, bias),
    `gelu(x1) * sigmoid(x2)` on the two channel halves, 1x1 conv 2D -> D - restated in token-major (B, L, C) form, the
    layout the CUDA path keeps throughout (the reference permutes to NCHW and back, models/ADNMUNet.py:158),
  * `Block.forward` (models/ADNMUNet.py:115-165) for any num_layers, including its quirks: `beta3`/`beta4` alias
    `beta1`/`beta2` (:145-146), `act`, `beta3`, `beta4` are declared and never read.
Parity pinning: the reference ships no golden vectors; this file is pinned against the unmodified reference
`Block` run in the build container (tests/golden/make_golden.py -> tests/golden/block_*.npz, tests/test_oracle_vs_golden.py).
Parameter names are the reference Block's state_dict keys.
"""
import math

import torch
import torch.nn.functional as F

from oracle import adnssd_oracle as AO

UNUSED_BLOCK_PARAMS = ("beta3", "beta4", "act.beta")     # declared (models/ADNMUNet.py:69-70,105), never read by forward


def rmsnorm_affine(x, weight, scale=None, shift=None, eps=1e-5):
    """README.md:22-30, then models/ADNMUNet.py:149 (scale / shift None: the bare module)."""
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight
    if scale is not None:
        y = scale * y
    if shift is not None:
        y = y + shift
    return y


def ffn_forward(p, x, H, W, prefix="ffns.0."):
    """models/model_untils.py:190-196 on token-major x (B, L, D)."""
    B, L, D = x.shape
    w_in, b_in = p[prefix + "project_in.conv.weight"], p[prefix + "project_in.conv.bias"]
    w_dw, b_dw = p[prefix + "dwconv.conv.weight"], p[prefix + "dwconv.conv.bias"]
    w_out, b_out = p[prefix + "project_out.conv.weight"], p[prefix + "project_out.conv.bias"]
    C4 = w_in.shape[0]
    h = x @ w_in.reshape(C4, D).t() + b_in                                  # project_in (1x1 conv)
    h = h.reshape(B, H, W, C4).permute(0, 3, 1, 2)
    h = F.conv2d(h, w_dw, b_dw, padding=1, groups=C4)                       # dwconv
    x1, x2 = h.chunk(2, dim=1)
    g = F.gelu(x1) * torch.sigmoid(x2)
    g = g.permute(0, 2, 3, 1).reshape(B, L, C4 // 2)
    return g @ w_out.reshape(D, C4 // 2).t() + b_out                        # project_out


def block_forward(p, x, H, W, headdim, d_state, num_layers=1, residual=None, features=None, norm_eps=1e-6):
    """models/ADNMUNet.py:115-165.  `p` maps the Block's state_dict keys to tensors."""
    if residual is not None:
        x = torch.cat((p["alpha1"] * x, p["alpha2"] * residual), dim=-1)
        if features is not None:
            x = x + torch.cat((p["alpha3"] * features, p["alpha4"] * features), dim=-1)
    elif features is not None:
        x = x + p["alpha3"] * features
    for i in range(num_layers):
        mp = {k[len(f"mixer_layers.{i}."):]: v for k, v in p.items() if k.startswith(f"mixer_layers.{i}.")}
        b1, b2 = p["beta1"][i], p["beta2"][i]
        xn = rmsnorm_affine(x, p[f"norm1_layers.{i}.weight"], p[f"scale1.{i}"], p[f"shift1.{i}"], norm_eps)
        x = b1 * x + b2 * AO.mixer_forward(mp, xn, H, W, headdim, d_state)
        xn = rmsnorm_affine(x, p[f"norm2_layers.{i}.weight"], p[f"scale2.{i}"], p[f"shift2.{i}"], norm_eps)
        x = b1 * x + b2 * ffn_forward(p, xn, H, W, prefix=f"ffns.{i}.")      # beta3 = beta1, beta4 = beta2 (:145-146)
    x = x * p["gamma"]
    if "out_proj.weight" in p:
        x = x @ p["out_proj.weight"].t() + p["out_proj.bias"]
    return x


def init_block_params(dim, out_dim, headdim=4, d_state=16, num_layers=1, seed=0, perturb=0.1, dtype=torch.float64):
    """Random parameters with the reference Block's shapes (models/ADNMUNet.py:49-113); every tensor perturbed so that no
    term of a parity check vanishes (scalars around 1, biases non-zero)."""
    g = torch.Generator().manual_seed(seed)

    def N(shape, std):
        return torch.randn(shape, generator=g, dtype=torch.float64) * std

    p = {}
    for k in ("alpha1", "alpha2", "alpha3", "alpha4"):
        p[k] = 1 + N((), perturb)
    for k in ("beta1", "beta2", "beta3", "beta4"):
        p[k] = 1 + N((num_layers,), perturb)
    hid2 = 4 * dim
    for i in range(num_layers):
        mp = AO.init_params(dim, headdim, d_state, seed=seed + 17 * (i + 1), perturb=perturb, dtype=torch.float64)
        for k, v in mp.items():
            p[f"mixer_layers.{i}.{k}"] = v
        p[f"norm1_layers.{i}.weight"] = 1 + N((dim,), perturb)
        p[f"norm2_layers.{i}.weight"] = 1 + N((dim,), perturb)
        pre = f"ffns.{i}."
        p[pre + "project_in.conv.weight"] = N((hid2, dim, 1, 1), 1 / math.sqrt(dim))
        p[pre + "project_in.conv.bias"] = N((hid2,), 0.2)
        p[pre + "dwconv.conv.weight"] = N((hid2, 1, 3, 3), 1 / 3)
        p[pre + "dwconv.conv.bias"] = N((hid2,), 0.2)
        p[pre + "project_out.conv.weight"] = N((dim, hid2 // 2, 1, 1), 1 / math.sqrt(hid2 // 2))
        p[pre + "project_out.conv.bias"] = N((dim,), 0.2)
        p[f"scale1.{i}"], p[f"shift1.{i}"] = 1 + N((), perturb), N((), perturb)
        p[f"scale2.{i}"], p[f"shift2.{i}"] = 1 + N((), perturb), N((), perturb)
    p["act.beta"] = torch.tensor(1.0, dtype=torch.float64)
    if dim != out_dim:
        p["out_proj.weight"] = N((out_dim, dim), 1 / math.sqrt(dim))
        p["out_proj.bias"] = N((out_dim,), 0.2)
    p["gamma"] = 1 + N((dim,), perturb)
    return {k: v.to(dtype) for k, v in p.items()}
