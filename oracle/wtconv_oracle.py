"""CPU oracle for the Haar-wavelet WTConv2d (TEST INFRASTRUCTURE - never on the product path).

Restates models/WTConv2d.py:100-153 of the reference with the db1 analysis/synthesis filters
(:9-29) written out as 2x2 butterflies instead of grouped (transposed) convolutions:
  quad (a b; c d) at rows 2i,2i+1 / cols 2j,2j+1
  LL = (a+b+c+d)/2   b1 = (a+b-c-d)/2   b2 = (a-b+c-d)/2   b3 = (a-b-c+d)/2        (:31-40)
  a = (LL+b1+b2+b3)/2  b = (LL+b1-b2-b3)/2  c = (LL-b1+b2-b3)/2  d = (LL-b1-b2+b3)/2  (:42-51)
Sub-band channel order after the DWT is c*4+band (:39,:122).  Backward is autograd through this
restatement (it is the checker, not the product).  Pinned by tests/golden/wtconv_*.npz, generated
from the reference itself by tests/golden/make_golden.py.
"""
import torch
import torch.nn.functional as F


def param_names(levels, bias=True):
    names = ["wt_filter", "iwt_filter", "base_conv.weight"] + (["base_conv.bias"] if bias else []) + ["base_scale.weight"]
    names += [f"wavelet_convs.{i}.weight" for i in range(levels)]        # models/WTConv2d.py:85-91 (ModuleList order)
    names += [f"wavelet_scale.{i}.weight" for i in range(levels)]
    return names


def haar_filters(C, dtype=torch.float32):
    """The frozen `wt_filter` / `iwt_filter` Parameters (4C,1,2,2) - identical for db1."""
    f = 0.5 * torch.tensor([[[1, 1], [1, 1]], [[1, 1], [-1, -1]], [[1, -1], [1, -1]], [[1, -1], [-1, 1]]], dtype=dtype)
    f = f[:, None].repeat(C, 1, 1, 1)
    return f, f.clone()


def init_params(C, k, levels, bias=True, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    wt, iwt = haar_filters(C, torch.float64)
    p = {"wt_filter": wt, "iwt_filter": iwt,
         "base_conv.weight": (torch.rand(C, 1, k, k, generator=g, dtype=torch.float64) * 2 - 1) / k,
         "base_scale.weight": 1 + 0.2 * torch.randn(1, C, 1, 1, generator=g, dtype=torch.float64)}
    if bias:
        p["base_conv.bias"] = 0.1 * torch.randn(C, generator=g, dtype=torch.float64)
    for i in range(levels):
        p[f"wavelet_convs.{i}.weight"] = (torch.rand(4 * C, 1, k, k, generator=g, dtype=torch.float64) * 2 - 1) / k
        p[f"wavelet_scale.{i}.weight"] = 0.1 + 0.05 * torch.randn(1, 4 * C, 1, 1, generator=g, dtype=torch.float64)
    return {n: p[n].to(dtype) for n in param_names(levels, bias)}


def haar_dwt(x):
    """(B,C,H,W), H and W even -> (B,C,4,H/2,W/2)."""
    a, b = x[..., 0::2, 0::2], x[..., 0::2, 1::2]
    c, d = x[..., 1::2, 0::2], x[..., 1::2, 1::2]
    return torch.stack([(a + b + c + d), (a + b - c - d), (a - b + c - d), (a - b - c + d)], 2) * 0.5


def haar_idwt(s):
    """(B,C,4,h,w) -> (B,C,2h,2w)."""
    ll, b1, b2, b3 = s[:, :, 0], s[:, :, 1], s[:, :, 2], s[:, :, 3]
    B, C, h, w = ll.shape
    out = s.new_empty(B, C, 2 * h, 2 * w)
    out[..., 0::2, 0::2] = (ll + b1 + b2 + b3) * 0.5
    out[..., 0::2, 1::2] = (ll + b1 - b2 - b3) * 0.5
    out[..., 1::2, 0::2] = (ll - b1 + b2 - b3) * 0.5
    out[..., 1::2, 1::2] = (ll - b1 - b2 + b3) * 0.5
    return out


def wtconv_forward(p, x, levels):
    B, C, H, W = x.shape
    k = p["base_conv.weight"].shape[-1]
    tags, shapes = [], []
    ll = x
    for i in range(levels):
        shapes.append(ll.shape)
        if ll.shape[2] % 2 or ll.shape[3] % 2:                      # models/WTConv2d.py:114-116
            ll = F.pad(ll, (0, ll.shape[3] % 2, 0, ll.shape[2] % 2))
        sub = haar_dwt(ll)
        ll = sub[:, :, 0]
        h, w = sub.shape[-2:]
        t = F.conv2d(sub.reshape(B, 4 * C, h, w), p[f"wavelet_convs.{i}.weight"], padding=k // 2, groups=4 * C)
        tags.append((p[f"wavelet_scale.{i}.weight"] * t).reshape(B, C, 4, h, w))
    nxt = 0
    for i in range(levels - 1, -1, -1):                             # models/WTConv2d.py:131-141
        t = tags[i]
        t = torch.cat([(t[:, :, 0] + nxt).unsqueeze(2), t[:, :, 1:4]], 2)
        nxt = haar_idwt(t)[:, :, :shapes[i][2], :shapes[i][3]]
    base = F.conv2d(x, p["base_conv.weight"], p.get("base_conv.bias"), padding=k // 2, groups=C)
    return p["base_scale.weight"] * base + nxt                      # :146-147


def wtconv_forward_backward(p, x, levels, dy):
    """Returns out, dx, {param grads}; frozen Haar filters get no gradient (requires_grad=False, :75-76)."""
    train = [n for n in p if n not in ("wt_filter", "iwt_filter")]
    q = {n: (v.detach().clone().requires_grad_(True) if n in train else v) for n, v in p.items()}
    xr = x.detach().clone().requires_grad_(True)
    out = wtconv_forward(q, xr, levels)
    gs = torch.autograd.grad(out, [xr] + [q[n] for n in train], dy)
    return out.detach(), gs[0], dict(zip(train, gs[1:]))
