"""CPU oracle for the ADN-SSD token mixer (TEST INFRASTRUCTURE - never on the product path).

A closed-form restatement, in plain PyTorch on the CPU, of `Mamba2.forward` from the
reference (models/ADNssd.py:302-462 with `non_casual_linear_attn` :252-299), plus an EXPLICIT
backward (no autograd) whose stage split is the one the CUDA kernels use.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it.

Parity pinning: the reference ships no golden vectors or tests, so the oracle is
pinned against outputs of the reference itself run in the build container
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`; `tests/test_oracle_vs_golden.py`).

Closed form, channel order of `in_proj` output = [z | x | B | C | dt]:
  raw  = u @ W_in^T                                           (models/ADNssd.py:309,315-317)
  act  = SiLU(dwconv3x3(raw[..., :2Di+2GN]; K))               (:329-372 and :388-390 collapse to one
         per-channel 3x3 kernel K: conv2d.weight[c/2] for even xBC channels, the outer product of
         the matching (3x1, 1x3) pair for odd ones, conv2d_z.weight for z)
  w    = softplus(raw_dt + dt_bias) * exp(A_log)              (:318,:310,:267-270 - the sign cancels)
  S'[b,j,c] = [j%2==c%2] * sum_l Bc[b,l,j] * w[b,l,hd(c)] * xc[b,l,c]      (:280, both parities)
  y    = Cc @ S' + D[hd(c)] * xc                              (:281-283,:406-411)
  out  = alpha1 * ( LN(y) @ W_out[:, :Di]^T + zc @ W_out[:, Di:]^T )        (:456-461)
with hd(c) = 2*((c//2)//headdim) + c%2.
"""
import math

import torch
import torch.nn.functional as F

PARAM_NAMES = (
    "dt_bias", "A_log", "D", "scale", "shift", "alpha1", "alpha2", "in_proj.weight",
    "conv_13_x1.weight", "conv_31_x1.weight", "conv_13_x2.weight", "conv_31_x2.weight",
    "conv_13_bc1.weight", "conv_31_bc1.weight", "conv_13_bc2.weight", "conv_31_bc2.weight",
    "conv2d.weight", "norm.weight", "norm.bias", "conv2d_z.weight", "out_proj.weight",
)
UNUSED_PARAMS = ("scale", "shift", "alpha2")  # models/ADNssd.py:227-228,246: declared, never read


def dims(d_model, headdim, d_state, ngroups=2, expand=2):
    Di = expand * d_model
    GN = ngroups * d_state
    nh = Di // headdim
    return dict(D=d_model, Di=Di, P=headdim, N=d_state, G=ngroups, GN=GN, nh=nh,
                Wd=Di + 2 * GN, CC=2 * Di + 2 * GN, dip=2 * Di + 2 * GN + nh)


def head_of_channel(Di, P):
    c = torch.arange(Di)
    return 2 * ((c // 2) // P) + (c % 2)


def parity_mask(GN, Di, dtype):
    j = torch.arange(GN)[:, None]
    c = torch.arange(Di)[None, :]
    return ((j % 2) == (c % 2)).to(dtype)


def init_params(d_model, headdim, d_state, seed=0, perturb=0.0, dtype=torch.float32):
    """Parameters with the reference's shapes (models/ADNssd.py:100-248).  `perturb` adds N(0,perturb^2)
    to every tensor so that no term of the parity check vanishes."""
    d = dims(d_model, headdim, d_state)
    g = torch.Generator().manual_seed(seed)
    Di, GN, nh = d["Di"], d["GN"], d["nh"]

    def U(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound

    dt = torch.exp(torch.rand(nh, generator=g, dtype=torch.float64) * (math.log(0.1) - math.log(0.001)) + math.log(0.001)).clamp(min=1e-4)
    p = {
        "dt_bias": dt + torch.log(-torch.expm1(-dt)),
        "A_log": torch.log(1 + 15 * torch.rand(nh, generator=g, dtype=torch.float64)),
        "D": torch.ones(nh, dtype=torch.float64),
        "scale": torch.tensor(1.0, dtype=torch.float64), "shift": torch.tensor(0.0, dtype=torch.float64),
        "alpha1": torch.tensor(1.0, dtype=torch.float64), "alpha2": torch.tensor(1.0, dtype=torch.float64),
        "in_proj.weight": torch.randn(d["dip"], d_model, generator=g, dtype=torch.float64).clamp(-2, 2) * 0.02,
        "conv2d.weight": U((d["Wd"] // 2, 1, 3, 3), 1 / 3), "conv2d_z.weight": U((Di, 1, 3, 3), 1 / 3),
        "norm.weight": torch.ones(Di, dtype=torch.float64), "norm.bias": torch.zeros(Di, dtype=torch.float64),
        "out_proj.weight": U((d_model, 2 * Di), 1 / math.sqrt(2 * Di) / math.sqrt(3)),
    }
    for tag, n in (("x1", Di // 4), ("x2", Di // 4), ("bc1", GN // 2), ("bc2", GN // 2)):
        p[f"conv_13_{tag}.weight"] = U((n, 1, 1, 3), 1 / math.sqrt(3))
        p[f"conv_31_{tag}.weight"] = U((n, 1, 3, 1), 1 / math.sqrt(3))
    if perturb:
        for k in p:
            p[k] = p[k] + perturb * torch.randn(p[k].shape, generator=g, dtype=torch.float64)
    return {k: p[k].to(dtype) for k in PARAM_NAMES}


def assemble_conv_kernels(p, d):
    """(CC,3,3) per-channel kernels in in_proj column order [z | x | B | C]; models/ADNssd.py:329-364,388-390."""
    Di, GN, Wd = d["Di"], d["GN"], d["Wd"]
    K = p["conv2d.weight"].new_zeros(d["CC"], 3, 3)
    K[:Di] = p["conv2d_z.weight"][:, 0]
    Kx = K[Di:]  # view over the Wd xBC channels
    Kx[0::2] = p["conv2d.weight"][:, 0]
    for off, tag in ((1, "1"), (3, "2")):
        w31 = torch.cat([p[f"conv_31_x{tag}.weight"], p[f"conv_31_bc{tag}.weight"]], 0)[:, 0, :, 0]  # (Wd/4,3) over rows
        w13 = torch.cat([p[f"conv_13_x{tag}.weight"], p[f"conv_13_bc{tag}.weight"]], 0)[:, 0, 0, :]  # (Wd/4,3) over cols
        Kx[off::4] = w31[:, :, None] * w13[:, None, :]
    return K


def scatter_conv_kernel_grads(p, d, dK):
    """dK (CC,3,3) -> gradients of the ten conv weight tensors (rank-1 chain rule for the 3x1/1x3 pairs)."""
    Di = d["Di"]
    g = {"conv2d_z.weight": dK[:Di, None].clone(), "conv2d.weight": dK[Di:][0::2, None].clone()}
    for off, tag in ((1, "1"), (3, "2")):
        w31 = torch.cat([p[f"conv_31_x{tag}.weight"], p[f"conv_31_bc{tag}.weight"]], 0)[:, 0, :, 0]
        w13 = torch.cat([p[f"conv_13_x{tag}.weight"], p[f"conv_13_bc{tag}.weight"]], 0)[:, 0, 0, :]
        dk = dK[Di:][off::4]
        d31 = (dk * w13[:, None, :]).sum(2)
        d13 = (dk * w31[:, :, None]).sum(1)
        nx = Di // 4
        g[f"conv_31_x{tag}.weight"] = d31[:nx, None, :, None].clone()
        g[f"conv_31_bc{tag}.weight"] = d31[nx:, None, :, None].clone()
        g[f"conv_13_x{tag}.weight"] = d13[:nx, None, None, :].clone()
        g[f"conv_13_bc{tag}.weight"] = d13[nx:, None, None, :].clone()
    return g


def _dwconv3x3(x_blc, K, H, W):
    B, L, C = x_blc.shape
    x = x_blc.reshape(B, H, W, C).permute(0, 3, 1, 2)
    y = F.conv2d(x, K[:, None], padding=1, groups=C)
    return y.permute(0, 2, 3, 1).reshape(B, L, C)


def mixer_forward(p, u, H, W, headdim, d_state, return_saved=False):
    B, L, Dm = u.shape
    assert L == H * W
    d = dims(Dm, headdim, d_state)
    Di, GN, CC = d["Di"], d["GN"], d["CC"]
    assert Di % 4 == 0 and GN % 4 == 0
    hd = head_of_channel(Di, headdim)
    raw = u @ p["in_proj.weight"].t()
    K = assemble_conv_kernels(p, d)
    pre = _dwconv3x3(raw[..., :CC], K, H, W)
    act = F.silu(pre)
    zc, xc, Bc, Cc = act[..., :Di], act[..., Di:2 * Di], act[..., 2 * Di:2 * Di + GN], act[..., 2 * Di + GN:]
    w = F.softplus(raw[..., CC:] + p["dt_bias"]) * torch.exp(p["A_log"])  # (B,L,nh)
    wc = w[..., hd]
    M = parity_mask(GN, Di, u.dtype)
    S = M * torch.einsum("blj,blc->bjc", Bc, wc * xc)
    y = torch.einsum("blj,bjc->blc", Cc, S) + p["D"][hd] * xc
    mu = y.mean(-1, keepdim=True)
    rstd = torch.rsqrt(y.var(-1, keepdim=True, unbiased=False) + 1e-5)
    yhat = (y - mu) * rstd
    yn = yhat * p["norm.weight"] + p["norm.bias"]
    Wo = p["out_proj.weight"]
    out = p["alpha1"] * (yn @ Wo[:, :Di].t() + zc @ Wo[:, Di:].t())
    if return_saved:
        return out, dict(raw=raw, pre=pre, act=act, S=S, y=y, yhat=yhat, rstd=rstd, yn=yn, w=w, K=K)
    return out


def mixer_backward(p, u, H, W, headdim, d_state, dout):
    """Explicit backward.  Stage split = the CUDA kernels' (phase B1: dout -> dy, dS', dCc, dzc;
    phase B2: dS' -> dxc, dBc, ddt; conv backward; in_proj backward)."""
    B, L, Dm = u.shape
    d = dims(Dm, headdim, d_state)
    Di, GN, CC, nh, P = d["Di"], d["GN"], d["CC"], d["nh"], headdim
    hd = head_of_channel(Di, P)
    _, sv = mixer_forward(p, u, H, W, headdim, d_state, return_saved=True)
    raw, pre, act, S, yhat, rstd, yn, w, K = (sv[k] for k in ("raw", "pre", "act", "S", "yhat", "rstd", "yn", "w", "K"))
    zc, xc, Bc, Cc = act[..., :Di], act[..., Di:2 * Di], act[..., 2 * Di:2 * Di + GN], act[..., 2 * Di + GN:]
    Wo = p["out_proj.weight"]
    a1 = p["alpha1"]
    grads = {}
    # ---- phase B1
    cat = torch.cat([yn, zc], -1)
    g = dout @ Wo                                   # (B,L,2Di)  d(out/alpha1)/d cat
    grads["alpha1"] = (g * cat).sum()
    grads["out_proj.weight"] = a1 * torch.einsum("bld,blj->dj", dout, cat)
    dyn, dzc = a1 * g[..., :Di], a1 * g[..., Di:]
    grads["norm.weight"] = (dyn * yhat).sum((0, 1))
    grads["norm.bias"] = dyn.sum((0, 1))
    dyh = dyn * p["norm.weight"]
    dy = rstd * (dyh - dyh.mean(-1, keepdim=True) - yhat * (dyh * yhat).mean(-1, keepdim=True))
    M = parity_mask(GN, Di, u.dtype)
    dS = M * torch.einsum("blj,blc->bjc", Cc, dy)
    dCc = torch.einsum("blc,bjc->blj", dy, S)
    grads["D"] = torch.zeros(nh, dtype=u.dtype).index_add_(0, hd, (dy * xc).sum((0, 1)))
    # ---- phase B2
    wc = w[..., hd]
    G = torch.einsum("blj,bjc->blc", Bc, dS)
    dxc = p["D"][hd] * dy + wc * G
    dBc = torch.einsum("blc,bjc->blj", wc * xc, dS)
    dw = torch.zeros(B, L, nh, dtype=u.dtype).index_add_(2, hd, xc * G)
    expA = torch.exp(p["A_log"])
    grads["A_log"] = (dw * w).sum((0, 1))
    ddt = dw * expA * torch.sigmoid(raw[..., CC:] + p["dt_bias"])
    grads["dt_bias"] = ddt.sum((0, 1))
    # ---- conv backward
    dact = torch.cat([dzc, dxc, dBc, dCc], -1)
    sg = torch.sigmoid(pre)
    dpre = dact * sg * (1 + pre * (1 - sg))
    dpre_img = dpre.reshape(B, H, W, CC).permute(0, 3, 1, 2)
    draw_c = F.conv_transpose2d(dpre_img, K[:, None], padding=1, groups=CC).permute(0, 2, 3, 1).reshape(B, L, CC)
    raw_pad = F.pad(raw[..., :CC].reshape(B, H, W, CC), (0, 0, 1, 1, 1, 1))
    dK = torch.stack([torch.stack([(dpre.reshape(B, H, W, CC) * raw_pad[:, a:a + H, b:b + W]).sum((0, 1, 2))
                                   for b in range(3)], -1) for a in range(3)], -2)  # (CC,3,3)
    grads.update(scatter_conv_kernel_grads(p, d, dK))
    # ---- in_proj backward
    draw = torch.cat([draw_c, ddt], -1)
    grads["in_proj.weight"] = torch.einsum("blj,bld->jd", draw, u)
    du = draw @ p["in_proj.weight"]
    return du, grads
