"""Drop-in for the reference's `Block` (models/ADNMUNet.py:49-165) and its `FeedForward` (models/model_untils.py:172-197).

Same constructor signature, same sub-module / parameter names, shapes and registration order (a reference state_dict loads
strictly and a same-seed construction consumes the RNG stream identically), same `forward(hidden_states, residual=None,
features=None, ...)`.  The forward keeps the activations token-major (B, L, C) from end to end and runs every stage in the
sm_100a library through the C ABI (include/adnb200.h):
    RMSNorm * scale + shift   adn_rmsnorm_forward / _backward          (README.md:22-30, models/ADNMUNet.py:149,155)
    mixer                     adnssd_forward / _backward               (models/ADNssd.py:302-462)
    beta1 x + beta2 y [* gamma]   adn_residual_forward / _backward      (models/ADNMUNet.py:152,158,161)
    FeedForward               adn_ffn_forward / _backward              (models/model_untils.py:190-196; no NCHW round trip)
    out_proj Linear           adn_linear_forward / _backward           (models/ADNMUNet.py:162-163)
The only PyTorch ops left are the optional skip-connection prologue (`cat` / `alpha` scaling, models/ADNMUNet.py:124-132).
"""
import math

import torch
import torch.nn as nn

from adnm_unet_b200 import _lib
from adnm_unet_b200.mixer import Mamba2
from adnm_unet_b200.rmsnorm import RMSNorm, rmsnorm_affine


def _f32(p):
    return p if p.dtype == torch.float32 and p.is_contiguous() else p.detach().float().contiguous()


class _ResidualFunction(torch.autograd.Function):
    """out = (beta1 * x + beta2 * y) * gamma;  beta1 / beta2 0-dim (or 1-element) fp32 tensors, gamma (D) or None."""

    @staticmethod
    def forward(ctx, x, y, beta1, beta2, gamma):
        _lib.require_cuda(x, "x")
        lib = _lib.load()
        x = x.contiguous()
        y = y.to(x.dtype).contiguous()
        D = x.shape[-1]
        tokens = x.numel() // D
        b1, b2 = _f32(beta1), _f32(beta2)
        ga = None if gamma is None else _f32(gamma)
        out = torch.empty_like(x)
        with _lib.on_device(x.device):
            _lib.check(lib.adn_residual_forward(_lib.ptr(x), _lib.ptr(y), _lib.ptr(b1), _lib.ptr(b2), _lib.ptr(ga), _lib.ptr(out), tokens, D,
                                                _lib.dtype_code(x), _lib.stream_ptr(x.device)), "adn_residual_forward")
        ctx.save_for_backward(x, y, b1, b2, ga)
        ctx.shapes = (beta1.shape, beta2.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, y, b1, b2, ga = ctx.saved_tensors
        D = x.shape[-1]
        tokens = x.numel() // D
        dout = dout.to(x.dtype).contiguous()
        dx, dy = torch.empty_like(x), torch.empty_like(x)
        flat = torch.empty(2 + D, dtype=torch.float32, device=x.device)
        ws = _lib.scratch(64, x.device)
        with _lib.on_device(x.device):
            _lib.check(lib.adn_residual_backward(_lib.ptr(x), _lib.ptr(y), _lib.ptr(dout), _lib.ptr(b1), _lib.ptr(b2), _lib.ptr(ga), _lib.ptr(dx),
                                                 _lib.ptr(dy), _lib.ptr(flat[0:1]), _lib.ptr(flat[1:2]), _lib.ptr(flat[2:]) if ga is not None else None,
                                                 _lib.ptr(ws), tokens, D, _lib.dtype_code(x), _lib.stream_ptr(x.device)), "adn_residual_backward")
        ni = ctx.needs_input_grad
        return (dx if ni[0] else None, dy if ni[1] else None, flat[0].reshape(ctx.shapes[0]) if ni[2] else None,
                flat[1].reshape(ctx.shapes[1]) if ni[3] else None, flat[2:] if ga is not None and ni[4] else None)


def residual_mix(x, y, beta1, beta2, gamma=None):
    return _ResidualFunction.apply(x, y, beta1, beta2, gamma)


_FFN_CACHE = {}


def _ffn_info(x, H, W, C4):
    B, L, D = x.shape
    if L != H * W:
        raise RuntimeError(f"ffn: L={L} != H*W={H * W}")
    key = (B, H, W, D, C4, x.dtype)
    hit = _FFN_CACHE.get(key)
    if hit is None:
        shape = _lib.AdnFfnShape(B=B, H=H, W=W, D=D, C4=C4, dtype=_lib.dtype_code(x))
        sv, fw, bw = (_lib.C.c_size_t() for _ in range(3))
        _lib.check(_lib.load().adn_ffn_workspace_bytes(shape, sv, fw, bw), "adn_ffn_workspace_bytes")
        hit = _FFN_CACHE[key] = (shape, sv.value, fw.value, bw.value)
    return hit


def _ffn_struct(tensors):
    s = _lib.AdnFfnWeights()
    for f, t in zip(_lib.FFN_FIELDS, tensors):
        setattr(s, f, t.data_ptr())
    return s


class _FfnFunction(torch.autograd.Function):
    """x (B, L, D) token-major, then the six FeedForward parameters in AdnFfnWeights order."""

    @staticmethod
    def forward(ctx, x, H, W, grad_mode, *params):
        _lib.require_cuda(x, "x")
        lib = _lib.load()
        x = x.contiguous()
        prepped = [_f32(p) for p in params]
        C4 = prepped[0].shape[0]
        shape, sv, fw, bw = _ffn_info(x, H, W, C4)
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)
        saved = _lib.scratch(sv, x.device) if need_grad else None
        ws = _lib.scratch(fw, x.device)
        y = torch.empty_like(x)
        wts = _ffn_struct(prepped)
        with _lib.on_device(x.device):
            _lib.check(lib.adn_ffn_forward(shape, wts, _lib.ptr(x), _lib.ptr(y), _lib.ptr(saved), _lib.ptr(ws), _lib.stream_ptr(x.device)),
                       "adn_ffn_forward")
        if need_grad:
            ctx.save_for_backward(x, saved, *prepped)
            ctx.cfg = (shape, bw, [p.shape for p in params], [p.dtype for p in params])
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, saved, *prepped = ctx.saved_tensors
        shape, bw, shapes, dtypes = ctx.cfg
        dy = dy.to(x.dtype).contiguous()
        sizes = [p.numel() for p in prepped]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=x.device)
        grads = list(flat.split(sizes))
        dx = torch.empty_like(x)
        ws = _lib.scratch(bw, x.device)
        with _lib.on_device(x.device):
            _lib.check(lib.adn_ffn_backward(shape, _ffn_struct(prepped), _lib.ptr(x), _lib.ptr(saved), _lib.ptr(dy), _lib.ptr(dx), _ffn_struct(grads),
                                            _lib.ptr(ws), _lib.stream_ptr(x.device)), "adn_ffn_backward")
        pg = tuple(g.view(s).to(dt) if need else None for g, s, dt, need in zip(grads, shapes, dtypes, ctx.needs_input_grad[4:]))
        return (dx if ctx.needs_input_grad[0] else None, None, None, None) + pg


class _LinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        _lib.require_cuda(x, "x")
        lib = _lib.load()
        x = x.contiguous()
        N, K = weight.shape
        tokens = x.numel() // K
        w, b = _f32(weight), None if bias is None else _f32(bias)
        nb = _lib.C.c_size_t()
        _lib.check(lib.adn_linear_workspace_bytes(tokens, K, N, _lib.dtype_code(x), nb), "adn_linear_workspace_bytes")
        ws = _lib.scratch(nb.value, x.device)
        y = torch.empty(x.shape[:-1] + (N,), dtype=x.dtype, device=x.device)
        with _lib.on_device(x.device):
            _lib.check(lib.adn_linear_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), tokens, K, N, _lib.dtype_code(x), _lib.ptr(ws),
                                              _lib.stream_ptr(x.device)), "adn_linear_forward")
        ctx.save_for_backward(x, w)
        ctx.cfg = (nb.value, bias is not None, weight.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x, w = ctx.saved_tensors
        nbytes, has_bias, wdtype = ctx.cfg
        N, K = w.shape
        tokens = x.numel() // K
        dy = dy.to(x.dtype).contiguous()
        dx = torch.empty_like(x)
        flat = torch.empty(N * K + N, dtype=torch.float32, device=x.device)
        dw, db = flat[:N * K].view(N, K), flat[N * K:]
        ws = _lib.scratch(nbytes, x.device)
        with _lib.on_device(x.device):
            _lib.check(lib.adn_linear_backward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db) if has_bias else None,
                                               tokens, K, N, _lib.dtype_code(x), _lib.ptr(ws), _lib.stream_ptr(x.device)), "adn_linear_backward")
        ni = ctx.needs_input_grad
        return dx if ni[0] else None, dw.to(wdtype) if ni[1] else None, db.to(wdtype) if has_bias and ni[2] else None


def linear_tokens(x, weight, bias=None):
    return _LinearFunction.apply(x, weight, bias)


class _ConvHolder(nn.Module):
    """Parameter container with the layout of the reference's `Conv2dLayer` (models/model_untils.py:73-93) for the three
    convolutions of FeedForward (no dropout / norm / activation there): `.conv` is the nn.Conv2d whose weight / bias the
    kernels read."""

    def __init__(self, cin, cout, k, pad, groups=1, bias=True):
        super().__init__()
        self.dropout = None
        self.conv = nn.Conv2d(cin, cout, k, (1, 1), pad, (1, 1), groups, bias)
        self.norm = None
        self.act = None


class FeedForward(nn.Module):
    """Mirror of models/model_untils.py:172-197; `forward_tokens` takes (B, L, D) + (H, W), `forward` the reference's NCHW."""

    def __init__(self, dim, ffn_expansion_factor=2, bias=True):
        super().__init__()
        if not bias:
            raise NotImplementedError("adnb200 FeedForward covers bias=True (the only configuration ADNM-UNet instantiates)")
        hidden = int(dim * ffn_expansion_factor)
        self.project_in = _ConvHolder(dim, hidden * 2, (1, 1), (0, 0), bias=bias)
        self.dwconv = _ConvHolder(hidden * 2, hidden * 2, (3, 3), (1, 1), groups=hidden * 2, bias=bias)
        self.project_out = _ConvHolder(hidden, dim, (1, 1), (0, 0), bias=bias)

    def _params(self):
        return (self.project_in.conv.weight, self.project_in.conv.bias, self.dwconv.conv.weight, self.dwconv.conv.bias,
                self.project_out.conv.weight, self.project_out.conv.bias)

    def forward_tokens(self, x, H, W):
        return _FfnFunction.apply(x, int(H), int(W), torch.is_grad_enabled(), *self._params())

    def forward(self, x):
        b, d, h, w = x.shape
        if torch.is_autocast_enabled("cuda"):
            x = x.to(torch.get_autocast_dtype("cuda"))
        y = self.forward_tokens(x.permute(0, 2, 3, 1).reshape(b, h * w, d), h, w)
        return y.view(b, h, w, d).permute(0, 3, 1, 2)


class Swish(nn.Module):
    """models/model_untils.py:162-169 (declared by Block as `act`, never called)."""

    def __init__(self, beta_init=1.0):
        super().__init__()
        self.beta = nn.Parameter(torch.tensor(beta_init, dtype=torch.float))

    def forward(self, x):
        return x * torch.sigmoid(self.beta * x)


class Block(nn.Module):
    """Mirror of models/ADNMUNet.py:49-165."""

    def __init__(self, dim, out_dim, mixer, norm_layer=None, fused_add_norm=False, residual_in_fp32=False, drop_path=0., drop=0.,
                 patches_resolution=(64, 64), mlp_ratio=4, num_layers=1, act_layer=nn.SiLU, attn=False):
        super().__init__()
        if drop_path > 0.:
            raise NotImplementedError("drop_path is 0 everywhere in ADNM-UNet")
        if norm_layer is None:
            raise NotImplementedError("pass norm_layer=partial(RMSNorm, eps=...) as create_block does")
        self.residual_in_fp32 = residual_in_fp32
        self.fused_add_norm = fused_add_norm
        self.dim = dim
        self.out_dim = out_dim
        self.num_layers = num_layers
        self.alpha1 = nn.Parameter(torch.tensor(1, dtype=torch.float))
        self.alpha2 = nn.Parameter(torch.tensor(1, dtype=torch.float))
        self.alpha3 = nn.Parameter(torch.tensor(1, dtype=torch.float))
        self.alpha4 = nn.Parameter(torch.tensor(1, dtype=torch.float))
        self.beta1 = nn.Parameter(torch.ones(num_layers))
        self.beta2 = nn.Parameter(torch.ones(num_layers))
        self.beta3 = nn.Parameter(torch.ones(num_layers))
        self.beta4 = nn.Parameter(torch.ones(num_layers))
        self.mixer_layers = nn.ModuleList([mixer() for _ in range(num_layers)])
        self.drop_path_layers = nn.ModuleList([nn.Identity() for _ in range(num_layers)])
        self.norm1_layers = nn.ModuleList([norm_layer(dim) for _ in range(num_layers)])
        self.ffns = nn.ModuleList([FeedForward(dim=dim, ffn_expansion_factor=2, bias=True) for _ in range(num_layers)])
        self.norm2_layers = nn.ModuleList([norm_layer(dim) for _ in range(num_layers)])
        self.scale1 = nn.ParameterList([nn.Parameter(torch.tensor(1.)) for _ in range(num_layers)])
        self.shift1 = nn.ParameterList([nn.Parameter(torch.tensor(0.)) for _ in range(num_layers)])
        self.scale2 = nn.ParameterList([nn.Parameter(torch.tensor(1.)) for _ in range(num_layers)])
        self.shift2 = nn.ParameterList([nn.Parameter(torch.tensor(0.)) for _ in range(num_layers)])
        self.act = Swish()
        if self.dim != self.out_dim:
            self.out_proj = nn.Linear(dim, out_dim)
        self.gamma = nn.Parameter(1 * torch.ones(dim))
        for n in list(self.norm1_layers) + list(self.norm2_layers):
            if not hasattr(n, "weight") or hasattr(n, "bias") and n.bias is not None:
                raise NotImplementedError("adnb200 Block fuses the standalone RMSNorm (weight only) of the reference README")

    def forward(self, hidden_states, residual=None, features=None, inference_params=None, use_checkpoint=False):
        b, l, d = hidden_states.shape
        h = w = int(math.sqrt(l))
        x = hidden_states
        if residual is not None:                                   # models/ADNMUNet.py:124-132
            x = torch.cat((self.alpha1 * x, self.alpha2 * residual), dim=-1)
            if features is not None:
                x = x + torch.cat((self.alpha3 * features, self.alpha4 * features), dim=-1)
        elif features is not None:
            x = x + self.alpha3 * features
        if torch.is_autocast_enabled("cuda"):
            x = x.to(torch.get_autocast_dtype("cuda"))
        x = x.contiguous()
        for i in range(self.num_layers):
            n1, n2 = self.norm1_layers[i], self.norm2_layers[i]
            b1, b2 = self.beta1[i], self.beta2[i]                  # beta3 / beta4 alias them (models/ADNMUNet.py:145-146)
            last = i == self.num_layers - 1
            # x also feeds the residual mix: take it from the norm's pass-through output so that both gradients meet in one kernel
            xn, x = rmsnorm_affine(x, n1.weight, self.scale1[i], self.shift1[i], n1.eps, passthrough=True)
            x = residual_mix(x, self.mixer_layers[i](xn, h, w), b1, b2)
            xn, x = rmsnorm_affine(x, n2.weight, self.scale2[i], self.shift2[i], n2.eps, passthrough=True)
            x = residual_mix(x, self.ffns[i].forward_tokens(xn, h, w), b1, b2, self.gamma if last else None)
        if self.num_layers == 0:
            x = x * self.gamma.view(1, 1, -1)
        if self.dim != self.out_dim:
            x = linear_tokens(x, self.out_proj.weight, self.out_proj.bias)
        return x


def make_block(d_model, out_dim, headdim=4, d_state=16, num_layers=1, norm_epsilon=1e-5, layer_idx=None):
    """`create_block` (models/ADNMUNet.py:243-292) with the sm_100a modules."""
    from functools import partial
    mixer = partial(Mamba2, layer_idx=layer_idx, d_model=d_model, headdim=headdim, linear_attn_duality=True, d_state=d_state)
    blk = Block(dim=d_model, out_dim=out_dim, mixer=mixer, num_layers=num_layers, norm_layer=partial(RMSNorm, eps=norm_epsilon),
                fused_add_norm=True, residual_in_fp32=True)
    blk.layer_idx = layer_idx
    return blk
