"""Drop-in for the reference's ADN-SSD mixer `models.ADNssd.Mamba2` (models/ADNssd.py:49-462).

Same constructor signature, same parameter names / shapes / registration order (so a reference state_dict loads
strictly), same `forward(u, H, W, seq_idx=None)`.  The whole forward and backward run in the
sm_100a library through the C ABI (include/adnb200.h: adnssd_forward / adnssd_backward).
"""
import math

import torch
import torch.nn as nn

from adnm_unet_b200 import _lib

# state_dict key -> C struct field, in AdnWeights order
_KEY_TO_FIELD = (
    ("dt_bias", "dt_bias"), ("A_log", "A_log"), ("D", "D"), ("scale", "scale"), ("shift", "shift"),
    ("alpha1", "alpha1"), ("alpha2", "alpha2"), ("in_proj.weight", "in_proj_w"),
    ("conv_13_x1.weight", "conv_13_x1_w"), ("conv_31_x1.weight", "conv_31_x1_w"),
    ("conv_13_x2.weight", "conv_13_x2_w"), ("conv_31_x2.weight", "conv_31_x2_w"),
    ("conv_13_bc1.weight", "conv_13_bc1_w"), ("conv_31_bc1.weight", "conv_31_bc1_w"),
    ("conv_13_bc2.weight", "conv_13_bc2_w"), ("conv_31_bc2.weight", "conv_31_bc2_w"),
    ("conv2d.weight", "conv2d_w"), ("norm.weight", "norm_w"), ("norm.bias", "norm_b"),
    ("conv2d_z.weight", "conv2d_z_w"), ("out_proj.weight", "out_proj_w"),
)
PARAM_KEYS = tuple(k for k, _ in _KEY_TO_FIELD)
UNUSED_KEYS = ("scale", "shift", "alpha2")          # declared, never read by the reference forward
USED_KEYS = tuple(k for k in PARAM_KEYS if k not in UNUSED_KEYS)


_SHAPE_CACHE = {}   # (B, H, W, D, Di, P, G, N, dtype) -> (AdnShape, saved bytes, fwd workspace bytes, bwd workspace bytes)


def _shape_info(u, H, W, headdim, d_state, d_inner, ngroups):
    B, L, D = u.shape
    if L != H * W:
        raise RuntimeError(f"adnssd: L={L} != H*W={H * W}")
    key = (B, H, W, D, d_inner, headdim, ngroups, d_state, u.dtype)
    hit = _SHAPE_CACHE.get(key)
    if hit is None:
        shape = _lib.AdnShape(B=B, H=H, W=W, D=D, Di=d_inner, P=headdim, G=ngroups, N=d_state, dtype=_lib.dtype_code(u), flags=0)
        sv, fw, bw = (_lib.C.c_size_t() for _ in range(3))
        _lib.check(_lib.load().adnssd_workspace_bytes(shape, sv, fw, bw), "adnssd_workspace_bytes")
        hit = _SHAPE_CACHE[key] = (shape, sv.value, fw.value, bw.value)
    return hit


def _weights_struct(cls, tensors):
    s = cls()
    for (key, field) in _KEY_TO_FIELD:
        t = tensors.get(key)
        setattr(s, field, None if t is None else t.data_ptr())
    return s


def _prep_param(p, device):
    if p.device != device:
        raise RuntimeError(f"adnssd: parameter on {p.device}, activations on {device}")
    if p.dtype == torch.float32 and p.is_contiguous():
        return p                      # only its data_ptr() is used
    return p.detach().float().contiguous()


STATS = {"forward_training": 0, "forward_inference": 0}   # which kernel variant ran (saved != NULL / saved == NULL)
_GRAD_NUMEL = {}   # tuple of parameter shapes -> (sizes, total): the 18 gradients live in ONE flat fp32 buffer


class _AdnSsdFunction(torch.autograd.Function):
    """u, then the 18 used parameters in USED_KEYS order.  `grad_mode` is torch.is_grad_enabled() sampled at the call
    site: inside forward() grad mode is always off and ctx.needs_input_grad ignores torch.no_grad(), so without it an
    eval-mode forward would allocate `saved` and run the training variant of the kernels."""

    @staticmethod
    def forward(ctx, u, H, W, headdim, d_state, d_inner, ngroups, grad_mode, *params):
        _lib.require_cuda(u, "u")
        lib = _lib.load()
        u = u.contiguous()
        shape, sv, fw, bw = _shape_info(u, H, W, headdim, d_state, d_inner, ngroups)
        tensors = {k: _prep_param(p, u.device) for k, p in zip(USED_KEYS, params)}
        wts = _weights_struct(_lib.AdnWeights, tensors)
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)
        saved = _lib.scratch(sv, u.device) if need_grad else None
        STATS["forward_training" if need_grad else "forward_inference"] += 1
        ws = _lib.scratch(fw, u.device)
        out = torch.empty_like(u)
        with _lib.on_device(u.device):
            _lib.check(lib.adnssd_forward(shape, wts, _lib.ptr(u), _lib.ptr(out), _lib.ptr(saved), _lib.ptr(ws),
                                          _lib.stream_ptr(u.device)), "adnssd_forward")
        if need_grad:
            ctx.save_for_backward(u, saved, *params)
            # the prepared fp32 tensors are kept alive next to their pointer struct (no-ops for fp32 contiguous parameters)
            ctx.cfg = (shape, bw, wts, tensors)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        u, saved, *params = ctx.saved_tensors
        shape, bw, wts, tensors = ctx.cfg
        dout = dout.to(u.dtype).contiguous()
        skey = tuple(t.shape for t in tensors.values())
        meta = _GRAD_NUMEL.get(skey)
        if meta is None:
            sizes = [t.numel() for t in tensors.values()]
            meta = _GRAD_NUMEL[skey] = (sizes, sum(sizes))
        flat = torch.empty(meta[1], dtype=torch.float32, device=u.device)
        grads = dict(zip(USED_KEYS, flat.split(meta[0])))
        gst = _weights_struct(_lib.AdnWeightGrads, grads)
        du = torch.empty_like(u)
        ws = _lib.scratch(bw, u.device)
        with _lib.on_device(u.device):
            _lib.check(lib.adnssd_backward(shape, wts, _lib.ptr(u), _lib.ptr(saved), _lib.ptr(dout), _lib.ptr(du), gst,
                                           _lib.ptr(ws), _lib.stream_ptr(u.device)), "adnssd_backward")
        pg = tuple((grads[k].view(p.shape) if p.dtype == torch.float32 else grads[k].to(p.dtype).view(p.shape)) if need else None
                   for k, p, need in zip(USED_KEYS, params, ctx.needs_input_grad[8:]))
        return (du if ctx.needs_input_grad[0] else None, None, None, None, None, None, None, None) + pg


def adnssd_mixer(u, H, W, params, headdim, d_state, ngroups=2, expand=2):
    """Functional form: `params` maps the reference's state_dict keys (PARAM_KEYS) to tensors."""
    d_inner = int(expand * u.shape[-1])
    return _AdnSsdFunction.apply(u, int(H), int(W), int(headdim), int(d_state), d_inner, int(ngroups),
                                 torch.is_grad_enabled(), *[params[k] for k in USED_KEYS])


class Mamba2(nn.Module):
    """Mirror of models/ADNssd.py:49-250 (constructor) and :302-462 (forward)."""

    def __init__(self, d_model, d_conv=3, conv_init=None, expand=2, headdim=8, ngroups=2, A_init_range=(1, 16),
                 dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4, dt_limit=(0.0, float("inf")),
                 learnable_init_states=False, bias=False, conv_bias=False, chunk_size=256, use_mem_eff_path=False,
                 layer_idx=None, device=None, dtype=None, linear_attn_duality=True, d_state=16, bimamba=True, **kwargs):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        if d_conv != 3 or bias or conv_bias or learnable_init_states or not linear_attn_duality:
            raise NotImplementedError("adnb200 Mamba2 covers the configuration ADNM-UNet instantiates: d_conv=3, no "
                                      "biases, linear_attn_duality=True, learnable_init_states=False")
        if conv_init is not None:
            raise NotImplementedError("conv_init must stay None (the reference's own branch references a missing attribute)")
        if not kwargs.get("ssd_positve_dA", True):
            raise NotImplementedError("ssd_positve_dA=False is not used by ADNM-UNet")
        self.bimamba, self.d_model, self.d_conv, self.conv_init, self.expand = bimamba, d_model, d_conv, conv_init, expand
        self.d_inner = int(expand * d_model)
        self.headdim, self.d_state = headdim, d_state
        if ngroups == -1:
            ngroups = self.d_inner // headdim
        self.ngroups = ngroups
        assert self.d_inner % headdim == 0
        if (self.d_inner // 2) % headdim != 0:
            # the reference's parity split `rearrange(x, 'b l (h p) -> ...')` on d_inner/2 channels raises in this case
            raise ValueError(f"(d_inner/2) % headdim != 0 (d_inner={self.d_inner}, headdim={headdim})")
        self.nheads = self.d_inner // headdim
        self.dt_limit, self.learnable_init_states = dt_limit, learnable_init_states
        self.chunk_size, self.use_mem_eff_path, self.layer_idx = chunk_size, use_mem_eff_path, layer_idx
        self.ssd_positve_dA = True
        self.linear_attn_duality = True
        self.kwargs = kwargs
        Di, GN = self.d_inner, ngroups * d_state
        d_in_proj = 2 * Di + 2 * GN + self.nheads
        # parameter containers, created in the reference's order (same RNG stream, same state_dict order)
        self.in_proj = nn.Linear(d_model, d_in_proj, bias=False, **fk)

        def dw(ch, k, pad):
            return nn.Conv2d(ch, ch, kernel_size=k, padding=pad, groups=ch, bias=False)

        self.conv_13_x1, self.conv_31_x1 = dw(Di // 4, (1, 3), (0, 1)), dw(Di // 4, (3, 1), (1, 0))
        self.conv_13_x2, self.conv_31_x2 = dw(Di // 4, (1, 3), (0, 1)), dw(Di // 4, (3, 1), (1, 0))
        self.conv_13_bc1, self.conv_31_bc1 = dw(2 * GN // 4, (1, 3), (0, 1)), dw(2 * GN // 4, (3, 1), (1, 0))
        self.conv_13_bc2, self.conv_31_bc2 = dw(2 * GN // 4, (1, 3), (0, 1)), dw(2 * GN // 4, (3, 1), (1, 0))
        self.conv2d = nn.Conv2d((Di + 2 * GN) // 2, (Di + 2 * GN) // 2, groups=(Di + 2 * GN) // 2, bias=False,
                                kernel_size=3, padding=1, **fk)
        dt = torch.exp(torch.rand(self.nheads, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = torch.clamp(dt, min=dt_init_floor)
        self.dt_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
        self.dt_bias._no_weight_decay = True
        assert A_init_range[0] > 0 and A_init_range[1] >= A_init_range[0]
        A = torch.empty(self.nheads, dtype=torch.float32, device=device).uniform_(*A_init_range)
        self.A_log = nn.Parameter(torch.log(A).to(dtype=dtype))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.nheads, device=device))
        self.D._no_weight_decay = True
        self.norm = nn.LayerNorm(Di)
        self.scale = nn.Parameter(torch.tensor(1.))
        self.shift = nn.Parameter(torch.tensor(0.))
        self.conv2d_z = nn.Conv2d(Di, Di, groups=Di, bias=False, kernel_size=3, padding=1, **fk)
        self.alpha1 = nn.Parameter(torch.tensor(1, dtype=torch.float))
        self.alpha2 = nn.Parameter(torch.tensor(1, dtype=torch.float))
        self.out_proj = nn.Linear(Di * 2, d_model, bias=False, **fk)

    def _param(self, key):
        obj = self
        for part in key.split("."):
            obj = getattr(obj, part)
        return obj

    def forward(self, u, H, W, seq_idx=None):
        if seq_idx is not None:
            raise NotImplementedError("seq_idx is always None in ADNM-UNet")
        if torch.is_autocast_enabled("cuda"):
            u = u.to(torch.get_autocast_dtype("cuda"))
        return _AdnSsdFunction.apply(u, int(H), int(W), self.headdim, self.d_state, self.d_inner, self.ngroups,
                                     torch.is_grad_enabled(), *[self._param(k) for k in USED_KEYS])
