"""Drop-in for the reference's `models.WTConv2d.WTConv2d` (models/WTConv2d.py:63-153): same constructor, same
state_dict (incl. the frozen `wt_filter` / `iwt_filter` Parameters), same `forward(x (B,C,H,W))`.
Forward and backward run in the sm_100a library (include/adnb200.h: wtconv_forward / wtconv_backward)."""
import torch
import torch.nn as nn

from adnm_unet_b200 import _lib


def haar_filters(C, dtype=torch.float):
    """db1 analysis / synthesis filters as the reference builds them (models/WTConv2d.py:9-29): (4C,1,2,2) each."""
    f = 0.5 * torch.tensor([[[1, 1], [1, 1]], [[1, 1], [-1, -1]], [[1, -1], [1, -1]], [[1, -1], [-1, 1]]], dtype=dtype)
    f = f[:, None].repeat(C, 1, 1, 1)
    return f, f.clone()


STATS = {"forward_training": 0, "forward_inference": 0}


def _structs(x, k, levels, base_w, base_b, base_s, conv_ws, scale_ws, cls):
    w = cls()
    w.base_conv_w = base_w.data_ptr()
    w.base_conv_b = None if base_b is None else base_b.data_ptr()
    w.base_scale_w = base_s.data_ptr()
    for i in range(levels):
        w.wavelet_conv_w[i] = conv_ws[i].data_ptr()
        w.wavelet_scale_w[i] = scale_ws[i].data_ptr()
    return w


class _WTConvFunction(torch.autograd.Function):
    """x, k, levels, has_bias, base_w, base_b (or None), base_scale, conv_w[0..levels), scale_w[0..levels)."""

    @staticmethod
    def forward(ctx, x, k, levels, grad_mode, base_w, base_b, base_s, *rest):
        _lib.require_cuda(x, "x")
        lib = _lib.load()
        x = x.contiguous()
        B, Cc, H, W = x.shape
        shape = _lib.WtShape(B=B, C=Cc, H=H, W=W, k=k, levels=levels, has_bias=int(base_b is not None), dtype=_lib.dtype_code(x))
        prep = lambda t: None if t is None else t.detach().float().contiguous()
        pw = [prep(t) for t in (base_w, base_b, base_s) + tuple(rest)]
        wts = _structs(x, k, levels, pw[0], pw[1], pw[2], pw[3:3 + levels], pw[3 + levels:], _lib.WtWeights)
        sv, fw, bw = (_lib.C.c_size_t() for _ in range(3))
        _lib.check(lib.wtconv_workspace_bytes(shape, sv, fw, bw), "wtconv_workspace_bytes")
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)   # needs_input_grad ignores torch.no_grad()
        saved = _lib.scratch(sv.value, x.device) if need_grad else None
        STATS["forward_training" if need_grad else "forward_inference"] += 1
        ws = _lib.scratch(fw.value, x.device)
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(lib.wtconv_forward(shape, wts, _lib.ptr(x), _lib.ptr(y), _lib.ptr(saved), _lib.ptr(ws),
                                          _lib.stream_ptr()), "wtconv_forward")
        if need_grad:
            ctx.save_for_backward(x, saved, *[t for t in (base_w, base_b, base_s) + tuple(rest) if t is not None])
            ctx.cfg = (k, levels, base_b is not None, bw.value)
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        k, levels, has_bias, bw = ctx.cfg
        x, saved, *ps = ctx.saved_tensors
        if not has_bias:
            ps.insert(1, None)
        B, Cc, H, W = x.shape
        shape = _lib.WtShape(B=B, C=Cc, H=H, W=W, k=k, levels=levels, has_bias=int(has_bias), dtype=_lib.dtype_code(x))
        prep = lambda t: None if t is None else t.detach().float().contiguous()
        pw = [prep(t) for t in ps]
        wts = _structs(x, k, levels, pw[0], pw[1], pw[2], pw[3:3 + levels], pw[3 + levels:], _lib.WtWeights)
        gr = [None if t is None else torch.empty_like(t) for t in pw]
        gst = _structs(x, k, levels, gr[0], gr[1], gr[2], gr[3:3 + levels], gr[3 + levels:], _lib.WtWeightGrads)
        dy = dy.to(x.dtype).contiguous()
        dx = torch.empty_like(x)
        ws = _lib.scratch(bw, x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.wtconv_backward(shape, wts, _lib.ptr(x), _lib.ptr(saved), _lib.ptr(dy), _lib.ptr(dx), gst,
                                           _lib.ptr(ws), _lib.stream_ptr()), "wtconv_backward")
        pg = tuple(None if (g is None or not need) else g.to(p.dtype).reshape(p.shape)
                   for g, p, need in zip(gr, ps, ctx.needs_input_grad[4:]))
        return (dx if ctx.needs_input_grad[0] else None, None, None, None) + pg


def wtconv2d(x, params, k, levels):
    """Functional form; `params` maps the reference's state_dict keys to tensors (Haar filter entries ignored)."""
    conv_ws = [params[f"wavelet_convs.{i}.weight"] for i in range(levels)]
    scale_ws = [params[f"wavelet_scale.{i}.weight"] for i in range(levels)]
    return _WTConvFunction.apply(x, int(k), int(levels), torch.is_grad_enabled(), params["base_conv.weight"], params.get("base_conv.bias"),
                                 params["base_scale.weight"], *conv_ws, *scale_ws)


class _ScaleModule(nn.Module):
    """models/WTConv2d.py:53-61 (parameter container; the multiply happens inside the fused kernel)."""

    def __init__(self, dims, init_scale=1.0, init_bias=0):
        super().__init__()
        self.dims = dims
        self.weight = nn.Parameter(torch.ones(*dims) * init_scale)
        self.bias = None


class WTConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=5, stride=1, bias=True, wt_levels=2, wt_type='db1'):
        super().__init__()
        assert in_channels == out_channels
        if wt_type not in ("db1", "haar"):
            raise NotImplementedError("adnb200 WTConv2d implements the db1 (Haar) wavelet ADNM-UNet uses")
        if stride != 1:
            raise NotImplementedError("stride > 1 is never used by ADNM-UNet")
        if kernel_size % 2 != 1:
            raise NotImplementedError("odd kernel sizes only (padding='same')")
        self.in_channels, self.wt_levels, self.stride, self.dilation = in_channels, wt_levels, stride, 1
        self.kernel_size = kernel_size
        wt, iwt = haar_filters(in_channels)
        self.wt_filter = nn.Parameter(wt, requires_grad=False)
        self.iwt_filter = nn.Parameter(iwt, requires_grad=False)
        self.base_conv = nn.Conv2d(in_channels, in_channels, kernel_size, padding='same', stride=1, dilation=1,
                                   groups=in_channels, bias=bias)
        self.base_scale = _ScaleModule([1, in_channels, 1, 1])
        self.wavelet_convs = nn.ModuleList(
            [nn.Conv2d(in_channels * 4, in_channels * 4, kernel_size, padding='same', stride=1, dilation=1,
                       groups=in_channels * 4, bias=False) for _ in range(wt_levels)])
        self.wavelet_scale = nn.ModuleList(
            [_ScaleModule([1, in_channels * 4, 1, 1], init_scale=0.1) for _ in range(wt_levels)])
        self.do_stride = None

    def forward(self, x):
        if torch.is_autocast_enabled("cuda"):
            x = x.to(torch.get_autocast_dtype("cuda"))
        return _WTConvFunction.apply(x, self.kernel_size, self.wt_levels, torch.is_grad_enabled(), self.base_conv.weight, self.base_conv.bias,
                                     self.base_scale.weight, *[c.weight for c in self.wavelet_convs],
                                     *[s.weight for s in self.wavelet_scale])
