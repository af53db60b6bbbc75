"""Drop-in for the reference's `StandardAttention` (models/ADNssd.py:26-47), the softmax attention inside the three
`Attention` bridges of ADNM-UNet (models/ADNMUNet.py:172-238): same constructor, same parameter names (`to_qkv.weight`,
`to_out.weight`, `to_out.bias`), same `forward(x, H, W)`.  to_qkv / to_out run on the library's GEMM (adn_linear_*), the
attention itself is one fused kernel per pass (adn_sdpa_forward / _backward): the B x heads x L x L score tensor of the
reference is never materialised."""
import torch
import torch.nn as nn

from adnm_unet_b200 import _lib
from adnm_unet_b200.block import linear_tokens


class _SdpaFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads, dh, scale, grad_mode):
        _lib.require_cuda(qkv, "qkv")
        lib = _lib.load()
        qkv = qkv.contiguous()
        B, L, three_inner = qkv.shape
        inner = heads * dh
        if three_inner != 3 * inner:
            raise RuntimeError(f"sdpa: last dimension {three_inner} != 3 * heads * dim_head = {3 * inner}")
        need_grad = bool(grad_mode) and ctx.needs_input_grad[0]
        out = torch.empty(B, L, inner, dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty(B, heads, L, dtype=torch.float32, device=qkv.device) if need_grad else None
        with _lib.on_device(qkv.device):
            _lib.check(lib.adn_sdpa_forward(_lib.ptr(qkv), _lib.ptr(out), _lib.ptr(lse), B, L, heads, dh, float(scale), _lib.dtype_code(qkv),
                                            _lib.stream_ptr(qkv.device)), "adn_sdpa_forward")
        if need_grad:
            ctx.save_for_backward(qkv, out, lse)
            ctx.cfg = (heads, dh, float(scale))
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        qkv, out, lse = ctx.saved_tensors
        heads, dh, scale = ctx.cfg
        B, L, _ = qkv.shape
        dout = dout.to(qkv.dtype).contiguous()
        dqkv = torch.empty_like(qkv)
        with _lib.on_device(qkv.device):
            _lib.check(lib.adn_sdpa_backward(_lib.ptr(qkv), _lib.ptr(out), _lib.ptr(lse), _lib.ptr(dout), _lib.ptr(dqkv), B, L, heads, dh, scale,
                                             _lib.dtype_code(qkv), _lib.stream_ptr(qkv.device)), "adn_sdpa_backward")
        return dqkv, None, None, None, None


def sdpa_packed(qkv, heads, dim_head, scale=None):
    """softmax(q k^T * scale) v from the packed (B, L, 3 * heads * dim_head) projection; returns (B, L, heads * dim_head)."""
    return _SdpaFunction.apply(qkv, int(heads), int(dim_head), dim_head ** -0.5 if scale is None else scale, torch.is_grad_enabled())


class StandardAttention(nn.Module):
    """Mirror of models/ADNssd.py:26-47."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0., **kwargs):
        super().__init__()
        if dropout != 0.:
            raise NotImplementedError("adnb200 StandardAttention covers dropout=0 (the only value ADNM-UNet passes)")
        if dim_head not in (4, 8, 16):
            raise NotImplementedError(f"adnb200 StandardAttention covers dim_head in (4, 8, 16); ADNM-UNet uses 4 (got {dim_head})")
        inner_dim = dim_head * heads
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Linear(inner_dim, dim)
        self.dropout = nn.Dropout(dropout)
        self.inner_dim = inner_dim

    def forward(self, x, H, W):
        if torch.is_autocast_enabled("cuda"):
            x = x.to(torch.get_autocast_dtype("cuda"))
        qkv = linear_tokens(x, self.to_qkv.weight, None)
        out = sdpa_packed(qkv, self.heads, self.dim_head, self.scale)
        return linear_tokens(out, self.to_out.weight, self.to_out.bias)
