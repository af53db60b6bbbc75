"""Drop-in for the standalone `RMSNorm` the reference README tells users to substitute for mamba_ssm's (README.md:22-30;
bound as `norm_layer` by create_block, models/ADNMUNet.py:278): same constructor, same single `weight` parameter, same
forward.  `rmsnorm_affine` additionally folds the Block's scalar `scale * norm(x) + shift` (models/ADNMUNet.py:149,155)
into the same pass.  Forward and backward are one kernel each in the sm_100a library (include/adnb200.h)."""
import torch
import torch.nn as nn

from adnm_unet_b200 import _lib


class _RmsNormFunction(torch.autograd.Function):
    """`passthrough`: also return x itself as a second output.  The Block feeds that alias to its residual mix
    (models/ADNMUNet.py:152), so the gradient of the residual path arrives HERE and is added to the norm's own dx inside the
    backward kernel - one pass and one rounding instead of autograd's separate accumulation of two bf16 tensors."""

    @staticmethod
    def forward(ctx, x, weight, scale, shift, eps, grad_mode, passthrough=False):
        ctx.set_materialize_grads(False)
        _lib.require_cuda(x, "x")
        lib = _lib.load()
        x = x.contiguous()
        D = x.shape[-1]
        tokens = x.numel() // D
        w = weight if weight.dtype == torch.float32 and weight.is_contiguous() else weight.detach().float().contiguous()
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)
        y = torch.empty_like(x)
        rstd = torch.empty(tokens, dtype=torch.float32, device=x.device) if need_grad else None
        with _lib.on_device(x.device):
            _lib.check(lib.adn_rmsnorm_forward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(y), _lib.ptr(rstd),
                                               tokens, D, float(eps), _lib.dtype_code(x), _lib.stream_ptr(x.device)), "adn_rmsnorm_forward")
        if need_grad:
            ctx.save_for_backward(x, w, scale, shift, rstd)
            ctx.wdtype = weight.dtype
        return (y, x.view_as(x)) if passthrough else y

    @staticmethod
    def backward(ctx, dy, dres=None):
        lib = _lib.load()
        x, w, scale, shift, rstd = ctx.saved_tensors
        D = x.shape[-1]
        tokens = x.numel() // D
        if dy is None:                       # only the pass-through output was used
            return (dres,) + (None,) * 6
        dy = dy.to(x.dtype).contiguous()
        if dres is not None:
            dres = dres.to(x.dtype).contiguous()
        dx = torch.empty_like(x)
        flat = torch.empty(D + 2, dtype=torch.float32, device=x.device)
        dw, dscale, dshift = flat[:D], flat[D:D + 1], flat[D + 1:]
        with _lib.on_device(x.device):
            _lib.check(lib.adn_rmsnorm_backward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(scale), _lib.ptr(rstd), _lib.ptr(dy), _lib.ptr(dres),
                                                _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(dscale) if scale is not None else None,
                                                _lib.ptr(dshift) if shift is not None else None, tokens, D, _lib.dtype_code(x),
                                                _lib.stream_ptr(x.device)), "adn_rmsnorm_backward")
        ni = ctx.needs_input_grad
        return (dx if ni[0] else None, dw.to(ctx.wdtype) if ni[1] else None,
                dscale.reshape(scale.shape) if scale is not None and ni[2] else None,
                dshift.reshape(shift.shape) if shift is not None and ni[3] else None, None, None, None)


def rmsnorm_affine(x, weight, scale=None, shift=None, eps=1e-5, passthrough=False):
    """scale * (x * rsqrt(mean(x^2, -1) + eps) * weight) + shift; scale / shift are 0-dim fp32 tensors or None.
    passthrough=True returns (y, x_alias): use x_alias for the residual branch (see _RmsNormFunction)."""
    return _RmsNormFunction.apply(x, weight, scale, shift, eps, torch.is_grad_enabled(), passthrough)


class RMSNorm(nn.Module):
    """README.md:22-30 of the reference; state_dict = {weight (d_model)} like mamba_ssm.ops.triton.layer_norm.RMSNorm."""

    def __init__(self, d_model: int, eps: float = 1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d_model, device=device, dtype=dtype))

    def forward(self, x):
        return rmsnorm_affine(x, self.weight, None, None, self.eps)
