"""Import the UNMODIFIED reference (kanyu369/ADNM-UNet) from /root/reference on CPU.

This file is test infrastructure used ONLY by `tests/golden/make_golden.py` (golden
generation, run in the build container where /root/reference is mounted) and by the
optional `-m "not gpu"` cross-checks that skip when /root/reference is absent.  Nothing
in the product path, in `-m gpu` tests, in `smoke()` or in `bench.py` imports it.

The reference cannot be imported as-is here:
  * `timm`, `pywt`, `mamba_ssm` are not installed -> tiny stand-in modules that export
    exactly the names the reference uses (models/ADNssd.py:5-9, models/model_untils.py:11-16,
    models/WTConv2d.py:4-5, models/ADNMUNet.py:11-16,27-32).  `mamba_ssm...layer_norm.RMSNorm`
    is the standalone class the reference README tells users to substitute (README.md:22-30).
  * `Mamba2.forward` does ten `torch.arange(..).to('cuda')` (models/ADNssd.py:329-382)
    -> `Tensor.to` is patched to ignore a bare 'cuda' target when no GPU is present.
No reference file is edited or copied.
"""
import contextlib
import importlib
import math
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("ADNM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "ADNssd.py"))


class _StandaloneRMSNorm(nn.Module):
    """README.md:22-30 of the reference (the 'no mamba_ssm' variant north_star names)."""

    def __init__(self, d_model: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d_model))

    def forward(self, x):
        output = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps)
        return output * self.weight.to(x.dtype)


def _install_shims():
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_adnm_shim", False):
        return

    def _mod(name):
        m = types.ModuleType(name)
        m._adnm_shim = True
        sys.modules[name] = m
        return m

    class DropPath(nn.Module):  # drop_path == 0 everywhere in ADNM-UNet -> identity
        def __init__(self, drop_prob=0.0, *a, **k):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            return x

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def to_ntuple(n):
        return lambda x: tuple(x) if isinstance(x, (tuple, list)) else tuple([x] * n)

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    class _Unused(nn.Module):
        def __init__(self, *a, **k):
            raise RuntimeError("timm stand-in: symbol imported but never used by ADNM-UNet")

    timm = _mod("timm")
    layers = _mod("timm.layers")
    models = _mod("timm.models")
    vit = _mod("timm.models.vision_transformer")
    timm.layers, timm.models, models.vision_transformer = layers, models, vit
    layers.DropPath, layers.to_2tuple, layers.to_ntuple, layers.trunc_normal_ = DropPath, to_2tuple, to_ntuple, trunc_normal_
    for n in ("AvgPool2dSame", "Mlp", "GlobalResponseNormMlp", "LayerNorm2d", "LayerNorm"):
        setattr(layers, n, _Unused)
    layers.create_conv2d = layers.get_act_layer = layers.make_divisible = lambda *a, **k: None
    models.register_model = lambda f: f
    vit._cfg = lambda **k: dict(k)
    vit._load_weights = lambda *a, **k: None

    pywt = _mod("pywt")
    _mod("pywt.data")
    s = 1.0 / math.sqrt(2.0)

    class Wavelet:  # db1 taps only (models/WTConv2d.py:10-12,20-21)
        def __init__(self, name):
            assert name in ("db1", "haar"), name
            self.dec_lo, self.dec_hi = [s, s], [-s, s]
            self.rec_lo, self.rec_hi = [s, s], [s, -s]

    pywt.Wavelet = Wavelet

    ms = _mod("mamba_ssm")
    ops = _mod("mamba_ssm.ops")
    tri = _mod("mamba_ssm.ops.triton")
    ms.ops, ops.triton = ops, tri

    def _dead(*a, **k):
        raise RuntimeError("mamba_ssm stand-in: dead symbol (linear_attn_duality=False branch)")

    for sub, names in (("ssd_combined", ("mamba_chunk_scan_combined", "mamba_split_conv1d_scan_combined")),
                       ("layernorm_gated", ("RMSNorm",)),
                       ("selective_state_update", ("selective_state_update",))):
        m = _mod("mamba_ssm.ops.triton." + sub)
        setattr(tri, sub, m)
        for n in names:
            setattr(m, n, _dead)
    ln = _mod("mamba_ssm.ops.triton.layer_norm")
    tri.layer_norm = ln
    ln.RMSNorm, ln.layer_norm_fn, ln.rms_norm_fn = _StandaloneRMSNorm, _dead, _dead


@contextlib.contextmanager
def cuda_to_is_noop():
    """Neutralise `.to('cuda')` of index vectors when running the reference on CPU."""
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.to

    def patched(self, *args, **kwargs):
        if args and isinstance(args[0], str) and args[0].startswith("cuda"):
            args = args[1:]
            if not args and not kwargs:
                return self
        return orig(self, *args, **kwargs)

    torch.Tensor.to = patched
    try:
        yield
    finally:
        torch.Tensor.to = orig


def load_reference():
    """Returns the reference `models` package namespace: .ADNssd, .WTConv2d, .ADNMUNet, .model_untils, .loss."""
    if not reference_available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    for name in ("WTConv2d", "model_untils", "ADNssd", "ADNMUNet", "loss"):
        setattr(ns, name, importlib.import_module("models." + name))
    return ns
