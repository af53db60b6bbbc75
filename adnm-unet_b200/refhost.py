"""Host the UNMODIFIED reference network (kanyu369/ADNM-UNet) around the B200 drop-ins.

The reference is pure Python; it is located at run time (never vendored into this package):
  $ADNM_REFERENCE_ROOT  ->  <repo>/baseline/_ref (git-ignored copy made by baseline/fetch_ref.py, travels to the GPU box)
  ->  /root/reference (build container only).
What this module does, and nothing more:
  * stand-ins for the three third-party packages the reference imports but this image lacks (`timm`, `pywt`,
    `mamba_ssm`): exactly the names the reference uses (models/ADNMUNet.py:11-16,27-32, models/ADNssd.py:5-9,
    models/model_untils.py:11-16, models/WTConv2d.py:4-5).  `mamba_ssm...layer_norm.RMSNorm` is the standalone class
    the reference README tells users to substitute (README.md:22-30) - the variant BASELINE.json names;
  * `Decoder.forward` hard-codes a 256 x 256 grid (models/ADNMUNet.py:634): for other image sizes the method is
    re-compiled from its own source with the literal replaced by sqrt(L) (no reference file is edited);
  * `build_adnm_unet(img_size, dropin)`: `VisionMamba` with the literals of `create_ADNMUNet(5, 20, 6)`
    (models/ADNMUNet.py:906-940) and, when `dropin`, the two module globals rebound to the sm_100a modules
    (SURVEY.md 8(b)): `models.ADNMUNet.Mamba2`, `models.model_untils.WTConv2d`, plus the Block-level names and the conv
    stages `models.ADNMUNet.WTLayer / PatchEmbed / OutProj`;
  * `reference_optimizer` / `reference_loss`: train_untils.py:29-43 without importing train_untils (it imports every
    baseline model and builds two of them at import).
"""
import contextlib
import importlib
import inspect
import math
import os
import sys
import textwrap
import types

import torch
import torch.nn as nn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.environ.get("ADNM_REFERENCE_ROOT"), os.path.join(_REPO, "baseline", "_ref"), "/root/reference")


def reference_root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "models", "ADNssd.py")):
            return c
    return None


def reference_available() -> bool:
    return reference_root() is not None


class StandaloneRMSNorm(nn.Module):
    """README.md:22-30 of the reference (the 'no mamba_ssm' variant)."""

    def __init__(self, d_model: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d_model))

    def forward(self, x):
        output = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps)
        return output * self.weight.to(x.dtype)


def install_shims():
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

    if "timm" not in sys.modules:
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

    if "pywt" not in sys.modules:
        pywt = _mod("pywt")
        _mod("pywt.data")
        s = 1.0 / math.sqrt(2.0)

        class Wavelet:  # db1 taps only (models/WTConv2d.py:10-12,20-21)
            def __init__(self, name):
                assert name in ("db1", "haar"), name
                self.dec_lo, self.dec_hi = [s, s], [-s, s]
                self.rec_lo, self.rec_hi = [s, s], [s, -s]

        pywt.Wavelet = Wavelet

    if "mamba_ssm" not in sys.modules:
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
        ln.RMSNorm, ln.layer_norm_fn, ln.rms_norm_fn = StandaloneRMSNorm, _dead, _dead


@contextlib.contextmanager
def cuda_to_is_noop(force=False):
    """Neutralise `.to('cuda')` of the index vectors (models/ADNssd.py:329-382) when running the reference on CPU.
    force=True: also on a box that HAS a GPU (the CPU baseline legs of bench.py run the reference on the host cores there)."""
    if torch.cuda.is_available() and not force:
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


_NS = None


def load_reference():
    """The reference `models` package as a namespace: .ADNssd .WTConv2d .ADNMUNet .model_untils .loss (imported once)."""
    global _NS
    if _NS is not None:
        return _NS
    root = reference_root()
    if root is None:
        raise FileNotFoundError("reference sources not found: run `python baseline/fetch_ref.py` in the build container "
                                "(copies the needed .py files of /root/reference into git-ignored baseline/_ref/)")
    install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)
    ns = types.SimpleNamespace(root=root)
    for name in ("WTConv2d", "model_untils", "ADNssd", "ADNMUNet", "loss"):
        setattr(ns, name, importlib.import_module("models." + name))
    ns.ref_Mamba2 = ns.ADNssd.Mamba2          # the reference's own classes, whatever the globals are rebound to later
    ns.ref_WTConv2d = ns.WTConv2d.WTConv2d
    ns.ref_Block, ns.ref_RMSNorm = ns.ADNMUNet.Block, ns.ADNMUNet.RMSNorm
    ns.ref_StandardAttention = ns.ADNssd.StandardAttention
    ns.ref_FeedForward = ns.model_untils.FeedForward
    ns.ref_stages = {n: getattr(ns.model_untils, n) for n in STAGE_NAMES}
    ns.ref_Conv2dLayer = ns.model_untils.Conv2dLayer
    _patch_decoder_size(ns.ADNMUNet)
    _NS = ns
    return ns


def _patch_decoder_size(mod):
    """models/ADNMUNet.py:634 `x.view(b,256,256,d)` -> sqrt(L): same code, the literal made size-generic."""
    if getattr(mod.Decoder.forward, "_adnm_size_generic", False):
        return
    src = textwrap.dedent(inspect.getsource(mod.Decoder.forward))
    lit = "x.view(b,256,256,d)"
    if lit not in src:
        raise RuntimeError("reference Decoder.forward changed: the 256-literal patch no longer applies")
    src = src.replace(lit, "x.view(b,int(math.sqrt(l)),int(math.sqrt(l)),d)")
    scope = {}
    exec(compile(src, mod.__file__ + ":Decoder.forward[size-generic]", "exec"), mod.__dict__, scope)
    fn = scope["forward"]
    fn._adnm_size_generic = True
    mod.Decoder.forward = fn


STAGE_NAMES = ("WTLayer", "PatchEmbed", "OutProj")      # SURVEY.md 8(f)2; resolved by Encoder / Decoder / Refiner from models.ADNMUNet's
                                                        # globals (`from .model_untils import *`, models/ADNMUNet.py:33)


@contextlib.contextmanager
def _bound(ns, dropin, mixer=True, wtconv=True, block=True, stages=True):
    """Rebind (or restore) the construction-time globals for the duration of a model build: the mixer and WTConv2d classes
     and, with `block`, the `Block` / `RMSNorm` names `create_block` resolves (models/ADNMUNet.py:277-291)
    and the `StandardAttention` name `Attention` resolves (models/ADNMUNet.py:181)."""
    old = (ns.ADNMUNet.Mamba2, ns.model_untils.WTConv2d, ns.ADNMUNet.Block, ns.ADNMUNet.RMSNorm, ns.ADNMUNet.StandardAttention,
           ns.model_untils.FeedForward)
    old_stages = {n: getattr(ns.ADNMUNet, n) for n in STAGE_NAMES}
    for n in STAGE_NAMES:
        setattr(ns.ADNMUNet, n, ns.ref_stages[n])
    old_conv_layer = ns.model_untils.Conv2dLayer
    ns.model_untils.Conv2dLayer = ns.ref_Conv2dLayer
    if dropin and stages and wtconv:
        from adnm_unet_b200 import convstage
        for n in STAGE_NAMES:
            setattr(ns.ADNMUNet, n, getattr(convstage, n))
        # the bridges' grouped convs: the reference's own Conv2dLayer, subclassed (EncoderToDecoder resolves the name at :623-673)
        if getattr(ns, "bridge_Conv2dLayer", None) is None:
            ns.bridge_Conv2dLayer = convstage.make_bridge_conv_layer(ns.ref_Conv2dLayer)
        ns.model_untils.Conv2dLayer = ns.bridge_Conv2dLayer
    if dropin:
        from adnm_unet_b200.mixer import Mamba2
        from adnm_unet_b200.wtconv import WTConv2d
        from adnm_unet_b200.block import Block
        from adnm_unet_b200.rmsnorm import RMSNorm
        if mixer:
            ns.ADNMUNet.Mamba2 = Mamba2
        if wtconv:
            ns.model_untils.WTConv2d = WTConv2d
        if block:
            from adnm_unet_b200.attention import StandardAttention
            from adnm_unet_b200.block import FeedForward
            ns.ADNMUNet.Block, ns.ADNMUNet.RMSNorm, ns.ADNMUNet.StandardAttention = Block, RMSNorm, StandardAttention
            ns.model_untils.FeedForward = FeedForward      # the seven FeedForwards of the EncoderToDecoder bridges (model_untils.py:738)
    else:
        ns.ADNMUNet.Mamba2, ns.model_untils.WTConv2d = ns.ref_Mamba2, ns.ref_WTConv2d
        ns.ADNMUNet.Block, ns.ADNMUNet.RMSNorm, ns.ADNMUNet.StandardAttention = ns.ref_Block, ns.ref_RMSNorm, ns.ref_StandardAttention
        ns.model_untils.FeedForward = ns.ref_FeedForward
    try:
        yield
    finally:
        (ns.ADNMUNet.Mamba2, ns.model_untils.WTConv2d, ns.ADNMUNet.Block, ns.ADNMUNet.RMSNorm, ns.ADNMUNet.StandardAttention,
         ns.model_untils.FeedForward) = old
        for n in STAGE_NAMES:
            setattr(ns.ADNMUNet, n, old_stages[n])
        ns.model_untils.Conv2dLayer = old_conv_layer


DEAD_BRIDGES = (3, 4, 5, 6)


def prune_dead_bridges(model):
    """`Decoder.forward` (models/ADNMUNet.py:603-606) runs all seven `EncoderToDecoder` bridges, but the outputs of e2ds[3..6]
    never reach the network's output: features[3] is not consumed at all (:608-630) and `WTLayer.forward` builds a `torch.cat`
    of its `features` argument and throws it away (models/model_untils.py:407-408) - which is also why those bridges'
    parameters are among the 307 tensors whose gradient stays None.  The reference still spends ~40 % of an
    inference forward on them (measured: 55 of 137 ms at B = 64, 256 x 256; e2ds.6 alone 31 ms at full resolution).
    This replaces the `forward` of those four bridge INSTANCES by one that returns its input (same shape and dtype as the
    real result, so the discarded `cat` in WTLayer still type-checks); modules, parameters and state_dict are untouched and
    every output / gradient of the network is bit-identical (tests/test_fullmodel_gpu.py runs with it).  Host-level dead-code
    elimination, applied to the drop-in variant only - the reference arm of every benchmark runs the network unmodified."""
    import types
    dec = getattr(model, "decoder", None)
    if dec is None or not hasattr(dec, "e2ds"):
        raise RuntimeError("prune_dead_bridges: not an ADNM-UNet (no decoder.e2ds)")
    for i in DEAD_BRIDGES:
        dec.e2ds[i].forward = types.MethodType(lambda self, x, res: x, dec.e2ds[i])
    return model


def build_adnm_unet(img_size=256, dropin=True, input_frames=5, output_frames=20, seed=0, mixer=True, wtconv=True, block=True,
                    prune_dead=None, stages=True):
    """`create_ADNMUNet(5, 20, 6)` (models/ADNMUNet.py:906-940) at `img_size`; seed -> identical init for both variants
    (the drop-in constructors consume the RNG stream exactly like the reference's: tests/test_abi_cpu.py)."""
    ns = load_reference()
    if seed is not None:
        torch.manual_seed(seed)
    with _bound(ns, dropin, mixer, wtconv, block, stages):
        model = ns.ADNMUNet.VisionMamba(
            img_size=img_size, depth=[1, 1, 1], refine_depth=[1, 1, 1, 1], refine_headdim=[4, 4, 4, 4],
            refine_dim=[32, 32, 32, 32] if output_frames > 5 else [32, 32, 16, 16],
            embed_dim=[32, 64, 128, 256, 512, 1024], headdim=4, channels=input_frames, out_channels=output_frames,
            ssm_cfg=None, norm_epsilon=1e-6, initializer_cfg=None, kernel=[5, 5, 5], ratio=[2, 2, 2, 2, 2, 2],
            wt_levels=[3, 2, 1], out_expand=2, InstanceNorm=True)
    if prune_dead is None:
        prune_dead = dropin and os.environ.get("ADNM_KEEP_DEAD_BRIDGES", "0") != "1"
    if prune_dead:
        prune_dead_bridges(model)
    return model


def build_block(dim, out_dim, dropin=True, headdim=4, seed=0, block=True):
    """One `Block` as `create_block` builds it (models/ADNMUNet.py:243-292), for the Block-level parity tests."""
    ns = load_reference()
    if seed is not None:
        torch.manual_seed(seed)
    with _bound(ns, dropin, block=block):
        blk = ns.ADNMUNet.create_block(dim, out_dim, headdim=headdim, norm_epsilon=1e-6, layer_idx=0)
    return blk


def reference_loss():
    """train_untils.py:43"""
    return load_reference().loss.enRainfallLoss(omega_t=0.57, alpha=0.25, gamma=0.)


def reference_optimizer(model):
    """train_untils.py:35-42"""
    return torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-9, weight_decay=1e-2, amsgrad=False)


ADAMW = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-9, weight_decay=1e-2)     # train_untils.py:29-42
CLIP_NORM = 0.025                                                              # train.py:79-94 (norm_max, epochs <= 4)
