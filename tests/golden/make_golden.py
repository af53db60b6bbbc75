"""Generate tests/golden/*.npz by running the UNMODIFIED reference (from /root/reference) on CPU in fp64.

Run in the build container only:   python tests/golden/make_golden.py
Parameters and inputs are float32-representable values (so they can be stored compactly and fed to
the fp32/bf16 CUDA paths unchanged); the reference is evaluated on them in float64.
Each .npz holds: `param/<state_dict key>`, `u`/`dout` (or `x`/`dy`), `out`, `du` (or `dx`), `grad/<key>`.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from ref_loader import cuda_to_is_noop, load_reference  # noqa: E402
from oracle import adnssd_oracle, block_oracle, convstage_oracle, wtconv_oracle  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def mixer_case(ref, name):
    D, P, N, B, g = cases.MIXER_CASES[name]
    p32 = adnssd_oracle.init_params(D, P, N, seed=11, perturb=0.3, dtype=torch.float32)
    m = ref.ADNssd.Mamba2(d_model=D, headdim=P, d_state=N).double()
    m.load_state_dict({k: v.double() for k, v in p32.items()}, strict=True)
    u32, g32 = cases.mixer_inputs(name, torch.float32)
    u = u32.double().requires_grad_(True)
    with cuda_to_is_noop():
        out = m(u, g, g)
    out.backward(g32.double())
    arrays = {"param/" + k: v.numpy() for k, v in p32.items()}
    big = g >= 64
    s = cases.SUBSAMPLE_STRIDE if big else 1
    if not big:
        arrays.update(u=u32.numpy(), dout=g32.numpy())
    arrays.update(out=out.detach()[:, ::s].numpy(), du=u.grad[:, ::s].numpy())
    for k, v in m.named_parameters():
        if v.grad is None:
            assert k in adnssd_oracle.UNUSED_PARAMS, k
        else:
            arrays["grad/" + k] = v.grad.numpy()
    save(name, **arrays)


def mixer_bf16_case(ref, name):
    """Model-scale parameters (init_params mirrors the reference init, models/ADNssd.py:201-222 + ADNMUNet.py:294-323)."""
    D, P, N, B, g, perturb = cases.MIXER_BF16_CASES[name]
    p32 = adnssd_oracle.init_params(D, P, N, seed=13, perturb=perturb, dtype=torch.float32)
    m = ref.ADNssd.Mamba2(d_model=D, headdim=P, d_state=N).double()
    m.load_state_dict({k: v.double() for k, v in p32.items()}, strict=True)
    u32, g32 = cases.mixer_bf16_inputs(name, torch.float32)
    u = u32.double().requires_grad_(True)
    with cuda_to_is_noop():
        out = m(u, g, g)
    out.backward(g32.double())
    arrays = {"param/" + k: v.numpy() for k, v in p32.items()}
    arrays.update(out=out.detach().float().numpy(), du=u.grad.float().numpy())
    for k, v in m.named_parameters():
        if v.grad is not None:
            arrays["grad/" + k] = v.grad.float().numpy()
    save(name, **arrays)


def mixer_corr_case(ref, name):
    """128 x 128 goldens with a correlated upstream gradient (cases.MIXER_CORR_CASES); out / du stored subsampled."""
    D, P, N, B, g, perturb = cases.MIXER_CORR_CASES[name]
    p32 = adnssd_oracle.init_params(D, P, N, seed=17, perturb=perturb, dtype=torch.float32)
    m = ref.ADNssd.Mamba2(d_model=D, headdim=P, d_state=N).double()
    m.load_state_dict({k: v.double() for k, v in p32.items()}, strict=True)
    u = cases.mixer_corr_u(name, torch.float32).double().requires_grad_(True)
    with cuda_to_is_noop():
        out = m(u, g, g)
    dout = cases.mixer_corr_dout(name, out.detach())
    # the tests rebuild dout from the ORACLE's forward: the two must agree far below fp32 resolution
    o2 = adnssd_oracle.mixer_forward({k: v.double() for k, v in p32.items()}, u.detach(), g, g, P, N)
    assert float((o2 - out.detach()).abs().max() / out.detach().abs().max()) < 1e-12
    out.backward(dout.double())
    s = cases.SUBSAMPLE_STRIDE
    arrays = {"param/" + k: v.numpy() for k, v in p32.items()}
    arrays.update(out=out.detach()[:, ::s].numpy(), du=u.grad[:, ::s].numpy())
    for k, v in m.named_parameters():
        if v.grad is not None:
            arrays["grad/" + k] = v.grad.numpy()
    save(name, **arrays)


def wtconv_case(ref, name):
    C, k, L, B, H, W, bias = cases.WTCONV_CASES[name]
    p32 = wtconv_oracle.init_params(C, k, L, bias=bias, seed=21, dtype=torch.float32)
    m = ref.WTConv2d.WTConv2d(C, C, kernel_size=k, bias=bias, wt_levels=L).double()
    m.load_state_dict({n: v.double() for n, v in p32.items()}, strict=True)
    x32, dy32 = cases.wtconv_inputs(name, torch.float32)
    x = x32.double().requires_grad_(True)
    out = m(x)
    out.backward(dy32.double())
    arrays = {"param/" + n: v.numpy() for n, v in p32.items()}
    arrays.update(x=x32.numpy(), dy=dy32.numpy(), out=out.detach().numpy(), dx=x.grad.numpy())
    for n, v in m.named_parameters():
        if v.requires_grad:
            arrays["grad/" + n] = v.grad.numpy()
    save(name, **arrays)


def block_case(ref, name):
    """The unmodified reference Block (models/ADNMUNet.py:49-165) via create_block, fp64, CPU."""
    dim, out_dim, B, g, skip = cases.BLOCK_CASES[name]
    p32 = block_oracle.init_block_params(dim, out_dim, headdim=4, d_state=16, seed=23, perturb=0.1, dtype=torch.float32)
    blk = ref.ADNMUNet.create_block(dim, out_dim, headdim=4, norm_epsilon=1e-6, layer_idx=0).double()
    blk.load_state_dict({k: v.double() for k, v in p32.items()}, strict=True)
    x, res, feat = (None if t is None else t.requires_grad_(True) for t in cases.block_inputs(name, torch.float64))
    with cuda_to_is_noop():
        out = blk(x, residual=res, features=feat)
    dout = cases.block_dout(name, out.detach())
    out.backward(dout.double())
    arrays = {"param/" + k: v.numpy() for k, v in p32.items()}
    arrays.update(out=out.detach().numpy(), dx=x.grad.numpy())
    if skip:
        arrays.update(dresidual=res.grad.numpy(), dfeatures=feat.grad.numpy())
    for k, v in blk.named_parameters():
        if v.grad is not None:
            arrays["grad/" + k] = v.grad.numpy()
    save(name, **arrays)


def convstage_case(ref, name):
    """The unmodified reference WTLayer / PatchEmbed / OutProj (models/model_untils.py:226-426,799-892), fp64, CPU."""
    kind, kw, B, g, skip = cases.CONVSTAGE_CASES[name]
    torch.manual_seed(31)
    m = getattr(ref.model_untils, kind)(**kw)
    p32 = convstage_oracle.perturb_params(m.state_dict(), seed=29)
    m = m.double()
    m.load_state_dict({k: v.double() for k, v in p32.items()}, strict=True)
    x, second = cases.convstage_inputs(name, torch.float64)
    x.requires_grad_(kind != "PatchEmbed")
    if second is not None and kind == "WTLayer":
        second.requires_grad_(True)
    if kind == "WTLayer":
        out = m(x, residual=second, features=second.detach() * 0.5 if skip else None)
    elif kind == "PatchEmbed":
        out, res = m(x)
        assert torch.equal(res, x.view(B, g, g, -1)[..., -1])
    else:
        out = m(x, second)
    dout = cases.convstage_dout(name, out.detach())
    out.backward(dout.double())
    arrays = {"param/" + k: v.numpy() for k, v in p32.items()}
    arrays.update(out=out.detach().numpy())
    if x.grad is not None:
        arrays.update(dx=x.grad.numpy())
    if second is not None and second.grad is not None:
        arrays.update(dsecond=second.grad.numpy())
    for k, v in m.named_parameters():
        if v.grad is not None:
            arrays["grad/" + k] = v.grad.numpy()
    save(name, **arrays)


def metrics_case():
    """Counts from the reference's own float2int / _cal_frame (datasets/Shanghai_metrics.py:45-47,105-114).
    The class constructor needs `lpips` (absent, downloads weights) so the two methods are called unbound."""
    import types
    sys.modules.setdefault("lpips", types.ModuleType("lpips"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_shanghai_metrics", "/root/reference/datasets/Shanghai_metrics.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    SimplifiedEvaluator = mod.SimplifiedEvaluator
    ev = object.__new__(SimplifiedEvaluator)
    ev.value_scale = 90
    obs, sim = cases.metric_inputs()
    o, s = ev.float2int(obs), ev.float2int(sim)
    table = np.zeros((4, 4), dtype=np.int64)
    for i, thr in enumerate([20, 30, 35, 40]):
        for b in range(obs.shape[0]):
            for t in range(obs.shape[1]):
                table[i] += np.array(ev._cal_frame(o[b][t], s[b][t], thr), dtype=np.int64)
    save(cases.METRIC_CASE[0], table=table, obs_int_checksum=np.int64(o.astype(np.int64).sum()),
         sim_int_checksum=np.int64(s.astype(np.int64).sum()))


def main():
    ref = load_reference()
    only = sys.argv[1:]
    for name in cases.MIXER_CASES:
        if not only or name in only:
            mixer_case(ref, name)
    for name in cases.MIXER_BF16_CASES:
        if not only or name in only:
            mixer_bf16_case(ref, name)
    for name in cases.MIXER_CORR_CASES:
        if not only or name in only:
            mixer_corr_case(ref, name)
    for name in cases.WTCONV_CASES:
        if not only or name in only:
            wtconv_case(ref, name)
    for name in cases.BLOCK_CASES:
        if not only or name in only:
            block_case(ref, name)
    for name in cases.CONVSTAGE_CASES:
        if not only or name in only:
            convstage_case(ref, name)
    if not only or cases.METRIC_CASE[0] in only:
        metrics_case()


if __name__ == "__main__":
    main()
