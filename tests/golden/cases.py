"""Golden-case definitions shared by make_golden.py (generator) and the tests (consumers).

Inputs are drawn from numpy's PCG64 streams (stable across platforms), never from /root/reference,
so every test can rebuild them on the GPU box.
"""
import numpy as np
import torch

# name -> (d_model, headdim, d_state, batch, grid)   [SURVEY.md §8(c): the four fp64-verified configs + a 128^2 one]
MIXER_CASES = {
    "mixer_d32_p4_n16_g16": (32, 4, 16, 2, 16),
    "mixer_d128_p4_n16_g8": (128, 4, 16, 2, 8),
    "mixer_d32_p4_n64_g8": (32, 4, 64, 2, 8),
    "mixer_d64_p8_n128_g6": (64, 8, 128, 2, 6),
    "mixer_d32_p4_n16_g128": (32, 4, 16, 1, 128),   # refiner shape at a 128x128 token grid; stored subsampled
}
SUBSAMPLE_STRIDE = 61  # token stride used to store `out`/`du` of the 128^2 case

# name -> (C, k, levels, batch, H, W, bias)   [instances of models/ADNMUNet.py at reduced size + the odd-size pad path]
WTCONV_CASES = {
    "wtconv_c5_k5_l3_64": (5, 5, 3, 2, 64, 64, True),
    "wtconv_c32_k5_l2_32": (32, 5, 2, 2, 32, 32, True),
    "wtconv_c64_k5_l1_16": (64, 5, 1, 1, 16, 16, True),
    "wtconv_c32_k5_l3_33": (32, 5, 3, 1, 33, 33, True),     # odd sizes at every level (models/WTConv2d.py:114-116)
    "wtconv_c8_k3_l2_20x28": (8, 3, 2, 2, 20, 28, False),   # non-square, k=3, no bias
}

METRIC_CASE = ("metrics_counts", (3, 20, 64, 64))


def rng_normal(seed, shape, dtype=torch.float64):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape)).to(dtype)


def rng_uniform(seed, shape, dtype=torch.float64):
    return torch.from_numpy(np.random.default_rng(seed).random(shape)).to(dtype)


def mixer_inputs(name, dtype=torch.float64):
    D, P, N, B, g = MIXER_CASES[name]
    seed = 1000 + sorted(MIXER_CASES).index(name)
    u = rng_normal(seed, (B, g * g, D), dtype)
    dout = rng_normal(seed + 500, (B, g * g, D), dtype)
    return u, dout


def wtconv_inputs(name, dtype=torch.float64):
    C, k, L, B, H, W, bias = WTCONV_CASES[name]
    seed = 2000 + sorted(WTCONV_CASES).index(name)
    return rng_normal(seed, (B, C, H, W), dtype), rng_normal(seed + 500, (B, C, H, W), dtype)


def metric_inputs():
    shape = METRIC_CASE[1]
    # uniform in [-0.1, 1.1]: exercises the clip, every threshold (20..40)/90, and values at exact k/90 boundaries
    obs = rng_uniform(3000, shape, torch.float32) * 1.2 - 0.1
    sim = rng_uniform(3001, shape, torch.float32) * 1.2 - 0.1
    obs.view(-1)[::97] = torch.tensor([20, 30, 35, 40], dtype=torch.float32).repeat(obs.numel())[: obs.view(-1)[::97].numel()] / 90
    sim.view(-1)[::89] = torch.tensor([40, 35, 30, 20], dtype=torch.float32).repeat(sim.numel())[: sim.view(-1)[::89].numel()] / 90
    return obs.numpy(), sim.numpy()

# bf16 parity regime: the model's own initialisation scale with a mild perturbation so that
# no term vanishes.  name -> (d_model, headdim, d_state, batch, grid, perturb)
MIXER_BF16_CASES = {
    "mixerinit_d32_p4_n16_g32": (32, 4, 16, 2, 32, 0.1),
    "mixerinit_d128_p4_n16_g16": (128, 4, 16, 2, 16, 0.1),
    "mixerinit_d32_p4_n64_g16": (32, 4, 64, 1, 16, 0.05),
}


def mixer_bf16_inputs(name, dtype=torch.float64):
    D, P, N, B, g, _ = MIXER_BF16_CASES[name]
    seed = 4000 + sorted(MIXER_BF16_CASES).index(name)
    return rng_normal(seed, (B, g * g, D), dtype), rng_normal(seed + 500, (B, g * g, D), dtype)


# Headline-shape goldens with an upstream gradient CORRELATED with the output (parity gaps 1-3): 128 x 128 token
# grids at the refiner width (d_model 32, the BASELINE configs[1] primary shape) and at the encoder4 width (d_model 128,
# configs[1] secondary), model-scale parameters.  dout = 0.1 * out / std(out) + N(0, 1): the correlated tenth keeps scalar
# gradients such as d alpha1 = <dout, out> / alpha1 from cancelling to ~0 (so they are graded with the common metric);
# u and dout are bf16-representable (identical inputs for the bf16 path and the fp64 reference);
# a FULLY correlated dout would make every gradient hypersensitive to the bf16 rounding of the inputs themselves
# (rounding u alone then moves du by 3e-2 at d_model 1024 - measured with an fp64 emulation).
# The test rebuilds dout from the oracle's fp64 forward (pinned to the reference at 1e-12), so only subsampled out / du
# and the parameter gradients are stored.    name -> (d_model, headdim, d_state, batch, grid, perturb)
MIXER_CORR_CASES = {
    "mixercorr_d32_p4_n16_g128": (32, 4, 16, 2, 128, 0.1),
    "mixercorr_d128_p4_n16_g128": (128, 4, 16, 1, 128, 0.05),
}
CORR_FRACTION = 0.1


def mixer_corr_u(name, dtype=torch.float64):
    D, P, N, B, g, _ = MIXER_CORR_CASES[name]
    seed = 5000 + sorted(MIXER_CORR_CASES).index(name)
    return bf16_exact(rng_normal(seed, (B, g * g, D), torch.float32)).to(dtype)


def bf16_exact(t):
    """Round to bf16-representable values (kept in the input dtype): the parity contract is about IDENTICAL inputs, so the
    activations handed to the bf16 path and to the fp64 reference / oracle are the same numbers."""
    return t.bfloat16().to(t.dtype)


def corr_dout(out_ref, seed, frac=CORR_FRACTION):
    """bf16-representable upstream gradient from an fp64 reference output (same formula in generator and tests)."""
    noise = rng_normal(seed, tuple(out_ref.shape), torch.float64)
    return bf16_exact((frac * out_ref.double() / out_ref.double().std() + noise).float())


def mixer_corr_dout(name, out_ref):
    seed = 5500 + sorted(MIXER_CORR_CASES).index(name)
    return corr_dout(out_ref, seed)


# Block goldens (SURVEY.md 8(c) golden (2)): the reference `Block` (models/ADNMUNet.py:49-165) as `create_block` builds it
# (headdim 4, d_state 16, RMSNorm eps 1e-6), perturbed parameters (oracle.block_oracle.init_block_params), fp64.
# name -> (dim, out_dim, batch, grid, skip): skip = called with `residual` and `features` (the decoder call, :124-129)
BLOCK_CASES = {
    "block_d32_o32_g16": (32, 32, 2, 16, False),
    "block_d64_o128_g8": (64, 128, 1, 8, False),
    "block_d64_o32_g8_skip": (64, 32, 2, 8, True),
}


def block_inputs(name, dtype=torch.float64):
    dim, out_dim, B, g, skip = BLOCK_CASES[name]
    seed = 6000 + 10 * sorted(BLOCK_CASES).index(name)
    half = dim // 2 if skip else dim
    x = bf16_exact(rng_normal(seed, (B, g * g, half), torch.float32)).to(dtype)
    res = bf16_exact(rng_normal(seed + 1, (B, g * g, half), torch.float32)).to(dtype) if skip else None
    feat = bf16_exact(rng_normal(seed + 2, (B, g * g, half), torch.float32)).to(dtype) if skip else None
    return x, res, feat


def block_dout(name, out_ref):
    seed = 6500 + sorted(BLOCK_CASES).index(name)
    return corr_dout(out_ref, seed, frac=1.0) / out_ref.numel()


# Conv-stage goldens: the reference `WTLayer`, `PatchEmbed`, `OutProj` (models/model_untils.py:226-426,799-892)
# with InstanceNorm=True and perturbed parameters (oracle.convstage_oracle.perturb_params), fp64, bf16-representable inputs.
# name -> (kind, ctor kwargs, batch, grid, skip)
CONVSTAGE_CASES = {
    "wtlayer_d32_o64_g16": ("WTLayer", dict(this_dim=32, next_dim=64, kernel=5, wt_levels=2), 2, 16, False),
    "wtlayer_d64_o32_g16_skip": ("WTLayer", dict(this_dim=64, next_dim=32, kernel=5, wt_levels=3, if_res=True), 2, 16, True),
    "patchembed_c5_e32_g32": ("PatchEmbed", dict(img_size=32, patch_size=2, in_channels=5, embed_dim=32, kernel=5, wt_levels=3), 2, 32, False),
    "outproj_e32_f20_g32": ("OutProj", dict(num_frames=20, embed_dim=32, img_size=[32, 32], wt_levels=3, out_expand=2), 2, 32, False),
}


def convstage_inputs(name, dtype=torch.float64):
    """(x, second): second = residual tokens (WTLayer skip), the residual frame (OutProj) or None."""
    kind, kw, B, g, skip = CONVSTAGE_CASES[name]
    seed = 7000 + 10 * sorted(CONVSTAGE_CASES).index(name)
    if kind == "WTLayer":
        c = kw["this_dim"] // 2 if skip else kw["this_dim"]
        x = bf16_exact(rng_normal(seed, (B, g * g, c), torch.float32)).to(dtype)
        second = bf16_exact(rng_normal(seed + 1, (B, g * g, c), torch.float32)).to(dtype) if skip else None
    elif kind == "PatchEmbed":
        x = bf16_exact(rng_uniform(seed, (B, g * g, kw["in_channels"]), torch.float32)).to(dtype)
        second = None
    else:
        x = bf16_exact(rng_normal(seed, (B, g * g, kw["embed_dim"]), torch.float32)).to(dtype)
        second = bf16_exact(rng_uniform(seed + 1, (B, g, g), torch.float32)).to(dtype)
    return x, second


def convstage_dout(name, out_ref):
    seed = 7500 + sorted(CONVSTAGE_CASES).index(name)
    return corr_dout(out_ref, seed, frac=1.0) / out_ref.numel()
