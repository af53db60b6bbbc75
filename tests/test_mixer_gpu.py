"""GPU parity of the ADN-SSD mixer (CUDA library through the C ABI) against the golden vectors produced by the
reference and against the CPU oracle.  Tolerances (BASELINE.json north_star), measured per tensor as
max|a-b| / max|b|:  1e-4 in fp32 check mode, 2e-2 in bf16."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import adnssd_oracle as AO

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}
# The fp64-verified goldens MIXER_CASES perturb EVERY parameter by N(0, 0.3^2) (in_proj 15x its init scale) so that no
# term of the check-mode comparison vanishes.  In that regime bf16 storage of the in_proj / conv outputs alone moves du by
# up to 2.1e-2 (CPU emulation with everything else in fp64, DESIGN.md "bf16 error budget"), so the bf16 bound there is a
# sanity bound; the 2e-2 contract is asserted on the model-scale goldens MIXER_BF16_CASES.
BF16_STRESS_TOL = 1e-1


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def load_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    params = {k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    return z, params, grads


def run_cuda(params, u, dout, g_h, g_w, P, N, dtype):
    import adnm_unet_b200 as A
    dev = torch.device("cuda:0")
    p = {k: v.to(dev).float().requires_grad_(k not in AO.UNUSED_PARAMS) for k, v in params.items()}
    ud = u.to(dev, dtype).requires_grad_(True)
    out = A.adnssd_mixer(ud, g_h, g_w, p, headdim=P, d_state=N)
    out.backward(dout.to(dev, dtype))
    torch.cuda.synchronize()
    return out, ud.grad, {k: v.grad for k, v in p.items() if v.grad is not None}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(cases.MIXER_CASES))
def test_mixer_matches_reference_golden(golden_dir, name, dtype):
    D, P, N, B, g = cases.MIXER_CASES[name]
    z, params, grads = load_case(golden_dir, name)
    u, dout = cases.mixer_inputs(name, torch.float32)
    out, du, pg = run_cuda(params, u, dout, g, g, P, N, dtype)
    s = cases.SUBSAMPLE_STRIDE if g >= 64 else 1
    tol = TOL[dtype] if dtype == torch.float32 else BF16_STRESS_TOL
    errs = {"out": rel(out[:, ::s], z["out"]), "du": rel(du[:, ::s], z["du"])}
    assert set(pg) == set(grads)
    for k, ref in grads.items():
        errs[k] = rel(pg[k].reshape(ref.shape), ref)
    # alpha1 in the bf16 stress regime: d alpha1 = <dout, out> / alpha1 with an INDEPENDENT random dout is a sum of signed
    # terms that cancels to ~1e-2 of its terms' magnitude, so even its sanity bound is wider; alpha1 is graded at 2e-2 with
    # the common metric wherever dout is correlated with the output (MIXER_CORR_CASES goldens, wide / row-kernel tests)
    bad = {k: v for k, v in errs.items() if not v < (tol if (k != "alpha1" or dtype == torch.float32) else 0.3)}
    assert not bad, f"{name} {dtype}: {bad} (all: {errs})"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(cases.MIXER_CORR_CASES))
def test_mixer_matches_reference_golden_headline_shapes(golden_dir, name, dtype):
    """BASELINE configs[1] shapes (128 x 128 tokens; d_model 32 = row kernels, d_model 128 = wide path) against goldens of
    the unmodified reference, upstream gradient correlated with the output: 2e-2 in bf16 and 1e-4 in fp32 on the output,
    du and EVERY parameter gradient incl. alpha1, all with the common metric."""
    D, P, N, B, g, _ = cases.MIXER_CORR_CASES[name]
    z, params, grads = load_case(golden_dir, name)
    u = cases.mixer_corr_u(name, torch.float32)
    out_ref = AO.mixer_forward({k: v.double() for k, v in params.items()}, u.double(), g, g, P, N)
    dout = cases.mixer_corr_dout(name, out_ref)
    out, du, pg = run_cuda(params, u, dout, g, g, P, N, dtype)
    s = cases.SUBSAMPLE_STRIDE
    errs = {"out": rel(out[:, ::s], z["out"]), "du": rel(du[:, ::s], z["du"]), "out_full_vs_oracle": rel(out, out_ref)}
    assert set(pg) == set(grads)
    for k, ref in grads.items():
        errs[k] = rel(pg[k].reshape(ref.shape), ref)
    bad = {k: v for k, v in errs.items() if not v < TOL[dtype]}
    assert not bad, f"{name} {dtype}: {bad} (all: {errs})"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(cases.MIXER_BF16_CASES))
def test_mixer_matches_reference_golden_model_scale(golden_dir, name, dtype):
    """The contract of BASELINE.json: <= 2e-2 in bf16, <= 1e-4 in fp32 check mode, per output / gradient tensor."""
    D, P, N, B, g, _ = cases.MIXER_BF16_CASES[name]
    z, params, grads = load_case(golden_dir, name)
    u, dout = cases.mixer_bf16_inputs(name, torch.float32)
    out, du, pg = run_cuda(params, u, dout, g, g, P, N, dtype)
    tol = TOL[dtype]
    errs = {"out": rel(out, z["out"]), "du": rel(du, z["du"])}
    assert set(pg) == set(grads)
    for k, ref in grads.items():
        errs[k] = rel(pg[k].reshape(ref.shape), ref)
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, f"{name} {dtype}: {bad} (all: {errs})"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [(16, 4, 8, 3, 5, 9), (32, 4, 16, 1, 33, 17), (48, 8, 16, 2, 7, 7), (256, 4, 16, 2, 4, 4),
                                 (32, 4, 32, 2, 16, 16), (32, 4, 64, 2, 8, 32)],   # d_state 32 / 64: tcgen05 tile kernels
                         ids=lambda c: "D%d_P%d_N%d_B%d_%dx%d" % c)
def test_mixer_matches_oracle_ragged_shapes(cfg, dtype):
    """Non-square and odd token grids, tiny and large d_model: compared with the CPU oracle in fp64."""
    D, P, N, B, H, W = cfg
    params = AO.init_params(D, P, N, seed=7, perturb=0.3, dtype=torch.float32)
    u = cases.rng_normal(11, (B, H * W, D), torch.float32)
    dout = cases.rng_normal(12, (B, H * W, D), torch.float32)
    p64 = {k: v.double() for k, v in params.items()}
    ref_out = AO.mixer_forward(p64, u.double(), H, W, P, N)
    ref_du, ref_g = AO.mixer_backward(p64, u.double(), H, W, P, N, dout.double())
    out, du, pg = run_cuda(params, u, dout, H, W, P, N, dtype)
    tol = TOL[dtype] if dtype == torch.float32 else BF16_STRESS_TOL   # perturb=0.3 stress regime, see above
    errs = {"out": rel(out, ref_out), "du": rel(du, ref_du)}
    for k, ref in ref_g.items():
        errs[k] = rel(pg[k].reshape(ref.shape), ref)
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, f"{cfg} {dtype}: {bad}"


@pytest.mark.parametrize("cfg", [(3, 5, 128), (2, 1, 128), (1, 2, 128), (5, 37, 128), (1, 3, 256), (2, 5, 256), (1, 2, 384)],
                         ids=lambda c: "B%d_%dx%d" % c)
def test_row_kernels_match_oracle(cfg):
    """Grids whose width is a multiple of 128 take the conv-as-GEMM row kernels (adnssd_rowconv.cuh): several samples per
    CTA range, H = 1 (no vertical taps), H = 2, a CTA range that ends mid-sample, and 256 / 384-wide grids (processed as
    128-wide strip images with neighbour tokens across the strip edges); model-scale weights, bf16 contract 2e-2."""
    B, H, W = cfg
    D, P, N = 32, 4, 16
    params = AO.init_params(D, P, N, seed=9, perturb=0.05, dtype=torch.float32)
    u = cases.bf16_exact(cases.rng_normal(21, (B, H * W, D), torch.float32))      # identical inputs for both sides
    p64 = {k: v.double() for k, v in params.items()}
    ref_out = AO.mixer_forward(p64, u.double(), H, W, P, N)
    dout = cases.corr_dout(ref_out, 22)      # correlated with the output: alpha1 is graded with the common metric
    ref_du, ref_g = AO.mixer_backward(p64, u.double(), H, W, P, N, dout.double())
    out, du, pg = run_cuda(params, u, dout, H, W, P, N, torch.bfloat16)
    errs = {"out": rel(out, ref_out), "du": rel(du, ref_du)}
    for k, ref in ref_g.items():
        errs[k] = rel(pg[k].reshape(ref.shape), ref)
    bad = {k: v for k, v in errs.items() if not v < 2e-2}
    assert not bad, f"{cfg}: {bad} (all: {errs})"


WIDE_CASES = [  # (D, P, N, B, H, W): the six encoder / decoder mixers of ADNM-UNet at 128^2 and 256^2 images + sweep corners
    (128, 4, 16, 3, 16, 16), (256, 4, 16, 2, 8, 8), (512, 4, 16, 2, 4, 4), (1024, 4, 16, 2, 4, 4), (1024, 4, 16, 1, 8, 8),
    (512, 4, 16, 1, 16, 16), (128, 4, 16, 1, 32, 32), (64, 8, 128, 2, 6, 6), (128, 4, 64, 2, 7, 9), (32, 4, 128, 2, 16, 16),
    (64, 4, 16, 1, 48, 48),
]


@pytest.mark.parametrize("cfg", WIDE_CASES, ids=lambda c: "D%d_P%d_N%d_B%d_%dx%d" % c)
def test_wide_path_matches_oracle(cfg):
    """d_model 64 ... 1024 and d_state up to 128 take the wide path (tcgen05 GEMMs, csrc/adnssd_wide.cuh): model-scale
    weights, bf16 contract 2e-2 on the output, du and EVERY parameter gradient with the common metric."""
    from adnm_unet_b200 import _lib
    D, P, N, B, H, W = cfg
    shape = _lib.AdnShape(B=B, H=H, W=W, D=D, Di=2 * D, P=P, G=2, N=N, dtype=_lib.ADN_BF16, flags=0)
    assert _lib.load().adnssd_kernel_family(shape) == 3, "shape is not served by the wide path"
    params = AO.init_params(D, P, N, seed=9, perturb=0.05, dtype=torch.float32)
    u = cases.bf16_exact(cases.rng_normal(31, (B, H * W, D), torch.float32))      # identical inputs for both sides
    p64 = {k: v.double() for k, v in params.items()}
    ref_out = AO.mixer_forward(p64, u.double(), H, W, P, N)
    dout = cases.corr_dout(ref_out, 32)
    ref_du, ref_g = AO.mixer_backward(p64, u.double(), H, W, P, N, dout.double())
    out, du, pg = run_cuda(params, u, dout, H, W, P, N, torch.bfloat16)
    errs = {"out": rel(out, ref_out), "du": rel(du, ref_du)}
    for k, ref in ref_g.items():
        errs[k] = rel(pg[k].reshape(ref.shape), ref)
    bad = {k: v for k, v in errs.items() if not v < 2e-2}
    assert not bad, f"{cfg}: {bad} (all: {errs})"


def test_every_mixer_of_the_reference_network_is_on_tensor_cores():
    """`sm100_supported` for all ten mixers of create_ADNMUNet (SURVEY.md 3.2 instance table) at 128^2 and
    256^2 images: the four refiner mixers on the fused d_model-32 kernels, the six encoder / decoder mixers on the wide path."""
    from adnm_unet_b200 import _lib
    lib = _lib.load()
    for img in (128, 256):
        inst = [(128, img // 8), (256, img // 16), (512, img // 32), (1024, img // 32), (1024, img // 16), (512, img // 8)] + [(32, img)] * 4
        for D, g in inst:
            shape = _lib.AdnShape(B=2, H=g, W=g, D=D, Di=2 * D, P=4, G=2, N=16, dtype=_lib.ADN_BF16, flags=0)
            fam = lib.adnssd_kernel_family(shape)
            assert fam == (2 if D == 32 else 3), (img, D, g, fam)


def test_module_is_a_drop_in(golden_dir):
    """Strict state_dict load of reference-shaped weights, forward(u, H, W), grads land on the same 18 parameters."""
    import adnm_unet_b200 as A
    name = "mixer_d32_p4_n16_g16"
    D, P, N, B, g = cases.MIXER_CASES[name]
    z, params, grads = load_case(golden_dir, name)
    m = A.Mamba2(d_model=D, headdim=P, d_state=N, layer_idx=3, linear_attn_duality=True)
    missing = m.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert list(m.state_dict()) == list(AO.PARAM_NAMES)
    m = m.cuda()
    u, dout = cases.mixer_inputs(name, torch.float32)
    ud = u.cuda().requires_grad_(True)
    out = m(ud, g, g)
    assert out.shape == ud.shape and out.is_contiguous()
    out.backward(dout.cuda())
    assert rel(out, z["out"]) < 1e-4 and rel(ud.grad, z["du"]) < 1e-4
    for k, p in m.named_parameters():
        if k in AO.UNUSED_PARAMS:
            assert p.grad is None, k
        else:
            assert rel(p.grad, grads[k]) < 1e-4, k
    # inference path (no saved buffers) gives the same output
    with torch.no_grad():
        out2 = m(u.cuda(), g, g)
    assert rel(out2, out.detach().cpu()) < 1e-5   # fp32 atomics in the state reduction: not bit-reproducible
    # autocast -> bf16 compute
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out3 = m(u.cuda(), g, g)
    assert out3.dtype == torch.bfloat16 and rel(out3, z["out"]) < 2e-2


def test_errors_are_loud():
    import adnm_unet_b200 as A
    params = AO.init_params(32, 4, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        A.adnssd_mixer(torch.zeros(1, 16, 32), 4, 4, params, headdim=4, d_state=16)
    pc = {k: v.cuda() for k, v in params.items()}
    with pytest.raises(RuntimeError, match="H\\*W"):
        A.adnssd_mixer(torch.zeros(1, 16, 32, device="cuda"), 4, 5, pc, headdim=4, d_state=16)
    with pytest.raises(RuntimeError, match="ngroups"):
        A.adnssd_mixer(torch.zeros(1, 16, 32, device="cuda"), 4, 4, pc, headdim=4, d_state=16, ngroups=4)
    with pytest.raises(RuntimeError, match="dtype"):
        A.adnssd_mixer(torch.zeros(1, 16, 32, device="cuda", dtype=torch.float16), 4, 4, pc, headdim=4, d_state=16)


def _rand_params(D, P, N, dev):
    return {k: v.to(dev).requires_grad_(k not in AO.UNUSED_PARAMS) for k, v in AO.init_params(D, P, N, seed=21, perturb=0.05).items()}


@pytest.mark.parametrize("B,g", [(16, 128), (2, 256)], ids=["bench_shape_B16_128x128", "refiner_256x256"])
def test_full_size_fast_path_agrees_with_check_mode(B, g):
    """BASELINE full sizes (configs[1] and the 256x256 refiner grid): the bf16 tcgen05 path against the fp32 check-mode
    path of the same library on the same inputs (the oracle itself would take minutes at this size), plus batch
    independence (a size-independent property of the path: every op is per-sample)."""
    import adnm_unet_b200 as A
    dev = torch.device("cuda:0")
    D, P, N = 32, 4, 16
    gen = torch.Generator().manual_seed(5)
    u = torch.randn(B, g * g, D, generator=gen)
    with torch.no_grad():      # check-mode forward first: the upstream gradient is correlated with the output (cases.corr_dout)
        out0 = A.adnssd_mixer(u.to(dev), g, g, _rand_params(D, P, N, dev), headdim=P, d_state=N).cpu()
    dout = (cases.CORR_FRACTION * out0 / out0.std() + torch.randn(B, g * g, D, generator=gen)).float()
    res = {}
    for dtype in (torch.float32, torch.bfloat16):
        p = _rand_params(D, P, N, dev)
        ud = u.to(dev, dtype).requires_grad_(True)
        out = A.adnssd_mixer(ud, g, g, p, headdim=P, d_state=N)
        out.backward(dout.to(dev, dtype))
        torch.cuda.synchronize()
        res[dtype] = (out.detach().float().cpu(), ud.grad.float().cpu(), {k: v.grad.float().cpu() for k, v in p.items() if v.grad is not None})
        del out, ud, p
        torch.cuda.empty_cache()
    ref, fast = res[torch.float32], res[torch.bfloat16]
    errs = {"out": rel(fast[0], ref[0]), "du": rel(fast[1], ref[1])}
    for k in ref[2]:
        errs[k] = rel(fast[2][k], ref[2][k])
    bad = {k: v for k, v in errs.items() if not v < 2e-2}
    assert not bad, f"{bad} (all: {errs})"
    # batch independence: sample 1 alone gives the same rows as inside the batch
    p = _rand_params(D, P, N, dev)
    with torch.no_grad():
        alone = A.adnssd_mixer(u[1:2].to(dev, torch.bfloat16), g, g, p, headdim=P, d_state=N).float().cpu()
    assert rel(alone[0], fast[0][1]) < 5e-3


def test_row_kernels_agree_with_tile_kernels():
    """The two sm_100a kernel families (conv-as-GEMM row kernels + warp-specialised backward vs. the halo-tile kernels)
    on the same 128-wide grid: independent implementations of the same math, selected with adn_set_option("rowconv", .) (the ADN_ROWCONV
    environment switch is read once at load, so tests flip the option through the ABI between whole fwd+bwd passes)."""
    B, H, W, D, P, N = 3, 9, 128, 32, 4, 16
    params = AO.init_params(D, P, N, seed=31, perturb=0.05, dtype=torch.float32)
    u = cases.rng_normal(41, (B, H * W, D), torch.float32)
    dout = cases.rng_normal(42, (B, H * W, D), torch.float32)
    res = {}
    from adnm_unet_b200 import _lib
    for flag in ("1", "0"):
        _lib.check(_lib.load().adn_set_option(b"rowconv", int(flag)), "adn_set_option")
        try:
            res[flag] = run_cuda(params, u, dout, H, W, P, N, torch.bfloat16)
        finally:
            _lib.load().adn_set_option(b"rowconv", 1)
    out1, du1, g1 = res["1"]
    out0, du0, g0 = res["0"]
    errs = {"out": rel(out1, out0.float().cpu()), "du": rel(du1, du0.float().cpu())}
    for k in g0:
        if k != "alpha1":      # random dout here: cancelling scalar, graded in test_row_kernels_match_oracle
            errs[k] = rel(g1[k], g0[k].float().cpu())
    bad = {k: v for k, v in errs.items() if not v < 2e-2}
    assert not bad, f"{bad} (all: {errs})"


def test_module_step_is_cuda_graph_capturable():
    """bench.py replays a captured fwd+bwd of the public module: a replay on new contents of the static input must
    reproduce the eager result (the C ABI enqueues on the caller's stream, allocates nothing and never synchronises)."""
    import adnm_unet_b200 as A
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    m = A.Mamba2(d_model=32, headdim=4, d_state=16).to(dev)
    params = [p for n, p in m.named_parameters() if n not in ("scale", "shift", "alpha2")]
    H, W = 4, 128
    u = torch.randn(2, H * W, 32, device=dev, dtype=torch.bfloat16, requires_grad=True)
    go = torch.randn_like(u)

    def clear():
        u.grad = None
        for p in params:
            p.grad = None

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):                     # warm-up outside capture
        for _ in range(2):
            clear()
            m(u, H, W).backward(go)
    torch.cuda.current_stream().wait_stream(s)
    clear()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = m(u, H, W)
        out.backward(go)
    new_u = torch.randn(2, H * W, 32, generator=torch.Generator().manual_seed(99)).to(dev, torch.bfloat16)
    with torch.no_grad():
        u.copy_(new_u)
    graph.replay()
    torch.cuda.synchronize()
    g_out, g_du, g_par = out.detach().clone(), u.grad.detach().clone(), [p.grad.detach().clone() for p in params]
    # eager reference on the same input
    ue = new_u.clone().requires_grad_(True)
    for p in params:
        p.grad = None
    oe = m(ue, H, W)
    oe.backward(go)
    torch.cuda.synchronize()
    # fp32 atomics in the state reduction make two runs differ by an ulp of the bf16 outputs (2^-8 of the value)
    assert rel(g_out, oe.detach().float().cpu()) < 4e-3
    assert rel(g_du, ue.grad.float().cpu()) < 4e-3
    for gp, p in zip(g_par, params):
        assert rel(gp, p.grad.float().cpu()) < 1e-2


def test_graphed_mixer_matches_eager_and_is_faster_to_issue():
    """adnm_unet_b200.graphed_mixer: make_graphed_callables around the module; same numbers as the eager call (up to the
    fp32-atomic ulp), gradients on the same parameters."""
    import adnm_unet_b200 as A
    torch.manual_seed(4)
    dev = torch.device("cuda:0")
    m = A.Mamba2(d_model=32, headdim=4, d_state=16).to(dev)
    H, W = 6, 128
    u0 = torch.randn(2, H * W, 32, device=dev, dtype=torch.bfloat16)
    f = A.graphed_mixer(m, u0, H, W)
    u = torch.randn_like(u0).requires_grad_(True)
    go = torch.randn_like(u0)
    for p in m.parameters():
        p.grad = None
    out_g = f(u)
    out_g.backward(go)
    torch.cuda.synchronize()
    g_du = u.grad.detach().clone()
    g_par = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    assert set(g_par) == set(n for n, _ in m.named_parameters()) - {"scale", "shift", "alpha2"}
    ue = u.detach().clone().requires_grad_(True)
    for p in m.parameters():
        p.grad = None
    out_e = m(ue, H, W)
    out_e.backward(go)
    torch.cuda.synchronize()
    assert rel(out_g, out_e.detach().float().cpu()) < 4e-3
    assert rel(g_du, ue.grad.float().cpu()) < 4e-3
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert rel(g_par[n], p.grad.float().cpu()) < 1e-2, n
