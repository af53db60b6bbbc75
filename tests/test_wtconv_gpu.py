"""GPU parity of the Haar WTConv2d kernels (through the C ABI) against reference-generated goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import wtconv_oracle as WO

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def run_cuda(params, x, dy, k, levels, dtype):
    import adnm_unet_b200 as A
    dev = torch.device("cuda:0")
    p = {n: v.to(dev).float().requires_grad_(n not in ("wt_filter", "iwt_filter")) for n, v in params.items()}
    xd = x.to(dev, dtype).requires_grad_(True)
    y = A.wtconv2d(xd, p, k, levels)
    y.backward(dy.to(dev, dtype))
    torch.cuda.synchronize()
    return y, xd.grad, {n: v.grad for n, v in p.items() if v.grad is not None}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("name", sorted(cases.WTCONV_CASES))
def test_wtconv_matches_reference_golden(golden_dir, name, dtype):
    C, k, L, B, H, W, bias = cases.WTCONV_CASES[name]
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    params = {n[6:]: torch.from_numpy(z[n]) for n in z.files if n.startswith("param/")}
    grads = {n[5:]: torch.from_numpy(z[n]) for n in z.files if n.startswith("grad/")}
    x, dy = cases.wtconv_inputs(name, torch.float32)
    y, dx, pg = run_cuda(params, x, dy, k, L, dtype)
    tol = TOL[dtype]
    errs = {"out": rel(y, z["out"]), "dx": rel(dx, z["dx"])}
    assert set(pg) == set(grads)
    for n, ref in grads.items():
        errs[n] = rel(pg[n], ref)
    bad = {n: v for n, v in errs.items() if not v < tol}
    assert not bad, f"{name} {dtype}: {bad}"


@pytest.mark.parametrize("cfg", [(3, 5, 3, 1, 9, 70), (4, 7, 2, 2, 37, 18), (2, 1, 1, 1, 5, 5), (16, 3, 4, 1, 128, 128)],
                         ids=lambda c: "C%d_k%d_L%d_B%d_%dx%d" % c)
def test_wtconv_matches_oracle_ragged(cfg):
    C, k, L, B, H, W = cfg
    params = WO.init_params(C, k, L, bias=True, seed=4, dtype=torch.float32)
    x = cases.rng_normal(31, (B, C, H, W), torch.float32)
    dy = cases.rng_normal(32, (B, C, H, W), torch.float32)
    ref_y, ref_dx, ref_g = WO.wtconv_forward_backward({n: v.double() for n, v in params.items()}, x.double(), L, dy.double())
    y, dx, pg = run_cuda(params, x, dy, k, L, torch.float32)
    errs = {"out": rel(y, ref_y), "dx": rel(dx, ref_dx)}
    for n, ref in ref_g.items():
        errs[n] = rel(pg[n], ref)
    bad = {n: v for n, v in errs.items() if not v < 1e-4}
    assert not bad, f"{cfg}: {bad}"


# every WTConv2d instance of ADNM-UNet at a 128 x 128 input (create_ADNMUNet(5, 20, 6): C, levels, plane, bias) and the two
# full-resolution ones at the reference's native 256 x 256, B = 2, k = 5
MODEL_INSTANCES = [(5, 3, 128, False), (32, 3, 128, False), (32, 2, 64, True), (64, 1, 32, True), (256, 1, 32, True), (128, 2, 64, True),
                   (64, 3, 128, True), (5, 3, 256, False), (64, 3, 256, True)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("cfg", MODEL_INSTANCES, ids=lambda c: "C%d_L%d_%d_bias%d" % c)
def test_wtconv_model_instance_sizes_match_oracle(cfg, dtype):
    C, L, G, bias = cfg
    B = 2 if G <= 128 else 1
    params = WO.init_params(C, 5, L, bias=bias, seed=6, dtype=torch.float32)
    x = cases.rng_normal(41, (B, C, G, G), torch.float32).to(dtype).float()
    dy = cases.rng_normal(42, (B, C, G, G), torch.float32).to(dtype).float()
    ref_y, ref_dx, ref_g = WO.wtconv_forward_backward({n: v.double() for n, v in params.items()}, x.double(), L, dy.double())
    y, dx, pg = run_cuda(params, x, dy, 5, L, dtype)
    errs = {"out": rel(y, ref_y), "dx": rel(dx, ref_dx)}
    for n, ref in ref_g.items():
        errs[n] = rel(pg[n], ref)
    bad = {n: v for n, v in errs.items() if not v < TOL[dtype]}
    assert not bad, f"{cfg} {dtype}: {bad}"


@pytest.mark.parametrize("cfg", [(3, 5, 3, 1, 9, 70), (4, 7, 2, 2, 37, 18), (6, 3, 2, 1, 40, 24), (16, 3, 4, 1, 128, 128)],
                         ids=lambda c: "C%d_k%d_L%d_B%d_%dx%d" % c)
def test_wtconv_matches_oracle_ragged_bf16(cfg):
    """bf16 storage on ragged planes (odd sizes -> the per-thread gather path; 16-byte aligned rows -> TMA tiles)."""
    C, k, L, B, H, W = cfg
    params = WO.init_params(C, k, L, bias=True, seed=4, dtype=torch.float32)
    x = cases.rng_normal(31, (B, C, H, W), torch.float32).bfloat16().float()
    dy = cases.rng_normal(32, (B, C, H, W), torch.float32).bfloat16().float()
    ref_y, ref_dx, ref_g = WO.wtconv_forward_backward({n: v.double() for n, v in params.items()}, x.double(), L, dy.double())
    y, dx, pg = run_cuda(params, x, dy, k, L, torch.bfloat16)
    errs = {"out": rel(y, ref_y), "dx": rel(dx, ref_dx)}
    for n, ref in ref_g.items():
        errs[n] = rel(pg[n], ref)
    bad = {n: v for n, v in errs.items() if not v < 2e-2}
    assert not bad, f"{cfg}: {bad}"


def test_wtconv_module_drop_in(golden_dir):
    import adnm_unet_b200 as A
    name = "wtconv_c32_k5_l3_33"
    C, k, L, B, H, W, bias = cases.WTCONV_CASES[name]
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    params = {n[6:]: torch.from_numpy(z[n]) for n in z.files if n.startswith("param/")}
    m = A.WTConv2d(C, C, kernel_size=k, bias=bias, wt_levels=L)
    m.load_state_dict(params, strict=True)
    m = m.cuda()
    x, dy = cases.wtconv_inputs(name, torch.float32)
    xd = x.cuda().requires_grad_(True)
    y = m(xd)
    y.backward(dy.cuda())
    assert rel(y, z["out"]) < 1e-4 and rel(xd.grad, z["dx"]) < 1e-4
    assert m.wt_filter.grad is None and m.iwt_filter.grad is None
    assert rel(m.base_conv.weight.grad, z["grad/base_conv.weight"]) < 1e-4
    with torch.no_grad():
        assert torch.equal(m(x.cuda()), y.detach())
