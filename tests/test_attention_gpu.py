"""GPU parity of the fused attention (adn_sdpa_*) and of the StandardAttention drop-in against the CPU oracle
(oracle/attention_oracle.py, pinned to the unmodified reference class)."""
import pytest
import torch

import cases
from oracle import attention_oracle as AT

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [(2, 256, 32, 4), (1, 1024, 8, 4), (3, 77, 5, 4), (2, 16, 256, 4), (1, 130, 3, 8), (2, 40, 2, 16)],
                         ids=lambda c: "B%d_L%d_h%d_dh%d" % c)
def test_sdpa_packed_matches_oracle(cfg, dtype):
    """models/ADNssd.py:41-46 from the packed projection; ragged L (not a multiple of the 128-key tile) included."""
    from adnm_unet_b200.attention import sdpa_packed
    B, L, heads, dh = cfg
    dev = torch.device("cuda:0")
    qkv = (cases.rng_normal(81, (B, L, 3 * heads * dh), torch.float32) * 1.5).to(dtype).float()
    dout = cases.rng_normal(82, (B, L, heads * dh), torch.float32).to(dtype).float()
    r = qkv.double().requires_grad_(True)
    ref = AT.sdpa_packed(r, heads, dh)
    ref.backward(dout.double())
    x = qkv.to(dev, dtype).requires_grad_(True)
    out = sdpa_packed(x, heads, dh)
    out.backward(dout.to(dev, dtype))
    torch.cuda.synchronize()
    errs = {"out": rel(out, ref), "dqkv": rel(x.grad, r.grad)}
    bad = {k: v for k, v in errs.items() if not v < TOL[dtype]}
    assert not bad, bad
    with torch.no_grad():
        assert torch.equal(sdpa_packed(x.detach(), heads, dh), out.detach())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
@pytest.mark.parametrize("cfg", [(128, 32, 4, 2, 256), (1024, 256, 4, 2, 16), (40, 10, 4, 1, 49)], ids=lambda c: "dim%d_h%d_dh%d_B%d_L%d" % c)
def test_standard_attention_module_matches_oracle(cfg, dtype):
    """The module drop-in (to_qkv -> fused attention -> to_out): output, input gradient, the three parameter gradients;
    a reference-shaped state_dict loads strictly."""
    from adnm_unet_b200.attention import StandardAttention
    dim, heads, dh, B, L = cfg
    dev = torch.device("cuda:0")
    p = AT.init_params(dim, heads, dh, seed=5)
    if dtype == torch.bfloat16:      # the tensor cores see bf16 weights: give the oracle the same numbers
        p = {k: (v.float().bfloat16().double() if k.endswith("weight") else v) for k, v in p.items()}
    m = StandardAttention(dim, heads=heads, dim_head=dh, dropout=0.)
    m.load_state_dict({k: v.float() for k, v in p.items()}, strict=True)
    m = m.to(dev)
    x = cases.rng_normal(91, (B, L, dim), torch.float32).to(dtype).float()
    dy = cases.rng_normal(92, (B, L, dim), torch.float32).to(dtype).float()
    pp = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    xr = x.double().requires_grad_(True)
    ref = AT.attention_forward(pp, xr, heads, dh)
    ref.backward(dy.double())
    xn = x.to(dev, dtype).requires_grad_(True)
    y = m(xn, 1, 1)
    y.backward(dy.to(dev, dtype))
    torch.cuda.synchronize()
    errs = {"out": rel(y, ref), "dx": rel(xn.grad, xr.grad)}
    for k, v in m.named_parameters():
        errs[k] = rel(v.grad, pp[k].grad)
    bad = {k: v for k, v in errs.items() if not v < TOL[dtype]}
    assert not bad, bad
