"""Diagnostic (not a test): bf16 fast path vs fp32 check mode at the bench shape, per-tensor relative errors.
Run with ADN_ROWCONV=0 to see the tile-kernel path."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_mixer_gpu import _rand_params, rel
import adnm_unet_b200 as A
dev = torch.device("cuda:0")
B, g, D, P, N = int(os.environ.get("DIAG_B", 16)), 128, 32, 4, 16
gen = torch.Generator().manual_seed(5)
u = torch.randn(B, g * g, D, generator=gen); dout = torch.randn(B, g * g, D, generator=gen)
res = {}
for dtype in (torch.float32, torch.bfloat16):
    p = _rand_params(D, P, N, dev)
    ud = u.to(dev, dtype).requires_grad_(True)
    out = A.adnssd_mixer(ud, g, g, p, headdim=P, d_state=N)
    out.backward(dout.to(dev, dtype)); torch.cuda.synchronize()
    res[dtype] = (out.detach().float().cpu(), ud.grad.float().cpu(), {k: v.grad.float().cpu() for k, v in p.items() if v.grad is not None})
ref, fast = res[torch.float32], res[torch.bfloat16]
errs = {"out": rel(fast[0], ref[0]), "du": rel(fast[1], ref[1])}
for k in ref[2]:
    errs[k] = rel(fast[2][k], ref[2][k])
print("ADN_ROWCONV=" + os.environ.get("ADN_ROWCONV", "1"), " ".join(f"{k.replace('.weight','')}={v:.1e}" for k, v in errs.items()))
rms = lambda a, b: ((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()
print("  rms-relative: out=%.2e du=%.2e" % (rms(fast[0], ref[0]), rms(fast[1], ref[1])))
