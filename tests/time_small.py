"""Diagnostic: wall time of small mixer calls (per-call overheads)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adnm_unet_b200 as A
from adnm_unet_b200 import _lib
from oracle import adnssd_oracle as AO
for (D, N, B, g) in [(32, 16, 2, 16), (32, 64, 2, 8), (128, 16, 2, 8)]:
    p = {k: v.cuda().requires_grad_(k not in AO.UNUSED_PARAMS) for k, v in AO.init_params(D, 4, N).items()}
    u = torch.randn(B, g * g, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with _lib.profile() as prof:
            out = A.adnssd_mixer(u, g, g, p, headdim=4, d_state=N)
            out.backward(torch.ones_like(out))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        top = sorted(prof.records, key=lambda r: -r[1])[:3]
        print(D, N, B, g, "iter", it, f"{dt*1e3:.1f} ms", top)
