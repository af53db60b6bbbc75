import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adnm_unet_b200 as A
from adnm_unet_b200 import _lib
D, N, g, B = 128, 16, 64, 16
m = A.Mamba2(d_model=D, headdim=4, d_state=N).cuda()
u = torch.randn(B, g * g, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
go = torch.randn_like(u)
for _ in range(2):
    m(u, g, g).backward(go)
with _lib.profile() as prof:
    m(u, g, g).backward(go)
tot = sum(t for _, t in prof.records)
print("total ms", tot)
for n, t in prof.records:
    print(f"{n:24s} {t*1e3:9.1f} us")
