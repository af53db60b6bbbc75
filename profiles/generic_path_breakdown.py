"""Per-launch device times (CUDA events recorded by the library, adn_prof_*) of one mixer forward + backward.
usage: python profiles/generic_path_breakdown.py [D] [d_state] [grid] [batch]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adnm_unet_b200 as A
from adnm_unet_b200 import _lib
D, N, g, B = (int(a) for a in (sys.argv[1:5] + ["128", "16", "128", "16"][len(sys.argv) - 1:]))
m = A.Mamba2(d_model=D, headdim=4, d_state=N).cuda()
u = torch.randn(B, g * g, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
go = torch.randn_like(u)
for _ in range(2):
    m(u, g, g).backward(go)
with _lib.profile() as prof:
    m(u, g, g).backward(go)
tot = sum(t for _, t in prof.records)
print(f"D={D} d_state={N} grid={g} B={B}: total {tot:.3f} ms over {len(prof.records)} launches")
for n, t in prof.records:
    print(f"{n:24s} {t*1e3:9.1f} us")
