"""Two forward+backward steps of the benchmark mixer shape (D=32, B=16, 128x128 tokens, bf16) - the command that
the ncu captures under profiles/ are taken on (never a timing source)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adnm_unet_b200 as A

D = int(os.environ.get("ADN_D", "32")); B = int(os.environ.get("ADN_B", "16")); G = int(os.environ.get("ADN_GRID", "128"))
torch.manual_seed(0)
m = A.Mamba2(d_model=D, headdim=4, d_state=16).cuda()
u = torch.randn(B, G * G, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
go = torch.randn_like(u)
for _ in range(int(os.environ.get("ADN_STEPS", "2"))):
    out = m(u, G, G)
    out.backward(go)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
