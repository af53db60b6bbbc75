"""Per-launch device times (library CUDA events) of one WTConv2d fwd+bwd on the profile shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adnm_unet_b200 as A
from adnm_unet_b200 import _lib

C, G, B, K, LV = 32, 128, int(os.environ.get("ADN_B", "64")), 5, 3
torch.manual_seed(0)
m = A.WTConv2d(C, C, kernel_size=K, wt_levels=LV).cuda()
x = torch.randn(B, C, G, G, device="cuda", dtype=torch.bfloat16, requires_grad=True)
gy = torch.randn_like(x)
for _ in range(3):
    y = m(x); y.backward(gy)
torch.cuda.synchronize()
with _lib.profile() as prof:
    y = m(x); y.backward(gy)
for name, t in prof.records:
    print(f"{name:16s} {t * 1e3:8.1f} us")
