"""WTConv2d fwd+bwd on the sweep shape of BASELINE configs[4] (C=32, k=5, levels=3, 128x128, B s.t. B*C*H*W = 2^27 elements
by default scaled down with ADN_B) - the command the WTConv ncu captures under profiles/ are taken on; also prints timings."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adnm_unet_b200 as A
from adnm_unet_b200 import _lib

C = int(os.environ.get("ADN_C", "32")); G = int(os.environ.get("ADN_GRID", "128")); B = int(os.environ.get("ADN_B", "64"))
K = int(os.environ.get("ADN_K", "5")); LV = int(os.environ.get("ADN_LEVELS", "3")); steps = int(os.environ.get("ADN_STEPS", "2"))
dt = torch.bfloat16 if os.environ.get("ADN_DTYPE", "bf16") == "bf16" else torch.float32
torch.manual_seed(0)
m = A.WTConv2d(C, C, kernel_size=K, wt_levels=LV).cuda()
x = torch.randn(B, C, G, G, device="cuda", dtype=dt, requires_grad=True)
gy = torch.randn_like(x)
for _ in range(steps):
    y = m(x); y.backward(gy)
torch.cuda.synchronize()
if os.environ.get("ADN_TIME"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with _lib.profile() as prof:
        e0.record()
        for _ in range(10):
            y = m(x); y.backward(gy)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    n = x.numel(); eb = x.element_size()
    print(f"wtconv C={C} k={K} L={LV} {G}x{G} B={B} {dt}: {ms*1e3:.1f} us fwd+bwd, algorithmic 5*N*e = {5*n*eb/1e6:.1f} MB -> {5*n*eb/ms/1e6:.1f} GB/s")
    per = {}
    for name, t in prof.records:
        per[name] = per.get(name, 0) + t / 10
    print({k: round(v * 1e3, 1) for k, v in per.items()})
print("ok", float(y.float().abs().mean()))
