"""Diagnostic driver for compute-sanitizer: one small fwd+bwd of the mixer (bf16 tcgen05 path and fp32 generic path),
one WTConv2d fwd+bwd and one threshold-count call."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adnm_unet_b200 as A
from oracle import adnssd_oracle as AO, wtconv_oracle as WO
for dtype in (torch.bfloat16, torch.float32):
    for (D, N, B, g) in [(32, 16, 2, 16), (32, 64, 1, 16)]:
        p = {k: v.cuda().requires_grad_(k not in AO.UNUSED_PARAMS) for k, v in AO.init_params(D, 4, N).items()}
        u = torch.randn(B, g * g, D, device="cuda", dtype=dtype, requires_grad=True)
        out = A.adnssd_mixer(u, g, g, p, headdim=4, d_state=N)
        out.backward(torch.ones_like(out))
        torch.cuda.synchronize()
        print("mixer", dtype, D, N, float(out.float().abs().mean()))
wp = {k: v.cuda().requires_grad_(k not in ("wt_filter", "iwt_filter")) for k, v in WO.init_params(8, 5, 3).items()}
x = torch.randn(2, 8, 33, 40, device="cuda", requires_grad=True)
y = A.wtconv2d(x, wp, 5, 3); y.backward(torch.ones_like(y)); torch.cuda.synchronize()
print("wtconv", float(y.abs().mean()))
print(A.threshold_counts(torch.rand(1000, device="cuda"), torch.rand(1000, device="cuda")).sum().item())
