"""Two forward+backward steps of the fused attention at the decoder.attn shape of config 4 (B=64, L=1024, 32 heads of 4,
bf16) - the command the attention ncu capture under profiles/ is taken on."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200.attention import sdpa_packed
from adnm_unet_b200 import _lib

B = int(os.environ.get("ADN_B", "64")); L = int(os.environ.get("ADN_L", "1024")); H = int(os.environ.get("ADN_HEADS", "32"))
torch.manual_seed(0)
qkv = torch.randn(B, L, 3 * H * 4, device="cuda", dtype=torch.bfloat16, requires_grad=True)
do = torch.randn(B, L, H * 4, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    o = sdpa_packed(qkv, H, 4)
    o.backward(do)
torch.cuda.synchronize()
with _lib.profile() as p:
    o = sdpa_packed(qkv, H, 4); o.backward(do)
print({n: round(t * 1e3, 1) for n, t in p.records})
print("ok", float(o.float().abs().mean()))
