"""Two forward+backward steps of one fused refiner Block (dim 32, 128x128 tokens, bf16) - the command the Block ncu
captures under profiles/ are taken on (never a timing source).  ADN_D / ADN_OUT / ADN_GRID / ADN_B override the shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200.block import make_block

D = int(os.environ.get("ADN_D", "32")); O = int(os.environ.get("ADN_OUT", str(D))); B = int(os.environ.get("ADN_B", "16")); G = int(os.environ.get("ADN_GRID", "128"))
torch.manual_seed(0)
blk = make_block(D, O, headdim=4, norm_epsilon=1e-6).cuda()
x = torch.randn(B, G * G, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
dy = torch.randn(B, G * G, O, device="cuda", dtype=torch.bfloat16)
for _ in range(int(os.environ.get("ADN_STEPS", "2"))):
    y = blk(x)
    y.backward(dy)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
