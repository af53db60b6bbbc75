"""Two forward+backward steps of one WTLayer drop-in at the network's decoder6 shape (64 -> 32 channels, 128x128 tokens, B = 32,
bf16 autocast) - the command the conv-stage ncu capture under profiles/ is taken on (never a timing source).
53 library launches per step: the capture skips the first step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import convstage  # noqa: E402

B = int(os.environ.get("ADN_B", "32"))
G = int(os.environ.get("ADN_GRID", "128"))
torch.manual_seed(0)
m = convstage.WTLayer(this_dim=64, next_dim=32, kernel=5, wt_levels=3, if_res=True).cuda()
x = torch.randn(B, G * G, 32, device="cuda", requires_grad=True)
r = torch.randn(B, G * G, 32, device="cuda", requires_grad=True)
dy = torch.randn(B, G * G, 32, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x, residual=r)
    y.backward(dy)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
