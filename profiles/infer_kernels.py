"""Kernel-time table of one inference forward of the hosted network (BASELINE configs[3]: B = 64, 256 x 256, bf16 autocast, eval +
no_grad), torch profiler, grouped by kernel name.  Usage: python profiles/infer_kernels.py [batch img] -> JSON."""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import refhost  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
img = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
model = refhost.build_adnm_unet(img, dropin=True, seed=0).to(dev).eval()
x = torch.rand(B, 5, 1, img, img, device=dev)


def fwd():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return model(x)


for _ in range(2):
    fwd()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fwd()
    torch.cuda.synchronize()
rows = sorted(((e.key[:100], e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0), key=lambda t: -t[1])
total = sum(t[1] for t in rows)
print(json.dumps({"batch": B, "img": img, "total_kernel_ms": total / 1e3, "n_launches": sum(t[2] for t in rows),
                  "top": [{"kernel": k, "ms": t / 1e3, "count": c, "share": t / total} for k, t, c in rows[:40]]}))
