"""Per-kernel device time of one fused Block forward + backward (library kernels only; the adn_prof_* event log).
Usage: python profiles/block_breakdown.py dim out_dim grid batch"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import _lib
from adnm_unet_b200.block import make_block

dim, out_dim, grid, B = (int(a) for a in sys.argv[1:5])
dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = make_block(dim, out_dim, headdim=4, norm_epsilon=1e-6).to(dev)
x = torch.randn(B, grid * grid, dim, device=dev, dtype=torch.bfloat16, requires_grad=True)
dy = torch.randn(B, grid * grid, out_dim, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    blk(x).backward(dy)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); y = blk(x); e1.record(); y.backward(dy); e2.record(); torch.cuda.synchronize()
print(f"dim={dim} out={out_dim} grid={grid} B={B}: fwd {e0.elapsed_time(e1):.3f} ms, bwd {e1.elapsed_time(e2):.3f} ms (events, eager launch)")
with _lib.profile() as p:
    blk(x).backward(dy)
tot = sum(ms for _, ms in p.records)
print(f"library kernels: {tot:.3f} ms over {len(p.records)} launches")
for name, ms in p.records:
    print(f"{name:28s} {ms * 1e3:8.1f} us")
