"""Per-kernel device time of one conv stage (WTLayer / PatchEmbed / OutProj drop-in) forward + backward (library kernels only;
the adn_prof_* event log), next to the unmodified reference module run eagerly under bf16 autocast on the same GPU.
Usage: python profiles/stage_breakdown.py kind grid batch [this_dim next_dim levels skip]
       kind in {wtlayer, patchembed, outproj}; defaults are the network's decoder6 / encoder1 / out_proj shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import _lib, convstage, refhost  # noqa: E402

kind, grid, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
extra = [int(a) for a in sys.argv[4:]]
dev = torch.device("cuda:0")
L = grid * grid
if kind == "wtlayer":
    this_dim, next_dim, levels, skip = extra if extra else (64, 32, 3, 1)
    kw = dict(this_dim=this_dim, next_dim=next_dim, kernel=5, wt_levels=levels, if_res=bool(skip))
    cin = this_dim // 2 if skip else this_dim
    args = lambda: (torch.randn(B, L, cin, device=dev, requires_grad=True),) + ((torch.randn(B, L, cin, device=dev, requires_grad=True),) if skip else ())
    cls = "WTLayer"
elif kind == "patchembed":
    kw = dict(img_size=grid, patch_size=2, in_channels=5, embed_dim=32, kernel=5, wt_levels=3)
    args = lambda: (torch.rand(B, 5, L, device=dev).transpose(1, 2),)
    cls = "PatchEmbed"
else:
    kw = dict(num_frames=20, embed_dim=32, img_size=[grid, grid], wt_levels=3, out_expand=2)
    args = lambda: (torch.randn(B, L, 32, device=dev, requires_grad=True), torch.rand(B, grid, grid, device=dev))
    cls = "OutProj"


def run(m, a):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(*a)
    y = y[0] if isinstance(y, tuple) else y
    return y


def timed(m):
    a = args()
    dy = None
    for _ in range(3):
        y = run(m, a)
        dy = torch.randn_like(y) if dy is None else dy
        y.backward(dy)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(); y = run(m, a); e1.record(); y.backward(dy); e2.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), e1.elapsed_time(e2), a, dy


torch.manual_seed(0)
new = getattr(convstage, cls)(**kw).to(dev)
f, b, a, dy = timed(new)
print(f"{cls} {kw} grid={grid} B={B}: drop-in fwd {f:.3f} ms, bwd {b:.3f} ms (events, eager launch, bf16 autocast)")
if refhost.reference_available():
    ns = refhost.load_reference()
    torch.manual_seed(0)
    ref = ns.ref_stages[cls](**kw).to(dev)
    fr, br, _, _ = timed(ref)
    print(f"{cls} reference (eager, its own WTConv2d): fwd {fr:.3f} ms, bwd {br:.3f} ms")
with _lib.profile() as p:
    run(new, a).backward(dy)
tot = sum(ms for _, ms in p.records)
print(f"library kernels: {tot:.3f} ms over {len(p.records)} launches")
agg = {}
for name, ms in p.records:
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += ms
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:28s} x{n:<3d} {ms * 1e3:9.1f} us")
print("launch order:")
for name, ms in p.records:
    print(f"  {name:28s} {ms * 1e3:8.1f} us")
