"""Probe: GPU time of the three live EncoderToDecoder bridges (reference code, models/model_untils.py:621-798) inside a training
step, eager and graph-replayed, and the kernel mix of one of their grouped convs.  Usage: python profiles/bridge_probe.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import refhost  # noqa: E402

dev = torch.device("cuda:0")
model = refhost.build_adnm_unet(128, dropin=True, seed=0).to(dev)
B = 32
dims = {0: (1024, 4), 1: (512, 8), 2: (256, 16)}


def run(i, x, r):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = model.decoder.e2ds[i](x=x, res=r)
    return y


S = torch.cuda.Stream()      # everything on one non-default stream (AccumulateGrad nodes must not live on the legacy stream)
torch.cuda.set_stream(S)
for i, (d, g) in dims.items():
    x = torch.randn(B, g * g, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    r = torch.randn(B, g * g, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
    y = run(i, x, r)
    dy = torch.randn_like(y)
    for _ in range(3):
        run(i, x, r).backward(dy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run(i, x, r).backward(dy)
    e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 5
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=S):
        run(i, x, r).backward(dy)
    gr.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"e2ds[{i}] dim {d} grid {g}: eager {eager:.3f} ms, graph replay {e0.elapsed_time(e1) / 10:.3f} ms per fwd+bwd")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(i, x, r).backward(dy)
        torch.cuda.synchronize()
    rows = sorted(((e.key[:90], e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0), key=lambda t: -t[1])
    print(f"  kernels: {sum(t[2] for t in rows)} launches, {sum(t[1] for t in rows) / 1e3:.3f} ms")
    for k, t, c in rows[:10]:
        print(f"    {t / 1e3:7.3f} ms x{c:<4d} {k}")
# one grouped conv alone
conv = torch.nn.Conv2d(512, 512, (3, 3), padding=(1, 1), groups=128).to(dev)
x = torch.randn(B, 512, 8, 8, device=dev, requires_grad=True)
with torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(2):
        conv(x).sum().backward()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        conv(x).sum().backward()
        torch.cuda.synchronize()
rows = sorted(((e.key[:90], e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0), key=lambda t: -t[1])
print(f"Conv2d(512, 512, 3x3, groups=128) 8x8 B=32 fwd+bwd: {sum(t[2] for t in rows)} launches, {sum(t[1] for t in rows) / 1e3:.3f} ms")
for k, t, c in rows[:8]:
    print(f"    {t / 1e3:7.3f} ms x{c:<4d} {k}")
