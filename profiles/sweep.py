"""BASELINE configs[4] sweep (evidence, not a bench line): ADN-SSD mixer fwd+bwd over token grids 32^2..256^2 and d_state
16..128 at a constant token count (B = 262144 / L), bf16, device-timed with CUDA events; and WTConv2d over the same grids.
Writes gpurun_out/r02_sweep.json (copied to profiles/ when committed).  Paths: W == 128, d_state 16 -> conv-as-GEMM row kernels; other L % 128 == 0 shapes of
d_model 32 with d_state in {16, 64} -> halo-tile tcgen05 kernels; d_model >= 64 and d_state 128 -> wide path (tcgen05 GEMMs, adnssd_wide.cuh)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adnm_unet_b200 as A
from adnm_unet_b200 import _lib

def time_fn(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

rows = []
TOK = 262144
for D in (32, 128):
    for N in (16, 32, 64, 128):
        for g in (32, 64, 128, 256):
            B = TOK // (g * g)
            torch.manual_seed(0)
            m = A.Mamba2(d_model=D, headdim=4, d_state=N).cuda()
            u = torch.randn(B, g * g, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
            go = torch.randn_like(u)
            def step():
                u.grad = None
                for p in m.parameters(): p.grad = None
                m(u, g, g).backward(go)
            with _lib.profile() as prof:
                step()
            names = sorted({n for n, _ in prof.records})
            path = "row" if "k_fconv" in names else ("tile" if "k_inproj" in names else ("wide" if "tcgemm_inproj" in names else "generic"))
            ms = time_fn(step, 10 if path != "generic" else 3)
            rows.append({"op": "adnssd_fwd_bwd", "d_model": D, "d_state": N, "grid": g, "batch": B, "path": path, "ms": ms,
                         "tokens_per_s": TOK / (ms * 1e-3)})
            print(rows[-1], flush=True)
            del m, u, go
            torch.cuda.empty_cache()
for C in (32, 64):
    for g in (32, 64, 128, 256):
        B = (1 << 27) // (C * g * g)
        m = A.WTConv2d(C, C, kernel_size=5, wt_levels=3).cuda()
        x = torch.randn(B, C, g, g, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        gy = torch.randn_like(x)
        def stepw():
            x.grad = None
            m(x).backward(gy)
        ms = time_fn(stepw, 5)
        n = x.numel()
        rows.append({"op": "wtconv_fwd_bwd", "C": C, "k": 5, "levels": 3, "grid": g, "batch": B, "ms": ms,
                     "algorithmic_GBps": 5 * n * 2 / (ms * 1e-3) / 1e9})
        print(rows[-1], flush=True)
        del m, x, gy
        torch.cuda.empty_cache()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "r02_sweep.json")
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump({"note": "eager launches (host-bound below ~0.4 ms); see bench.py for the graph-replayed headline", "rows": rows}, open(out, "w"), indent=1)
