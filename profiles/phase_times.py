"""Diagnostic: in-kernel phase timers (adn_phase_*) for one fwd+bwd of the benchmark mixer shape."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adnm_unet_b200 as A
from adnm_unet_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
m = A.Mamba2(d_model=32, headdim=4, d_state=16).cuda()
u = torch.randn(16, 128 * 128, 32, device="cuda", dtype=torch.bfloat16, requires_grad=True)
go = torch.randn_like(u)
for _ in range(2):
    m(u, 128, 128).backward(go)
torch.cuda.synchronize()
lib.adn_phase_enable(1)
m(u, 128, 128).backward(go)
buf = (C.c_ulonglong * 64)()
lib.adn_phase_read(buf)
lib.adn_phase_enable(0)
names = {0: "k_bwd1_ws epilogue t0 [wait mma1 | y + stats | bar.sync x2 | yhat + g_y sums | dy, dz stores + sd | wait mma2 | epi2 + loop | .]", 1: "k_bwd2", 2: "k_bconv_du producer [wait empty | issue]", 3: "k_bconv_du epilogue t0 [wait slot_full | tmem ld + release | exchange + bar | shuffle + store | . | . | . | loop]",
         4: "k_fconv MMA thread [wait u | wait acc_empty | issue conv | wait st_full | issue state | . | . | loop]",
         5: "k_bconv_du MMA thread [wait stage | wait acc_empty | issue | . | . | . | . | loop]",
         6: "k_bconv_wg MMA thread [wait u | wait A | issue | (epi) wait done | (epi) contract+atomics | CTA prologue | CTA body | loop]",
         7: "k_fconv epilogue t0+t128 [wait acc_full | wait st_empty | compute half0 | compute half1 | release | . | . | loop]"}
for k, n in names.items():
    row = [buf[k * 8 + i] for i in range(8)]
    tot = sum(row) or 1
    print(n, "total Mcycles(thread0 sum)", round(tot / 1e6, 2), [f"{100 * v / tot:.0f}%" for v in row], "kcycles/CTA(148):", [round(v / 148e3, 1) for v in row])

ct = (C.c_ulonglong * 1920)()
lib.adn_cta_times_read(ct)
for kid, n in enumerate(("k_fconv", "k_bconv_du", "k_bconv_wg")):
    rows = [[ct[(kid * 160 + c) * 4 + j] for j in range(4)] for c in range(160)]
    rows = [r for r in rows if r[1] > r[0] > 0]
    if not rows:
        continue
    t0 = min(r[0] for r in rows)
    med = lambda v: sorted(v)[len(v) // 2] / 1e3
    print(f"{n}: {len(rows)} CTAs, first start -> last end {(max(r[1] for r in rows) - t0) / 1e3:.1f} us; per CTA (median, us): "
          f"prologue {med([r[2] - r[0] for r in rows]):.1f}, MMA loop ends at {med([r[3] - r[0] for r in rows]):.1f}, "
          f"CTA ends at {med([r[1] - r[0] for r in rows]):.1f} (max {max(r[1] - r[0] for r in rows) / 1e3:.1f})")
