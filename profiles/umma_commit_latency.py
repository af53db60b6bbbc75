import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import _lib
lib = _lib.load()
cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
for it in (1, 1, 2, 4, 16, 64, 256):
    for N in (32, 208):
        _lib.check(lib.adn_bench_umma(0, N, 128, 0, it, 1, _lib.ptr(cyc), _lib.stream_ptr()), "bench")
        torch.cuda.synchronize()
        print(f"iters={it} N={N}: total {cyc.item()} cycles")
