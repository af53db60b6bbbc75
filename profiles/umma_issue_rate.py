"""Diagnostic: cycles per tcgen05.mma (128 x N x 16, bf16, SWIZZLE_NONE operands in shared memory) as a function of N,
operand major-ness, row pitch and start shift.  Feeds the cost model in DESIGN.md (small-N MMAs are not free)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import _lib
lib = _lib.load()
iters = 2000
for ctas in (1, 148):
    cyc = torch.zeros(ctas, dtype=torch.int64, device="cuda")
    for mode in (0, 1):
        for pitch, shift in ((128, 0), (130, 0), (130, 1)):
            row = []
            for N in (16, 32, 64, 96, 128, 208, 256):
                _lib.check(lib.adn_bench_umma(mode, N, pitch, shift, iters, ctas, _lib.ptr(cyc), _lib.stream_ptr()), "bench")
                torch.cuda.synchronize()
                row.append(f"N={N}:{cyc.float().mean().item() / iters:6.1f}")
            print(f"ctas={ctas} mode={'K-major' if mode == 0 else 'MN-major'} pitch={pitch} shift={shift}  " + "  ".join(row))
