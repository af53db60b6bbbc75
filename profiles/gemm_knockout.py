"""Where does a tcgen05 GEMM launch spend its time?  in_proj-shaped problem (M = 262144, N = 640, K = 128, bf16 out) timed with the
kernel's knock-out switches (adn_set_option("gemm_dbg")): 1 = no epilogue stores, 2 = no operand loads, 4 = no MMAs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adnm_unet_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
def run(M, N, K, a_mn, b_mn, c_mode, splitk=1, iters=10):
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    B = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    C = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if c_mode == 0 else torch.float32)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    def go():
        _lib.check(lib.adn_selftest_gemm(M, N, K, 0, a_mn, b_mn, _lib.ptr(A), A.shape[1], 0, _lib.ptr(B), B.shape[1], 0, None, 8, 0, None, 8, 0,
                                         _lib.ptr(C), N, 0, c_mode, 1, splitk, None, 0, _lib.ptr(st), _lib.stream_ptr()), "gemm")
    res = {}
    for dbg in (0, 1, 2, 4, 3, 7, 7 + 8, 7 + 16, 7 + 8 + 16, 7 + 8 + 16 + 32):
        lib.adn_set_option(b"gemm_dbg", dbg)
        for _ in range(2): go()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): go()
        e1.record(); torch.cuda.synchronize()
        res[dbg] = e0.elapsed_time(e1) / iters * 1e3
    lib.adn_set_option(b"gemm_dbg", 0)
    return res
for name, args in (("inproj 262144x640x128 bf16", (262144, 640, 128, 0, 0, 0)), ("g 262144x512x128 f32", (262144, 512, 128, 0, 1, 1)),
                   ("du 262144x128x640 bf16", (262144, 128, 640, 0, 1, 0)), ("dWin 640x128x262144 atomic", (640, 128, 262144, 1, 1, 2, 59)),
                   ("ffn_in 524288x128x32 bf16", (524288, 128, 32, 0, 0, 0)), ("square 8192^3 bf16", (8192, 8192, 8192, 0, 0, 0))):
    r = run(*args)
    flops = 2.0 * args[0] * args[1] * args[2]
    print(f"{name:32s} full {r[0]:8.1f} us ({flops / r[0] / 1e6:6.1f} TF/s) | no stores {r[1]:8.1f} | no loads {r[2]:8.1f} | no MMA {r[4]:8.1f} | no loads+stores {r[3]:8.1f} | nothing {r[7]:8.1f} | +no fence {r[15]:8.1f} | +no math {r[23]:8.1f} | +neither {r[31]:8.1f} | +no tmem ld {r[63]:8.1f}")
