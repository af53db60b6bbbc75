"""Diagnostic (not a test): per-tensor relative errors of the CUDA mixer vs golden for every case / dtype."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
from test_mixer_gpu import load_case, run_cuda, rel
gd = os.path.join(ROOT, "tests", "golden")
for name in sorted(cases.MIXER_CASES):
    D, P, N, B, g = cases.MIXER_CASES[name]
    z, params, grads = load_case(gd, name)
    u, dout = cases.mixer_inputs(name, torch.float32)
    for dtype in (torch.float32, torch.bfloat16):
        out, du, pg = run_cuda(params, u, dout, g, g, P, N, dtype)
        s = cases.SUBSAMPLE_STRIDE if g >= 64 else 1
        errs = {"out": rel(out[:, ::s], z["out"]), "du": rel(du[:, ::s], z["du"])}
        for k, ref in grads.items():
            errs[k] = rel(pg[k].reshape(ref.shape), ref)
        print(name, dtype, " ".join(f"{k.replace('.weight','')}={v:.1e}" for k, v in errs.items()))
for name in sorted(cases.MIXER_BF16_CASES):
    D, P, N, B, g, _ = cases.MIXER_BF16_CASES[name]
    z, params, grads = load_case(gd, name)
    u, dout = cases.mixer_bf16_inputs(name, torch.float32)
    for dtype in (torch.float32, torch.bfloat16):
        out, du, pg = run_cuda(params, u, dout, g, g, P, N, dtype)
        errs = {"out": rel(out, z["out"]), "du": rel(du, z["du"])}
        for k, ref in grads.items():
            errs[k] = rel(pg[k].reshape(ref.shape), ref)
        print(name, dtype, " ".join(f"{k.replace('.weight','')}={v:.1e}" for k, v in errs.items()))
