"""Build the C-ABI CUDA library in-tree:  python -m adnm_unet_b200.build   (nvcc cross-compiles without a GPU).
One object per translation unit, compiled in parallel and only when stale; the link step produces lib/libadnb200.so."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libadnb200.so")
SOURCES = ["adnssd_api.cu", "adnssd_sm100.cu", "wtconv.cu", "metrics.cu", "optim.cu", "rmsnorm.cu", "block.cu", "sdpa.cu", "convstage.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(PKG, "..", "include", "adnb200.h")]


def _obj(src, tag):
    return os.path.join(LIBDIR, "obj", os.path.splitext(src)[0] + tag + ".o")


def _stale(target, deps):
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False, phase_timing=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = ".pt" if phase_timing else ""
    hdrs = _headers()
    os.makedirs(os.path.join(LIBDIR, "obj"), exist_ok=True)
    todo = [s for s in SOURCES if force or _stale(_obj(s, tag), [os.path.join(CSRC, s)] + hdrs)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-DADN_PHASE_TIMING"] if phase_timing else []) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", _obj(src, tag), os.path.join(CSRC, src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for src, r in ex.map(compile_one, todo):
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
            if verbose:
                print(r.stderr)
    objs = [_obj(s, tag) for s in SOURCES]
    marker = os.path.join(LIBDIR, "obj", ".variant")
    variant = open(marker).read() if os.path.isfile(marker) else None
    if todo or variant != tag or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
        with open(marker, "w") as f:
            f.write(tag)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, phase_timing="--phase-timing" in sys.argv))
