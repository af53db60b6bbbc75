"""Copy the reference files the full-model harness hosts into git-ignored `baseline/_ref/` so that they travel to the GPU
box with the `gpurun` snapshot.  Run in the build container, where /root/reference is
mounted:   python baseline/fetch_ref.py        (also done by `__graft_entry__.build()`).

Nothing from `baseline/_ref/` is committed (it is in .gitignore) and nothing under it is edited: the files are the
UNMODIFIED reference, used (i) as the host network around the sm_100a drop-ins (adnm_unet_b200.refhost), (ii) as the
like-for-like baseline of the parity tests and of `bench.py --impl reference` / `gpu_baseline`."""
import os
import shutil
import sys

SRC = os.environ.get("ADNM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("models/__init__.py", "models/ADNssd.py", "models/ADNMUNet.py", "models/WTConv2d.py", "models/model_untils.py",
         "models/loss.py", "models/MLA.py", "datasets/Shanghai_metrics.py", "README.md")


def fetch(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "models", "ADNssd.py")):
        if verbose:
            print(f"fetch_ref: {SRC} not mounted; keeping whatever is in {DST}")
        return os.path.isdir(DST)
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    if verbose:
        print(f"fetch_ref: {len(FILES)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if fetch() else 1)
