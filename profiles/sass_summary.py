"""Per-kernel SASS evidence of the built library (no GPU needed): instruction counts that show which hardware path a kernel
uses - UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG / UTMASTG / UTMAREDG (TMA tensor loads / stores / reduce-adds),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), FFMA2 (packed fp32), MUFU - plus registers per thread.
    python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "adnm-unet_b200", "lib", "libadnb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+)", line)
    if m and cur:
        regs[cur] = int(m.group(1))
KEYS = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "FFMA2", "FFMA", "MUFU", "LDG", "STG", "LDS", "STS", "ATOM", "RED")
rows = []
cur, cnt, total = None, None, 0
def flush():
    if cur:
        rows.append((cur, total, dict(cnt)))
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        cur, cnt, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        total += 1
        op = m.group(1).split(".")[0]
        for k in KEYS:
            if op == k:
                cnt[k] += 1
flush()
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(lib, ROOT)}: {len(rows)} kernels; columns = SASS instruction counts (static), REG = registers per thread")
print(f"{'kernel':70s} {'REG':>4s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
for (name, total, cnt), d in sorted(zip(rows, dem), key=lambda t: t[1]):
    short = re.sub(r"\(.*", "", d).replace("void ", "")[:70]
    print(f"{short:70s} {regs.get(name, 0):4d} {total:6d} " + " ".join(f"{cnt.get(k, 0):8d}" for k in KEYS))
