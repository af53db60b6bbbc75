"""`ncu --csv --log-file X.csv --metrics ...` (long format: one row per launch and metric) -> the small JSON summaries under
profiles/ (same schema as summarize_ncu.py, which reads a .ncu-rep: those exceed the 64 MiB gpurun return limit for the wide
captures, so round 2 converts on the GPU box and brings back only the CSV).

    python profiles/summarize_ncu_csv.py gpurun_out/r2ncu/block_d32.csv profiles/r02_block_kernels.json "command label"
"""
import csv
import json
import re
import sys

from summarize_ncu import METRICS, UNIT_SCALE


def main(src, out, command=""):
    lines = open(src, newline="").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    launches, order = {}, []
    for r in rows:
        i = int(r["ID"])
        if i not in launches:
            name = re.sub(r"^void ", "", r["Kernel Name"])
            short = re.match(r"(?:\w+::)*(\w+)", name).group(1)
            launches[i] = {"launch": len(order), "kernel": short, "signature": name[:120]}
            order.append(i)
        key = METRICS.get(r["Metric Name"])
        if key and r["Metric Value"] not in ("", "n/a"):
            launches[i][key] = float(r["Metric Value"].replace(",", "")) * UNIT_SCALE.get(r["Metric Unit"], 1.0)
    ls = [launches[i] for i in order]
    kernels = {}
    for e in ls:
        k = kernels.setdefault(e["kernel"], {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        k["launches"] += 1
        for key in ("time_us", "dram_read_bytes", "dram_write_bytes"):
            k[key] += e.get(key, 0.0)
    json.dump({"source": src, "command": command, "ncu": "--metrics <the METRICS of summarize_ncu.py> --clock-control none --csv",
               "note": "`launches`: every captured launch in order; `kernels`: sums over the launches of each kernel name",
               "launches": ls, "kernels": kernels}, open(out, "w"), indent=1)
    for e in ls:
        print(f"{e['launch']:3d} {e['kernel']:22s} {e.get('time_us', 0):8.1f} us  dram {(e.get('dram_read_bytes', 0) + e.get('dram_write_bytes', 0)) / 1e6:7.1f} MB"
              f"  {e.get('dram_pct_of_peak', 0):5.1f}% dram  {e.get('warps_active_pct', 0):5.1f}% warps  {e.get('issue_active_pct', 0):4.1f}% issue"
              f"  L1 {e.get('l1_hit_pct', 0):4.1f}%  regs {e.get('registers_per_thread', 0):.0f}")


if __name__ == "__main__":
    sys.path.insert(0, __file__.rsplit("/", 1)[0])
    main(*sys.argv[1:4])
