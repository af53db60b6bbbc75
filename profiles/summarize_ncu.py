"""ncu report -> small JSON summary committed under profiles/ (run where `ncu` is on PATH; no GPU needed).

    python profiles/summarize_ncu.py gpurun_out/prof_r1b.ncu-rep profiles/r01_mixer_kernels.json ["command label"]
"""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
              "usecond": 1.0, "nsecond": 1e-3}


def main(rep, out, command="python profiles/run_mixer_once.py (D=32, B=16, 128x128 tokens, bf16), second step"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    launches, kernels = [], {}
    for n, r in enumerate(rows[2:]):
        name = re.sub(r"^void ", "", r[idx["Kernel Name"]])
        short = re.match(r"(?:\w+::)*(\w+)", name).group(1)
        e = {"launch": n, "kernel": short, "signature": name[:120]}
        for m, key in METRICS.items():
            if m in idx and r[idx[m]] not in ("", "n/a"):
                v = float(r[idx[m]].replace(",", ""))
                e[key] = v * UNIT_SCALE.get(units[idx[m]], 1.0)
        launches.append(e)      # EVERY launch, in order (round 1 keyed by kernel name and kept only the last launch of each)
        k = kernels.setdefault(short, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        k["launches"] += 1
        for key in ("time_us", "dram_read_bytes", "dram_write_bytes"):
            k[key] += e.get(key, 0.0)
    json.dump({"source": rep, "command": command, "ncu": "--set full --clock-control none --import-source on",
               "note": "`launches`: every captured launch in order; `kernels`: sums over the launches of each kernel name",
               "launches": launches, "kernels": kernels}, open(out, "w"), indent=1)
    for e in launches:
        print(f"{e['launch']:3d} {e['kernel']:20s} {e.get('time_us', 0):8.1f} us  dram {(e.get('dram_read_bytes', 0) + e.get('dram_write_bytes', 0)) / 1e6:7.1f} MB"
              f"  {e.get('dram_pct_of_peak', 0):5.1f}% dram  {e.get('warps_active_pct', 0):5.1f}% warps  {e.get('tensor_pipe_active_pct', 0):4.1f}% tensor"
              f"  {e.get('issue_active_pct', 0):4.1f}% issue")


if __name__ == "__main__":
    main(*sys.argv[1:4])
