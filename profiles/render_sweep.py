"""profiles/r02_sweep.json -> the markdown tables of DESIGN.md section 6."""
import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_sweep.json"))
rows = d["rows"]
mix = [r for r in rows if r["op"] == "adnssd_fwd_bwd"]
print("| d_model | d_state | path | 32² | 64² | 128² | 256² |\n|---|---|---|---|---|---|---|")
for D in sorted({r["d_model"] for r in mix}):
    for N in sorted({r["d_state"] for r in mix}):
        sel = {r["grid"]: r for r in mix if r["d_model"] == D and r["d_state"] == N}
        if sel:
            paths = " / ".join(dict.fromkeys(sel[g]["path"] for g in sorted(sel)))
            print(f"| {D} | {N} | {paths} | " + " | ".join(f"{sel[g]['ms']:.2f}" for g in sorted(sel)) + " |")
wt = [r for r in rows if r["op"] != "adnssd_fwd_bwd"]
if wt:
    print("\n| C | 32² | 64² | 128² | 256² |\n|---|---|---|---|---|")
    for C in sorted({r["C"] for r in wt}):
        sel = {r["grid"]: r for r in wt if r["C"] == C}
        print(f"| {C} | " + " | ".join(f"{sel[g]['ms']:.2f} ms ({sel[g].get('algorithmic_GBps', 0):.0f} GB/s)" for g in sorted(sel)) + " |")
