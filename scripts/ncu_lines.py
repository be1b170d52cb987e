#!/usr/bin/env python
"""Per-source-line hot spots of an `ncu --set full --import-source on` capture (compiled with -lineinfo):
stall samples by reason, warp instructions and active threads per CUDA source line.

    python scripts/ncu_lines.py gpurun_out/r2a_vps.ncu-rep [top=40] > profiles/r2a_vps_lines.txt
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines, cur_file, hdr = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r):
            hdr.setdefault(h, i)                 # duplicate header names ("Source"): first = CUDA source, second = SASS
        nh = len(r)
    elif hdr and len(r) >= nh and r[0] != "":
        # asm strings in a source line break the CSV quoting: take the numeric columns from the END of the row
        src = ",".join(r[1:len(r) - nh + 2])
        lines.append((cur_file, [r[0], src] + r[len(r) - nh + 2:]))
ci = hdr
STALLS = ["stall_long_sb", "stall_lg", "stall_mio", "stall_short_sb", "stall_wait", "stall_math", "stall_not_selected",
          "stall_selected", "stall_branch_resolving", "stall_no_inst", "stall_dispatch", "stall_barrier", "stall_drain"]
tot = sum(int(r[ci["# Samples"]] or 0) for _, r in lines)
tot_inst = sum(int(r[ci["Instructions Executed"]] or 0) for _, r in lines)
agg = {s: sum(int(r[ci[s]] or 0) for _, r in lines) for s in STALLS}
print(f"# {rep}: {tot} samples, {tot_inst} warp instructions")
print("# stall totals: " + ", ".join(f"{k[6:]} {100.0 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print("file:line | samples % | top stalls | warp inst % | avg thr | source")
lines.sort(key=lambda fr: -int(fr[1][ci["# Samples"]] or 0))
for f, r in lines[:top]:
    s = int(r[ci["# Samples"]] or 0)
    n = int(r[ci["Instructions Executed"]] or 0)
    st = sorted(((int(r[ci[k]] or 0), k[6:]) for k in STALLS), reverse=True)[:3]
    print(f"{f}:{r[0]} | {100.0 * s / tot:.2f} | " + " ".join(f"{k}={v}" for v, k in st if v) +
          f" | {100.0 * n / tot_inst:.2f} | {r[ci['Avg. Threads Executed']]} | {r[1].strip()[:110]}")
