#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/<tag>_*) into the small text
summaries committed under profiles/: the launch list, the raw metric page of each full
capture, and a per-region SASS hot-spot table (instructions, active threads, stall samples).

    python scripts/summarize_profiles.py r6
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEEP = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]

shutil.copy(os.path.join(src, f"{tag}_launches.csv"), os.path.join(dst, f"{tag}_launches.csv"))
shutil.copy(os.path.join(src, f"{tag}_bench.json"), os.path.join(dst, f"{tag}_bench.json"))
for name in ("twostate", "vps"):
    rep = os.path.join(src, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(os.path.join(dst, f"{tag}_{name}_metrics.csv"), "w") as f:
        f.write("metric,unit,value\n")
        f.write(f"kernel,,\"{vals[hdr.index('Kernel Name')]}\"\n")
        for k in KEEP:
            if k in hdr:
                f.write(f"{k},{units[hdr.index(k)]},{vals[hdr.index(k)]}\n")
    # DRAM traffic per launch of this kernel, consumed by bench.py (roofline.traffic)
    import json
    tpath = os.path.join(dst, "traffic.json")
    tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
    rd, wr = float(vals[hdr.index("dram__bytes_read.sum")]), float(vals[hdr.index("dram__bytes_write.sum")])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd *= scale[units[hdr.index("dram__bytes_read.sum")]]; wr *= scale[units[hdr.index("dram__bytes_write.sum")]]
    key = "twostate_primal" if name == "twostate" else "volpathsimple_adjoint"
    tj[key] = {"dram_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "capture": f"profiles/{tag}_{name}_metrics.csv"}
    json.dump(tj, open(tpath, "w"), indent=1)
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(sass.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    chunks = collections.OrderedDict()
    total = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        a = int(r[ci["Address"]], 16) // 0x400 * 0x400
        n, t, s = int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)
        c = chunks.setdefault(a, [0, 0, 0])
        c[0] += n; c[1] += t; c[2] += s
        total += n
    with open(os.path.join(dst, f"{tag}_{name}_sass_regions.csv"), "w") as f:
        f.write("sass_region_1KiB,warp_instructions,share_pct,avg_active_threads,stall_samples\n")
        for a, c in chunks.items():
            if c[0]:
                f.write(f"0x{a & 0xfffff:05x},{c[0]},{100.0 * c[0] / total:.2f},{c[1] / c[0]:.1f},{c[2]}\n")
    print(name, "ok")
