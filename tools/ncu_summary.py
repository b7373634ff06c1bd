#!/usr/bin/env python
"""Text summary of kernels in an .ncu-rep (via `ncu -i rep --page raw --csv`).  usage: ncu_summary.py raw.csv [kernel-substring]"""
import csv
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = sys.argv[2] if len(sys.argv) > 2 else ""


def num(s):
    return float(s.replace(",", ""))


for r in rows[2:]:
    if want not in r[idx["Kernel Name"]]:
        continue
    for k in KEYS:
        if k in idx:
            print(f"{k} = {r[idx[k]]} {units[idx[k]]}")
    stalls = [(h, num(r[i])) for h, i in idx.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and r[i] not in ("", "n/a")]
    for h, v in sorted(stalls, key=lambda x: -x[1])[:8]:
        print(f"{h} = {int(v)} warp")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    t = sum(num(r[idx[k]]) * scale[units[idx[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    print(f"traffic (dram read + write per launch) = {int(t)} bytes\n")
