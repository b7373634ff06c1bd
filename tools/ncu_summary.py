#!/usr/bin/env python3
"""Compact text summary of an `ncu --page raw --csv` export (one block per profiled launch): the metrics the
profiles/ notes quote.  Usage: python tools/ncu_summary.py raw.csv [> profiles/NAME.txt]"""
import csv
import sys

KEYS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    col = {n: i for i, n in enumerate(hdr)}
    for r in data:
        if not r:
            continue
        print(f"Kernel Name = {r[col['Kernel Name']]}")
        for k in KEYS:
            if k in col:
                print(f"{k} = {r[col[k]]} {units[col[k]]}")
        stalls = sorted(((float(r[i]), n) for n, i in col.items() if n.startswith("smsp__pcsamp_warps_issue_stalled_")
                         and not n.endswith("_not_issued") and r[i].replace(".", "").isdigit()), reverse=True)
        for v, n in stalls[:8]:
            print(f"{n} = {v:.0f} warp")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
