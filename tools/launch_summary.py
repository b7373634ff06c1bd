#!/usr/bin/env python
"""Summarise one bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list:
the launches between the last two mac_kernel launches.  usage: launch_summary.py launches.csv "<header note>" """
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
launches = []
for r in rows:
    if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui].strip(), 1e-6)
    name = r[ki].split("(")[0]
    launches.append((name, v * scale))
macs = [i for i, (n, _) in enumerate(launches) if "mac_kernel" in n]
lo, hi = macs[-2] + 1, macs[-1] + 1
agg = OrderedDict()
for n, ms in launches[lo:hi]:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# one step = the launches between two consecutive mac_kernel launches")
for n, (c, ms) in agg.items():
    print(f"{n:60s} n={c:3d} {ms:8.3f} ms {100 * ms / tot:5.1f}%")
print(f"total {tot:.3f} ms over {hi - lo} launches (ncu --metrics gpu__time_duration.sum --clock-control none; serialized, cold-cache)")
