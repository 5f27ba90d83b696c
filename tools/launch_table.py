#!/usr/bin/env python
"""Markdown table (share, device time, launches per kernel) from an `ncu --metrics gpu__time_duration.sum --csv`
launch list: python tools/launch_table.py launches.csv [calls]  (calls = forward calls captured, default 5)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 5
hdr, rows = None, []
for r in csv.reader(open(path, errors="ignore")):
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        rows.append(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[ki]
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", "")) / 1e3
total = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {total / 1e3:.1f} ms summed device time over {calls} calls = "
      f"{len(rows) / calls:.0f} launches and {total / 1e3 / calls:.2f} ms per call\n")
print("| share | us per call | launches per call | kernel |")
print("|---:|---:|---:|---|")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if us / total < 0.0005:
        continue
    print(f"| {100 * us / total:.1f} % | {us / calls:.0f} | {n / calls:.1f} | `{name[:110]}` |")
