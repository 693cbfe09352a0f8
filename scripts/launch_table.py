#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV): python scripts/launch_table.py file.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(ln for ln in open(sys.argv[1]) if not ln.startswith("==")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
hdr, acc = None, collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        v = float(r[hdr["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    unit = r[hdr["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    a = acc.setdefault(r[hdr["Kernel Name"]], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in acc.values())
for k, a in sorted(acc.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{a[1]:9.3f} ms {a[0]:3d}x {100 * a[1] / tot:5.1f}%  {k[:100]}")
print(f"{tot:9.3f} ms total, {sum(a[0] for a in acc.values())} launches")
