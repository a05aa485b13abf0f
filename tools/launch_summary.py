#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
start = rows.index(hdr)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start + 1:]:
    n = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':58s} {'n':>4s} {'total ms':>9s} {'avg us':>10s} {'share':>6s}")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{n[:58]:58s} {a[0]:4d} {a[1] / 1e3:9.2f} {a[1] / a[0]:10.1f} {a[1] / tot * 100:5.1f}%")
