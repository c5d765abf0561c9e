#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the step)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':28s} {'n':>4s} {'total_us':>10s} {'share':>7s} {'max_us':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:28s} {len(v):4d} {sum(v):10.1f} {100 * sum(v) / tot:6.1f}% {max(v):9.1f}")
print(f"{'TOTAL':28s} {sum(len(v) for v in agg.values()):4d} {tot:10.1f}")
