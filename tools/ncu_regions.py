#!/usr/bin/env python3
"""Where a kernel's warp instructions and stall samples go, from an ncu report with source (--import-source on):
usage: tools/ncu_regions.py report.ncu-rep [chunk]   -- SASS in chunks of `chunk` instructions + the top stall sites"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, ii, it, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ii]), int(r[it]), int(r[ismp]), r[isrc]))
    except (ValueError, IndexError):
        pass
tot, tots = sum(d[0] for d in data), sum(d[2] for d in data)
print("total warp inst", tot, "samples", tots, "sass instructions", len(data))
for i in range(0, len(data), chunk):
    seg = data[i:i + chunk]
    wi, ti, sm = sum(d[0] for d in seg), sum(d[1] for d in seg), sum(d[2] for d in seg)
    if wi / tot > 0.004 or sm / tots > 0.004:
        print("%5d  warp inst %5.1f%%  lanes %4.1f  samples %5.1f%%   %s" % (i, 100 * wi / tot, ti / max(wi, 1), 100 * sm / tots, seg[0][3][:60]))
print("top stall sites:")
for d in sorted(data, key=lambda d: -d[2])[:20]:
    print("%6d samples  %9d exec  %4.1f lanes  %s" % (d[2], d[0], d[1] / max(d[0], 1), d[3][:90]))
