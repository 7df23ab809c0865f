#!/usr/bin/env python
"""Per-kernel mean device time from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg.setdefault(r[ki].split("(")[0], []).append(v)
for k, v in agg.items():
    print(f"{k[:64]:64s} n={len(v):3d} mean={sum(v)/len(v)/1000:8.1f} us")
