#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total ms, average ms."""
import collections
import csv
import gzip
import sys

path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
hdr, agg = None, collections.OrderedDict()
for r in csv.reader(op(path, "rt")):
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(d["Metric Unit"], 1e-6)
    a = agg.setdefault(d["Kernel Name"][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(t for _, t in agg.values())
for k, (n, t) in agg.items():
    print(f"{n:5d} {t:10.3f} ms  avg {t / n:8.4f}  {100 * t / tot:5.1f}%  {k}")
