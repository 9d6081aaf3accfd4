#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.reader(open(sys.argv[1])):
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = d["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        k = d["Kernel Name"][:70]
        agg[k][0] += 1
        agg[k][1] += v
total = sum(v[1] for v in agg.values())
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{us / 1e3:10.3f} ms {100 * us / total:5.1f} % {n:6d} x  {k}")
