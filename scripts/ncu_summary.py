#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean and share."""
import collections, csv, sys
path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    k = row["Kernel Name"][:86]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else v * 1000 if row["Metric Unit"] == "ms" else v
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{n - skip} launches, {tot / 1000:.3f} ms total")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{a[1] / tot * 100:5.1f}%  {a[1] / a[0]:9.1f} us x{a[0]:4d}  {k}")
