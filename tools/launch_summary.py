#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. Dev tool."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    m = re.match(r"(?:void )?(?:mmcm::)?(\w+)(<[^>]*>)?", name)
    key = (m.group(1) + (m.group(2) or "")) if m else name
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:48s} n={n:5d} total={t:10.1f} us  avg={t / n:8.1f} us  share={t / tot * 100:5.1f}%")
print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
