#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  Usage: launch_table.py file.csv [title]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("# per-launch device times under ncu are cold-cache and serialised: compare SHARES, not absolutes")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} launches {a[0]:4d}  total {a[1] / 1e3:10.1f} us  {100 * a[1] / tot:5.1f} %")
print(f"{'total':60s} launches {sum(a[0] for a in agg.values()):4d}  total {tot / 1e3:10.1f} us")
