#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python scripts/launch_list.py launches.csv [last_n_launches_to_print]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
start = rows.index(hdr)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    try:
        seq.append((r[ki].split("(")[0][-70:], float(r[vi].replace(",", "")) / 1000.0))
    except ValueError:
        pass
agg = collections.OrderedDict()
for n, v in seq:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
print(f"{len(seq)} launches")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us total {c:5d} launches {t / c:9.1f} us avg  {n}")
if len(sys.argv) > 2:
    for n, v in seq[-int(sys.argv[2]):]:
        print(f"{v:9.1f} us  {n}")
