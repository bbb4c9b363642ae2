"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (per step).
usage: python tools/agg_ncu.py launches.csv [steps]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e3 / steps:.1f} us per step over {steps} step(s)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:46s} n={v[0] // steps:4d} us/step={v[1] / 1e3 / steps:9.1f} avg={v[1] / 1e3 / v[0]:7.1f} {v[1] / tot * 100:5.1f}%")
