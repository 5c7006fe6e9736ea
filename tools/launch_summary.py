#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)
    short = re.sub(r"<.*", "", r["Kernel Name"].replace("<unnamed>::", "")).split("::")[-1].replace("void ", "")
    agg[short][0] += 1
    agg[short][1] += v
    seq.append((short, v))
tot = sum(v for _, v in agg.values())
print(f"{'kernel':34s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} {c:8d} {v / 1e3:10.2f} {v / c:10.1f} {100 * v / tot:6.1f}%")
print(f"{'total':34s} {sum(c for c, _ in agg.values()):8d} {tot / 1e3:10.2f}")
if len(sys.argv) > 2:
    for name in sys.argv[2:]:
        vals = [v for k, v in seq if k == name]
        if vals:
            n = len(vals)
            print(name, "first/25%/50%/75%/last us:", [round(vals[i], 1) for i in (0, n // 4, n // 2, 3 * n // 4, n - 1)])
