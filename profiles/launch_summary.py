"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by (kernel, grid).
usage: python profiles/launch_summary.py gpurun_out/launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
d = defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = r[ki].split("(")[0].replace("<unnamed>::", "")
    d[(name, r[gi] if gi is not None else "")].append(v * scale)
tot = sum(sum(v) for v in d.values())
print(f"{'kernel':34s} {'grid':>14s} {'n':>5s} {'mean us':>10s} {'share':>7s}")
for (k, g), v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:34s} {g:>14s} {len(v):5d} {sum(v)/len(v):10.1f} {100*sum(v)/tot:6.1f}%")
