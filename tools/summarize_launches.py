"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*\)$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if unit in ("nsecond", "ns"):
        v /= 1e3
    elif unit in ("msecond", "ms"):
        v *= 1e3
    agg[name][0] += 1
    agg[name][1] += v
    total += v
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>10s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} {n:8d} {t:12.1f} {100 * t / total:6.1f}% {t / n:10.1f}")
print(f"{'TOTAL':70s} {sum(v[0] for v in agg.values()):8d} {total:12.1f}")
