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
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    agg[name][0] += 1
    agg[name][1] += v
    total += v
print(f"{'kernel':60s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {n:8d} {t / 1e6:10.3f} {100 * t / total:6.1f}%")
print(f"{'TOTAL':60s} {sum(v[0] for v in agg.values()):8d} {total / 1e6:10.3f}")
