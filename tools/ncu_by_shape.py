"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid size): tells the
tap-GEMM shapes of a step apart without instrumenting the run."""
import csv
import re
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("t2h::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
    key = (name[:44], r.get("Grid Size", "?"), r.get("Block Size", "?"))
    agg[key][0] += 1
    agg[key][1] += v
    total += v
print(f"{'kernel':44s} {'grid':>16s} {'n':>6s} {'total ms':>9s} {'us each':>8s} {'share':>6s}")
for (name, grid, blk), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{name:44s} {grid:>16s} {n:6d} {t / 1e6:9.3f} {t / n / 1e3:8.2f} {100 * t / total:5.1f}%")
print(f"TOTAL {total / 1e6:.3f} ms")
