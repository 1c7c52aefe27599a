"""Developer aid: condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into (kernel, launches, total us)
of the LAST forward in the log (tools/profile_config.py runs two; the first one also prepares the derived weights).
Usage: python tools/launch_summary.py out.csv [launches per forward, default: half of the log] [order]"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        rows.append((r["Kernel Name"], us))
n_last = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else len(rows) // 2
half = rows[-n_last:]
agg = OrderedDict()
for k, us in half:
    k = k.split("(")[0]
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + us)
print(f"{len(half)} launches, {sum(t for _, t in agg.values()):.1f} us (cold caches, serialised)")
if sys.argv[-1] == "order":      # in launch order
    for k, us in half:
        print(f"  {us:8.1f}  {k.split('(')[0]}")
else:
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {t:9.1f} us  {n:4d} x  {k}")
