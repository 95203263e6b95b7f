"""Summarises an ncu `--metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total and mean
duration, share of the step. Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/x.md"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    agg[name][0] += 1
    agg[name][1] += v_us
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |")
print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | | |")
