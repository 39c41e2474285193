#!/usr/bin/env python
"""Condense gpurun_out/<tag>/ncu_*.csv (scripts/gpu_ncu_configs.sh) into profiles/<tag>_ncu_kernels.csv: one row per
(config, kernel name) with the median of every metric over that kernel's profiled launches."""
import csv
import glob
import os
import statistics
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rows_out = []
metrics_seen = []
for path in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", tag, "ncu_*.csv"))):
    cfg = os.path.basename(path)[4:-4]
    lines = [l for l in open(path) if l.startswith('"')]
    rd = list(csv.DictReader(lines))
    per = defaultdict(lambda: defaultdict(list))
    for r in rd:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        m = r["Metric Name"] + " [" + r["Metric Unit"] + "]"
        if m not in metrics_seen:
            metrics_seen.append(m)
        per[name][m].append(v)
    for name, ms in per.items():
        row = {"config": cfg, "kernel": name, "launches": max(len(v) for v in ms.values())}
        for m, vals in ms.items():
            row[m] = statistics.median(vals)
        rows_out.append(row)
dst = os.path.join(ROOT, "profiles", tag + "_ncu_kernels.csv")
with open(dst, "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=["config", "kernel", "launches"] + metrics_seen)
    w.writeheader()
    for r in rows_out:
        w.writerow(r)
print("wrote", dst, len(rows_out), "rows")
