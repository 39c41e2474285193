#!/usr/bin/env python
"""Summarise a gpurun_out/<tag>/ directory into profiles/ (tracked):
   profiles/<tag>_launches.csv        -- per-kernel launch count / mean / share from the ncu launch list
   profiles/<tag>_rock_step_ncu.csv   -- selected raw metrics of the `ncu --set full` capture, per launch
   profiles/rock_step_ncu_summary.json-- dram bytes per launch (read by bench.py as roofline.traffic)
Usage: python scripts/ncu_summary.py r01b [report-name]"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep_name = sys.argv[2] if len(sys.argv) > 2 else "rock_step"
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

KEYS = ["lts__t_sectors_srcunit_tex.sum", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]

# ---- launch list
lp = os.path.join(src, "launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    iname, ival = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows:
        agg[r[iname]].append(float(r[ival].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(dst, tag + "_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "mean_ns", "total_ns", "share_of_listed_gpu_time"])
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            w.writerow([k[:140], len(v), "%.0f" % (sum(v) / len(v)), "%.0f" % sum(v), "%.4f" % (sum(v) / total)])
    print("wrote", tag + "_launches.csv")

# ---- full capture
rp = os.path.join(src, rep_name + ".ncu-rep")
if os.path.exists(rp):
    out = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h in KEYS or h in ("Kernel Name", "ID")]
    with open(os.path.join(dst, "%s_%s_ncu.csv" % (tag, rep_name)), "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in data:
            w.writerow([r[i] for i in cols])
    def col(name):
        i = hdr.index(name)
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(units[i], 1.0)
        return [float(r[i].replace(",", "")) * scale for r in data]
    rd, wr, t = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
    summ = {"tag": tag, "report": rep_name, "launches_captured": len(data), "kernel": data[0][hdr.index("Kernel Name")][:120],
            "dram_bytes_read_per_launch": sum(rd) / len(rd), "dram_bytes_write_per_launch": sum(wr) / len(wr),
            "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
            "gpu_time_us_under_ncu": sum(t) / len(t),
            "l2_bytes_from_sms_per_launch": (32.0 * sum(col("lts__t_sectors_srcunit_tex.sum")) / len(rd)
                                             if "lts__t_sectors_srcunit_tex.sum" in hdr else None),
            "note": "writes still resident in the 126 MB L2 at kernel end are not in dram__bytes_write; "
                    "l2_bytes_from_sms = 32 B x lts__t_sectors_srcunit_tex.sum (every byte the SMs moved through L2)"}
    name = "rock_step_ncu_summary.json" if rep_name == "rock_step" else rep_name + "_ncu_summary.json"
    with open(os.path.join(dst, name), "w") as f:
        json.dump(summ, f, indent=1)
    print(json.dumps(summ))
