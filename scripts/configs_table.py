#!/usr/bin/env python
"""Markdown table of a profiles/<tag>_configs.json (scripts/bench_configs.py output).  python scripts/configs_table.py r01r"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(os.path.join(ROOT, "profiles", sys.argv[1] + "_configs.json")))
print("| config | kernel | B/unit | µs/launch | units/s | GB/s | % of measured peak |")
print("|---|---|---|---|---|---|---|")
for r in d["rows"]:
    if r["kernel"].startswith("rollout"):
        continue
    print("| %s | %s | %d | %.1f | %.3g | %.0f | %.1f |" % (r["config"].replace("^", "^"), r["kernel"], r["bytes_per_unit"], r["us_per_launch"],
                                                         r["units_per_s"], r["achieved_gbs"], 100 * r["frac_of_peak"]))
print()
print("| config | fused rollout T=32: env-steps/s | µs/launch | env-steps per launch |")
print("|---|---|---|---|")
for r in d["rows"]:
    if r["kernel"].startswith("rollout"):
        print("| %s | %.3g | %.0f | %.3g |" % (r["config"], r["units_per_s"], r["us_per_launch"], r["env_steps_per_launch"]))
