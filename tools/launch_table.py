#!/usr/bin/env python3
"""Per-kernel durations of the LAST MSM in an ncu launch list (csv of gpu__time_duration.sum).
usage: tools/launch_table.py launches.csv [first_kernel_substring]"""
import csv, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    try:
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]))
    except (KeyError, ValueError):
        pass
first = sys.argv[2] if len(sys.argv) > 2 else "digits"
starts = [i for i, r in enumerate(rows) if first in r[0] and "hist" in r[0]]
last = rows[starts[-1]:] if starts else rows
tot = 0
for name, v, unit in last:
    us = v / 1000.0 if unit in ("ns", "nsecond") else v
    tot += us
    print(f"{us:10.1f} us  {name[:90]}")
print(f"{tot:10.1f} us  total of {len(last)} launches")
