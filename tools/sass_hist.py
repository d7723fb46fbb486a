#!/usr/bin/env python3
"""Opcode histogram of one kernel's SASS: tools/sass_hist.py <lib.so> <function-substring>"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
on = False
hist = collections.Counter()
ins = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    if not on:
        continue
    m = ins.match(line)
    if not m:
        continue
    parts = m.group(1).split(".")
    key = parts[0]
    if key == "IMAD" and len(parts) > 1 and parts[1] in ("WIDE", "MOV", "IADD", "SHL", "HI", "X"):
        key = "IMAD." + parts[1]
    hist[key] += 1
tot = sum(hist.values())
for k, v in hist.most_common():
    print(f"{v:6d} {100.0 * v / tot:5.1f}% {k}")
print(f"{tot:6d} TOTAL")
