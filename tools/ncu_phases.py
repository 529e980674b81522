#!/usr/bin/env python3
"""Instruction share per source-line range of an `ncu --page source --print-source cuda,sass --csv` dump
(SASS rows are attributed to the CUDA line they follow).
usage: ncu_phases.py dump.csv file.cuh:lo-hi=name ..."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
ranges = []
for a in sys.argv[2:]:
    loc, name = a.split("=")
    f, r = loc.split(":")
    lo, hi = r.split("-")
    ranges.append((f, int(lo), int(hi), name))
ins = defaultdict(int); smp = defaultdict(int)
fname = ""; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= i_i:
        continue
    if r[0].strip().isdigit():
        ln = int(r[0]); lfile = fname
        continue
    try:
        n = int(r[i_i]); s = int(r[i_s])
    except ValueError:
        continue
    key = "other:" + lfile
    for f, lo, hi, name in ranges:
        if f == lfile and lo <= ln <= hi:
            key = name; break
    ins[key] += n; smp[key] += s
ti = sum(ins.values()) or 1; ts = sum(smp.values()) or 1
print(f"warp instructions {ti}, samples {ts}")
for k, v in sorted(ins.items(), key=lambda kv: -kv[1]):
    print(f"{100*v/ti:5.1f}% ins {100*smp[k]/ts:5.1f}% smp  {v:14d}  {k}")
