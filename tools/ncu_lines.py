#!/usr/bin/env python3
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv > x.csv; ncu_lines.py x.csv [top]"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    samples = defaultdict(int)
    insts = defaultdict(int)
    text = {}
    hdr = None
    fname = ""
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            i_s = hdr.index("# Samples")
            i_i = hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) <= i_i:
            continue
        if r[0].strip().isdigit():
            key = (fname, int(r[0]))
            text[key] = r[1]
        try:
            samples[key] += int(r[i_s])
            insts[key] += int(r[i_i])
        except (ValueError, UnboundLocalError):
            pass
    tot = sum(samples.values()) or 1
    toti = sum(insts.values()) or 1
    print(f"total samples {tot}, warp instructions {toti}")
    for key, s in sorted(samples.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{100*s/tot:5.1f}% smp {100*insts[key]/toti:5.1f}% ins  {key[0]}:{key[1]:4d}  {text.get(key,'').strip()[:100]}")


if __name__ == "__main__":
    main()
