#!/usr/bin/env python3
"""profiles/search_kernel_traffic.json from an `ncu --set full` capture of the bench's dominant kernel.

usage: tools/capture_to_json.py <round tag> <report.ncu-rep> <stats.json written by tools/profile_batch.py>
Fails (exit 1) when the captured kernel is not the tile / search kernel or the launch is not the bench's
(searches, MPA) — bench.py only uses the figures when they describe its own launch."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag, rep, stats_path = sys.argv[1:4]
    st = json.load(open(stats_path))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    best = None
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        name = d.get("Kernel Name", "")
        if "search_tile_kernel" in name or "search_kernel" in name:
            if best is None or float(d["gpu__time_duration.sum"].replace(",", "")) > float(best["gpu__time_duration.sum"].replace(",", "")):
                best = d
    if best is None:
        sys.exit("no search kernel in the capture")
    units = dict(zip(hdr, rows[1]))

    def num(k):
        return float(best[k].replace(",", ""))

    def to_bytes(k):
        u = units[k].lower()
        return num(k) * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    tu = units["gpu__time_duration.sum"].lower().replace("second", "s")
    t_ms = num("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[tu]
    out = {
        "what": f"one launch of {best['Kernel Name'].split('(')[0]} from the ncu --set full capture profiles/{tag}_search_ncu.md "
                "(tools/gpu_round.sh: the bench's own 512-scenario records, auto launch shape)",
        "round": tag, "kernel": best["Kernel Name"].split("(")[0], "searches": st["searches"], "mpa": st["mpa"],
        "shape": st["shape"], "pops": st["pops"], "escalated": st["escalated"],
        "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
        "warp_instructions": num("smsp__inst_executed.sum"),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "sms": 148, "kernel_ms_under_ncu": t_ms,
    }
    if st["searches"] != 358400:
        sys.exit(f"capture is of {st['searches']} searches, the bench launches 358400 (512 scenarios x 35 steps x 20 vehicles)")
    json.dump(out, open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
