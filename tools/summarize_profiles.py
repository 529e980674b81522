#!/usr/bin/env python3
"""Turn the raw ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

usage: tools/summarize_profiles.py <round tag, e.g. r01> [launches.csv] [name=report.ncu-rep ...]
"""
import csv
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(path, out):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        if "search_kernel" in name:
            name = "pdmpc::search_kernel<...> block " + r["Block Size"]
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ns
        a[2] = max(a[2], ns)
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write("# ncu launch list summary (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
        f.write(f"source: {os.path.basename(path)}, {sum(a[0] for a in agg.values())} launches\n\n")
        f.write("| kernel | launches | total ms | share | longest ms |\n|---|---|---|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e6:.3f} |\n")
    print("wrote", out)


def report(name, rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary: {name} ({os.path.basename(rep)})\n\n")
        for vals in rows[2:]:
            kn = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
            f.write(f"kernel: `{kn[:120]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            d = dict(zip(hdr, zip(vals, units)))
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
            f.write("\n")
    print("wrote", out)


def main():
    tag = sys.argv[1]
    for a in sys.argv[2:]:
        if "=" in a:
            name, rep = a.split("=", 1)
            report(name, rep, os.path.join(ROOT, "profiles", f"{tag}_{name}_ncu.md"))
        else:
            launches(a, os.path.join(ROOT, "profiles", f"{tag}_launches_summary.md"))


if __name__ == "__main__":
    main()
