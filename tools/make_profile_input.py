#!/usr/bin/env python3
"""Pre-roll road-network scenarios on the CPU (oracle as planner) and save the flat
search records to build/road_rN.npz so that profiling runs on the GPU box skip
the closed-loop generation.  build/ is git-ignored but travels with gpurun."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_py  # noqa: E402
from pdmpc_b200 import scenario  # noqa: E402
from pdmpc_b200.mpa import get_mpa  # noqa: E402
from pdmpc_b200.records import SearchBatch  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    mpa_type = sys.argv[2] if len(sys.argv) > 2 else "triple_speed"
    mpa = get_mpa(mpa_type, non_convex=True)
    bs = []
    for seed in range(1, n + 1):
        sc = scenario.commonroad_scenario(mpa, 20, seed=seed)
        bs.append(scenario.roll_out(sc, lambda b: oracle_py.plan_batch(mpa, b, 8), 35))
    b = SearchBatch.concat(bs)
    out = os.path.join(ROOT, "build", f"road_{mpa_type}_r{n}.npz")
    b.save(out)
    print(out, b.n, "searches")


if __name__ == "__main__":
    main()
