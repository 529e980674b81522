#!/usr/bin/env python3
"""Run the staged search kernel a few times on a pre-rolled record file (for ncu)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pdmpc_b200 import capi  # noqa: E402
if os.environ.get('PDMPC_LIB'):
    capi.LIB_PATH = os.environ['PDMPC_LIB']
    capi.load_library.__defaults__ = (capi.LIB_PATH,)
from pdmpc_b200.mpa import get_mpa  # noqa: E402
from pdmpc_b200.records import SearchBatch  # noqa: E402


def main():
    path = sys.argv[1]
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    tile = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    mpa_type = "triple_speed" if "triple" in path else "single_speed"
    mpa = get_mpa(mpa_type, non_convex=True)
    b = SearchBatch.load(path)
    if reps > 1:
        b = SearchBatch.concat([b] * reps)
    p = capi.Planner(0)
    p.upload_mpa(mpa)
    p.set_variant(tile)
    if os.environ.get('PDMPC_LANE_LIMITS'):
        nodes, pops = (int(x) for x in os.environ['PDMPC_LANE_LIMITS'].split(','))
        p.set_lane_limits(nodes, pops)
    p.stage(b)
    for i in range(runs):
        p.run_staged()
        p.sync()
        st = p.stats()
        print(f"variant {tile} run {i}: {b.n} searches kernel {st.kernel_ms:.3f} ms -> {b.n / st.kernel_ms * 1e3:.0f} plans/s"
              f" (lane stage {st.lanes_ms:.3f} ms, handed over {st.handed_over})")
    r = p.fetch()
    st = p.stats()
    print("pops", st.total_pops, "nodes", st.total_nodes, "cols", st.total_obstacle_cols,
          "exhausted", int(r.is_exhausted.sum()), "max pops", int(r.n_pops.max()))


if __name__ == "__main__":
    main()
