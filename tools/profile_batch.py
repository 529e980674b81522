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
    import glob
    files = sorted(glob.glob(path))
    b = SearchBatch.concat([SearchBatch.load(f) for f in files]) if len(files) > 1 else SearchBatch.load(files[0])
    if reps > 1:
        b = SearchBatch.concat([b] * reps)
    p = capi.Planner(0)
    p.upload_mpa(mpa)
    if os.environ.get('PDMPC_SORT_BY_POPS'):   # experiment: longest searches first (needs PDMPC_NO_REORDER=1)
        import numpy as np
        r0 = p.plan_batch(b)
        key = r0.n_pops.astype(np.int64)
        mode = os.environ['PDMPC_SORT_BY_POPS']
        order = np.argsort(-key, kind='stable') if mode == 'desc' else np.random.default_rng(0).permutation(b.n)
        b = b.select(order)
        print('sorted by pops', mode, 'max', key.max())
    p.set_variant(tile)
    p.set_cta_queue(True)   # as bench.py
    if os.environ.get('PDMPC_TILE_POINTS'):
        p.set_tile_points(int(os.environ['PDMPC_TILE_POINTS']))
    p.stage(b)
    for i in range(runs):
        p.run_staged()
        p.sync()
        st = p.stats()
        print(f"variant {tile} run {i}: {b.n} searches kernel {st.kernel_ms:.3f} ms -> {b.n / st.kernel_ms * 1e3:.0f} plans/s"
              f" (shape {st.shape}, handed over {st.handed_over})")
    r = p.fetch()
    st = p.stats()
    if os.environ.get('PDMPC_CHECK'):   # same answers as the one-search-per-warp shape
        import numpy as np
        p.set_variant(1)
        p.run_staged()
        r1 = p.fetch()
        for f in ("pop_hash", "n_pops", "n_expanded", "is_exhausted", "trims", "y_predicted", "g_path", "h_path", "shape_x"):
            a, b_ = getattr(r, f), getattr(r1, f)
            assert np.array_equal(a, b_, equal_nan=(a.dtype.kind == "f")), f
        print("identical to shape 1 on", b.n, "searches")
    if os.environ.get('PDMPC_STATS_JSON'):
        import json
        json.dump({"searches": int(b.n), "mpa": mpa_type, "shape": int(st.shape), "pops": int(st.total_pops),
                   "nodes": int(st.total_nodes), "escalated": int(st.escalated)}, open(os.environ['PDMPC_STATS_JSON'], "w"))
    print("pops", st.total_pops, "nodes", st.total_nodes, "cols", st.total_obstacle_cols,
          "exhausted", int(r.is_exhausted.sum()), "max pops", int(r.n_pops.max()))


if __name__ == "__main__":
    main()
