#!/usr/bin/env python3
"""Throughput of the sampled optimizer on a pre-rolled record file (staged inputs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from pdmpc_b200 import capi  # noqa: E402
from pdmpc_b200.mpa import get_mpa  # noqa: E402
from pdmpc_b200.records import SearchBatch  # noqa: E402

path = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n_max = int(sys.argv[3]) if len(sys.argv) > 3 else 250
mpa = get_mpa("triple_speed" if "triple" in path else "single_speed", non_convex=True)
b = SearchBatch.load(path)
if reps > 1:
    b = SearchBatch.concat([b] * reps)
seeds = (np.arange(b.n) % 35 + 2).astype(np.uint32)
p = capi.Planner(0)
p.upload_mpa(mpa)
p.stage(b)
for i in range(3):
    p.mcts_run_staged(seeds, n_max)
    p.sync()
    st = p.stats()
    print(f"mcts run {i}: {b.n} searches, n_max {n_max}: kernel {st.kernel_ms:.3f} ms -> {b.n / st.kernel_ms * 1e3:.0f} plans/s")
r = p.fetch()
st = p.stats()
print("traversal steps", st.total_pops, "nodes", st.total_nodes, "cols", st.total_obstacle_cols,
      "exhausted", int(r.is_exhausted.sum()), "expansions/search", float(r.n_expanded.mean()))
