#!/usr/bin/env python3
"""A few two-vehicle joint searches (crossing courses) - the launch ncu captures for joint_search_kernel.
Keep it SMALL under ncu: every replay pass saves and restores the node arenas (slots x capacity x 144 B); a
148-search launch with 2^21-node arenas (42 GB) did not finish 40 passes within 400 s (round 1, session 9)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import straight_iter  # noqa: E402
from pdmpc_b200 import capi  # noqa: E402
from pdmpc_b200.mpa import get_mpa  # noqa: E402
from pdmpc_b200.records import CHECKER_SAT, SearchBatch  # noqa: E402

mpa = get_mpa("single_speed", non_convex=False)
rows = []
rng = np.random.default_rng(1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    gap, off = rng.uniform(0.2, 0.6), rng.uniform(0.4, 0.9)
    rows += [straight_iter(mpa, x=0.0, y=0.0, yaw=0.0), straight_iter(mpa, x=off, y=-gap, yaw=np.pi / 2)]
b = SearchBatch.from_iters(rows, mpa.Hp, CHECKER_SAT, mpa.dt_seconds)
p = capi.Planner(0)
p.upload_mpa(mpa)
p.set_node_capacity(1 << 19)
r = p.joint_plan_batch(b, 2, False)
st = p.stats()
print(f"{len(rows) // 2} joint searches x 2 vehicles: kernel {st.kernel_ms:.1f} ms, pops {st.total_pops}, nodes {st.total_nodes}")
