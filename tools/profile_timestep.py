#!/usr/bin/env python3
"""One pdmpc_plan_timestep call over the committed road time-step fixture taken as ONE batch (8 time steps x 20
vehicles, 275 predecessor edges) - the launch ncu captures for the dependency-ordered CTA kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_golden_timesteps  # noqa: E402
from test_timestep_gpu import concat_timesteps  # noqa: E402
from pdmpc_b200 import capi  # noqa: E402

mpa, steps = load_golden_timesteps("timestep_road_triple_speed")
batch, deps, exp = concat_timesteps(steps)
p = capi.Planner(0)
p.upload_mpa(mpa)
for _ in range(3):
    r = p.plan_timestep(batch, deps, False)
st = p.stats()
assert (r.pop_hash == exp.pop_hash).all()
print(f"{batch.n} searches, {deps.pred_idx.size} predecessor edges: kernel {st.kernel_ms:.3f} ms, pops {st.total_pops}")
