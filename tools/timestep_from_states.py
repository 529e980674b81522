#!/usr/bin/env python3
"""Wall time of the device side of one 20-vehicle time step: pdmpc_plan_timestep_from_states (one call) against
pdmpc_sample_inputs + pdmpc_assemble_obstacles + pdmpc_plan_timestep_closed_loop (three calls, the batch assembled on the
host in between).  Both closed loops end in the same poses (asserted)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from pdmpc_b200 import capi, scenario  # noqa: E402
from pdmpc_b200.mpa import get_mpa  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 35
SEEDS = [int(s) for s in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 4]
mpa = get_mpa("triple_speed", non_convex=True)
p = capi.Planner(0)
p.upload_mpa(mpa)
p.set_cta_queue(True)
p.upload_reachable_sets(scenario.local_reachable_sets_conv(mpa))
hl, hw = scenario.VEH_LENGTH / 2 + 0.01, scenario.VEH_WIDTH / 2 + 0.01
t_one, t_three = [], []


def timed(bucket, fn):
    def wrapped(*a):
        t0 = time.perf_counter()
        r = fn(*a)
        bucket[-1] += (time.perf_counter() - t0) * 1e3
        return r
    return wrapped


for seed in SEEDS:
    sc = scenario.commonroad_scenario(mpa, 20, seed=seed)
    p.upload_road(scenario.road_tables([sc]))
    p.closed_loop_reset(20, hl, hw)
    one = scenario.ScenarioRunner(sc, None, states_fn=timed(t_one, lambda *a: p.plan_timestep_from_states(*a, raise_on_search_error=False)))
    for _ in range(STEPS):
        t_one.append(0.0)
        one.step_timestep()
    p.closed_loop_reset(20, hl, hw)
    three = scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 20, seed=seed), None,
                                    inputs_fn=timed(t_three, p.sample_inputs), obstacles_fn=timed(t_three, p.assemble_obstacles),
                                    closed_loop_fn=timed(t_three, lambda b, d, s, st: p.plan_timestep_closed_loop(b, d, s, st, False)))
    for _ in range(STEPS):
        t_three.append(0.0)
        three.step_timestep()
    assert np.array_equal(one.pose, three.pose) and np.array_equal(one.trim, three.trim)
for name, t in (("one call (pdmpc_plan_timestep_from_states)", t_one), ("three calls (inputs, obstacles, planning)", t_three)):
    t = np.array(t[5:])
    print(f"{name}: p50 {np.percentile(t, 50):.3f} ms  p99 {np.percentile(t, 99):.3f} ms  max {t.max():.3f} ms  over {t.size} time steps")
