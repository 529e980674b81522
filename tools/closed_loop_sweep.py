#!/usr/bin/env python3
"""BASELINE configs[4] in CLOSED LOOP: R road-network scenarios advanced in lockstep, every time step of all of
them as ONE pdmpc_plan_timestep call (R x 20 searches with their predecessor DAGs; the hand-over of the
predecessors' areas happens on the device).  Only the optimizer calls are timed — the scenario logic around
them (reference trajectories, coupling, priorities) is the Python stand-in of the MATLAB callers.

  python tools/closed_loop_sweep.py [--scenarios 16,64,128] [--steps 12] [--check 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", default="16,64,128")
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--mpa", default="triple_speed")
    ap.add_argument("--variant", type=int, default=0, help="pdmpc_set_variant: 0 auto, 1 warp per search, 4 CTA per search")
    ap.add_argument("--check", type=int, default=2, help="scenarios re-run alone (one call per time step) and compared")
    args = ap.parse_args()
    from pdmpc_b200 import capi, scenario
    from pdmpc_b200.mpa import get_mpa
    mpa = get_mpa(args.mpa, non_convex=True)
    planner = capi.Planner(0)
    planner.upload_mpa(mpa)
    planner.set_variant(args.variant)
    rows = []
    for R in [int(x) for x in args.scenarios.split(",")]:
        runners = [scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 20, seed=1 + s), None,
                                           timestep_fn=lambda b, d: planner.plan_timestep(b, d, False)) for s in range(R)]
        call_ms, kern_ms, pops = [], [], 0

        def timed(b, d):
            t0 = time.perf_counter()
            r = planner.plan_timestep(b, d, False)
            call_ms.append((time.perf_counter() - t0) * 1e3)
            st = planner.stats()
            kern_ms.append(st.kernel_ms)
            return r

        t_all = time.perf_counter()
        for k in range(args.steps):
            res = scenario.lockstep_step(runners, timed)
            pops += int(res.n_pops.sum())
        t_all = time.perf_counter() - t_all
        for s in range(min(args.check, R)):     # the same scenario alone must end in the same state
            a = scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 20, seed=1 + s), None,
                                        timestep_fn=lambda b, d: planner.plan_timestep(b, d, False))
            a.run(args.steps)
            assert np.array_equal(a.pose, runners[s].pose) and np.array_equal(a.trim, runners[s].trim)
        n = R * 20
        warm = call_ms[1:] if len(call_ms) > 1 else call_ms
        rows.append({"variant": args.variant, "scenarios": R, "searches_per_call": n, "steps": args.steps,
                     "call_ms_p50": float(np.percentile(warm, 50)), "call_ms_max": float(np.max(warm)),
                     "kernel_ms_p50": float(np.percentile(kern_ms[1:] or kern_ms, 50)),
                     "plans_per_s_planning_calls": n * len(warm) / (sum(warm) * 1e-3),
                     "pops_per_plan": pops / (n * args.steps),
                     "wall_s_including_python_host_logic": round(t_all, 1)})
        print(json.dumps(rows[-1]), flush=True)
    planner.close()


if __name__ == "__main__":
    main()
