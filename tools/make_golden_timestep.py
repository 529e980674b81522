#!/usr/bin/env python3
"""Golden fixtures for pdmpc_plan_timestep: tests/golden/timestep_*.npz.

Each fixture holds the MPA tables and, for a run of consecutive time steps of a closed loop,
the ONE-CALL inputs (every vehicle's base iter_v as a flat batch, the predecessor lists, the
fallback areas) with the expected outputs.  Expected outputs = the reference's way through a
time step (computation levels one after the other, host-side obstacle assembly:
scenario.plan_timestep_by_levels) planned by the C oracle; the file is only written after the
same time steps planned by the SECOND restatement (oracle/matlab_literal.py, matrix form, the
reference's own priority-queue source) give the same flags, node counts, trims, poses and areas
for every vehicle — including the hand-over of predecessors' areas between levels.

Run here (build container):  python tools/make_golden_timestep.py
"""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import matlab_literal as ml  # noqa: E402
from oracle import oracle_py  # noqa: E402
from pdmpc_b200 import scenario  # noqa: E402
from pdmpc_b200.mpa import get_mpa  # noqa: E402
from pdmpc_b200.records import BatchResult  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ml_plan_batch(mpa, batch):
    """plan_fn of the matrix-form restatement (the fields the time-step hand-over and the check read)."""
    r = BatchResult.empty(batch.n, batch.Hp)
    for i, it in enumerate(batch.to_iters()):
        info = ml.do_graph_search(it, mpa, batch.checker, trig="spec", use_reference_pq=True)
        r.status[i] = 0
        r.is_exhausted[i] = info.is_exhausted
        r.n_expanded[i] = info.n_expanded
        r.trims[i, 0] = it.trim_indices
        r.y_predicted[i] = np.nan
        if not info.is_exhausted:
            r.trims[i, 1:] = info.predicted_trims
            r.y_predicted[i] = info.y_predicted
            for k, shp in enumerate(info.shapes):
                r.shape_npts[i, k] = shp.shape[1]
                r.shape_x[i, k, :shp.shape[1]] = shp[0]
                r.shape_y[i, k, :shp.shape[1]] = shp[1]
    return r


def make(name, mpa, sc, steps, keep):
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    runner = scenario.ScenarioRunner(sc, None, timestep_fn=lambda b, d: scenario.plan_timestep_by_levels(plan, b, d))
    runner.run(steps)
    recs = runner.timestep_records[-keep:]
    d = {}
    for f in dataclasses.fields(mpa):
        d["mpa__" + f.name] = np.asarray(getattr(mpa, f.name))
    n_pred = n_exh = 0
    for s, (_k, batch, deps, ref) in enumerate(recs):
        second = scenario.plan_timestep_by_levels(lambda b: ml_plan_batch(mpa, b), batch, deps)
        for fld in ("is_exhausted", "n_expanded", "trims", "shape_npts"):
            assert np.array_equal(getattr(second, fld), getattr(ref, fld)), (name, s, fld)
        for fld in ("y_predicted", "shape_x", "shape_y"):
            assert np.array_equal(getattr(second, fld), getattr(ref, fld), equal_nan=True), (name, s, fld)
        n_pred += deps.pred_idx.size
        n_exh += int(ref.is_exhausted.sum())
        for prefix, obj in ((f"s{s}__in__", batch), (f"s{s}__deps__", deps), (f"s{s}__out__", ref)):
            for f in dataclasses.fields(obj):
                d[prefix + f.name] = np.asarray(getattr(obj, f.name))
    d["n_steps"] = np.asarray(len(recs))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: {len(recs)} time steps, {sum(r[1].n for r in recs)} searches, {n_pred} predecessor edges, "
          f"{n_exh} exhausted; both restatements agree")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    mpa = get_mpa("triple_speed", non_convex=True)
    make("timestep_road_triple_speed", mpa, scenario.commonroad_scenario(mpa, 20, seed=1), 14, 8)
    mpa = get_mpa("single_speed", non_convex=False)
    make("timestep_circle_single_speed", mpa, scenario.circle_scenario(mpa, 4), 22, 12)
