#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

The reference (MATLAB) cannot run in this image and ships no golden vectors for
the search itself, so the fixtures are produced by the C oracle
and are only written after the SECOND, independent restatement
(oracle/matlab_literal.py, matrix-form numpy driving the reference's own
unmodified priority-queue source compiled into oracle/_ref) reproduces every
search of the fixture: pop sequence, node count, flags, trims and bit-identical
poses.  Each .npz holds the MPA tables, the flat search records and the expected
outputs, so tests on the GPU box need neither /root/reference nor scipy.

Also writes kat_lanelet1.npz: lanelets{1} of the reference's LabMapCommonRoad.xml
(the geometry the reference's own unit tests use,
tests/unittests/hlc/intersect_unittest.m:8-36).

Run here (build container):  python tools/make_golden.py
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
from pdmpc_b200.records import SearchBatch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def roll(sc, mpa, steps):
    """Closed loop with the oracle as planner; returns (iters, batch)."""
    iters = []
    orig = SearchBatch.from_iters

    def capture(its, Hp, checker, dt):
        iters.extend(its)
        return orig(its, Hp, checker, dt)

    SearchBatch.from_iters = staticmethod(capture)
    try:
        recs = scenario.ScenarioRunner(sc, lambda b: oracle_py.plan_batch(mpa, b)).run(steps)
    finally:
        SearchBatch.from_iters = staticmethod(orig)
    return iters, SearchBatch.concat([r.batch for r in recs])


def cross_check(mpa, iters, batch, ref):
    for i, it in enumerate(iters):
        info = ml.do_graph_search(it, mpa, batch.checker, trig="spec", use_reference_pq=True)
        assert info.is_exhausted == bool(ref.is_exhausted[i]), i
        assert info.n_expanded == int(ref.n_expanded[i]), i
        assert info.pops == list(oracle_py.plan_trace(mpa, batch, i)), i
        if not info.is_exhausted:
            assert info.predicted_trims == list(ref.trims[i, 1:]), i
            assert info.tree_path == list(ref.tree_path[i]), i
            assert np.array_equal(info.y_predicted.view(np.uint64), ref.y_predicted[i].view(np.uint64)), i
            for k, shp in enumerate(info.shapes):
                n = int(ref.shape_npts[i, k])
                assert shp.shape[1] == n
                assert np.array_equal(shp[0].view(np.uint64), ref.shape_x[i, k, :n].view(np.uint64))
                assert np.array_equal(shp[1].view(np.uint64), ref.shape_y[i, k, :n].view(np.uint64))


def save(name, mpa, batch, ref):
    d = {}
    for f in dataclasses.fields(mpa):
        d["mpa__" + f.name] = np.asarray(getattr(mpa, f.name))
    for f in dataclasses.fields(batch):
        d["in__" + f.name] = np.asarray(getattr(batch, f.name))
    for f in dataclasses.fields(ref):
        d["out__" + f.name] = np.asarray(getattr(ref, f.name))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{path}: {batch.n} searches, {int(ref.n_pops.sum())} pops, {int(ref.is_exhausted.sum())} exhausted, "
          f"{os.path.getsize(path)} bytes")


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = [
        ("circle_sat_single_speed", "circle", "single_speed", False, 4, 0, 20),
        ("road_interx_single_speed", "road", "single_speed", True, 12, 2, 8),
        ("road_interx_triple_speed", "road", "triple_speed", True, 20, 1, 6),
    ]
    for name, kind, mpa_type, non_convex, amount, seed, steps in cases:
        mpa = get_mpa(mpa_type, non_convex=non_convex)
        sc = scenario.circle_scenario(mpa, amount) if kind == "circle" else \
            scenario.commonroad_scenario(mpa, amount, seed=seed)
        iters, batch = roll(sc, mpa, steps)
        ref = oracle_py.plan_batch(mpa, batch)
        cross_check(mpa, iters, batch, ref)
        save(name, mpa, batch, ref)
    # lanelets{1} for the reference's intersect_lanelets KATs
    road = scenario.road_map()
    np.savez_compressed(os.path.join(OUT, "kat_lanelet1.npz"), lanelet=road.lanelets[0])
    print("lanelet 1:", road.lanelets[0].shape, "left starts", road.lanelets[0][0, 2:4], "right starts",
          road.lanelets[0][0, 0:2])


if __name__ == "__main__":
    main()
