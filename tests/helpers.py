"""Shared builders for the test-suite (seeded, deterministic)."""
from __future__ import annotations

import functools

import numpy as np

from oracle import oracle_py
from pdmpc_b200 import scenario
from pdmpc_b200.mpa import get_mpa
from pdmpc_b200.records import CHECKER_INTERX, CHECKER_SAT, IterationData, SearchBatch


@functools.lru_cache(maxsize=None)
def circle_records(steps: int = 30, amount: int = 4):
    """BASELINE configs[0]: circle, 4 vehicles, constant priority, single_speed MPA, SAT."""
    mpa = get_mpa("single_speed", non_convex=False)
    sc = scenario.circle_scenario(mpa, amount)
    batch = scenario.roll_out(sc, lambda b: oracle_py.plan_batch(mpa, b), steps)
    return mpa, batch


@functools.lru_cache(maxsize=None)
def road_records(mpa_type: str = "triple_speed", steps: int = 8, amount: int = 20, seed: int = 1):
    """BASELINE configs[1]: road network, 20 vehicles, coloring priorities, InterX."""
    mpa = get_mpa(mpa_type, non_convex=True)
    sc = scenario.commonroad_scenario(mpa, amount, seed=seed)
    batch = scenario.roll_out(sc, lambda b: oracle_py.plan_batch(mpa, b), steps)
    return mpa, batch


def rect(cx, cy, hw, hh):
    """closed axis-aligned rectangle [2, 5]"""
    return np.array([[cx - hw, cx - hw, cx + hw, cx + hw, cx - hw],
                     [cy - hh, cy + hh, cy + hh, cy - hh, cy - hh]], dtype=np.float64)


def straight_iter(mpa, x=0.0, y=0.0, yaw=0.0, speed=None, **kw) -> IterationData:
    """One vehicle at the origin heading +x with a straight reference line."""
    Hp = mpa.Hp
    v = float(mpa.get_straight_speeds_of_mpa().max()) if speed is None else speed
    d = v * mpa.dt_seconds * np.arange(1, Hp + 1)
    ref = np.column_stack([x + np.cos(yaw) * d, y + np.sin(yaw) * d])
    return IterationData(x0=np.array([x, y, yaw, 0.0]), trim_indices=mpa.trim_from_values(0.0, 0.0),
                         reference_trajectory_points=ref, v_ref=np.full(Hp, v), **kw)


GOLDEN_DIR = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")
GOLDEN_CASES = ("circle_sat_single_speed", "road_interx_single_speed", "road_interx_triple_speed")


def load_golden(name: str):
    """tests/golden/<name>.npz (tools/make_golden.py) -> (mpa, batch, expected BatchResult)."""
    import dataclasses
    import os

    from pdmpc_b200.mpa import MotionPrimitiveAutomaton
    from pdmpc_b200.records import BatchResult

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def build(cls, prefix):
        kw = {}
        for f in dataclasses.fields(cls):
            a = z[prefix + f.name]
            kw[f.name] = a.item() if a.ndim == 0 else a
        return cls(**kw)

    mpa = build(MotionPrimitiveAutomaton, "mpa__")
    batch = build(SearchBatch, "in__")
    exp = build(BatchResult, "out__")
    return mpa, batch, exp


def load_golden_mcts(name: str):
    """tests/golden/mcts_expected.npz (tools/make_golden_mcts.py) -> (seeds, n_max, expected BatchResult)
    of the sampled optimizer on the records of fixture `name`."""
    import dataclasses
    import os

    from pdmpc_b200.records import BatchResult

    z = np.load(os.path.join(GOLDEN_DIR, "mcts_expected.npz"))
    kw = {}
    for f in dataclasses.fields(BatchResult):
        a = z[f"{name}__out__{f.name}"]
        kw[f.name] = a.item() if a.ndim == 0 else a
    return z[name + "__seeds"], int(z[name + "__n_max"]), BatchResult(**kw)


def load_golden_timesteps(name: str):
    """tests/golden/timestep_*.npz (tools/make_golden_timestep.py) ->
    (mpa, [(batch, deps, expected BatchResult) per time step])."""
    import dataclasses
    import os

    from pdmpc_b200.mpa import MotionPrimitiveAutomaton
    from pdmpc_b200.records import BatchResult, TimestepDeps

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))

    def build(cls, prefix):
        kw = {}
        for f in dataclasses.fields(cls):
            a = z[prefix + f.name]
            kw[f.name] = a.item() if a.ndim == 0 else a
        return cls(**kw)

    mpa = build(MotionPrimitiveAutomaton, "mpa__")
    steps = [(build(SearchBatch, f"s{s}__in__"), build(TimestepDeps, f"s{s}__deps__"), build(BatchResult, f"s{s}__out__"))
             for s in range(int(z["n_steps"]))]
    return mpa, steps
