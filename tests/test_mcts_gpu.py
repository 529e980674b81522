"""Parity of the CUDA sampled optimizer (pdmpc_mcts_plan_batch, MonteCarloTreeSearch.m)
against the committed fixtures and the CPU oracle: bit-exact roll-out trace, counts, flags,
trims; bit-identical poses, costs and shapes."""
import dataclasses

import numpy as np
import pytest

from oracle import oracle_py, parity
from pdmpc_b200 import capi
from pdmpc_b200.mpa import get_mpa
from pdmpc_b200.records import CHECKER_SAT, SearchBatch

from helpers import GOLDEN_CASES, load_golden, load_golden_mcts, rect, road_records, straight_iter

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mcts_golden_fixture(planner, name):
    """No oracle at run time: expected outputs were written after both CPU restatements agreed."""
    mpa, batch, _ = load_golden(name)
    seeds, n_max, exp = load_golden_mcts(name)
    planner.upload_mpa(mpa)
    dev = planner.mcts_plan_batch(batch, seeds, n_max)
    info = parity.compare(dev, exp)
    assert info["n"] == batch.n


@pytest.mark.parametrize("n_max", [1, 7, 250, 600])
def test_mcts_against_oracle_budgets(planner, n_max):
    mpa, batch = road_records("triple_speed", 8)
    seeds = (np.arange(batch.n) % 35 + 2).astype(np.uint32)
    planner.upload_mpa(mpa)
    dev = planner.mcts_plan_batch(batch, seeds, n_max)
    ref = oracle_py.mcts_plan_batch(mpa, batch, seeds, n_max, 4)
    parity.compare(dev, ref)
    assert (dev.n_expanded < n_max + mpa.Hp).all()


def test_mcts_sat_checker_and_realistic_mpa(planner):
    mpa = get_mpa("single_speed", non_convex=False)
    _, b = road_records("single_speed", 4)
    b = dataclasses.replace(b, checker=CHECKER_SAT)
    seeds = np.full(b.n, 11, dtype=np.uint32)
    planner.upload_mpa(mpa)
    parity.compare(planner.mcts_plan_batch(b, seeds, 90), oracle_py.mcts_plan_batch(mpa, b, seeds, 90, 4))
    mpa2, b2 = road_records("realistic", 3, amount=10, seed=5)
    seeds2 = (np.arange(b2.n) + 2).astype(np.uint32)
    planner.upload_mpa(mpa2)
    parity.compare(planner.mcts_plan_batch(b2, seeds2, 250), oracle_py.mcts_plan_batch(mpa2, b2, seeds2, 250, 4))


def test_mcts_edge_cases(planner):
    """Empty batch; a vehicle walled in (every root edge invalid -> exhausted after one expansion per
    root successor); seed 0 (MATLAB maps it to 5489); staged form equals the host-buffer form."""
    mpa = get_mpa("single_speed", non_convex=False)
    planner.upload_mpa(mpa)
    Hp = mpa.Hp
    empty = SearchBatch.from_iters([], Hp, CHECKER_SAT, mpa.dt_seconds)
    r = planner.mcts_plan_batch(empty, np.zeros(0, dtype=np.uint32), 10)
    assert r.status.size == 0
    its = [straight_iter(mpa), straight_iter(mpa, obstacles=[rect(0.1, 0.0, 2.0, 2.0)]),
           straight_iter(mpa, x=1.0, y=2.0, yaw=0.7)]
    b = SearchBatch.from_iters(its, Hp, CHECKER_SAT, mpa.dt_seconds)
    seeds = np.array([0, 3, 5489], dtype=np.uint32)
    dev = planner.mcts_plan_batch(b, seeds, 30)
    ref = oracle_py.mcts_plan_batch(mpa, b, seeds, 30)
    parity.compare(dev, ref)
    assert dev.is_exhausted.tolist() == [0, 1, 0]
    assert dev.n_expanded[1] == int(mpa.transition[0, its[1].trim_indices - 1].sum())
    planner.stage(b)
    planner.mcts_run_staged(seeds, 30)
    parity.compare(planner.fetch(), ref)
    with pytest.raises(capi.PdmpcError):
        planner.mcts_plan_batch(b, seeds, 0)
    with pytest.raises(capi.PdmpcError):
        planner.mcts_plan_batch(b, seeds, 5000)


def test_mcts_large_batch_properties(planner):
    """At bench scale (no oracle): same seed -> same plan; exhausted <=> no trims; node ids grow
    along the path; every successful plan's first pose is one maneuver away from the start."""
    mpa, batch = road_records("triple_speed", 8)
    big = SearchBatch.concat([batch] * 8)
    seeds = np.tile((np.arange(batch.n) % 35 + 2).astype(np.uint32), 8)
    planner.upload_mpa(mpa)
    a = planner.mcts_plan_batch(big, seeds, 250)
    assert (a.status == 0).all()
    for f in ("pop_hash", "trims", "tree_path", "n_expanded"):
        v = getattr(a, f).reshape(8, batch.n, -1)
        assert (v == v[0]).all(), f
    ok = a.is_exhausted == 0
    assert ((a.trims[:, 1:] > 0).all(axis=1) == ok).all()
    assert (np.diff(a.tree_path[ok], axis=1) > 0).all()
    d = np.hypot(a.y_predicted[ok, 0, 0] - big.x0[ok], a.y_predicted[ok, 0, 1] - big.y0[ok])
    assert d.max() < 0.5
