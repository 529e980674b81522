"""pdmpc_plan_timestep (one dependency-ordered launch for a whole time step) against its
specification: the level-by-level loop with host-side obstacle assembly, planned by the CPU oracle
(scenario.plan_timestep_by_levels; PrioritizedController.m:297-324, :449-506).  Bit-exact on every
output field, including the pop-order hash."""
import dataclasses

import numpy as np
import pytest

from oracle import oracle_py, parity
from pdmpc_b200 import capi, scenario
from pdmpc_b200.mpa import get_mpa
from pdmpc_b200.records import CHECKER_INTERX, CHECKER_SAT, SearchBatch, TimestepDeps

from helpers import rect, straight_iter

pytestmark = pytest.mark.gpu


def closed_loop(planner, mpa, sc, steps):
    """Closed loop driven by the device's one-call time step; every step checked against the oracle."""
    planner.upload_mpa(mpa)
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    runner = scenario.ScenarioRunner(sc, None, timestep_fn=lambda b, d: planner.plan_timestep(b, d, False))
    totals = {"searches": 0, "preds": 0, "exhausted": 0}
    for _ in range(steps):
        dev = runner.step_timestep()
        _k, batch, deps, _ = runner.timestep_records[-1]
        ref = scenario.plan_timestep_by_levels(plan, batch, deps)
        parity.compare(dev, ref)
        totals["searches"] += batch.n
        totals["preds"] += int(deps.pred_idx.size)
        totals["exhausted"] += int(ref.is_exhausted.sum())
    return runner, totals


def test_circle_config0_sat_closed_loop(planner):
    mpa = get_mpa("single_speed", non_convex=False)
    _, tot = closed_loop(planner, mpa, scenario.circle_scenario(mpa, 4), 30)
    assert tot["preds"] > 0


@pytest.mark.parametrize("mpa_type,steps", [("single_speed", 12), ("triple_speed", 10)])
def test_road_config1_interx_closed_loop(planner, mpa_type, steps):
    mpa = get_mpa(mpa_type, non_convex=True)
    runner, tot = closed_loop(planner, mpa, scenario.commonroad_scenario(mpa, 20, seed=2), steps)
    assert tot["preds"] > 20 and tot["searches"] == 20 * steps


def collect_timesteps(mpa, sc, steps):
    """(batch, deps) of every time step of an oracle-driven closed loop."""
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    runner = scenario.ScenarioRunner(sc, None, timestep_fn=lambda b, d: scenario.plan_timestep_by_levels(plan, b, d))
    runner.run(steps)
    return [(b, d, r) for _k, b, d, r in runner.timestep_records]


def concat_timesteps(items):
    """Many time steps / scenarios as ONE call: predecessor indices shifted per block."""
    batch = SearchBatch.concat([b for b, _d, _r in items])
    off, ptr, idx = 0, [np.zeros(1, dtype=np.int32)], []
    for b, d, _r in items:
        ptr.append(d.pred_ptr[1:] + ptr[-1][-1])
        idx.append(d.pred_idx + off)
        off += b.n
    deps = TimestepDeps(np.concatenate(ptr).astype(np.int32), np.concatenate(idx).astype(np.int32),
                        np.concatenate([d.fb_npts for _b, d, _r in items]),
                        np.concatenate([d.fb_x for _b, d, _r in items]),
                        np.concatenate([d.fb_y for _b, d, _r in items]))
    ref = dataclasses.replace(items[0][2], **{
        f.name: np.concatenate([getattr(r, f.name) for _b, _d, r in items])
        for f in dataclasses.fields(items[0][2]) if isinstance(getattr(items[0][2], f.name), np.ndarray)})
    return batch, deps, ref


def test_many_timesteps_in_one_call_more_searches_than_sms(planner):
    """12 time steps x 20 vehicles = 240 searches > 148 CTAs: work items are handed out in a
    topological order, so a CTA never waits for a search that has not been taken yet."""
    mpa = get_mpa("triple_speed", non_convex=True)
    items = collect_timesteps(mpa, scenario.commonroad_scenario(mpa, 20, seed=3), 12)
    batch, deps, ref = concat_timesteps(items)
    assert batch.n == 240
    planner.upload_mpa(mpa)
    dev = planner.plan_timestep(batch, deps, False)
    parity.compare(dev, ref)
    # the same searches in reverse order (predecessors now have HIGHER indices than their successors)
    perm = np.arange(batch.n)[::-1].copy()
    inv = np.empty_like(perm)
    inv[perm] = np.arange(batch.n)
    rb = batch.select(perm)
    preds = [inv[deps.preds(int(i))] for i in perm]
    rdeps = TimestepDeps.build(preds, [deps.fallback_shapes(int(i)) for i in perm], batch.Hp)
    rdev = planner.plan_timestep(rb, rdeps, False)
    rref = dataclasses.replace(ref, **{f.name: getattr(ref, f.name)[perm] for f in dataclasses.fields(ref)
                                       if isinstance(getattr(ref, f.name), np.ndarray)})
    parity.compare(rdev, rref)


def test_equals_level_by_level_device_calls(planner):
    """One call == the drop-in's previous call pattern (one pdmpc_plan_batch per computation level)."""
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    for batch, deps, _ref in collect_timesteps(mpa, scenario.commonroad_scenario(mpa, 20, seed=5), 6):
        by_levels = scenario.plan_timestep_by_levels(lambda b: planner.plan_batch(b, False), batch, deps)
        parity.compare(planner.plan_timestep(batch, deps, False), by_levels)


@pytest.mark.parametrize("checker", [CHECKER_SAT, CHECKER_INTERX])
def test_chain_and_fallback_areas_of_exhausted_predecessor(planner, checker):
    """A chain 0 <- 1 <- 2 ... of vehicles on parallel lines; vehicle 0 starts inside an obstacle
    (exhausted) and publishes fallback areas that lie across vehicle 1's line."""
    mpa = get_mpa("single_speed", non_convex=checker == CHECKER_INTERX)
    Hp = mpa.Hp
    n = 24
    iters = [straight_iter(mpa, x=0.0, y=1.0 * i) for i in range(n)]
    iters[0].obstacles.append(rect(0.05, 0.0, 0.02, 0.3))      # crosses every first maneuver of vehicle 0
    batch = SearchBatch.from_iters(iters, Hp, checker, mpa.dt_seconds)
    preds = [[]] + [[i - 1] for i in range(1, n)]
    preds[5] = [4, 2, 0]                                         # several predecessors, unordered
    fb = [None] * n
    fb[0] = [rect(0.5, 1.0, 0.05, 0.3)] * Hp                     # on vehicle 1's line
    deps = TimestepDeps.build(preds, fb, Hp)
    planner.upload_mpa(mpa)
    dev = planner.plan_timestep(batch, deps, False)
    ref = scenario.plan_timestep_by_levels(lambda b: oracle_py.plan_batch(mpa, b), batch, deps)
    parity.compare(dev, ref)
    assert ref.is_exhausted[0] == 1
    free = oracle_py.plan_batch(mpa, SearchBatch.from_iters([iters[1]], Hp, checker, mpa.dt_seconds))
    assert ref.n_expanded[1] != free.n_expanded[0] or not np.array_equal(ref.trims[1], free.trims[0])


def test_no_dependencies_equals_plan_batch_and_empty_batch(planner):
    mpa = get_mpa("single_speed", non_convex=True)
    planner.upload_mpa(mpa)
    batch, deps, _ = collect_timesteps(mpa, scenario.commonroad_scenario(mpa, 20, seed=1), 1)[0]
    none = TimestepDeps.build([[] for _ in range(batch.n)], [None] * batch.n, batch.Hp)
    parity.compare(planner.plan_timestep(batch, none, False), planner.plan_batch(batch, False))
    empty = batch.select(np.zeros(0, dtype=np.int64))
    r = planner.plan_timestep(empty, TimestepDeps.build([], [], batch.Hp))
    assert r.status.size == 0


def test_bad_relations_rejected(planner):
    mpa = get_mpa("single_speed", non_convex=True)
    planner.upload_mpa(mpa)
    Hp = mpa.Hp
    batch = SearchBatch.from_iters([straight_iter(mpa, y=float(i)) for i in range(3)], Hp, CHECKER_INTERX, mpa.dt_seconds)
    for preds in ([[1], [2], [0]], [[0], [], []], [[], [7], []], [[], [-1], []]):
        with pytest.raises(capi.PdmpcError) as e:
            planner.plan_timestep(batch, TimestepDeps.build(preds, [None] * 3, Hp))
        assert e.value.code == capi.PDMPC_ERR_BAD_INPUT
    too_many = [[]] + [[0] * (2048 // (Hp * 8) + 1)] + [[]]
    with pytest.raises(capi.PdmpcError):
        planner.plan_timestep(batch, TimestepDeps.build(too_many, [None] * 3, Hp))
    d = TimestepDeps.build([[], [0], [1]], [None] * 3, Hp)
    d.fb_npts[0, 0] = 1
    with pytest.raises(capi.PdmpcError):
        planner.plan_timestep(batch, d)
    # the planner is still usable afterwards
    ok = planner.plan_timestep(batch, TimestepDeps.build([[], [0], [1]], [None] * 3, Hp))
    assert int(ok.status.max()) == 0


def test_realistic_mpa_takes_the_warp_kernel(planner):
    """Search trees beyond 32768 nodes do not fit the one-CTA-per-search kernel: the one-call time step
    then runs one warp per search (same hand-over, per-slot scratch in HBM) — same answers."""
    mpa = get_mpa("realistic", non_convex=True)
    planner.upload_mpa(mpa)
    items = collect_timesteps(mpa, scenario.commonroad_scenario(mpa, 12, seed=2), 3)
    for batch, deps, ref in items:
        parity.compare(planner.plan_timestep(batch, deps, False), ref)


def test_valid_only_queue_of_the_time_step(planner):
    """pdmpc_set_cta_queue(1): the one-call time step (and small plan_batch calls) run the CTA shape with its
    valid-only queue.  Every output equals the specification's; pop_hash then covers the valid pops only,
    which the oracle reproduces in its hash_valid_pops_only mode."""
    from helpers import load_golden_timesteps
    try:
        planner.set_cta_queue(True)
        for name in ("timestep_road_triple_speed", "timestep_circle_single_speed"):
            mpa, steps = load_golden_timesteps(name)
            planner.upload_mpa(mpa)
            for batch, deps, exp in steps:
                dev = planner.plan_timestep(batch, deps, False)
                assert planner.stats().shape == 5
                parity.compare(dev, exp, skip=("pop_hash",))
                # the level-by-level specification with the valid-pop hash
                ref5 = scenario.plan_timestep_by_levels(
                    lambda x: oracle_py.plan_batch(mpa, x, hash_valid_pops_only=True), batch, deps)
                parity.compare(dev, ref5)
                small = planner.plan_batch(batch.select(np.arange(min(batch.n, 5))), False)
                assert planner.stats().shape == 5 and small.status.max() == 0
    finally:
        planner.set_cta_queue(False)


@pytest.mark.parametrize("variant", [1, 4])
def test_both_launch_shapes_of_the_time_step(planner, variant):
    """pdmpc_set_variant forces the warp-per-search (1) or the CTA-per-search (4) shape of pdmpc_plan_timestep:
    fixtures step by step and as one 160-search call, circle/SAT and road/InterX, plus the 40-vehicle
    reachable-set case (many predecessors' columns next to large own obstacle polylines)."""
    from helpers import load_golden_timesteps
    try:
        planner.set_variant(variant)
        for name in ("timestep_road_triple_speed", "timestep_circle_single_speed"):
            mpa, steps = load_golden_timesteps(name)
            planner.upload_mpa(mpa)
            for batch, deps, exp in steps:
                parity.compare(planner.plan_timestep(batch, deps, False), exp)
            batch, deps, exp = concat_timesteps(steps)
            parity.compare(planner.plan_timestep(batch, deps, False), exp)
        mpa = get_mpa("triple_speed", non_convex=True)
        planner.upload_mpa(mpa)
        plan = lambda b: oracle_py.plan_batch(mpa, b)
        runner = scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 40, seed=3, allow_shared_paths=True), None,
                                         max_num_CLs=2, timestep_fn=lambda b, d: planner.plan_timestep(b, d, False))
        for _ in range(4):
            dev = runner.step_timestep()
            _k, batch, deps, _ = runner.timestep_records[-1]
            parity.compare(dev, scenario.plan_timestep_by_levels(plan, batch, deps))
    finally:
        planner.set_variant(0)


def test_thousands_of_searches_in_one_call(planner):
    """64 scenarios x 20 vehicles x 2 time steps = 2560 searches with their predecessor DAGs in ONE call
    in forward and in shuffled index order, auto shape and both forced shapes."""
    mpa = get_mpa("single_speed", non_convex=True)
    planner.upload_mpa(mpa)
    items = []
    for seed in range(1, 65):
        items.extend(collect_timesteps(mpa, scenario.commonroad_scenario(mpa, 20, seed=seed), 2))
    batch, deps, ref = concat_timesteps(items)
    assert batch.n == 2560
    dev = planner.plan_timestep(batch, deps, False)
    parity.compare(dev, ref)
    try:
        for variant in (1, 4):
            planner.set_variant(variant)
            parity.compare(planner.plan_timestep(batch, deps, False), ref)
    finally:
        planner.set_variant(0)
    rng = np.random.default_rng(5)
    perm = rng.permutation(batch.n)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(batch.n)
    rdeps = TimestepDeps.build([inv[deps.preds(int(i))] for i in perm], [deps.fallback_shapes(int(i)) for i in perm], batch.Hp)
    rdev = planner.plan_timestep(batch.select(perm), rdeps, False)
    rref = dataclasses.replace(ref, **{f.name: getattr(ref, f.name)[perm] for f in dataclasses.fields(ref)
                                       if isinstance(getattr(ref, f.name), np.ndarray)})
    parity.compare(rdev, rref)


@pytest.mark.parametrize("name", ["timestep_road_triple_speed", "timestep_circle_single_speed"])
def test_golden_timestep_fixture(planner, name):
    """One-call time steps against the committed fixtures (tools/make_golden_timestep.py, written after
    both CPU restatements agreed on every vehicle of every step): no oracle at run time."""
    from helpers import load_golden_timesteps
    mpa, steps = load_golden_timesteps(name)
    planner.upload_mpa(mpa)
    for batch, deps, exp in steps:
        parity.compare(planner.plan_timestep(batch, deps, False), exp)
    # all steps of the fixture as ONE call
    batch, deps, exp = concat_timesteps(steps)
    parity.compare(planner.plan_timestep(batch, deps, False), exp)


def test_explorative_priorities_config2_all_permutations_in_one_call(planner):
    """BASELINE configs[2]: every priority permutation of a time step (n_CL x 20 searches, one DAG per
    permutation) as ONE pdmpc_plan_timestep call; closed loop identical to the oracle-driven one."""
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    dev = scenario.ExplorativeRunner(scenario.commonroad_scenario(mpa, 20, seed=2),
                                     lambda b, d: planner.plan_timestep(b, d, False))
    ref = scenario.ExplorativeRunner(scenario.commonroad_scenario(mpa, 20, seed=2),
                                     lambda b, d: scenario.plan_timestep_by_levels(plan, b, d))
    dev.run(6)
    ref.run(6)
    assert np.array_equal(dev.pose, ref.pose) and np.array_equal(dev.trim, ref.trim)
    for a, b in zip(dev.explorative_records, ref.explorative_records):
        assert np.array_equal(a["chosen"], b["chosen"]) and np.array_equal(a["solution_cost"], b["solution_cost"])
        parity.compare(a["result_local"], b["result_local"])
    assert max(e["n_permutations"] for e in ref.explorative_records) >= 3


@pytest.mark.parametrize("max_cls", [1, 2, 4])
def test_config3_40_vehicles_reachable_set_obstacles(planner, max_cls):
    """BASELINE configs[3] (computation-level-limited planning): 40 vehicles, at most max_cls computation
    levels, the other predecessors enter as reachable-set obstacles (polygons of up to 25 vertices, hundreds
    of obstacle columns per search).  One-call time steps in closed loop against the oracle, then every
    search of the run (with the predecessors' areas assembled on the host) through all launch shapes."""
    from test_parity_gpu import check
    mpa = get_mpa("triple_speed", non_convex=True)
    sc = scenario.commonroad_scenario(mpa, 40, seed=1, allow_shared_paths=True)
    planner.upload_mpa(mpa)
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    runner = scenario.ScenarioRunner(sc, None, max_num_CLs=max_cls,
                                     timestep_fn=lambda b, d: planner.plan_timestep(b, d, False))
    level_batches = []

    def capture(b):
        level_batches.append(b)
        return plan(b)

    for _ in range(5):
        dev = runner.step_timestep()
        _k, batch, deps, _ = runner.timestep_records[-1]
        parity.compare(dev, scenario.plan_timestep_by_levels(capture, batch, deps))
    flat = SearchBatch.concat(level_batches)
    assert flat.n == 200 and np.diff(flat.poly_ptr).max() >= 15
    info, _dev, _ref = check(planner, mpa, flat)
    assert info["exhausted"] > 0


def test_lockstep_scenarios_one_call_per_time_step(planner):
    """BASELINE configs[4] in closed loop: several scenarios advanced together, every time step of all of
    them as ONE pdmpc_plan_timestep call; each ends exactly where it ends when advanced alone."""
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    ts = lambda b, d: planner.plan_timestep(b, d, False)
    seeds = (1, 2, 3, 4)
    together = [scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 20, seed=s), None, timestep_fn=ts) for s in seeds]
    for _ in range(4):
        res = scenario.lockstep_step(together, ts)
        assert res.status.size == 80 and int(res.status.max()) == 0
    for s, r in zip(seeds, together):
        alone = scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 20, seed=s), None, timestep_fn=ts)
        alone.run(4)
        assert np.array_equal(alone.pose, r.pose) and np.array_equal(alone.trim, r.trim)


def test_pack_plan_rows_on_device(planner):
    """pdmpc_pack_plan_rows: costs + plans of the last call as rows in DEVICE memory (what the ranks all_gather
    for BASELINE configs[2]); exhausted searches take the caller's fallback rows."""
    import torch
    from helpers import load_golden_timesteps
    mpa, steps = load_golden_timesteps("timestep_road_triple_speed")
    planner.upload_mpa(mpa)
    Hp = mpa.Hp
    L = 2 + 21 * Hp
    batch, deps, exp = concat_timesteps(steps[:3])
    n_veh = steps[0][0].n
    res = planner.plan_timestep(batch, deps, False)
    parity.compare(res, exp)
    assert res.is_exhausted.any() and not res.is_exhausted.all()
    rng = np.random.default_rng(0)
    fb = rng.normal(size=(n_veh, L))
    dst = torch.full((batch.n, L), -7.0, dtype=torch.float64, device="cuda:0")
    planner.pack_plan_rows(batch.n, n_veh, fb, dst.data_ptr())
    rows = dst.cpu().numpy()
    for r in range(batch.n):
        if res.is_exhausted[r]:
            want = fb[r % n_veh].copy()
            want[1] = 1.0
        else:
            want = np.concatenate([[res.g_path[r, Hp], 0.0], res.trims[r, 1:], res.y_predicted[r].reshape(-1),
                                   res.shape_npts[r], res.shape_x[r].reshape(-1), res.shape_y[r].reshape(-1)])
        assert np.array_equal(rows[r].view(np.uint64), want.view(np.uint64)), r
    # a prefix of the rows, no fallback rows: exhausted searches give zeros (flag 1)
    dst.fill_(-7.0)
    planner.pack_plan_rows(5, n_veh, None, dst.data_ptr())
    rows = dst.cpu().numpy()
    assert (rows[5:] == -7.0).all()
    with pytest.raises(capi.PdmpcError):
        planner.pack_plan_rows(batch.n + 1, n_veh, None, dst.data_ptr())
