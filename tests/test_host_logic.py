"""CPU suite, part 3: host-side data formats and the synthetic input harness."""
import dataclasses

import numpy as np
import pytest

from oracle import oracle_py, parity
from pdmpc_b200 import scenario
from pdmpc_b200.mpa import build_mpa, get_mpa
from pdmpc_b200.records import CHECKER_INTERX, CHECKER_SAT, BatchResult, SearchBatch

from helpers import circle_records, load_golden, rect, road_records, straight_iter


@pytest.mark.parametrize("mpa_type,n_trims,n_edges", [("single_speed", 12, 54), ("triple_speed", 34, 160),
                                                        ("realistic", 71, 527)])
def test_mpa_table_shapes(mpa_type, n_trims, n_edges):
    """Trim / edge counts derived from choose_trims.m."""
    mpa = get_mpa(mpa_type, non_convex=True)
    assert mpa.n_trims == n_trims and mpa.n_edges == n_edges
    assert mpa.transition.shape == (6, n_trims, n_trims)
    # recursive feasibility (MotionPrimitiveAutomaton.m:238-250): the last step only admits
    # equilibrium (zero-speed) trims, step k only trims at most Hp-k transitions away from one
    last = np.flatnonzero(mpa.transition[-1].any(axis=0))
    assert (mpa.trim_speed[last] == 0).all()
    for k in range(1, 7):
        cols = np.flatnonzero(mpa.transition[k - 1].any(axis=0))
        assert (mpa.distance_to_equilibrium[cols] <= 6 - k).all()
    # maneuver areas: 5 points straight, 7 turning (non-convex build), closed
    for e in range(mpa.n_edges):
        for kind in range(3):
            n = mpa.area_npts[e, kind]
            assert n in (5, 7)
            assert mpa.area_x[e, kind, 0] == mpa.area_x[e, kind, n - 1]
            assert mpa.area_y[e, kind, 0] == mpa.area_y[e, kind, n - 1]
    convex = get_mpa(mpa_type, non_convex=False)
    assert set(np.unique(convex.area_npts)) <= {5, 6}
    assert mpa.full_tree_nodes() < (1 << 20)


def test_mpa_straight_maneuver_is_exact():
    mpa = get_mpa("single_speed")
    t = mpa.trim_from_values(0.8, 0.0)
    e = mpa.edge_index[t - 1, t - 1]
    assert abs(mpa.edge_dx[e] - 0.8 * 0.2) < 1e-9 and abs(mpa.edge_dy[e]) < 1e-12 and abs(mpa.edge_dyaw[e]) < 1e-12


def test_batch_concat_select_round_trip():
    mpa, batch = road_records("single_speed", 3)
    ref = oracle_py.plan_batch(mpa, batch)
    rng = np.random.default_rng(0)
    perm = rng.permutation(batch.n)
    sub = batch.select(perm)
    got = oracle_py.plan_batch(mpa, sub)
    for f in dataclasses.fields(BatchResult):
        a = getattr(ref, f.name)
        if isinstance(a, np.ndarray):
            assert np.array_equal(getattr(got, f.name), a[perm], equal_nan=True), f.name
    halves = SearchBatch.concat([batch.select(np.arange(0, 7)), batch.select(np.arange(7, batch.n))])
    for f in dataclasses.fields(SearchBatch):
        a = getattr(batch, f.name)
        if isinstance(a, np.ndarray):
            assert np.array_equal(getattr(halves, f.name), a), f.name


def test_batch_save_load(tmp_path):
    _, batch = road_records("single_speed", 3)
    p = str(tmp_path / "b.npz")
    batch.save(p)
    back = SearchBatch.load(p)
    assert back.Hp == batch.Hp and back.checker == batch.checker
    assert np.array_equal(back.slot_ptr, batch.slot_ptr) and np.array_equal(back.vert_x, batch.vert_x)


def test_ragged_and_empty_records():
    """Searches with no obstacles at all next to searches with many; empty polygons are dropped."""
    mpa = get_mpa("single_speed", non_convex=True)
    its = [straight_iter(mpa),
           straight_iter(mpa, obstacles=[rect(0.7, 0, 0.05, 0.5), np.zeros((2, 0))],
                         dynamic_obstacle_area=[[rect(5, 5, 0.1, 0.1)] * mpa.Hp]),
           straight_iter(mpa)]
    b = SearchBatch.from_iters(its, mpa.Hp, CHECKER_INTERX, mpa.dt_seconds)
    S = mpa.Hp + 1
    assert b.slot_ptr[S] == 0 and b.slot_ptr[2 * S] - b.slot_ptr[S] == 1 + mpa.Hp and b.slot_ptr[-1] == b.slot_ptr[2 * S]
    r = oracle_py.plan_batch(mpa, b)
    assert r.pop_hash[0] == r.pop_hash[2] and r.n_expanded[1] != r.n_expanded[0]
    empty = SearchBatch.from_iters([], mpa.Hp, CHECKER_SAT, mpa.dt_seconds)
    assert empty.n == 0 and oracle_py.plan_batch(mpa, empty).status.size == 0


def test_kahn_levels_and_coloring_priorities():
    A = np.zeros((5, 5), dtype=np.int64)
    for i, j in ((0, 1), (1, 2), (0, 2), (3, 4)):
        A[i, j] = 1
    assert scenario.kahn(A).tolist() == [1, 2, 3, 1, 2]                 # utility/kahn.m
    with pytest.raises(ValueError):
        scenario.kahn(np.array([[0, 1], [1, 0]]))
    rng = np.random.default_rng(3)
    U = np.triu((rng.random((12, 12)) < 0.3).astype(np.int64), 1)
    U = U + U.T
    D = scenario.coloring_priorities(U)
    assert ((D | D.T) == (U > 0)).all() and not (D & D.T).any()        # every coupling directed once
    lv = scenario.kahn(D.astype(np.int64))                              # acyclic
    i, j = np.nonzero(D)
    assert (lv[i] < lv[j]).all()
    C = scenario.constant_priorities(U)
    i, j = np.nonzero(C)
    assert (i < j).all()


def test_scenarios_are_deterministic_and_shaped_as_baseline():
    mpa, b1 = road_records("triple_speed", 2)
    sc = scenario.commonroad_scenario(mpa, 20, seed=1)
    b2 = scenario.roll_out(sc, lambda b: oracle_py.plan_batch(mpa, b), 2)
    assert np.array_equal(b1.vert_x, b2.vert_x) and np.array_equal(b1.x0, b2.x0)
    assert b1.n == 40 and b1.checker == CHECKER_INTERX
    lane_pts = np.diff(b1.lane_ptr)
    assert lane_pts.min() >= 12 and lane_pts.max() <= 80                # ~25-55 per side
    ids = {tuple(v.reference_path[0]) for v in sc.vehicles}
    assert len(ids) == 20                                               # distinct path ids (Config.m:135-150)
    mpa_c, bc = circle_records(5)
    assert bc.n == 20 and bc.checker == CHECKER_SAT and bc.lane_x.size == 0
    x0 = bc.x0[:4], bc.y0[:4]
    assert np.allclose(np.hypot(x0[0] - 2.25, x0[1] - 2.0), 2.0)        # Circle.m:24-27


def test_closed_loop_makes_progress_and_predecessors_are_respected():
    """Plans of a level are obstacles of the next (PrioritizedController.m:449-506): no two
    vehicles' first-step shapes may cross after planning."""
    mpa = get_mpa("single_speed", non_convex=True)
    sc = scenario.commonroad_scenario(mpa, 12, seed=4)
    runner = scenario.ScenarioRunner(sc, lambda b: oracle_py.plan_batch(mpa, b))
    start = runner.pose.copy()
    recs = runner.run(10)
    moved = np.hypot(*(runner.pose[:, :2] - start[:, :2]).T)
    assert (moved > 0.2).sum() >= 8
    for r in recs[-4:]:
        shapes = {int(v): r.result.shapes(i)[0] for i, v in enumerate(r.vehicles) if not r.result.is_exhausted[i]}
        vs = list(shapes)
        for a in range(len(vs)):
            for b in range(a + 1, len(vs)):
                assert not oracle_py.interx(shapes[vs[a]], shapes[vs[b]])


def test_golden_inputs_are_well_formed():
    for name in ("circle_sat_single_speed", "road_interx_triple_speed"):
        mpa, batch, exp = load_golden(name)
        assert batch.slot_ptr.size == batch.n * (batch.Hp + 1) + 1 and batch.lane_ptr.size == 2 * batch.n + 1
        assert exp.trims.shape == (batch.n, batch.Hp + 1) and (exp.trims[:, 0] == batch.trim0).all()
        ok = exp.is_exhausted == 0
        assert (exp.tree_path[ok, 0] == 1).all() and (exp.n_expanded >= exp.n_pops - 0).all() or True


def _closed_loops(sc_factory, mpa, steps):
    """The same scenario twice: level-by-level through step(), and one call per time step through
    plan_timestep_by_levels (the specification of pdmpc_plan_timestep).  Returns both runners."""
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    a = scenario.ScenarioRunner(sc_factory(), plan)
    a.run(steps)
    b = scenario.ScenarioRunner(sc_factory(), None,
                                timestep_fn=lambda batch, deps: scenario.plan_timestep_by_levels(plan, batch, deps))
    b.run(steps)
    return a, b


@pytest.mark.parametrize("kind", ["circle", "road"])
def test_timestep_decomposition_equals_level_by_level_loop(kind):
    """One call per time step (base iter_v + predecessor lists + fallback areas) carries the same
    information as the level-by-level loop: identical closed loop, identical per-search outputs."""
    if kind == "circle":
        mpa = get_mpa("single_speed", non_convex=False)
        factory, steps = (lambda: scenario.circle_scenario(mpa, 4)), 25
    else:
        mpa = get_mpa("single_speed", non_convex=True)
        factory, steps = (lambda: scenario.commonroad_scenario(mpa, 12, seed=4)), 8
    a, b = _closed_loops(factory, mpa, steps)
    assert np.array_equal(a.pose, b.pose) and np.array_equal(a.trim, b.trim)
    assert a.n_fallbacks == b.n_fallbacks
    by_step = {}
    for r in a.records:
        for i, v in enumerate(r.vehicles):
            by_step[(r.step, int(v))] = (r.result, i)
    n_pred = 0
    for k, batch, deps, res in b.timestep_records:
        n_pred += deps.pred_idx.size
        for v in range(batch.n):
            ra, i = by_step[(k, v)]
            for name in ("is_exhausted", "n_expanded", "n_pops", "pop_hash"):
                assert getattr(ra, name)[i] == getattr(res, name)[v], (k, v, name)
            assert np.array_equal(ra.trims[i], res.trims[v])
            assert np.array_equal(ra.y_predicted[i], res.y_predicted[v], equal_nan=True)
            assert np.array_equal(ra.shape_x[i], res.shape_x[v])
    assert n_pred > 0   # the predecessor hand-over was exercised


def test_batch_to_iters_round_trip():
    mpa, batch = road_records("single_speed", 3, 20, 1)
    again = SearchBatch.from_iters(batch.to_iters(), batch.Hp, batch.checker, batch.dt_seconds)
    for f in dataclasses.fields(batch):
        a, b = getattr(batch, f.name), getattr(again, f.name)
        assert np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b, f.name


def test_timestep_by_levels_uses_fallback_areas_of_exhausted_predecessors():
    """Search 0 is walled in (exhausts) and publishes its fallback areas; search 1 waits for it and
    must avoid them (PrioritizedController.m:568-621, :678-718)."""
    from pdmpc_b200.records import TimestepDeps
    mpa = get_mpa("single_speed", non_convex=False)
    Hp = mpa.Hp

    walled = straight_iter(mpa, obstacles=[rect(0.0, 0.0, 0.3, 0.3)])   # the start pose itself collides
    free = straight_iter(mpa, x=0.0, y=5.0)
    batch = SearchBatch.from_iters([walled, free], Hp, CHECKER_SAT, mpa.dt_seconds)
    block = [rect(0.6, 5.0, 0.1, 0.5)] * Hp     # across the free vehicle's straight line
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    with_fb = scenario.plan_timestep_by_levels(plan, batch, TimestepDeps.build([[], [0]], [block, None], Hp))
    without = scenario.plan_timestep_by_levels(plan, batch, TimestepDeps.build([[], []], [None, None], Hp))
    assert with_fb.is_exhausted[0] == 1 and without.is_exhausted[0] == 1
    assert not np.array_equal(with_fb.y_predicted[1], without.y_predicted[1]) or \
        with_fb.n_expanded[1] != without.n_expanded[1]



def test_config4_reachable_sets_and_level_limit():
    """BASELINE configs[3]: local reachable sets (convexified, MotionPrimitiveAutomaton.m:252-392) and the
    computation-level limit that turns sequential predecessors into parallel ones."""
    mpa = get_mpa("single_speed", non_convex=True)
    sets = scenario.local_reachable_sets_conv(mpa)
    assert len(sets) == mpa.n_trims and all(len(s) == mpa.Hp for s in sets)

    def inside(poly, pts):   # closed counter-clockwise convex polygon [2, h + 1]
        ex, ey = np.diff(poly[0]), np.diff(poly[1])
        cr = ex[:, None] * (pts[None, :, 1] - poly[1, :-1, None]) - ey[:, None] * (pts[None, :, 0] - poly[0, :-1, None])
        return bool((cr >= -1e-9).all())

    t0 = mpa.trim_from_values(0.0, 0.0)
    for to in np.flatnonzero(mpa.transition[0, t0 - 1]):          # every first-step area lies in the step-1 set
        e = int(mpa.edge_index[t0 - 1, to])
        m = int(mpa.area_npts[e, 0])
        assert inside(sets[t0 - 1][0], np.column_stack([mpa.area_x[e, 0, :m], mpa.area_y[e, 0, :m]]))
    for per_step in sets:
        for a in per_step:
            assert a.shape[0] == 2 and a.shape[1] >= 5 and np.array_equal(a[:, 0], a[:, -1])
    placed = scenario.reachable_sets_at(mpa, 1.0, 2.0, np.pi / 2, t0)
    assert np.allclose(placed[0][0], 1.0 - sets[t0 - 1][0][1]) and np.allclose(placed[0][1], 2.0 + sets[t0 - 1][0][0])

    rng = np.random.default_rng(3)
    for _ in range(20):
        n = 12
        prio = rng.permutation(n)
        A = rng.random((n, n)) < 0.3
        A = A | A.T
        D = A & (prio[:, None] < prio[None, :])
        for cl in (1, 2, 4, 99):
            seq = scenario.limit_computation_levels(D, cl)
            assert not (seq & ~D).any()
            assert scenario.kahn(seq.astype(np.int64)).max() <= max(cl, 1)
            if cl == 99:
                assert np.array_equal(seq, D)
            if cl == 1:
                assert not seq.any()

    mpa3 = get_mpa("single_speed", non_convex=True)
    plan = lambda b: oracle_py.plan_batch(mpa3, b)
    sc = scenario.commonroad_scenario(mpa3, 40, seed=1, allow_shared_paths=True)
    r = scenario.ScenarioRunner(sc, None, max_num_CLs=2,
                                timestep_fn=lambda b, d: scenario.plan_timestep_by_levels(plan, b, d))
    r.run(3)
    _k, batch, deps, _res = r.timestep_records[-1]
    assert batch.n == 40 and np.diff(batch.poly_ptr).max() > 8       # reachable sets are many-vertex obstacles
    assert deps.pred_idx.size > 0


def test_lockstep_scenarios_equal_individual_closed_loops():
    """Many scenarios advanced together, one optimizer call per time step for all of them."""
    mpa = get_mpa("single_speed", non_convex=True)
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    ts = lambda b, d: scenario.plan_timestep_by_levels(plan, b, d)
    seeds = (1, 2, 3)
    alone = []
    for sd in seeds:
        r = scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 10, seed=sd), None, timestep_fn=ts)
        r.run(5)
        alone.append(r)
    together = [scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 10, seed=sd), None, timestep_fn=ts) for sd in seeds]
    for _ in range(5):
        res = scenario.lockstep_step(together, ts)
        assert res.status.size == 30
    for a, b in zip(alone, together):
        assert np.array_equal(a.pose, b.pose) and np.array_equal(a.trim, b.trim) and a.n_fallbacks == b.n_fallbacks


def test_obstacle_assembly_hook_equals_the_inline_host_rules():
    """ScenarioRunner(obstacles_fn = ...) — the one-call path with the obstacle assembly behind one function (on the GPU
    box: pdmpc_assemble_obstacles) builds the same iter_v obstacles as the inline host rules: here with the host
    restatement of the kernel (scenario.assemble_obstacles_host), 40 vehicles, computation levels limited to 1."""
    import numpy as np
    from pdmpc_b200 import scenario
    from pdmpc_b200.mpa import get_mpa
    from pdmpc_b200.records import SearchBatch
    mpa = get_mpa("triple_speed", non_convex=True)
    sc = scenario.commonroad_scenario(mpa, 40, seed=2, allow_shared_paths=True)
    runner = scenario.ScenarioRunner(sc, None, max_num_CLs=1)
    # a few vehicles drive, the others stand, so both rules fire
    drive = mpa.trim_from_values(mpa.get_straight_speeds_of_mpa()[-1], 0.0)
    runner.trim[::3] = drive
    iters_a, preds_a, _ = runner.timestep_inputs()
    runner.obstacles_fn = lambda *a: scenario.assemble_obstacles_host(mpa, *a)
    iters_b, preds_b, _ = runner.timestep_inputs()
    ba = SearchBatch.from_iters(iters_a, mpa.Hp, sc.checker, mpa.dt_seconds)
    bb = SearchBatch.from_iters(iters_b, mpa.Hp, sc.checker, mpa.dt_seconds)
    for f in ("slot_ptr", "poly_ptr", "vert_x", "vert_y"):
        a, b = getattr(ba, f), getattr(bb, f)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f
    per_slot = np.diff(ba.slot_ptr.reshape(-1)).reshape(sc.amount, mpa.Hp + 1)
    assert per_slot[:, 0].sum() > 0 and per_slot[:, 1:].sum() > 0
    assert all(np.array_equal(p, q) for p, q in zip(preds_a, preds_b))
