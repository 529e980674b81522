"""The input side of a time step on the device (pdmpc_upload_road / pdmpc_sample_inputs):
reference-trajectory sampling, predicted lanelets and lanelet boundaries of every vehicle in one call, bit for bit
against the host restatement of the reference's functions in pdmpc_b200/scenario.py
(sample_reference_trajectory.m, get_arc_distance_to_endpoint.m, projection_2d.m, get_predicted_lanelets.m,
get_lanelets_boundary.m), on start poses, on poses off the path and along closed loops."""
import numpy as np
import pytest

from oracle import oracle_py, parity
from pdmpc_b200 import capi, scenario
from pdmpc_b200.mpa import get_mpa


def host_inputs(mpa, veh, road, x, y, trim):
    pts, v_ref, idx, cur = scenario.get_reference_trajectory(mpa, veh, x, y, trim, mpa.dt_seconds)
    lan = scenario.get_predicted_lanelets(veh, idx)
    left, right = scenario.get_lanelets_boundary(lan, road, veh)
    return pts, v_ref, idx, cur, lan, left, right


def check_rows(mpa, vehicles, road, poses, trims, o):
    for i, veh in enumerate(vehicles):
        pts, v_ref, idx, cur, lan, left, right = host_inputs(mpa, veh, road, poses[i, 0], poses[i, 1], int(trims[i]))
        assert np.array_equal(o["ref_x"][i].view(np.uint64), np.ascontiguousarray(pts[:, 0]).view(np.uint64)), i
        assert np.array_equal(o["ref_y"][i].view(np.uint64), np.ascontiguousarray(pts[:, 1]).view(np.uint64)), i
        assert np.array_equal(o["v_ref"][i], v_ref) and o["ref_index"][i].tolist() == idx.tolist()
        assert int(o["current_index"][i]) == cur
        got = o["predicted_lanelets"][i]
        assert got[got > 0].tolist() == lan.tolist()
        l0, l1, l2 = (int(v) for v in o["lane_ptr"][2 * i:2 * i + 3])
        assert np.array_equal(np.vstack([o["lane_x"][l0:l1], o["lane_y"][l0:l1]]), left)
        assert np.array_equal(np.vstack([o["lane_x"][l1:l2], o["lane_y"][l1:l2]]), right)


def test_road_tables_shape():
    mpa = get_mpa("single_speed", non_convex=True)
    scs = [scenario.commonroad_scenario(mpa, 20, seed=s) for s in (1, 2)]
    t = scenario.road_tables(scs)
    assert t["path_ptr"].size == 41 and t["bound_ptr"].size == 2 * scs[0].road.n + 1
    assert t["lan_ptr"][-1] == t["lanelets_index"].size == t["points_index"].size
    v = scs[1].vehicles[3]
    p = 20 + 3
    assert np.array_equal(t["path_x"][t["path_ptr"][p]:t["path_ptr"][p + 1]], v.reference_path[:, 0])
    assert t["reference_speed"][p] == v.reference_speed


@pytest.mark.gpu
def test_sample_inputs_matches_the_host_functions(planner):
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    scs = [scenario.commonroad_scenario(mpa, 20, seed=s) for s in (1, 2, 3)]
    planner.upload_road(scenario.road_tables(scs))
    vehicles = [v for sc in scs for v in sc.vehicles]
    n = len(vehicles)
    rng = np.random.default_rng(0)
    poses = np.array([[v.x_start, v.y_start] for v in vehicles])
    trims = np.full(n, mpa.trim_from_values(0.0, 0.0))
    speed = lambda tr: np.array([mpa.trim_speed[t - 1] for t in tr])
    o = planner.sample_inputs(np.arange(n), poses[:, 0], poses[:, 1], speed(trims), mpa.dt_seconds)
    check_rows(mpa, vehicles, scs[0].road, poses, trims, o)
    # anywhere along the paths, off the centre line, any trim, rows in any order / repeated
    for _ in range(4):
        rows = rng.integers(0, n, size=97)
        poses = np.zeros((rows.size, 2))
        for r, p in enumerate(rows):
            path = vehicles[p].reference_path
            j = rng.integers(0, path.shape[0])
            poses[r] = path[j] + rng.normal(scale=0.03, size=2)
        trims = rng.integers(1, mpa.trim_speed.size + 1, size=rows.size)
        o = planner.sample_inputs(rows, poses[:, 0], poses[:, 1], speed(trims), mpa.dt_seconds)
        check_rows(mpa, [vehicles[p] for p in rows], scs[0].road, poses, trims, o)
    # exactly ON path points (ties of the closest-point search, lambda == 0 / 1)
    rows = np.arange(n)
    poses = np.array([vehicles[p].reference_path[rng.integers(1, vehicles[p].reference_path.shape[0] - 1)] for p in rows])
    o = planner.sample_inputs(rows, poses[:, 0], poses[:, 1], speed(trims[:n]), mpa.dt_seconds)
    check_rows(mpa, vehicles, scs[0].road, poses, trims[:n], o)
    # errors are loud
    with pytest.raises(capi.PdmpcError) as e:
        planner.sample_inputs(np.array([n]), poses[:1, 0], poses[:1, 1], speed(trims[:1]), mpa.dt_seconds)
    assert e.value.code == capi.PDMPC_ERR_BAD_INPUT
    with pytest.raises(capi.PdmpcError) as e:
        planner.sample_inputs(rows, poses[:, 0], poses[:, 1], speed(trims[:n]), mpa.dt_seconds, lane_capacity=100)
    assert e.value.code == capi.PDMPC_ERR_CAPACITY


@pytest.mark.gpu
def test_closed_loop_with_device_side_inputs(planner):
    """ScenarioRunner(inputs_fn = Planner.sample_inputs): the whole time step (inputs + one-call planning) on the device
    drives the same closed loop as the host input functions + oracle planner."""
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    scs = [scenario.commonroad_scenario(mpa, 20, seed=s) for s in (5, 6)]
    planner.upload_road(scenario.road_tables(scs))
    for k, sc in enumerate(scs):
        dev = scenario.ScenarioRunner(sc, None, timestep_fn=lambda b, d: planner.plan_timestep(b, d, False),
                                      inputs_fn=planner.sample_inputs, path_id0=20 * k)
        ref = scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 20, seed=5 + k),
                                      lambda b: oracle_py.plan_batch(mpa, b, 4))
        for _ in range(8):
            dev.step()
            ref.step()
            assert np.array_equal(dev.pose.view(np.uint64), ref.pose.view(np.uint64)) and np.array_equal(dev.trim, ref.trim)
        _k, tb, td, got = dev.timestep_records[-1]
        parity.compare(got, scenario.plan_timestep_by_levels(lambda x: oracle_py.plan_batch(mpa, x), tb, td))


def test_host_sincos_follows_the_arithmetic_specification():
    """scenario.sincos_spec (areas placed on the host: standstill rectangles) == the oracle's / the device's sin/cos."""
    import math
    rng = np.random.default_rng(0)
    for x in list(rng.uniform(-60, 60, 3000)) + [0.0, math.pi / 2, math.pi, -math.pi, 1e-300, 3 * math.pi / 2, 2.5, -2.5]:
        assert scenario.sincos_spec(x) == oracle_py.sincos(x), x


@pytest.mark.gpu
def test_closed_loop_with_fallback_plans_on_the_device(planner):
    """pdmpc_plan_timestep_closed_loop: the fallback plan of an exhausted vehicle (standstill, or the previous plan
    shifted by one step: PrioritizedController.m:568-621, :678-718) is built on the device from the plans of the previous
    time step kept there; together with pdmpc_sample_inputs a time step needs the host only for coupling and priorities.
    Same closed loop as the host fallback logic driven by the oracle, with exhausted searches along the way."""
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    seeds = (1, 7)
    scs = [scenario.commonroad_scenario(mpa, 20, seed=s) for s in seeds]
    planner.upload_road(scenario.road_tables(scs))
    planner.closed_loop_reset(40, scenario.VEH_LENGTH / 2 + 0.01, scenario.VEH_WIDTH / 2 + 0.01)
    fallbacks = 0
    for k, sc in enumerate(scs):
        dev = scenario.ScenarioRunner(sc, None, inputs_fn=planner.sample_inputs, path_id0=20 * k,
                                      closed_loop_fn=lambda b, d, s, st: planner.plan_timestep_closed_loop(b, d, s, st, False))
        ref = scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 20, seed=seeds[k]),
                                      lambda b: oracle_py.plan_batch(mpa, b, 4))
        for step in range(14):
            dev.step()
            ref.step()
            assert np.array_equal(dev.pose.view(np.uint64), ref.pose.view(np.uint64)), (k, step)
            assert np.array_equal(dev.trim, ref.trim)
            for i in range(20):   # what every vehicle published: planned or fallback areas
                for a, b in zip(dev.prev_shapes[i], ref.prev_shapes[i]):
                    assert np.array_equal(a, b)
        assert dev.n_fallbacks == ref.n_fallbacks
        fallbacks += dev.n_fallbacks
    assert fallbacks > 5
    # a fresh scenario in used slots needs a reset: stale plans would be shifted into it
    planner.closed_loop_reset(40, scenario.VEH_LENGTH / 2 + 0.01, scenario.VEH_WIDTH / 2 + 0.01)
    with pytest.raises(capi.PdmpcError):
        b = dev.timestep_records[-1][1]
        planner.plan_timestep_closed_loop(b, dev.timestep_records[-1][2], np.zeros(b.n, dtype=np.int32), np.zeros(b.n, dtype=np.uint8))


def _obstacle_csr(batch):
    return batch.slot_ptr, batch.poly_ptr, batch.vert_x, batch.vert_y


@pytest.mark.gpu
@pytest.mark.parametrize("max_cls", [1, 2])
def test_assemble_obstacles_equals_the_host_assembly(planner, max_cls):
    """pdmpc_assemble_obstacles: standstill rectangles of standing successors
    (PrioritizedController.m:508-540, get_occupied_areas.m:19-25) and reachable sets of parallel predecessors
    (:391-407, MotionPrimitiveAutomaton.m:649-687) of all 40 vehicles in one call — the obstacle CSR equals the host
    assembly bit for bit at every time step of a closed loop that runs on the device-assembled obstacles."""
    from pdmpc_b200.records import SearchBatch
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    planner.upload_reachable_sets(scenario.local_reachable_sets_conv(mpa))
    sc = scenario.commonroad_scenario(mpa, 40, seed=2, allow_shared_paths=True)
    runner = scenario.ScenarioRunner(sc, None, max_num_CLs=max_cls,
                                     timestep_fn=lambda b, d: planner.plan_timestep(b, d, False),
                                     obstacles_fn=planner.assemble_obstacles)
    seen_static = seen_dynamic = 0
    for _ in range(6):
        fn, runner.obstacles_fn = runner.obstacles_fn, None
        iters_h, preds_h, _ = runner.timestep_inputs()
        runner.obstacles_fn = fn
        iters_d, preds_d, _ = runner.timestep_inputs()
        bh = SearchBatch.from_iters(iters_h, mpa.Hp, sc.checker, mpa.dt_seconds)
        bd = SearchBatch.from_iters(iters_d, mpa.Hp, sc.checker, mpa.dt_seconds)
        for a, b_ in zip(_obstacle_csr(bh), _obstacle_csr(bd)):
            assert a.dtype == b_.dtype and np.array_equal(a.view(np.uint8), b_.view(np.uint8))
        assert all(np.array_equal(p, q) for p, q in zip(preds_h, preds_d))
        sp = bh.slot_ptr.reshape(-1)
        per_slot = (sp[1:] - sp[:-1]).reshape(sc.amount, mpa.Hp + 1)
        seen_static += int(per_slot[:, 0].sum())
        seen_dynamic += int(per_slot[:, 1:].sum())
        dev = runner.step_timestep()                      # the loop advances on the device-assembled obstacles
        _k, batch, deps, _ = runner.timestep_records[-1]
        parity.compare(dev, scenario.plan_timestep_by_levels(lambda x: oracle_py.plan_batch(mpa, x), batch, deps))
    assert seen_static > 0 and seen_dynamic > 0


@pytest.mark.gpu
def test_assemble_obstacles_errors_and_empty_rows(planner):
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    sets = scenario.local_reachable_sets_conv(mpa)
    planner.upload_reachable_sets(sets)
    n, Hp = 3, mpa.Hp
    x, y, yaw = np.array([0.5, 1.0, 2.0]), np.array([0.25, 1.5, 3.0]), np.array([0.0, 1.25, -2.5])
    speed, trim = np.array([0.0, 0.5, 0.005]), np.array([1, 5, 2], dtype=np.int32)
    # row 0 sees rows 1 (driving: no obstacle) and 2 (standing) as successors; row 2 has row 1 as parallel predecessor
    o = planner.assemble_obstacles(x, y, yaw, speed, trim, [[1, 2], [], []], [[], [], [1]], 0.12, 0.0625)
    sp = o["slot_ptr"]
    assert sp[1] - sp[0] == 1 and all(sp[k + 1] - sp[k] == 0 for k in range(1, 2 * (Hp + 1)))
    assert all(sp[2 * (Hp + 1) + k + 1] - sp[2 * (Hp + 1) + k] == 1 for k in range(1, Hp + 1))
    rect = scenario.occupied_area(x[2], y[2], yaw[2], offset=0.0)   # half sizes passed explicitly below
    s_, c = scenario.sincos_spec(yaw[2])
    xl = np.array([-1.0, -1.0, 1.0, 1.0, -1.0]) * 0.12
    yl = np.array([-1.0, 1.0, 1.0, -1.0, -1.0]) * 0.0625
    assert rect.shape == (2, 5)
    assert np.array_equal(o["vert_x"][:5], c * xl - s_ * yl + x[2]) and np.array_equal(o["vert_y"][:5], s_ * xl + c * yl + y[2])
    want = scenario.reachable_sets_at(mpa, x[1], y[1], yaw[1], 5)
    pp = o["poly_ptr"]
    for k in range(Hp):
        p = sp[2 * (Hp + 1) + k + 1]
        assert np.array_equal(o["vert_x"][pp[p]:pp[p + 1]], want[k][0]) and np.array_equal(o["vert_y"][pp[p]:pp[p + 1]], want[k][1])
    with pytest.raises(capi.PdmpcError) as e:      # capacities are checked, nothing is truncated
        planner.assemble_obstacles(x, y, yaw, speed, trim, [[1, 2], [], []], [[], [], [1]], 0.12, 0.0625, poly_capacity=2, vert_capacity=4096)
    assert e.value.code == capi.PDMPC_ERR_CAPACITY
    with pytest.raises(capi.PdmpcError):
        planner.assemble_obstacles(x, y, yaw, speed, trim, [[7], [], []], [[], [], []], 0.12, 0.0625)
    with pytest.raises(capi.PdmpcError):            # open polygon
        bad = [[a.copy() for a in row] for row in sets]
        bad[0][0] = bad[0][0][:, :-1]
        planner.upload_reachable_sets(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("amount,max_cls,seed", [(20, 99, 7), (40, 2, 3)])
def test_whole_time_step_from_states_in_one_call(planner, amount, max_cls, seed):
    """pdmpc_plan_timestep_from_states: inputs, obstacle assembly, dependency-ordered planning and fallback plans of a
    time step chained on the device (the host only decides coupling and priorities) drive the same closed loop as the
    three separate device calls and as the host functions + oracle; every row of every step equals the separate calls."""
    mpa = get_mpa("triple_speed", non_convex=True)
    planner.upload_mpa(mpa)
    mk = lambda: scenario.commonroad_scenario(mpa, amount, seed=seed, allow_shared_paths=amount > 33)
    sc = mk()
    planner.upload_road(scenario.road_tables([sc]))
    planner.upload_reachable_sets(scenario.local_reachable_sets_conv(mpa))
    hl, hw = scenario.VEH_LENGTH / 2 + 0.01, scenario.VEH_WIDTH / 2 + 0.01
    # (a) one call per time step
    planner.closed_loop_reset(amount, hl, hw)
    one = scenario.ScenarioRunner(sc, None, max_num_CLs=max_cls,
                                  states_fn=lambda *a: planner.plan_timestep_from_states(*a, raise_on_search_error=False))
    res_one = [one.step_timestep() for _ in range(7)]
    poses_one, trims_one = one.pose.copy(), one.trim.copy()
    # (b) the three separate device calls (own closed-loop state: reset)
    planner.closed_loop_reset(amount, hl, hw)
    sep = scenario.ScenarioRunner(mk(), None, max_num_CLs=max_cls, inputs_fn=planner.sample_inputs,
                                  obstacles_fn=planner.assemble_obstacles,
                                  closed_loop_fn=lambda b, d, s, st: planner.plan_timestep_closed_loop(b, d, s, st, False))
    for k in range(7):
        got = sep.step_timestep()
        parity.compare(res_one[k], got)
    assert np.array_equal(poses_one.view(np.uint64), sep.pose.view(np.uint64)) and np.array_equal(trims_one, sep.trim)
    assert one.n_fallbacks == sep.n_fallbacks
    # (c) host functions + oracle, level by level
    ref = scenario.ScenarioRunner(mk(), lambda b: oracle_py.plan_batch(mpa, b, 4), max_num_CLs=max_cls) if max_cls >= amount else None
    if ref is not None:
        for _ in range(7):
            ref.step()
        assert np.array_equal(poses_one.view(np.uint64), ref.pose.view(np.uint64)) and np.array_equal(trims_one, ref.trim)
    with pytest.raises(capi.PdmpcError):     # slots must be distinct
        planner.plan_timestep_from_states([0, 1], [1.0, 2.0], [1.0, 2.0], [0.0, 0.0], [0.0, 0.0], [1, 1], [[], []], [[], []],
                                          [[], []], [3, 3], hl, hw, mpa.dt_seconds, sc.checker)
