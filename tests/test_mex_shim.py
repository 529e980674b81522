"""The MEX shim (p-dmpc_b200/matlab/pdmpc_b200_mex.cpp) compiled UNMODIFIED against a fake MATLAB
C Matrix API and driven the way GraphSearchCuda.m drives it.  CPU part: it builds, dispatches
commands and reports a missing device loudly.  GPU part: full marshalling parity against the
golden fixtures, one run_optimizer-equivalent call per search."""
import numpy as np
import pytest

import mexfake
from helpers import GOLDEN_CASES, load_golden, load_golden_mcts
from test_oracle_cpu import _iters_from_batch


@pytest.fixture(scope="module")
def matlab(built):
    return mexfake.FakeMatlab()


def test_shim_builds_and_fails_loudly_without_device(matlab):
    import torch
    with pytest.raises(mexfake.MexError, match="pdmpc:usage"):
        matlab.call(0)
    with pytest.raises(mexfake.MexError, match="pdmpc:handle"):
        matlab.call(0, mexfake.PLAN, 99.0)
    if not torch.cuda.is_available():
        with pytest.raises(mexfake.MexError, match="no CPU fallback"):
            matlab.call(1, mexfake.CREATE, 0.0)


def test_matlab_side_sources_present():
    """The MATLAB class keeps the reference's call surface (OptimizerInterface.m:14)."""
    import os
    src = open(os.path.join(mexfake.ROOT, "p-dmpc_b200", "matlab", "GraphSearchCuda.m")).read()
    assert "classdef GraphSearchCuda < OptimizerInterface" in src
    assert "function info = run_optimizer(obj, ~, iter, mpa, options, ~)" in src
    assert "create_control_results_info_from_mex" in src
    assert "with_hdv_reachable_sets" in src and "hdv_adjacency" in src
    smp = open(os.path.join(mexfake.ROOT, "p-dmpc_b200", "matlab", "MonteCarloTreeSearchCuda.m")).read()
    assert "classdef MonteCarloTreeSearchCuda < OptimizerInterface" in smp
    assert "function info_v = run_optimizer(obj, vehicle_index, iter, mpa, options, time_step)" in smp
    assert "time_step + vehicle_index" in smp
    comp = open(os.path.join(mexfake.ROOT, "p-dmpc_b200", "matlab", "compile_pdmpc_b200.m")).read()
    assert "mex(" in comp and "pdmpc_b200_mex.cpp" in comp


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mex_path_matches_golden(matlab, name):
    mpa, batch, exp = load_golden(name)
    (h,) = matlab.call(1, mexfake.CREATE, 0.0)
    try:
        trans, man = mexfake.matlab_mpa(mpa)
        matlab.call(0, mexfake.UPLOAD_MPA, float(h), trans, man)
        Hp = mpa.Hp
        step = max(1, batch.n // 40)
        picks = sorted(set(range(0, batch.n, step)) | set(np.flatnonzero(exp.is_exhausted).tolist()))
        for i in picks:
            it = _iters_from_batch(batch, i)
            exh, n_exp, trims, ypred, g, hh, shapes = matlab.call(
                7, mexfake.PLAN, float(h), it.x0[:3], float(it.trim_indices), it.reference_trajectory_points,
                it.v_ref, list(it.obstacles), [list(r) for r in it.dynamic_obstacle_area] or [],
                it.predicted_lanelet_boundary[0], it.predicted_lanelet_boundary[1], float(batch.checker),
                float(batch.dt_seconds))
            assert exh == bool(exp.is_exhausted[i]) and int(np.asarray(n_exp).reshape(-1)[0]) == exp.n_expanded[i]
            if not exh:
                assert trims.reshape(-1).astype(int).tolist() == exp.trims[i].tolist()
                assert np.array_equal(ypred.T.view(np.uint64), exp.y_predicted[i].view(np.uint64))
                assert np.array_equal(g.reshape(-1).view(np.uint64), exp.g_path[i].view(np.uint64))
                for k in range(Hp):
                    n = int(exp.shape_npts[i, k])
                    assert shapes[k].shape == (2, n)
                    assert np.array_equal(shapes[k][0].view(np.uint64), exp.shape_x[i, k, :n].view(np.uint64))
        (st,) = matlab.call(1, mexfake.STATS, float(h), raw=True)
        assert matlab.field(st, "total_pops").reshape(-1)[0] >= 1
    finally:
        matlab.call(0, mexfake.DESTROY, float(h))
        matlab.clear_mex()


@pytest.mark.gpu
def test_mex_sampled_path_matches_golden(matlab):
    """PLAN_SAMPLED (MonteCarloTreeSearchCuda.m) through the shim against the committed fixture."""
    name = "road_interx_triple_speed"
    mpa, batch, _ = load_golden(name)
    seeds, n_max, exp = load_golden_mcts(name)
    (h,) = matlab.call(1, mexfake.CREATE, 0.0)
    try:
        trans, man = mexfake.matlab_mpa(mpa)
        matlab.call(0, mexfake.UPLOAD_MPA, float(h), trans, man)
        for i in range(0, batch.n, 7):
            it = _iters_from_batch(batch, i)
            exh, n_exp, trims, ypred, g, hh, shapes = matlab.call(
                7, mexfake.PLAN_SAMPLED, float(h), it.x0[:3], float(it.trim_indices), it.reference_trajectory_points,
                it.v_ref, list(it.obstacles), [list(r) for r in it.dynamic_obstacle_area] or [],
                it.predicted_lanelet_boundary[0], it.predicted_lanelet_boundary[1], float(batch.checker),
                float(batch.dt_seconds), float(seeds[i]), float(n_max))
            assert exh == bool(exp.is_exhausted[i]) and int(np.asarray(n_exp).reshape(-1)[0]) == exp.n_expanded[i]
            if not exh:
                assert trims.reshape(-1).astype(int).tolist() == exp.trims[i].tolist()
                assert np.array_equal(ypred.T.view(np.uint64), exp.y_predicted[i].view(np.uint64))
                assert g.reshape(-1)[-1] == exp.g_path[i, -1]
    finally:
        matlab.call(0, mexfake.DESTROY, float(h))
        matlab.clear_mex()


def _timestep_args(batch, deps):
    """The PLAN_TIMESTEP argument list (MATLAB layouts) of one time step's flat batch."""
    n, Hp = batch.n, batch.Hp
    iters = batch.to_iters()
    S = max((len(it.obstacles) for it in iters), default=0)
    R = max((len(it.dynamic_obstacle_area) for it in iters), default=0)
    obstacles = [[it.obstacles[s] if s < len(it.obstacles) else None for s in range(max(S, 1))] for it in iters]
    dyn = [[(it.dynamic_obstacle_area[r][k] if r < len(it.dynamic_obstacle_area) and
             it.dynamic_obstacle_area[r][k].shape[1] else None)
            for k in range(Hp) for r in range(max(R, 1))] for it in iters]
    coupling = np.zeros((n, n))
    for i in range(n):
        coupling[deps.preds(i), i] = 1.0
    fb = [[s if s.shape[1] else None for s in deps.fallback_shapes(i)] for i in range(n)]
    return (np.stack([batch.x0, batch.y0, batch.yaw0], axis=1), batch.trim0.astype(np.float64).reshape(-1, 1),
            np.stack([batch.ref_x.reshape(n, Hp), batch.ref_y.reshape(n, Hp)], axis=2), batch.v_ref.reshape(n, Hp),
            obstacles, dyn, [it.predicted_lanelet_boundary[0] for it in iters],
            [it.predicted_lanelet_boundary[1] for it in iters], float(batch.checker), float(batch.dt_seconds),
            coupling, fb)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["timestep_road_triple_speed", "timestep_circle_single_speed"])
def test_mex_timestep_path_matches_golden(matlab, name):
    """PLAN_TIMESTEP (plan_timestep_cuda.m: all vehicles of a time step in one call) through the shim
    against the committed time-step fixtures."""
    from helpers import load_golden_timesteps
    mpa, steps = load_golden_timesteps(name)
    (h,) = matlab.call(1, mexfake.CREATE, 0.0)
    try:
        trans, man = mexfake.matlab_mpa(mpa)
        matlab.call(0, mexfake.UPLOAD_MPA, float(h), trans, man)
        for batch, deps, exp in steps[::3]:
            exh, n_exp, trims, y, shapes, g, hh = matlab.call(7, mexfake.PLAN_TIMESTEP, float(h), *_timestep_args(batch, deps))
            n, Hp = batch.n, batch.Hp
            assert g.shape == (n, Hp + 1) and hh.shape == (n, Hp + 1)
            assert exh.reshape(-1).astype(int).tolist() == exp.is_exhausted.tolist()
            assert n_exp.reshape(-1).astype(int).tolist() == exp.n_expanded.tolist()
            y = y.reshape(3, Hp, n, order="F")
            for i in range(n):
                if exp.is_exhausted[i]:
                    continue
                assert trims[i].astype(int).tolist() == exp.trims[i].tolist()
                assert np.array_equal(y[:, :, i].T.view(np.uint64), exp.y_predicted[i].view(np.uint64))
                # costs along the path (tree.g / tree.h of next_nodes): what compute_solution_cost compares
                assert np.array_equal(np.ascontiguousarray(g[i]).view(np.uint64), exp.g_path[i].view(np.uint64))
                assert np.array_equal(np.ascontiguousarray(hh[i]).view(np.uint64), exp.h_path[i].view(np.uint64))
                for k in range(Hp):
                    m = int(exp.shape_npts[i, k])
                    sh = shapes[i + n * k]
                    assert sh.shape == (2, m) and np.array_equal(sh[0].view(np.uint64), exp.shape_x[i, k, :m].view(np.uint64))
    finally:
        matlab.call(0, mexfake.DESTROY, float(h))
        matlab.clear_mex()


def test_timestep_matlab_helper_present():
    import os
    src = open(os.path.join(mexfake.ROOT, "p-dmpc_b200", "matlab", "plan_timestep_cuda.m")).read()
    assert "PLAN_TIMESTEP = 6" in src and "directed_coupling_sequential" in src
    assert "create_control_results_info_from_mex" in src
    assert "g(i, k + 1), h(i, k + 1)" in src and "with_hdv_reachable_sets" in src


@pytest.mark.gpu
def test_mex_joint_path_matches_oracle(matlab):
    """PLAN_JOINT (GraphSearchCuda.run_optimizer with iter.amount > 1, the centralized controller's call)
    through the shim against the oracle's joint search."""
    from oracle import oracle_py
    from pdmpc_b200.mpa import get_mpa
    from test_joint import joint_cases
    mpa = get_mpa("single_speed", non_convex=False)
    (h,) = matlab.call(1, mexfake.CREATE, 0.0)
    try:
        trans, man = mexfake.matlab_mpa(mpa)
        matlab.call(0, mexfake.UPLOAD_MPA, float(h), trans, man)
        from pdmpc_b200.records import TimestepDeps
        for name, batch, nV in joint_cases(mpa)[:1] + joint_cases(mpa)[3:4]:
            n, Hp = batch.n, batch.Hp
            ref = oracle_py.joint_plan_batch(mpa, batch, nV, max_nodes=1 << 17)
            args = _timestep_args(batch, TimestepDeps.build([[] for _ in range(n)], [None] * n, Hp))[:10]
            exh, n_exp, trims, y, shapes, g, hh = matlab.call(7, mexfake.PLAN_JOINT, float(h), *args)
            assert exh.reshape(-1).astype(int).tolist() == ref.is_exhausted.tolist(), name
            assert n_exp.reshape(-1).astype(int).tolist() == ref.n_expanded.tolist()
            y = y.reshape(3, Hp, n, order="F")
            for i in range(n):
                assert trims[i].astype(int).tolist() == ref.trims[i].tolist()
                assert np.array_equal(y[:, :, i].T.view(np.uint64), ref.y_predicted[i].view(np.uint64))
            assert np.array_equal(g.reshape(-1).view(np.uint64), ref.g_path[0].view(np.uint64))
    finally:
        matlab.call(0, mexfake.DESTROY, float(h))
        matlab.clear_mex()


@pytest.mark.gpu
def test_mex_sample_inputs_matches_the_host_functions(matlab):
    """UPLOAD_ROAD + SAMPLE_INPUTS (sample_inputs_cuda.m): reference trajectories and lanelet boundaries of all vehicles of
    a time step through the shim, bit for bit against the host restatement of the reference's functions."""
    from pdmpc_b200 import scenario
    from pdmpc_b200.mpa import get_mpa
    mpa = get_mpa("triple_speed", non_convex=True)
    sc = scenario.commonroad_scenario(mpa, 20, seed=2)
    (h,) = matlab.call(1, mexfake.CREATE, 0.0)
    try:
        trans, man = mexfake.matlab_mpa(mpa)
        matlab.call(0, mexfake.UPLOAD_MPA, float(h), trans, man)
        bounds = [[np.ascontiguousarray(l), np.ascontiguousarray(r)] for l, r in sc.road.boundary]      # {nL x 2} of n x 2
        veh = sc.vehicles
        matlab.call(0, mexfake.UPLOAD_ROAD, float(h), bounds, [v.reference_path for v in veh],
                    [np.asarray(v.lanelets_index, dtype=np.float64) for v in veh],
                    [np.asarray(v.points_index, dtype=np.float64) for v in veh],
                    np.array([float(v.is_loop) for v in veh]), np.array([v.reference_speed for v in veh]))
        rng = np.random.default_rng(5)
        n = len(veh)
        poses = np.array([v.reference_path[rng.integers(0, v.reference_path.shape[0])] + rng.normal(scale=0.02, size=2) for v in veh])
        trims = rng.integers(1, mpa.trim_speed.size + 1, size=n)
        speed = np.array([mpa.trim_speed[t - 1] for t in trims])
        ref, v_ref, pidx, cur, pred, left, right = matlab.call(7, mexfake.SAMPLE_INPUTS, float(h), np.arange(1, n + 1, dtype=np.float64),
                                                               poses[:, 0], poses[:, 1], speed, mpa.dt_seconds)
        assert ref.shape == (n, mpa.Hp, 2) and v_ref.shape == (n, mpa.Hp)
        for i, v in enumerate(veh):
            pts, vr, idx, c = scenario.get_reference_trajectory(mpa, v, poses[i, 0], poses[i, 1], int(trims[i]), mpa.dt_seconds)
            lan = scenario.get_predicted_lanelets(v, idx)
            lft, rgt = scenario.get_lanelets_boundary(lan, sc.road, v)
            assert np.array_equal(np.ascontiguousarray(ref[i]).view(np.uint64), np.ascontiguousarray(pts).view(np.uint64))
            assert np.array_equal(v_ref[i], vr) and pidx[i].astype(int).tolist() == idx.tolist() and int(cur[i, 0]) == c
            assert pred[i].reshape(-1).astype(int).tolist() == lan.tolist()
            assert np.array_equal(left[i], lft) and np.array_equal(right[i], rgt)
        with pytest.raises(mexfake.MexError):
            matlab.call(1, mexfake.SAMPLE_INPUTS, float(h), np.array([n + 1.0]), poses[:1, 0], poses[:1, 1], speed[:1], mpa.dt_seconds)
    finally:
        matlab.call(0, mexfake.DESTROY, float(h))
        matlab.clear_mex()


def test_sample_inputs_matlab_helper_present():
    import os
    src = open(os.path.join(mexfake.ROOT, "p-dmpc_b200", "matlab", "sample_inputs_cuda.m")).read()
    assert "UPLOAD_ROAD = 8" in src and "SAMPLE_INPUTS = 9" in src and "predicted_lanelet_boundary" in src


@pytest.mark.gpu
def test_mex_assemble_obstacles_matches_the_host_assembly(matlab):
    """UPLOAD_REACHABLE_SETS + ASSEMBLE_OBSTACLES (assemble_obstacles_cuda.m): standing successors' areas and parallel
    predecessors' reachable sets of all vehicles through the shim, bit for bit against the host restatement."""
    from pdmpc_b200 import scenario
    from pdmpc_b200.mpa import get_mpa
    mpa = get_mpa("triple_speed", non_convex=True)
    (h,) = matlab.call(1, mexfake.CREATE, 0.0)
    try:
        trans, man = mexfake.matlab_mpa(mpa)
        matlab.call(0, mexfake.UPLOAD_MPA, float(h), trans, man)
        sets = scenario.local_reachable_sets_conv(mpa)
        matlab.call(0, mexfake.UPLOAD_REACHABLE_SETS, float(h), [[np.ascontiguousarray(a) for a in row] for row in sets])
        rng = np.random.default_rng(11)
        n = 6
        x0 = np.column_stack([rng.uniform(0.5, 4.0, n), rng.uniform(0.5, 3.5, n), rng.uniform(-3.0, 3.0, n),
                              np.array([0.0, 0.6, 0.0, 0.3, 0.005, 0.9])])
        trims = rng.integers(1, mpa.n_trims + 1, size=n)
        succ = [[1, 2], [2, 4], [], [4, 5], [], []]
        par = [[], [0], [0, 1], [], [3], [0, 3, 4]]
        obs, dyn = matlab.call(2, mexfake.ASSEMBLE_OBSTACLES, float(h), x0, trims.astype(np.float64),
                               [np.array(s, dtype=np.float64) + 1 for s in succ], [np.array(p, dtype=np.float64) + 1 for p in par],
                               0.12, 0.0625)
        want = scenario.assemble_obstacles_host(mpa, x0[:, 0], x0[:, 1], x0[:, 2], x0[:, 3], trims, succ, par, 0.12, 0.0625)
        sp, pp = want["slot_ptr"], want["poly_ptr"]
        poly = lambda p: np.vstack([want["vert_x"][pp[p]:pp[p + 1]], want["vert_y"][pp[p]:pp[p + 1]]])
        for i in range(n):
            s0 = i * (mpa.Hp + 1)
            got_static = list(obs[i])                       # {s_i x 1} cell
            assert len(got_static) == sp[s0 + 1] - sp[s0]
            for q, a in enumerate(got_static):
                assert np.array_equal(a, poly(sp[s0] + q))
            rows = len(par[i])
            assert len(dyn[i]) == rows * mpa.Hp             # {p_i x Hp} cell, column-major
            for q in range(rows):
                for k in range(mpa.Hp):
                    assert np.array_equal(dyn[i][q + rows * k], poly(sp[s0 + 1 + k] + q))
        with pytest.raises(mexfake.MexError):
            matlab.call(2, mexfake.ASSEMBLE_OBSTACLES, float(h), x0, trims.astype(np.float64),
                        [np.array([n + 1.0])] + [np.zeros(0)] * (n - 1), [np.zeros(0)] * n, 0.12, 0.0625)
    finally:
        matlab.call(0, mexfake.DESTROY, float(h))
        matlab.clear_mex()


def test_assemble_obstacles_matlab_helper_present():
    import os
    src = open(os.path.join(mexfake.ROOT, "p-dmpc_b200", "matlab", "assemble_obstacles_cuda.m")).read()
    assert "UPLOAD_REACHABLE_SETS = 10" in src and "ASSEMBLE_OBSTACLES = 11" in src and "local_reachable_sets_conv" in src
