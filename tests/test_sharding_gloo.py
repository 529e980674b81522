"""N>1 host logic on CPU: world_size-2 gloo (one process per would-be GPU)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import oracle_py
    from pdmpc_b200 import sharding
    from helpers import road_records
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # (1) scenario sharding: disjoint cover, no collective on the data path
        mine = sharding.shard_block_cyclic(7, rank, world)
        # (2) every rank plans ITS searches only; results must equal the single-process run
        mpa, batch = road_records("single_speed", 3)
        idx = sharding.shard_block_cyclic(batch.n, rank, world, block=4)
        res = oracle_py.plan_batch(mpa, batch.select(idx))      # CPU stand-in for the per-rank planner
        counters = sharding.gather_counters([idx.size, int(res.n_pops.sum())])
        # (3) permutation choice: 5 permutations over 2 ranks, 6 vehicles in 2 sub-graphs
        rng = np.random.default_rng(7)
        cost = rng.random((5, 6))
        cost[3] = cost[1]                                        # exact tie between permutations 1 and 3
        cost[:, 3:] += 1e-10 * rng.random((5, 3))                # differences below the 1e-8 rounding
        bel = np.array([1, 1, 1, 2, 2, 2])
        pid = sharding.shard_block_cyclic(5, rank, world)
        chosen, sc = sharding.choose_permutation(cost[pid], pid, 5, bel)
        plans = np.arange(5 * 6 * 3, dtype=np.float64).reshape(5, 6, 3)
        win = sharding.gather_winner_plans(plans[pid], pid, 5, chosen, bel)
        q.put((rank, mine.tolist(), idx.tolist(), res.pop_hash.tolist(), counters.tolist(), chosen.tolist(),
               sc.tolist(), win.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_permutation_choice():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle_py
    from pdmpc_b200 import sharding
    from helpers import road_records
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, mine0, idx0, h0, c0, ch0, sc0, w0), (r1, mine1, idx1, h1, c1, ch1, sc1, w1) = out
    assert sorted(mine0 + mine1) == list(range(7)) and not set(mine0) & set(mine1)
    mpa, batch = road_records("single_speed", 3)
    ref = oracle_py.plan_batch(mpa, batch)
    assert sorted(idx0 + idx1) == list(range(batch.n))
    assert h0 == ref.pop_hash[idx0].tolist() and h1 == ref.pop_hash[idx1].tolist()
    assert c0 == c1 and sum(row[0] for row in c0) == batch.n and sum(row[1] for row in c0) == int(ref.n_pops.sum())
    # both ranks reach the same decision, equal to the single-process computation
    rng = np.random.default_rng(7)
    cost = rng.random((5, 6))
    cost[3] = cost[1]
    cost[:, 3:] += 1e-10 * rng.random((5, 3))
    bel = np.array([1, 1, 1, 2, 2, 2])
    chosen, sc = sharding.choose_permutation(cost, np.arange(5), 5, bel)
    assert ch0 == ch1 == chosen.tolist() and sc0 == sc1 == sc.tolist()
    plans = np.arange(5 * 6 * 3, dtype=np.float64).reshape(5, 6, 3)
    assert w0 == w1 == sharding.gather_winner_plans(plans, np.arange(5), 5, chosen, bel).tolist()


def test_first_minimum_and_rounding_rule():
    from pdmpc_b200 import sharding
    # PrioritizedExplorativeController.m:153-154: ties after round(., 8) go to the FIRST permutation
    cost = np.array([[0.5, 0.2], [0.5 - 4e-9, 0.2], [0.4, 0.1 + 1e-12], [0.4 + 3e-9, 0.1]])
    chosen, sc = sharding.choose_permutation(cost, np.arange(4), 4, np.array([1, 2]))
    assert chosen.tolist() == [2, 2]
    assert sharding.matlab_round(np.array([0.123456785, -0.123456785, 2.5e-9]), 8).tolist() == \
        pytest.approx([0.12345679, -0.12345679, 0.0], abs=1e-15)
    with pytest.raises(ValueError):
        sharding.choose_permutation(cost[:2], [0, 1], 4, np.array([1, 2]))


def _explorative_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import oracle_py
    from pdmpc_b200 import scenario
    from pdmpc_b200.mpa import get_mpa
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        mpa = get_mpa("single_speed", non_convex=True)
        plan = lambda b: oracle_py.plan_batch(mpa, b)
        r = scenario.ExplorativeRunner(scenario.commonroad_scenario(mpa, 12, seed=4),
                                       lambda b, d: scenario.plan_timestep_by_levels(plan, b, d), rank=rank, world=world)
        r.run(5)
        q.put((rank, r.pose.tolist(), r.trim.tolist(), [e["chosen"].tolist() for e in r.explorative_records],
               [e["solution_cost"].tolist() for e in r.explorative_records],
               [e["searches_local"] for e in r.explorative_records]))
    finally:
        dist.destroy_process_group()


def test_explorative_priorities_two_ranks_equal_one_process():
    """BASELINE configs[2]: the permutations of a time step solved by two ranks (one all_gather of the
    costs, one of the winners' plans) drive the same closed loop as one process solving all of them."""
    import torch.multiprocessing as mp
    from oracle import oracle_py
    from pdmpc_b200 import scenario
    from pdmpc_b200.mpa import get_mpa
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_explorative_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    mpa = get_mpa("single_speed", non_convex=True)
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    one = scenario.ExplorativeRunner(scenario.commonroad_scenario(mpa, 12, seed=4),
                                     lambda b, d: scenario.plan_timestep_by_levels(plan, b, d))
    one.run(5)
    for rank, pose, trim, chosen, cost, local in out:
        assert pose == one.pose.tolist() and trim == one.trim.tolist()
        assert chosen == [e["chosen"].tolist() for e in one.explorative_records]
        assert cost == [e["solution_cost"].tolist() for e in one.explorative_records]
    # the two ranks split the permutations of every step between them
    for a, b, e in zip(out[0][5], out[1][5], one.explorative_records):
        assert a + b == e["searches_local"] and a > 0
    # some step chose a permutation other than the base prioritisation
    assert any(any(c) for c in out[0][3])
    # and differs from the plain prioritized loop
    plainr = scenario.ScenarioRunner(scenario.commonroad_scenario(mpa, 12, seed=4), plan)
    plainr.run(5)
    assert not np.array_equal(plainr.pose, one.pose)


def test_computation_level_permutations_form_a_latin_square():
    from pdmpc_b200 import scenario
    for n in (1, 2, 3, 5, 8):
        for seed in (1, 2, 35):
            r = scenario.computation_level_permutations(n, seed)
            assert r[0].tolist() == list(range(1, n + 1))
            assert all(sorted(row) == list(range(1, n + 1)) for row in r)
            assert all(sorted(col) == list(range(1, n + 1)) for col in r.T)
    assert np.array_equal(scenario.computation_level_permutations(5, 7), scenario.computation_level_permutations(5, 7))
    D = np.zeros((5, 5), dtype=bool)
    D[0, 1] = D[3, 4] = True
    assert scenario.weak_components(D).tolist() == [1, 1, 2, 3, 3]


def test_fixed_permutation_set_pads_and_truncates():
    """BASELINE configs[2] fixes 8 permutations per time step: the reference's Latin-square rows first, then rows
    of further squares (distinct while n_CL! allows), truncated when n_CL > 8."""
    import math
    from pdmpc_b200 import scenario
    for n in (1, 2, 3, 4, 5, 8, 9):
        for seed in (1, 17):
            base = scenario.computation_level_permutations(n, seed)
            r = scenario.fixed_permutation_set(n, seed, 8)
            assert r.shape == (8, n)
            m = min(n, 8)
            assert np.array_equal(r[:m], base[:m])                       # the reference's rows come first
            assert all(sorted(row) == list(range(1, n + 1)) for row in r)
            distinct = len({tuple(row) for row in r})
            assert distinct == min(8, math.factorial(n))
            assert np.array_equal(r, scenario.fixed_permutation_set(n, seed, 8))


def test_explorative_exchange_hook_equals_host_exchange():
    """ExplorativeRunner(exchange=...) — the bench's device-side exchange starts from rows
    [cost, fallback flag, trims, poses, shape sizes, shapes] per (permutation, vehicle) as pdmpc_pack_plan_rows
    writes them; fed with the same rows built on the host it must drive the same closed loop as the built-in
    host exchange, with 8 fixed permutations per step."""
    from oracle import oracle_py
    from pdmpc_b200 import scenario, sharding
    from pdmpc_b200.mpa import get_mpa
    mpa = get_mpa("single_speed", non_convex=True)
    Hp = mpa.Hp
    plan = lambda b: oracle_py.plan_batch(mpa, b)
    last = {}

    def ts(b, d):
        last["res"] = scenario.plan_timestep_by_levels(plan, b, d)
        return last["res"]

    def exchange(n_rows, n, fb_rows, mine, P, belonging):
        res = last["res"]
        L = 2 + 21 * Hp
        rows = np.zeros((n_rows, L))
        for r in range(n_rows):
            if res.is_exhausted[r]:
                rows[r] = fb_rows[r % n]
                rows[r, 1] = 1.0
            else:
                rows[r] = np.concatenate([[res.g_path[r, Hp], 0.0], res.trims[r, 1:], res.y_predicted[r].reshape(-1),
                                          res.shape_npts[r], res.shape_x[r].reshape(-1), res.shape_y[r].reshape(-1)])
        rows = rows.reshape(P, n, L)
        chosen, cost = sharding.solution_costs(rows[:, :, 0], belonging)
        return chosen, cost, np.stack([rows[chosen[belonging[v] - 1], v, 1:] for v in range(n)])

    a = scenario.ExplorativeRunner(scenario.commonroad_scenario(mpa, 12, seed=4), ts, fixed_permutations=8, exchange=exchange)
    b = scenario.ExplorativeRunner(scenario.commonroad_scenario(mpa, 12, seed=4), ts, fixed_permutations=8)
    a.run(4)
    b.run(4)
    assert np.array_equal(a.pose, b.pose) and np.array_equal(a.trim, b.trim)
    for x, y in zip(a.explorative_records, b.explorative_records):
        assert x["n_permutations"] == 8 and np.array_equal(x["chosen"], y["chosen"])
        assert np.array_equal(x["solution_cost"], y["solution_cost"])
