"""bench.py --config explorative: BASELINE configs[2] — simultaneous multiple prioritizations,
20 vehicles x 8 priority permutations per time step, the permutations sharded over the GPUs (one per GPU at 8).

    PrioritizedExplorativeController.m:21-176   permutations of the computation levels, all solved, the
                                                cheapest per weakly connected sub-graph applied
    computation_level_permutations :241-309     Latin square (fixed to 8 rows here: scenario.fixed_permutation_set)

Per time step a rank plans its permutations as ONE pdmpc_plan_timestep call (each permutation is a complete
20-vehicle time step with its own predecessor DAG), packs costs + plans into rows in device memory
(pdmpc_pack_plan_rows) and the ranks exchange them with ONE NCCL all_gather on the device buffers; the choice
(sum per sub-graph in vehicle order, round(., 8), first minimum) is then made identically on every rank.
The closed loop must equal the one a single rank computes with all 8 permutations in one call (asserted).

A bench "step" = one time step of the closed loop.  Timed: the planning call and the exchange (the hot path and
its collective); the replicated host logic around them (reference trajectory, coupling, priorities — Python) is
reported separately and is not part of the metric.
"""
from __future__ import annotations

import json
import time

import numpy as np


def run(args, rank: int, local_rank: int, world: int, ClockSampler) -> None:
    import torch
    import torch.distributed as dist
    from . import capi, scenario, sharding
    from .mpa import get_mpa

    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    device = torch.device(f"cuda:{dev}")
    planner = capi.Planner(dev)
    mpa = get_mpa(args.mpa, non_convex=True)
    planner.upload_mpa(mpa)
    planner.set_cta_queue(True)
    Hp, n, P = mpa.Hp, args.vehicles, args.permutations
    L = 2 + 21 * Hp
    per_rank = (P + world - 1) // world
    local_rows = torch.zeros((per_rank * n, L), dtype=torch.float64, device=device)
    gathered = torch.zeros((world, per_rank * n, L), dtype=torch.float64, device=device)
    host_rows = torch.zeros((world, per_rank * n, L), dtype=torch.float64).pin_memory()
    call_ms, exch_ms, coll_ms = [], [], []

    def timed_call(b, d):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = planner.plan_timestep(b, d, False)
        call_ms.append((time.perf_counter() - t0) * 1e3)
        return r

    def exchange(n_rows, n_veh, fb_rows, mine, n_perm, belonging):
        """Costs + plans of this rank's permutations -> every rank: device rows, one all_gather, one copy."""
        t0 = time.perf_counter()
        planner.pack_plan_rows(n_rows, n_veh, fb_rows, local_rows.data_ptr())
        t1 = time.perf_counter()
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), local_rows.view(-1))
        else:
            gathered[0].copy_(local_rows)
        host_rows.copy_(gathered, non_blocking=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        rows_all = host_rows.numpy().reshape(world, per_rank, n_veh, L)
        # permutation p was planned by rank p % world as its (p // world)-th (sharding.shard_block_cyclic)
        rows = np.stack([rows_all[p % world, p // world] for p in range(n_perm)])
        chosen, solution_cost = sharding.solution_costs(rows[:, :, 0], belonging)
        plans = np.stack([rows[chosen[belonging[v] - 1], v, 1:] for v in range(n_veh)])
        exch_ms.append((time.perf_counter() - t0) * 1e3)
        coll_ms.append((t2 - t1) * 1e3)
        return chosen, solution_cost, plans

    total_steps = args.warmup + args.steps

    def closed_loop(rk, ws, ex, timestep_fn):
        """total_steps time steps over scenarios seed 1, 2, ... (sim_steps each); returns the runners."""
        runners, done, seed = [], 0, 1
        host_ms = []
        while done < total_steps:
            r = scenario.ExplorativeRunner(scenario.commonroad_scenario(mpa, n, seed=seed), timestep_fn, rank=rk, world=ws,
                                           device=None, fixed_permutations=P, exchange=ex)
            for _ in range(min(args.sim_steps, total_steps - done)):
                t0 = time.perf_counter()
                r.step()
                host_ms.append((time.perf_counter() - t0) * 1e3)
                done += 1
            runners.append(r)
            seed += 1
        return runners, host_ms

    # untimed pass: allocations, NCCL channels, library buffers
    closed_loop(rank, world, exchange, timed_call)
    call_ms.clear(); exch_ms.clear(); coll_ms.clear()
    sampler = ClockSampler(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    runners, step_ms = closed_loop(rank, world, exchange, timed_call)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    st = planner.stats()
    W = args.warmup
    plan = np.array(call_ms[W:]) + np.array(exch_ms[W:])
    t_local = torch.tensor([plan.sum()], dtype=torch.float64, device=device)
    lat = torch.tensor(plan, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_local, op=dist.ReduceOp.MAX)
        dist.all_reduce(lat, op=dist.ReduceOp.MAX)     # a time step is planned when the slowest rank is done
    lat = lat.cpu().numpy()
    searches_per_step = P * n
    value = searches_per_step * args.steps / (float(t_local.item()) * 1e-3)

    if rank == 0:
        same = None
        single = None
        if world > 1:
            # the same closed loop on ONE rank, all 8 permutations in one call, host-side exchange (no collective)
            single_ms = []

            def single_call(b, d):
                t0 = time.perf_counter()
                r = planner.plan_timestep(b, d, False)
                single_ms.append((time.perf_counter() - t0) * 1e3)
                return r
            ones, _ = closed_loop(0, 1, None, single_call)
            same = all(np.array_equal(a.pose, b.pose) and np.array_equal(a.trim, b.trim) and
                       all(np.array_equal(x["chosen"], y["chosen"]) and np.array_equal(x["solution_cost"], y["solution_cost"])
                           for x, y in zip(a.explorative_records, b.explorative_records))
                       for a, b in zip(ones, runners))
            single = {"p50": float(np.percentile(single_ms[W:], 50)), "p99": float(np.percentile(single_ms[W:], 99))}
        recs = [e for r in runners for e in r.explorative_records]
        line = {
            "metric": "vehicle-plans/sec", "value": value, "unit": "plans/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(t_local.item()) / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[2]: CPM Lab road network, {n} vehicles, {args.mpa} MPA, Hp {Hp}, InterX, "
                                   f"explorative priorities: {P} permutations of the computation levels per time step "
                                   f"(reference: n_CL rows of one Latin square; padded / truncated to {P}, "
                                   f"scenario.fixed_permutation_set), sharded block-cyclically over {world} GPU(s), closed loop",
                       "searches_per_step": searches_per_step, "permutations_per_gpu": per_rank,
                       "l2": "not flushed: a closed-loop time step (latency workload), inputs from host buffers every step"},
            "e2e": {"value": value, "unit": "plans/s", "h2d_bytes_per_step": int(st.h2d_bytes),
                    "d2h_bytes_per_step": int(st.d2h_bytes) + int(host_rows.numel() * 8),
                    "note": "the timed region IS end to end: host buffers in, plans of the chosen permutation on every "
                            "rank's host out (planning call + pack + all_gather + copy)"},
            "gpu_launches": int(args.steps * 2),   # per time step: the dependency-ordered search kernel + the pack kernel
            "clocks": clocks,
            "latency_ms_per_timestep": {"p50": float(np.percentile(lat, 50)), "p99": float(np.percentile(lat, 99)),
                                        "max": float(lat.max()), "n": int(lat.size),
                                        "what": "planning call + exchange of one time step, max over ranks"},
            "planning_call_ms": {"p50": float(np.percentile(call_ms[W:], 50)), "p99": float(np.percentile(call_ms[W:], 99))},
            "exchange_ms": {"p50": float(np.percentile(exch_ms[W:], 50)), "p99": float(np.percentile(exch_ms[W:], 99)),
                            "all_gather_and_copy_p50": float(np.percentile(coll_ms[W:], 50)),
                            "share_of_planning": float(np.sum(exch_ms[W:]) / np.sum(plan)),
                            "bytes_gathered_per_rank": int(gathered.numel() * 8),
                            "what": "pdmpc_pack_plan_rows (device) + ONE all_gather on the device rows (NCCL) + one "
                                    "device-to-host copy + the choice"},
            "single_rank_call_ms": single,
            "closed_loop_equals_single_rank": same,
            "permutation_other_than_base_chosen_in_steps": int(sum(bool(np.any(e["chosen"])) for e in recs)),
            "computation_levels": {"mean": float(np.mean([e["n_levels"] for e in recs])), "max": int(max(e["n_levels"] for e in recs))},
            "fallbacks": int(sum(r.n_fallbacks for r in runners)),
            "host_logic_ms_per_step_p50": float(np.percentile(step_ms[W:], 50)),
            "roofline": None, "cpu_baseline": None,
            "note": "latency-bound closed loop: `value` counts the searches of all permutations per planning time; the "
                    "replicated Python host logic (host_logic_ms_per_step) is outside the metric",
        }
        print(json.dumps(line))
        assert same is not False, "sharded closed loop differs from the single-rank one"
    planner.close()
