#!/usr/bin/env python3
"""BASELINE configs[2] on N GPUs: simultaneous multiple prioritizations, permutations sharded over the ranks.

  python tools/explorative_run.py [--steps 35] [--seed 1]                       # one GPU: all permutations in one call
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/explorative_run.py                                                  # N GPUs, NCCL exchange

Every rank runs the (replicated) closed loop; per time step it solves ITS permutations with one
pdmpc_plan_timestep call, then the ranks exchange the cost matrix and the winners' plans (two NCCL
all_gathers, sharding.py).  Rank 0 prints one JSON line: per-time-step wall time (planning call + exchange),
and a check that the closed loop equals the one a single rank computes with all permutations in one call.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=35)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--vehicles", type=int, default=20)
    ap.add_argument("--mpa", default="triple_speed")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    import torch
    from pdmpc_b200 import capi, scenario
    from pdmpc_b200.mpa import get_mpa
    torch.cuda.set_device(local)
    device = torch.device(f"cuda:{local}")
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    mpa = get_mpa(args.mpa, non_convex=True)
    planner = capi.Planner(local)
    planner.upload_mpa(mpa)
    call_ms = []

    def timed_call(b, d):
        t0 = time.perf_counter()
        r = planner.plan_timestep(b, d, False)
        call_ms.append((time.perf_counter() - t0) * 1e3)
        return r

    def run(rk, ws, dev):
        r = scenario.ExplorativeRunner(scenario.commonroad_scenario(mpa, args.vehicles, seed=args.seed), timed_call,
                                       rank=rk, world=ws, device=dev)
        step_ms = []
        for _ in range(args.steps):
            if ws > 1:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            r.step()
            step_ms.append((time.perf_counter() - t0) * 1e3)
        return r, np.array(step_ms)

    run(rank, world, device if world > 1 else None)        # warm-up (allocations, NCCL channels)
    call_ms.clear()
    r, step_ms = run(rank, world, device if world > 1 else None)
    sharded_calls = np.array(call_ms)
    line = None
    if rank == 0:
        call_ms.clear()
        one, one_ms = (r, step_ms) if world == 1 else run(0, 1, None)
        same = bool(np.array_equal(one.pose, r.pose) and np.array_equal(one.trim, r.trim) and all(
            np.array_equal(a["chosen"], b["chosen"]) for a, b in zip(one.explorative_records, r.explorative_records)))
        P = [e["n_permutations"] for e in r.explorative_records]
        line = {"config": "BASELINE configs[2]: road network, %d vehicles, %s MPA, explorative priorities, %d steps"
                          % (args.vehicles, args.mpa, args.steps),
                "n_gpus": world, "permutations_per_step": {"mean": float(np.mean(P)), "max": int(max(P))},
                "searches_per_step_all_ranks": float(np.mean(P)) * args.vehicles,
                "planning_call_ms": {"p50": float(np.percentile(sharded_calls, 50)),
                                     "p99": float(np.percentile(sharded_calls, 99))},
                "single_rank_call_ms": {"p50": float(np.percentile(call_ms, 50)), "p99": float(np.percentile(call_ms, 99))}
                if world > 1 else None,
                "closed_loop_equals_single_rank": same,
                "permutation_other_than_base_chosen_in_steps": int(sum(bool(np.any(e["chosen"])) for e in r.explorative_records)),
                "fallbacks": int(r.n_fallbacks),
                "note": "step wall time is dominated by the replicated Python host logic (reference-trajectory sampling, "
                        "coupling, priorities); planning_call_ms is the optimizer call the ranks shard"}
        print(json.dumps(line))
        assert same
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    planner.close()


if __name__ == "__main__":
    main()
