"""Multi-GPU plumbing of the optimizer path: one process per GPU.

Searches are independent given their inputs, so the data path needs NO collective:
whole scenarios go to different ranks (`shard_block_cyclic`).  The single real
exchange step of the reference is the choice among simultaneously solved priority
permutations (BASELINE config 3):

  PrioritizedExplorativeController.compute_solution_cost  :94-109   cost = tree.get_cost(goal) per vehicle
  receive_solution_cost                                   :124-144  summed per weakly connected sub-graph
  choose_solution                                         :146-176  round(., 8), first minimum

`choose_permutation` reproduces it across ranks when every rank solved a subset of
the permutations: one all_gather of the (tiny) per-vehicle cost rows, then a
deterministic reduction in vehicle order on every rank (an all_reduce would leave
the summation order to the backend; the reference's result before rounding depends
on it), then `gather_winner_plans` moves the chosen plans with one more all_gather.
Works with any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests);
with no process group it degenerates to the single-process computation.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def shard_block_cyclic(n_items: int, rank: int, world: int, block: int = 1) -> np.ndarray:
    """Indices of the items (scenarios, permutations) rank `rank` owns."""
    idx = np.arange(n_items)
    return idx[(idx // block) % world == rank]


def matlab_round(x: np.ndarray, digits: int = 8) -> np.ndarray:
    """MATLAB round(x, n): scale, round half away from zero, unscale."""
    s = 10.0 ** digits
    y = np.asarray(x, dtype=np.float64) * s
    return np.sign(y) * np.floor(np.abs(y) + 0.5) / s


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except ImportError:
        pass
    return None


def _all_gather_rows(local: np.ndarray, device=None, collective: bool = True) -> np.ndarray:
    """Concatenate equally shaped float64 arrays of all ranks along a new leading axis.
    collective=False: this caller owns everything (a single-rank computation inside a multi-rank job)."""
    dist = _dist() if collective else None
    if dist is None:
        return local[None]
    import torch
    t = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])


def solution_costs(cost: np.ndarray, belonging_vector: np.ndarray):
    """cost[p, v] -> (chosen permutation per sub-graph, rounded cost matrix [P, n_graphs]): summed per weakly
    connected sub-graph in vehicle order (receive_solution_cost :124-144), round(., 8) and first minimum
    (choose_solution :153-154).  Deterministic on every rank: the order of the additions is fixed here, not
    left to a reduction collective."""
    belonging_vector = np.asarray(belonging_vector)
    n_graphs = int(belonging_vector.max())
    solution_cost = np.zeros((cost.shape[0], n_graphs))
    for g in range(1, n_graphs + 1):
        for v in np.flatnonzero(belonging_vector == g):
            solution_cost[:, g - 1] = solution_cost[:, g - 1] + cost[:, v]
    solution_cost = matlab_round(solution_cost, 8)
    return np.argmin(solution_cost, axis=0), solution_cost


def choose_permutation(cost_local: np.ndarray, perm_ids_local: Sequence[int], n_permutations: int,
                       belonging_vector: np.ndarray, device=None, collective: bool = True):
    """cost_local[p, v]: cost-to-come of vehicle v's goal node in the p-th permutation THIS rank
    solved (perm_ids_local[p] = its global index); belonging_vector[v] = sub-graph (1-based) of v.
    Returns (chosen permutation per sub-graph [n_graphs], rounded cost matrix [n_permutations, n_graphs])."""
    belonging_vector = np.asarray(belonging_vector)
    n_veh = belonging_vector.size
    full_local = np.zeros((n_permutations, n_veh))
    mask_local = np.zeros((n_permutations, n_veh))
    for row, p in zip(np.asarray(cost_local, dtype=np.float64).reshape(len(perm_ids_local), n_veh), perm_ids_local):
        full_local[p] = row
        mask_local[p] = 1.0
    gathered = _all_gather_rows(np.stack([full_local, mask_local]), device, collective)     # [world, 2, P, V]
    cost = np.zeros((n_permutations, n_veh))
    owned = np.zeros((n_permutations, n_veh))
    for r in range(gathered.shape[0]):           # every permutation is owned by exactly one rank
        cost += gathered[r, 0] * gathered[r, 1]
        owned += gathered[r, 1]
    if not np.all(owned == 1.0):
        raise ValueError("every permutation must be solved by exactly one rank")
    chosen, solution_cost = solution_costs(cost, belonging_vector)
    return chosen, solution_cost


def gather_winner_plans(plans_local: np.ndarray, perm_ids_local: Sequence[int], n_permutations: int,
                        chosen: np.ndarray, belonging_vector: np.ndarray, device=None, collective: bool = True) -> np.ndarray:
    """plans_local[p, v, :] = flat plan (trims, poses, shapes ...) of vehicle v in local permutation p.
    Returns [n_veh, plan_len]: for every vehicle the plan of its sub-graph's chosen permutation
    (the role of publish_predictions with permutation index 0, :156-162)."""
    belonging_vector = np.asarray(belonging_vector)
    n_veh = belonging_vector.size
    plans_local = np.asarray(plans_local, dtype=np.float64).reshape(len(perm_ids_local), n_veh, -1)
    full = np.zeros((n_permutations, n_veh, plans_local.shape[2]))
    for row, p in zip(plans_local, perm_ids_local):
        full[p] = row
    gathered = _all_gather_rows(full, device, collective).sum(axis=0)        # disjoint ownership: sum == select
    out = np.zeros((n_veh, plans_local.shape[2]))
    for v in range(n_veh):
        out[v] = gathered[chosen[belonging_vector[v] - 1], v]
    return out


def gather_counters(local: Sequence[float], device=None) -> np.ndarray:
    """Config 5 needs only a final gather of per-rank counters (plans, pops, device ms)."""
    return _all_gather_rows(np.asarray(local, dtype=np.float64), device)
