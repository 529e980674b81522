#!/usr/bin/env python3
"""Centralized (joint) search timing: J two-vehicle searches (random crossing geometries) in one
pdmpc_joint_plan_batch call, and the first time steps of the 3-vehicle circle; C oracle timed beside it
(one thread) and used as the checker."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import straight_iter  # noqa: E402
from oracle import oracle_py, parity  # noqa: E402
from pdmpc_b200 import capi, scenario  # noqa: E402
from pdmpc_b200.mpa import get_mpa  # noqa: E402
from pdmpc_b200.records import CHECKER_SAT, SearchBatch  # noqa: E402

mpa = get_mpa("single_speed", non_convex=False)
p = capi.Planner(0)
p.upload_mpa(mpa)
CAP = 1 << 23
p.set_node_capacity(CAP)
MODE = sys.argv[1] if len(sys.argv) > 1 else "full_hash"   # full_hash | valid_hash (pdmpc_set_cta_queue 1)
p.set_cta_queue(MODE == "valid_hash")
HV = MODE == "valid_hash"
print("mode", MODE)
rng = np.random.default_rng(1)
for J in (1, 64, 512):
    rows = []
    for _ in range(J):
        gap, off = rng.uniform(0.2, 0.6), rng.uniform(0.4, 0.9)
        rows += [straight_iter(mpa, x=0.0, y=0.0, yaw=0.0), straight_iter(mpa, x=off, y=-gap, yaw=np.pi / 2)]
    b = SearchBatch.from_iters(rows, mpa.Hp, CHECKER_SAT, mpa.dt_seconds)
    p.set_node_capacity(1 << 20 if J > 64 else CAP)
    cap = 1 << 20 if J > 64 else CAP
    p.joint_plan_batch(b, 2, False)
    t0 = time.perf_counter()
    r = p.joint_plan_batch(b, 2, False)
    wall = time.perf_counter() - t0
    st = p.stats()
    t0 = time.perf_counter()
    ref = oracle_py.joint_plan_batch(mpa, b, 2, max_nodes=cap, hash_valid_pops_only=HV)
    t_cpu = time.perf_counter() - t0
    parity.compare(r, ref)
    print(json.dumps({"joint_searches": J, "vehicles": 2, "nodes_per_search": float(ref.n_expanded[::2].mean()),
                      "pops_per_search": float(ref.n_pops[::2].mean()), "kernel_ms": st.kernel_ms, "call_ms": wall * 1e3,
                      "gpu_joint_plans_per_s": J / (st.kernel_ms * 1e-3), "oracle_1thread_joint_plans_per_s": J / t_cpu,
                      "rerun_exact": int(st.handed_over)}))
p.set_node_capacity(CAP)
sc = scenario.circle_scenario(mpa, 3)
run = scenario.CentralizedRunner(sc, None)
iters = [run._iter_for(i) for i in range(3)]
b = SearchBatch.from_iters(iters, mpa.Hp, CHECKER_SAT, mpa.dt_seconds)
p.joint_plan_batch(b, 3, False)
r = p.joint_plan_batch(b, 3, False)
st = p.stats()
t0 = time.perf_counter()
ref = oracle_py.joint_plan_batch(mpa, b, 3, max_nodes=CAP, hash_valid_pops_only=HV)
t_cpu = time.perf_counter() - t0
parity.compare(r, ref)
print(json.dumps({"joint_searches": 1, "vehicles": 3, "nodes": int(ref.n_expanded[0]), "pops": int(ref.n_pops[0]),
                  "kernel_ms": st.kernel_ms, "oracle_1thread_ms": t_cpu * 1e3}))

# --- vehicles a little off their reference lines (no exact left/right mirror ties) and the circle closed loop ----
rows = []
for _ in range(64):
    gap, off = rng.uniform(0.2, 0.6), rng.uniform(0.4, 0.9)
    e = rng.normal(scale=2e-3, size=6)
    a, c = straight_iter(mpa, x=0.0, y=0.0, yaw=0.0), straight_iter(mpa, x=off, y=-gap, yaw=np.pi / 2)
    a.x0[:3] += e[:3]
    c.x0[:3] += e[3:]
    rows += [a, c]
for J in (1, 64):
    b = SearchBatch.from_iters(rows[: 2 * J], mpa.Hp, CHECKER_SAT, mpa.dt_seconds)
    p.joint_plan_batch(b, 2, False)
    r = p.joint_plan_batch(b, 2, False)
    st = p.stats()
    t0 = time.perf_counter()
    ref = oracle_py.joint_plan_batch(mpa, b, 2, max_nodes=CAP, hash_valid_pops_only=HV)
    t_cpu = time.perf_counter() - t0
    parity.compare(r, ref)
    print(json.dumps({"perturbed": True, "joint_searches": J, "vehicles": 2, "nodes_per_search": float(ref.n_expanded[::2].mean()),
                      "pops_per_search": float(ref.n_pops[::2].mean()), "kernel_ms": st.kernel_ms,
                      "oracle_1thread_ms": t_cpu * 1e3, "rerun_exact": int(st.handed_over)}))
for amount, steps in ((2, 10), (3, 4)):
    kms, reruns, cpu_ms = [], [], []

    def jp(b, n):
        r = p.joint_plan_batch(b, n, False)
        st = p.stats()
        kms.append(st.kernel_ms)
        reruns.append(int(st.handed_over))
        t0 = time.perf_counter()
        oracle_py.joint_plan_batch(mpa, b, n, max_nodes=CAP)
        cpu_ms.append((time.perf_counter() - t0) * 1e3)
        return r
    run = scenario.CentralizedRunner(scenario.circle_scenario(mpa, amount), jp)
    run.run(steps)
    print(json.dumps({"closed_loop_circle": amount, "kernel_ms_per_step": [round(x, 2) for x in kms], "rerun_exact": reruns,
                      "oracle_1thread_ms_per_step": [round(x, 2) for x in cpu_ms]}))
