#!/usr/bin/env python3
"""Small CTA-shape workload for compute-sanitizer (racecheck / memcheck / synccheck): one road time step of the
fixtures through pdmpc_plan_timestep (DEPS instance, exact and valid-only queue), the same searches level-free through
pdmpc_plan_batch (non-DEPS instance, one master per CTA), and 160 searches at once (several masters per CTA)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from helpers import load_golden_timesteps  # noqa: E402
from oracle import parity  # noqa: E402
from pdmpc_b200 import capi  # noqa: E402
from pdmpc_b200.records import SearchBatch, TimestepDeps  # noqa: E402

mpa, steps = load_golden_timesteps("timestep_road_triple_speed")
p = capi.Planner(0)
p.upload_mpa(mpa)
batch, deps, exp = steps[2]
for vo in (False, True):
    p.set_cta_queue(vo)
    p.set_variant(4)
    parity.compare(p.plan_timestep(batch, deps, False), exp, skip=("pop_hash",) if vo else ())
    r = p.plan_batch(batch, False)          # non-DEPS instance, one master per CTA (no predecessors' areas: other answers)
    assert r.status.max() == 0
    big = SearchBatch.concat([s[0] for s in steps])
    r = p.plan_batch(big, False)            # 160 searches > #SMs: several masters per CTA
    assert r.status.max() == 0
    bd = TimestepDeps.concat([s[1] for s in steps], [s[0].n for s in steps])
    r = p.plan_timestep(big, bd, False)     # DEPS instance with several masters per CTA
    assert r.status.max() == 0
print("sanitize_cta: done", int(r.n_pops.sum()), "pops")
