#!/usr/bin/env python3
"""Experiment: how much of the throughput launch is the tail of its longest searches?
Times the tile shape on all records, on the records below a pop threshold, and the CTA shape on the rest."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from pdmpc_b200 import capi  # noqa: E402
if os.environ.get('PDMPC_LIB'):
    capi.LIB_PATH = os.environ['PDMPC_LIB']
    capi.load_library.__defaults__ = (capi.LIB_PATH,)
from pdmpc_b200.mpa import get_mpa  # noqa: E402
from pdmpc_b200.records import SearchBatch  # noqa: E402


def timed(p, b, variant, runs=3):
    p.set_variant(variant)
    p.stage(b)
    ts = []
    for _ in range(runs):
        p.run_staged()
        p.sync()
        ts.append(p.stats().kernel_ms)
    return min(ts), p.stats().shape


def main():
    path = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    mpa = get_mpa("triple_speed", non_convex=True)
    import glob
    files = sorted(glob.glob(path))
    b = SearchBatch.concat([SearchBatch.load(f) for f in files]) if len(files) > 1 else SearchBatch.load(files[0])
    if reps > 1:
        b = SearchBatch.concat([b] * reps)
    p = capi.Planner(0)
    p.upload_mpa(mpa)
    p.set_variant(2)
    r = p.plan_batch(b)
    pops = r.n_pops.astype(np.int64)
    t_all, sh = timed(p, b, 2)
    print(f"all {b.n}: shape {sh} {t_all:.2f} ms -> {b.n / t_all / 1e3:.3f} M plans/s; max pops {pops.max()}")
    for vo in (True,):
        p.set_cta_queue(vo)
        for esc in [int(v) for v in os.environ.get("PDMPC_ESC_SWEEP", "0,2560").split(",")]:
            for gate in [int(v) for v in os.environ.get("PDMPC_ESC_GATES", "-1").split(",")]:
                p.set_escalation(esc, gate)
                t, sh = timed(p, b, int(os.environ.get("PDMPC_SHAPE", "2")))
                p.fetch()
                st = p.stats()
                print(f"valid-only {vo} escalation {esc} short-list gate {gate}: {t:.2f} ms -> {b.n / t / 1e3:.3f} M plans/s "
                      f"(escalated {st.escalated}, launches {st.kernel_launches})")
    p.set_cta_queue(False)
    p.set_escalation(0)
    for thr in ():
        lo = b.select(np.nonzero(pops <= thr)[0])
        hi = b.select(np.nonzero(pops > thr)[0])
        t_lo, _ = timed(p, lo, 2)
        t_hi5, s5 = timed(p, hi, 5)
        t_hi4, s4 = timed(p, hi, 4)
        print(f"thr {thr}: <= {lo.n} searches {t_lo:.2f} ms | > {hi.n} searches ({pops[pops > thr].sum()} pops) "
              f"shape {s5} {t_hi5:.2f} ms, shape {s4} {t_hi4:.2f} ms | sum {t_lo + t_hi5:.2f} ms "
              f"-> {b.n / (t_lo + t_hi5) / 1e3:.3f} M plans/s")


if __name__ == "__main__":
    main()
