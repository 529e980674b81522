#!/usr/bin/env python3
"""Phase breakdown (PDMPC_PROFILE build) of the single longest search of a record file, run alone."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pdmpc_b200 import capi  # noqa: E402
if os.environ.get('PDMPC_LIB'):
    capi.LIB_PATH = os.environ['PDMPC_LIB']
    capi.load_library.__defaults__ = (capi.LIB_PATH,)
from pdmpc_b200.mpa import get_mpa  # noqa: E402
from pdmpc_b200.records import SearchBatch  # noqa: E402
import numpy as np  # noqa: E402

path = sys.argv[1]
mpa = get_mpa("triple_speed" if "triple" in path else "single_speed", non_convex=True)
b = SearchBatch.load(path)
p = capi.Planner(0)
p.upload_mpa(mpa)
p.set_variant(1)
r = p.plan_batch(b)
order = np.argsort(r.n_pops)[::-1]
variants = [int(v) for v in os.environ.get("PDMPC_VARIANTS", "1,4").split(",")]
for rank in (0, len(order) // 2):
    i = int(order[rank])
    one = b.select([i])
    for variant in variants:
        p.set_variant(variant)
        p.stage(one)
        for _ in range(3):
            p.run_staged()
        p.sync()
        st = p.stats()
        rr = p.fetch()
        assert rr.n_pops[0] == r.n_pops[i] and rr.n_expanded[0] == r.n_expanded[i]
        assert variant == 5 or rr.pop_hash[0] == r.pop_hash[i]
        print(f"variant {variant} search {i}: pops {int(rr.n_pops[0])} nodes {int(rr.n_expanded[0])} exhausted {int(rr.is_exhausted[0])} "
              f"polys {int(one.poly_ptr.size - 1)} verts {int(one.vert_x.size)} lane pts {int(one.lane_x.size)} "
              f"kernel {st.kernel_ms:.3f} ms -> {st.kernel_ms * 1e3 / max(int(rr.n_pops[0]), 1):.2f} us/pop"
              f" (re-run exact: {p.stats().handed_over})")
# one whole time step's worth: the 20 longest searches of the file as one batch
top = b.select(order[:20])
for variant in variants:
    p.set_variant(variant)
    p.stage(top)
    for _ in range(3):
        p.run_staged()
    p.sync()
    print(f"variant {variant}: 20 longest searches as one batch: kernel {p.stats().kernel_ms:.3f} ms")
