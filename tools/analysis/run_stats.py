"""Offline search statistics on a recorded batch (design aid). usage: run_stats.py batch.npz [first count [K]]"""
import ctypes as C, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pdmpc_b200 import capi
from pdmpc_b200.mpa import get_mpa
from pdmpc_b200.records import SearchBatch
b = SearchBatch.load(sys.argv[1])
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
count = int(sys.argv[3]) if len(sys.argv) > 3 else b.n
K = int(sys.argv[4]) if len(sys.argv) > 4 else 32
mpa = get_mpa("triple_speed", non_convex=True)
d, keep = capi.mpa_desc(mpa)
bi = capi.batch_in(b)
L = C.CDLL(os.path.join(ROOT, "build", "libstats.so"))
L.stats_set_front(K)
out = np.zeros(64, dtype=np.int64)
n = L.stats_batch(C.byref(d), C.byref(bi), first, count, out.ctypes.data_as(C.POINTER(C.c_int64)), 64)
names = ["pops", "valid_pops", "goal_pops", "expansions", "nodes", "ties"] + [f"qlen2^{i}" for i in range(16)] + \
    ["segs", "segs_bbox", "segs_c2row", "pairs_hit", "checks", "checks_any_bbox", "front_pops", "back_pops",
     "back_pushes", "front_inserts", "evictions", "child_is_next", "invalid_children", "children"]
for k, v in zip(names, out[:n]):
    print(f"{k:18s} {v}")
