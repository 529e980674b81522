import json, os, sys, time
import numpy as np
ROOT = "/root/repo" if os.path.exists("/root/repo/tests") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import straight_iter
from oracle import oracle_py, parity
from pdmpc_b200 import capi
if os.environ.get('PDMPC_LIB'):
    capi.LIB_PATH = os.environ['PDMPC_LIB']
    capi.load_library.__defaults__ = (capi.LIB_PATH,)
NCASE = int(os.environ.get('NCASE', '8'))
from pdmpc_b200.mpa import get_mpa
from pdmpc_b200.records import CHECKER_SAT, SearchBatch
mpa = get_mpa("single_speed", non_convex=False)
p = capi.Planner(0); p.upload_mpa(mpa); CAP = 1 << 23; p.set_node_capacity(CAP)
rng = np.random.default_rng(1)
for mode in os.environ.get("MODES", "cta_valid_only,cta_exact,warp").split(","):
    p.set_variant(1 if mode == "warp" else 0); p.set_cta_queue(mode == "cta_valid_only")
    rng = np.random.default_rng(1)
    out = []
    for j in range(NCASE):
        gap, off = rng.uniform(0.2, 0.6), rng.uniform(0.4, 0.9)
        e = rng.normal(scale=2e-3, size=6)
        a, c = straight_iter(mpa, x=0.0, y=0.0, yaw=0.0), straight_iter(mpa, x=off, y=-gap, yaw=np.pi / 2)
        a.x0[:3] += e[:3]; c.x0[:3] += e[3:]
        b = SearchBatch.from_iters([a, c], mpa.Hp, CHECKER_SAT, mpa.dt_seconds)
        p.joint_plan_batch(b, 2, False)
        r = p.joint_plan_batch(b, 2, False); st = p.stats()
        t0 = time.perf_counter(); ref = oracle_py.joint_plan_batch(mpa, b, 2, max_nodes=CAP, hash_valid_pops_only=(mode == "cta_valid_only")); tc = (time.perf_counter() - t0) * 1e3
        parity.compare(r, ref)
        out.append((int(ref.n_expanded[0]), int(ref.n_pops[0]), round(st.kernel_ms, 2), round(tc, 2), int(st.handed_over)))
    print(mode, "(nodes, pops, kernel ms, oracle ms, rerun):", out)
