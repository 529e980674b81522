#!/usr/bin/env python3
"""Expected outputs of the sampled optimizer (MonteCarloTreeSearch.m) on the committed
fixtures' search records -> tests/golden/mcts_expected.npz.

Like tools/make_golden.py: the C oracle's outputs are only written after the matrix-form
restatement (oracle/matlab_literal.do_mcts: numpy's own MT19937, the reference's
unmodified priority-queue source, MATLAB-style `children` bookkeeping) reproduces every
search — roll-out trace, counts, flags, trims, bit-identical poses and shapes.
Seeds follow the reference's rule time_step + vehicle_index (:31), here a fixed pattern.

Run here (build container):  python tools/make_golden_mcts.py
"""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle_py  # noqa: E402
from helpers import GOLDEN_CASES, load_golden  # noqa: E402
import test_mcts_cpu as T  # noqa: E402

N_MAX = {"circle_sat_single_speed": 40, "road_interx_single_speed": 120, "road_interx_triple_speed": 250}


def main():
    d = {}
    for name in GOLDEN_CASES:
        mpa, batch, _ = load_golden(name)
        seeds = (2 + (np.arange(batch.n) * 7) % 53).astype(np.uint32)
        got = T._compare_with_literal(mpa, batch, range(batch.n), seeds, N_MAX[name])
        d[name + "__seeds"] = seeds
        d[name + "__n_max"] = np.int32(N_MAX[name])
        for f in dataclasses.fields(got):
            d[name + "__out__" + f.name] = np.asarray(getattr(got, f.name))
        print(name, batch.n, "searches, exhausted", int(got.is_exhausted.sum()), "expansions", int(got.n_expanded.sum()))
    out = os.path.join(ROOT, "tests", "golden", "mcts_expected.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
