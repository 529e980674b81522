"""Parity comparison between the CUDA path and the CPU oracle (TEST INFRASTRUCTURE).

Bar (BASELINE.json north_star): bit-exact trim sequences, feasibility flags and
expanded-node counts; states and costs within 1e-9 relative.  Because oracle and
device share the arithmetic spec we additionally require the doubles
to be BIT-IDENTICAL and the FNV hash of the pop order to match.
"""
from __future__ import annotations

import numpy as np

REL_TOL = 1e-9   # north_star tolerance for states and costs


def _bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def compare(dev, ref, require_bit_identical: bool = True, skip=()) -> dict:
    """Raises AssertionError on the first violated field; returns summary counts.
    skip: field names left out (pop_hash of launch shape 5 against a fixture that hashed every pop)."""
    n = ref.status.size
    assert dev.status.size == n
    for name in ("status", "is_exhausted", "n_expanded", "n_pops", "pop_hash", "trims", "tree_path",
                 "shape_npts"):
        if name in skip:
            continue
        a, b = getattr(dev, name), getattr(ref, name)
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)[0]
            raise AssertionError(f"{name} differs at {tuple(bad)}: device {a[tuple(bad)]} vs oracle {b[tuple(bad)]}")
    for name in ("y_predicted", "g_path", "h_path", "shape_x", "shape_y"):
        a, b = getattr(dev, name), getattr(ref, name)
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        assert np.array_equal(nan_a, nan_b), f"{name}: NaN pattern differs"
        fa, fb = a[~nan_a], b[~nan_b]
        if fa.size:
            err = np.abs(fa - fb) / np.maximum(np.abs(fb), 1e-300)
            err[(fa == fb)] = 0.0
            assert err.max() <= REL_TOL, f"{name}: max relative error {err.max():.3e} > {REL_TOL}"
        if require_bit_identical and not np.array_equal(_bits(a), _bits(b)):
            bad = np.argwhere(_bits(a) != _bits(b))[0]
            raise AssertionError(f"{name} not bit-identical at {tuple(bad)}: {a[tuple(bad)]!r} vs {b[tuple(bad)]!r}")
    return {"n": int(n), "exhausted": int(ref.is_exhausted.sum()), "pops": int(ref.n_pops.sum()),
            "nodes": int(ref.n_expanded.sum())}
