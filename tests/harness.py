"""Shared helpers of the parity tests: netlist fixtures and the parity predicate of SURVEY.md §8(d)."""
import json
import os

import numpy as np

from ftspice_b200 import netlist as nl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "reference_netlists.json")

RTOL = 1e-9   # north_star: node voltages / branch currents within 1e-9 relative ...
ATOL = 1e-12  # ... + 1e-12 absolute (applies to voltage rows too)


_golden = None


def golden() -> dict:
    """The reference's 17 example netlists + the oracle's results (tests/golden/make_golden.py)."""
    global _golden
    if _golden is None:
        with open(GOLDEN_PATH) as f:
            _golden = json.load(f)
    return _golden


def netlist_names():
    return sorted(golden().keys())


def load(name: str) -> nl.Circuit:
    return nl.load(golden()[name]["netlist"])


def fl(rows):
    """repr strings -> float64 array"""
    return np.array([[float(v) for v in r] for r in rows], np.float64) if rows and isinstance(rows[0], list) \
        else np.array([float(v) for v in rows], np.float64)


def assert_close(got, want, what="", spread=None, spread_factor=8.0):
    """|got - want| <= 1e-9*|want| + 1e-12 everywhere.

    `spread` (sparse [[i, j, value], ...] from the golden file) lists the entries where the ORACLE's own result
    moves by more than a tenth of that tolerance when its arithmetic changes in the last bit (FMA contraction,
    another libm): ill-conditioned unknowns that the reference cannot reproduce to 1e-9 across platforms either.
    There the allowance is widened by spread_factor x the oracle's own spread.  Returns how many values needed it.
    """
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    both_nan = np.isnan(got) & np.isnan(want)
    err = np.abs(got - want)
    tol = RTOL * np.abs(want) + ATOL
    strict_bad = ~(err <= tol) & ~both_nan
    if spread:
        tol = tol.copy()
        for i, j, v in spread:
            if got.ndim == 2 and i < got.shape[0]:
                tol[i, j] += spread_factor * float(v)
    bad = ~(err <= tol) & ~both_nan
    if bad.any():
        idx = np.argwhere(bad)[0]
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} values differ; first at {tuple(idx)}: "
                             f"got {got[tuple(idx)]!r} want {want[tuple(idx)]!r} (|d|={err[tuple(idx)]:.3e})")
    return int(strict_bad.sum())
