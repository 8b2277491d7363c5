"""The search kernels select, for an agent whose expanded actions all carry the same action value, the
least-visited action with integer compares only (select_action in mamcts_b200/csrc/search_kernel.cuh). The
claim behind it - the FIRST maximum of the reference's FP64 UCB scores (uct_statistic.h:223-244, strict >,
initial maximum DBL_MIN) is the FIRST minimum of the counts whenever N >= 2, 0.01 <= 2c <= 1e6, n_a <= 2^20 and bound
estimation is on - is checked here numerically with the same IEEE operations (numpy float64 division and
square root are correctly rounded, math.log is the host libm the kernels' log table comes from)."""
from __future__ import annotations

import math
import sys

import numpy as np


def _reference_first_max(q: float, counts, N: int, two_c: float) -> int:
    # calculate_bounds_based_on_stat with equal values (:122-132): (0, max) or (0, 1)
    lo, hi = 0.0, (q if q != 0.0 else 1.0)
    rng = np.float64(hi) - np.float64(lo)
    log2n = np.float64(2.0) * np.float64(math.log(float(N)))
    best, max_value = 0, np.float64(sys.float_info.min)
    for a, n in enumerate(counts):
        norm = (np.float64(q) - np.float64(lo)) / rng
        v = norm + np.float64(two_c) * np.sqrt(log2n / np.float64(n))
        if v > max_value:
            max_value, best = v, a
    return best


def _first_min(counts) -> int:
    return int(np.argmin(np.asarray(counts)))   # numpy returns the first minimum


def test_first_maximum_of_ucb_is_first_minimum_of_counts():
    rng = np.random.default_rng(12345)
    checked = 0
    for _ in range(20000):
        k = int(rng.integers(1, 9))
        regime = int(rng.integers(0, 4))
        if regime == 0:      # small counts, many ties
            counts = rng.integers(1, 4, size=k)
        elif regime == 1:    # adjacent large counts (the hardest case for the proof)
            base = int(rng.integers(1, (1 << 20) - 8))
            counts = base + rng.integers(0, 3, size=k)
        elif regime == 2:    # wide range
            counts = rng.integers(1, 1 << 20, size=k)
        else:                # near the upper limit
            counts = (1 << 20) - rng.integers(0, 4, size=k)
        counts = [int(c) for c in counts]
        N = max(2, min(1 << 20, int(sum(counts) + rng.integers(0, 3))))
        q = float(rng.choice([0.0, -0.0, 1.0, -1000.0, 100.0, 0.4601886772625664, -3.15e-7]))
        two_c = float(rng.choice([1.4, 0.01, 0.7, 20.0, 1e6]))
        assert _reference_first_max(q, counts, N, two_c) == _first_min(counts), (q, counts, N, two_c)
        checked += 1
    assert checked == 20000


def test_the_gate_is_needed():
    # N = 1: 2 ln N = 0, every score is equal -> the reference keeps the FIRST action, not the least visited
    assert _reference_first_max(0.0, [5, 1], 1, 1.4) == 0 and _first_min([5, 1]) == 1


def test_the_upper_gate_on_the_exploration_constant_is_needed():
    # a huge exploration constant overflows every bonus to +inf: all scores tie and the reference's strict >
    # keeps the FIRST expanded action - the shortcut (2c <= 1e6 in select_action) must not apply
    with np.errstate(over="ignore"):
        assert _reference_first_max(0.0, [5, 1], 1000, 1.5e308) == 0 and _first_min([5, 1]) == 1
    # at the gate itself (and N = 2, the smallest N of the shortcut) the claim still holds
    assert _reference_first_max(0.0, [2, 1], 2, 1e6) == 1
