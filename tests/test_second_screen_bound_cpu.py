"""The second screening stage of the UCB arg-max (search_kernel.cuh: exact_ucb_argmax with try_fast) decides a close
call from scores computed as (Q - lo) * rcp(range) + 2c * sqrt(L) * rsqrt(n), where rcp / rsqrt are the hardware's
approximations (relative error <= 2^-22, PTX ISA: rcp.approx.ftz.f64, rsqrt.approx.ftz.f64) refined by two Newton
steps, and only when the best score leads by more than 1e-11 * (max(|norm| + |bonus|) + 1). The kernel's comment claims
|score' - score| <= 1e-14 * (|norm| + |bonus|) against the reference's correctly rounded evaluation
(norm + 2c * sqrt(2 ln N / n), uct_statistic.h:223-244). Checked here numerically, with the approximations' seeds placed
at the edges of their documented error band, in the arithmetic the device uses (float64, fused multiply-adds emulated
in higher precision)."""
import math
from fractions import Fraction

import numpy as np


def _fma(a, b, c):
    # one rounding, like __fma_rn: exact product and sum in rationals, then to the nearest double
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _fast_rcp(x, seed_err):
    y = (1.0 / x) * (1.0 + seed_err)
    for _ in range(2):
        e = _fma(-x, y, 1.0)
        y = _fma(y, e, y)
    return y


def _fast_rsqrt(x, seed_err):
    y = (1.0 / math.sqrt(x)) * (1.0 + seed_err)
    h = 0.5 * x
    for _ in range(2):
        e = _fma(-(h * y), y, 0.5)
        y = _fma(y, e, y)
    return y


def test_newton_refined_reciprocals_reach_2_pow_minus_50():
    rng = np.random.default_rng(11)
    worst_rcp = worst_rsq = 0.0
    for _ in range(4000):
        x = float(rng.uniform(1e-6, 1e6)) if rng.random() < 0.5 else float(rng.integers(1, 70000))
        for s in (2.0 ** -22, -(2.0 ** -22)):
            worst_rcp = max(worst_rcp, abs(Fraction(_fast_rcp(x, s)) * Fraction(x) - 1))
            r = _fast_rsqrt(x, s)
            # relative error of r against 1/sqrt(x): r^2 * x = (1 + e)^2
            worst_rsq = max(worst_rsq, abs(float(Fraction(r) ** 2 * Fraction(x)) - 1.0) / 2)
    assert worst_rcp <= 2.0 ** -50, float(worst_rcp)
    assert worst_rsq <= 2.0 ** -50, worst_rsq


def test_fast_scores_stay_within_the_claimed_bound_of_the_reference_scores():
    rng = np.random.default_rng(12)
    worst = 0.0
    for _ in range(3000):
        lo = float(rng.uniform(-1000.0, 0.0))
        rng_w = float(rng.uniform(1e-3, 1100.0))
        q = lo + float(rng.uniform(0.0, 1.0)) * rng_w
        N = int(rng.integers(3, 60000))
        n = int(rng.integers(1, N + 1))
        two_c = float(rng.choice([0.01, 0.7, 1.4, 10.0, 1000.0]))
        log2N = 2.0 * math.log(float(N))
        # the reference: five correctly rounded operations
        norm = (q - lo) / rng_w
        ref = norm + two_c * math.sqrt(log2N / float(n))
        for s1 in (2.0 ** -22, -(2.0 ** -22)):
            for s2 in (2.0 ** -22, -(2.0 ** -22)):
                ir = _fast_rcp(rng_w, s1)
                sq = two_c * (log2N * _fast_rsqrt(log2N, s2))
                nrm = (q - lo) * ir
                bonus = sq * _fast_rsqrt(float(n), -s2)
                fast = nrm + bonus
                bound = 1e-14 * (abs(nrm) + abs(bonus))
                assert abs(fast - ref) <= bound, (fast, ref, bound, q, lo, rng_w, N, n, two_c)
                worst = max(worst, abs(fast - ref) / (abs(nrm) + abs(bonus)))
    # and the decision margin (1e-11) is three orders of magnitude above what was observed
    assert worst < 1e-14
