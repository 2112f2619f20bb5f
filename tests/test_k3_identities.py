"""Arithmetic identities k3_rdp (aruco3_b200/csrc/k3_contours.cu) relies on, checked on the CPU in the same f64 arithmetic.

The reference splits a span iff `dmax > eps` with dmax = numerator / sqrt(a^2 + b^2) in f64 (imageproc's
approximate_polygon_dp; oracle/a3ref.c).  For frames below 2^14 pixels a side the kernel decides that from the
squares — best^2 against eps^2 (a^2 + b^2) — whenever the two sides differ by more than 2^-40 relatively, and only
otherwise takes the square root and the division; and it takes the first index of the largest numerator as the
reference's "first strict maximum of the distance", which needs different numerators to give different quotients."""
import numpy as np

HI, LO = 1.0000000000009095, 0.9999999999990905  # 1 + 2^-40, 1 - 2^-40 as written in the kernel


def _decide(best, a, b, eps):
    """The kernel's fast decision: True / False when decided, None when it falls back to the exact expression."""
    d2 = np.float64(a * a + b * b)
    lhs = np.float64(best) * np.float64(best)
    rhs = (np.float64(eps) * np.float64(eps)) * d2
    if lhs > rhs * HI:
        return True
    if lhs < rhs * LO:
        return False
    return None


def _exact(best, a, b, eps):
    den = np.sqrt(np.float64(a) * np.float64(a) + np.float64(b) * np.float64(b))
    with np.errstate(divide="ignore"):
        return bool(np.float64(best) / den > np.float64(eps))


def test_constants():
    assert HI == 1.0 + 2.0 ** -40 and LO == 1.0 - 2.0 ** -40


def test_squared_comparison_never_contradicts_the_quotient():
    rng = np.random.default_rng(3)
    undecided = 0
    cases = 0
    for _ in range(20000):
        a, b = (int(v) for v in rng.integers(-16383, 16384, size=2))
        if a == 0 and b == 0:
            continue
        n = int(rng.integers(4, 200000))
        eps = np.float64(n) * np.float64(0.05)
        den = float(np.hypot(a, b))
        # numerators around the threshold (where a wrong decision could hide) and anywhere
        around = int(round(float(eps) * den))
        for best in {max(1, around + d) for d in (-2, -1, 0, 1, 2)} | {int(rng.integers(1, 1 << 29))}:
            got = _decide(best, a, b, eps)
            cases += 1
            if got is None:
                undecided += 1
            else:
                assert got == _exact(best, a, b, eps), (best, a, b, n)
    assert undecided < cases // 100  # the exact path is the rare one


def test_exactly_representable_ties_fall_back():
    # 3-4-5 triangles: den is exact, so best = eps * den makes both sides equal and the fast test must not decide
    for k in (1, 7, 100, 3000):
        a, b = 3 * k, 4 * k
        for n in (20, 100, 4000):
            eps = np.float64(n) * np.float64(0.05)
            best = float(eps) * 5 * k
            if best == int(best) and best >= 1:
                assert _decide(int(best), a, b, eps) is None
                assert _exact(int(best), a, b, eps) is False  # dmax == eps is not a split


def test_different_numerators_give_different_quotients():
    rng = np.random.default_rng(4)
    for _ in range(20000):
        a, b = (int(v) for v in rng.integers(-16383, 16384, size=2))
        if a == 0 and b == 0:
            continue
        den = np.sqrt(np.float64(a * a + b * b))
        t = int(rng.integers(2, 1 << 34))
        assert np.float64(t - 1) / den != np.float64(t) / den
