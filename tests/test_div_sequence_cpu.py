"""The division sequence of the GPTQ row chain (csrc/gptq.cu: refined_rcp / div_rn_by), emulated with
exact rational arithmetic: with the reciprocal prepared off the chain,

    r0 = rcp.approx(d)  (MUFU.RCP, within 1 ulp);  r = fma(fma(-d, r0, 1), r0, r0)
    q0 = rn(n * r);     q = fma(r, fma(-d, q0, n), q0)

must be the correctly rounded float32 quotient n / d — what NumPy computes on the reference's side
(gptq.py:186, :197) — for every operand pair inside the range the kernel sends down this path
(1e-18 < |n|, |d| < 1e18).  The device side is covered by the bit-exact GPTQ parity tests; this pins
the argument itself, including a reciprocal estimate that is off by an ulp in either direction."""
from fractions import Fraction

import numpy as np

from tests.helpers import stable_seed


def _rn32(x: Fraction) -> Fraction:
    """Round a rational to the nearest float32 (ties to even), normal range."""
    if x == 0:
        return Fraction(0)
    sign = -1 if x < 0 else 1
    a = abs(x)
    e = a.numerator.bit_length() - a.denominator.bit_length()      # 2^(e-1) <= a < 2^(e+1)
    if Fraction(2) ** e > a:
        e -= 1                                                     # now 2^e <= a < 2^(e+1)
    scale = Fraction(2) ** (e - 23)
    m = a / scale                                                  # in [2^23, 2^24)
    lo = m.numerator // m.denominator
    frac = m - lo
    if frac > Fraction(1, 2) or (frac == Fraction(1, 2) and lo % 2 == 1):
        lo += 1
    assert -126 <= e <= 127
    return sign * lo * scale


def _fma(a: Fraction, b: Fraction, c: Fraction) -> Fraction:
    return _rn32(a * b + c)


def _f(x) -> Fraction:
    return Fraction(float(np.float32(x)))


def _quotient(n: Fraction, d: Fraction, ulps: int) -> Fraction:
    r0 = _rn32(1 / d)
    if ulps:                                                       # a MUFU.RCP result one ulp off
        r0 = _f(np.nextafter(np.float32(float(r0)), np.float32(np.inf if ulps > 0 else -np.inf)))
    r = _fma(_fma(-d, r0, Fraction(1)), r0, r0)
    q0 = _rn32(n * r)
    return _fma(r, _fma(-d, q0, n), q0)


def test_three_fma_quotient_is_the_correctly_rounded_division():
    rng = np.random.default_rng(stable_seed("div_rn_by"))
    cases = []
    # magnitudes of the path: weights / scales, quantization errors / diagonal of U, and the range ends
    for lo, hi, count in ((-4.0, 1.0, 3000), (-17.9, 17.9, 3000)):
        n = (10.0 ** rng.uniform(lo, hi, count) * rng.choice([-1.0, 1.0], count)).astype(np.float32)
        d = (10.0 ** rng.uniform(lo, hi, count) * rng.choice([-1.0, 1.0], count)).astype(np.float32)
        cases += list(zip(n, d))
    # quotients next to rounding boundaries: n = rn(d * (k + 1/2 +- tiny)) for small integers k
    d = (10.0 ** rng.uniform(-3, 0, 2000)).astype(np.float32)
    k = rng.integers(-8, 8, 2000).astype(np.float32)
    n = (d * (k + np.float32(0.5))).astype(np.float32)
    n = np.nextafter(n, np.where(rng.random(2000) < 0.5, np.float32(np.inf), np.float32(-np.inf))).astype(np.float32)
    cases += list(zip(n, d))
    bad = 0
    for i, (n_i, d_i) in enumerate(cases):
        if not (1e-18 < abs(float(n_i)) < 1e18 and 1e-18 < abs(float(d_i)) < 1e18):
            continue
        want = _f(np.float32(n_i) / np.float32(d_i))
        got = _quotient(_f(n_i), _f(d_i), ulps=(i % 3) - 1)
        bad += int(got != want)
    assert bad == 0
