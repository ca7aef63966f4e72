"""The branch-free division / square root of csrc/gms_common.cuh (gms_div_rn_normal, gms_sqrt_rn_normal: used by k_adam_sh,
the expansion and the preprocess kernels) restated in EXACT rational arithmetic: starting from any reciprocal / reciprocal
square root seed within the error bound of the hardware approximation (rcp.approx: 1 ulp; rsqrt.approx: 2^-22.9), the
Newton-corrected sequences return the correctly rounded IEEE result for normal operands -- which is why they can replace
`/` and sqrtf bit for bit (the -m gpu test test_adam_sh_factored_abi_survives_denormal_second_moments checks the same on
the device: 0 of 196 752 elements differ from the sqrtf / division build)."""
import math
from fractions import Fraction

import numpy as np


def _round_f32(x: Fraction) -> Fraction:
    """Round an exact rational to the nearest float32 (ties to even), normal range only."""
    if x == 0:
        return Fraction(0)
    s = -1 if x < 0 else 1
    a = abs(x)
    e = math.floor(math.log2(a)) if a.numerator.bit_length() - a.denominator.bit_length() < 1000 else 0
    while Fraction(2) ** e > a:
        e -= 1
    while Fraction(2) ** (e + 1) <= a:
        e += 1
    q = Fraction(2) ** (e - 23)                     # ulp
    n, r = divmod(a, q)
    n = int(n)
    if r * 2 > q or (r * 2 == q and n % 2 == 1):
        n += 1
    return s * n * q


def _f(x) -> Fraction:
    return Fraction(float(np.float32(x)))


def _fma(a, b, c):
    return _round_f32(a * b + c)


def _ulp_step(x: Fraction, k: int) -> Fraction:
    v = np.float32(float(x))
    return Fraction(float((v.view(np.int32) + np.int32(k)).view(np.float32)))


def _div_rn_normal(n, d, seed_err):
    r = _ulp_step(_round_f32(1 / d), seed_err)      # rcp.approx.ftz within `seed_err` ulp
    r = _fma(_fma(-d, r, Fraction(1)), r, r)
    q = _round_f32(n * r)
    return _fma(_fma(-d, q, n), r, q)


def _sqrt_rn_normal(x, seed_err):
    exact_rsqrt = Fraction(1 / math.sqrt(float(x)))              # (double precision: far inside a float32 ulp)
    y = _ulp_step(_round_f32(exact_rsqrt), seed_err)
    s = _round_f32(x * y)
    h = _round_f32(Fraction(1, 2) * y)
    return _fma(_fma(-s, s, x), h, s)


def _sqrt_correct(x: Fraction) -> Fraction:
    """Correctly rounded float32 square root from integer arithmetic."""
    e = math.floor(math.log2(float(x)))
    e -= e % 2                                       # x = m * 2^e with m in [1, 4)
    m = x / Fraction(2) ** e
    scale = 1 << 120
    root = math.isqrt(int(m * scale))                # floor(sqrt(m) * 2^60)
    lo = _round_f32(Fraction(root, 1 << 60) * Fraction(2) ** (e // 2))
    # the 2^-60 truncation cannot cross a float32 rounding boundary except at exact ties, which sqrt never produces
    return lo


def test_division_sequence_is_correctly_rounded_for_any_admissible_seed():
    rng = np.random.default_rng(3)
    bad = 0
    for _ in range(4000):
        d = _f(np.exp(rng.uniform(np.log(1e-15), np.log(1e4))))
        n = _f(np.exp(rng.uniform(np.log(1e-20), np.log(1e4))) * rng.choice([-1.0, 1.0]))
        ref = _round_f32(n / d)
        if abs(ref) < Fraction(2) ** -120:
            continue
        for k in (-1, 0, 1):
            bad += _div_rn_normal(n, d, k) != ref
    assert bad == 0


def test_sqrt_sequence_is_correctly_rounded_for_any_admissible_seed():
    rng = np.random.default_rng(4)
    bad = 0
    for _ in range(4000):
        x = _f(np.exp(rng.uniform(np.log(1e-30), np.log(1e8))))
        ref = _sqrt_correct(x)
        for k in (-2, -1, 0, 1, 2):                  # rsqrt.approx: maximum relative error 2^-22.9
            bad += _sqrt_rn_normal(x, k) != ref
    assert bad == 0
