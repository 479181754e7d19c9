"""Shared helpers for the parity tests."""
import math
from fractions import Fraction

import numpy as np

VARIANTS_SUM = [(0, False), (2, False), (3, False), (4, False), (5, False), (6, False), (7, False), (8, False),
                (4, True), (6, True), (8, True)]
VARIANTS_DOT = [(0, False), (3, False), (4, False), (8, False), (4, True), (6, True), (8, True)]


def cpu41_to_gpu39(limbs41):
    from oracle.oracle import cpu41_to_gpu39 as f
    return f(limbs41)


def exact_sum_fraction(a):
    return sum((Fraction(float(x)) for x in a), Fraction(0))


def exact_dot_fraction(a, b):
    return sum((Fraction(float(x)) * Fraction(float(y)) for x, y in zip(a, b)), Fraction(0))


def round_fraction(fr: Fraction) -> float:
    """correctly rounded (nearest-even) double of an exact rational"""
    if fr == 0:
        return 0.0
    return float(fr)          # CPython's Fraction.__float__ -> int/int true division is correctly rounded


def limbs_from_fraction(fr: Fraction, nlimbs=39, fwords=20):
    """normal form of the exact value, python ints"""
    t = fr * (1 << (52 * fwords))
    assert t.denominator == 1, "value has bits below the accumulator LSB"
    t = t.numerator
    out = []
    for _ in range(nlimbs - 1):
        out.append(t & ((1 << 52) - 1))
        t >>= 52
    out.append(t)
    return np.array(out, dtype=np.int64)


def ref_round_model(limbs, fwords=20):
    """executable model of the reference Round() on normalised limbs."""
    D, MASK = 52, (1 << 52) - 1
    acc = [int(v) for v in limbs]
    nl = len(acc)
    negative = acc[-1] < 0
    i = nl - 1
    while i >= 0 and acc[i] == 0:
        i -= 1
    if negative:
        while i >= 0 and (acc[i] & MASK) == MASK:
            i -= 1
    if i < 0:
        return 0.0
    hiword = (MASK - acc[i]) if negative else acc[i]
    rounded = float(hiword)
    hi = math.ldexp(rounded, (i - fwords) * D)
    if i == 0:
        return -hi if negative else hi
    hiword -= int(rounded)
    mid = math.ldexp(float(hiword), (i - fwords) * D)
    sticky = 0
    for j in range(0, i - 1):
        sticky |= ((1 << D) - acc[j]) if negative else acc[j]
    loword = ((1 << D) - acc[i - 1]) if negative else acc[i - 1]
    loword |= 1 if sticky else 0
    lo = math.ldexp(float(loword), (i - 1 - fwords) * D)
    assert mid == 0
    hi = hi + lo
    return -hi if negative else hi


def same_double(x, y):
    """bit equality (+0.0 and -0.0 differ), with NaN == NaN"""
    if isinstance(x, float) and isinstance(y, float) and math.isnan(x) and math.isnan(y):
        return True
    return bool(np.float64(x).view(np.uint64) == np.float64(y).view(np.uint64))
