"""Input generators: host-side mirror of the reference's include/common.hpp / src/common/common.cpp.

Same names and distributions as the reference (`init_naive` common.cpp:147, `init_fpuniform`
common.cpp:30, `init_ill_cond` common.cpp:113, `init_lognormal` common.cpp:66) plus `gen_dot`
(Ogita-Rump-Oishi Alg. 6.1, which the reference cites in common.hpp:140-150 but only half
implements).  The reference draws from libc rand() / std::random_device, which is neither portable
nor reproducible; here every generator is a pure function of (seed, element index) built on a
counter-based integer hash (splitmix64), so that

* numpy (host) and torch (device) produce bit-identical vectors from the same seed, and
* any slice [lo, hi) of a 2^30-element vector can be regenerated on its own (CPU oracle on a slice,
  per-rank shards in the multi-GPU bench) without materialising the rest.

Doubles are assembled from integer fields (sign | exponent | mantissa), never through floating
point arithmetic that could differ between libraries.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np

_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB


# ---- splitmix64 on numpy uint64 / torch int64 -----------------------------------------------------
def _mix_np(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(_GOLD)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
        return z ^ (z >> np.uint64(31))


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr_t(z, k: int):
    """logical shift right on torch int64"""
    return (z >> k) & ((1 << (64 - k)) - 1)


def _mix_t(z):
    z = z + _s64(_GOLD)
    z = (z ^ _lsr_t(z, 30)) * _s64(_C1)
    z = (z ^ _lsr_t(z, 27)) * _s64(_C2)
    return z ^ _lsr_t(z, 31)


def _stream_key(seed: int, stream: int) -> int:
    return (seed * 0xD1342543DE82EF95 + stream * 0xA0761D6478BD642F + 0x2545F4914F6CDD1D) & _M64


def _bits(seed: int, stream: int, lo: int, hi: int, device=None):
    """64 random bits per index in [lo, hi): numpy uint64 (device None) or torch int64."""
    key = _stream_key(seed, stream)
    if device is None:
        idx = np.arange(lo, hi, dtype=np.uint64)
        with np.errstate(over="ignore"):
            return _mix_np(_mix_np(idx ^ np.uint64(key)))
    import torch
    idx = torch.arange(lo, hi, dtype=torch.int64, device=device)
    return _mix_t(_mix_t(idx ^ _s64(key)))


def _assemble(sign, efield, mant, device):
    """sign (0/1), biased exponent field, 52-bit mantissa -> float64 array/tensor."""
    if device is None:
        bits = (sign.astype(np.uint64) << np.uint64(63)) | (efield.astype(np.uint64) << np.uint64(52)) | mant
        return bits.view(np.float64)
    import torch
    bits = (sign << 63) | (efield << 52) | mant
    return bits.view(torch.float64)


def _mod(r, m: int, device):
    """non-negative r mod m for uint64 numpy / int64 torch (r already reduced to < 2^62)."""
    if device is None:
        return r % np.uint64(m)
    return r % m


def _u62(r, device):
    if device is None:
        return r >> np.uint64(2)
    return _lsr_t(r, 2)


def _mant(r, device):
    if device is None:
        return r & np.uint64((1 << 52) - 1)
    return r & ((1 << 52) - 1)


# ---- the reference's generators ---------------------------------------------------------------------
def init_naive(n: int, lo: int = 0, hi: int | None = None, device=None):
    """common.cpp:147-150: every element is 1.1"""
    hi = n if hi is None else hi
    if device is None:
        return np.full(hi - lo, 1.1, dtype=np.float64)
    import torch
    return torch.full((hi - lo,), 1.1, dtype=torch.float64, device=device)


def init_fpuniform(n: int, range_: int, emax: int, seed: int = 1, neg_ratio: int = 1, lo: int = 0,
                   hi: int | None = None, device=None):
    """common.cpp:18-33 randDouble/init_fpuniform: mantissa uniform in [1,2), exponent uniform integer
    in [emax-range, emax), all positive when neg_ratio <= 1 (the reference always passes 1), else
    negative with probability 1/neg_ratio.  BASELINE's "log-uniform 1e-100..1e100" is
    init_fpuniform(n, 664, 332)."""
    hi = n if hi is None else hi
    r1 = _bits(seed, 1, lo, hi, device)
    r2 = _bits(seed, 2, lo, hi, device)
    mant = _mant(r1, device)
    e = _mod(_u62(r2, device), range_, device) + (emax - range_ + 1023)
    if neg_ratio > 1:
        r3 = _bits(seed, 3, lo, hi, device)
        sign = (_mod(_u62(r3, device), neg_ratio, device) == 0)
        sign = sign.astype(np.uint64) if device is None else sign.long()
    else:
        sign = (mant * 0) if device is None else mant * 0
    return _assemble(sign, e, mant, device)


def _pow2(e, device):
    """exact 2^e (|e| < 1022) assembled from the exponent field"""
    if device is None:
        return ((e + 1023).astype(np.uint64) << np.uint64(52)).view(np.float64)
    import torch
    return ((e + 1023) << 52).view(torch.float64)


def init_ill_cond(n: int, c: float, seed: int = 1, lo: int = 0, hi: int | None = None, device=None):
    """common.cpp:113-145 init_ill_cond(n, a, c) with c taken as a double: elements (2x-1)*2^e with
    x uniform in [0,1); first half: e = round(U*b/2) with e[0] = round(b/2)+1; second half: e rising
    linearly from 0 to b/2, e[n-1] = 0; b = log2(c).  (The x vector of Ogita-Rump-Oishi Alg. 6.1.)
    (2x-1) is a signed 53-bit integer times 2^-52, so every element is assembled exactly."""
    hi = n if hi is None else hi
    b = math.log2(c)
    n2 = n // 2
    r1 = _bits(seed, 11, lo, hi, device)
    r2 = _bits(seed, 12, lo, hi, device)
    step = (b / 2) / max(n - n2, 1)
    if device is None:
        idx = np.arange(lo, hi, dtype=np.int64)
        m = (r1 >> np.uint64(11)).astype(np.int64) - (1 << 52)          # uniform integer in [-2^52, 2^52)
        u = (r2 >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        e_first = np.rint(u * (b / 2)).astype(np.int64)
        e_first = np.where(idx == 0, int(round(b / 2)) + 1, e_first)
        e_second = np.floor(step * (idx - n2).astype(np.float64)).astype(np.int64)
        e = np.where(idx < n2, e_first, e_second)
        e = np.where(idx == n - 1, 0, e)
        return m.astype(np.float64) * _pow2(e - 52, None)
    import torch
    idx = torch.arange(lo, hi, dtype=torch.int64, device=device)
    m = _lsr_t(r1, 11) - (1 << 52)
    u = _lsr_t(r2, 11).double() * (1.0 / (1 << 53))
    e_first = torch.round(u * (b / 2)).long()
    e_first = torch.where(idx == 0, torch.full_like(idx, int(round(b / 2)) + 1), e_first)
    e_second = torch.floor(step * (idx - n2).double()).long()
    e = torch.where(idx < n2, e_first, e_second)
    e = torch.where(idx == n - 1, torch.zeros_like(e), e)
    return m.double() * _pow2(e - 52, device)


def init_lognormal(n: int, mean: float, stddev: float, seed: int = 1):
    """common.cpp:66-73 (host only; the reference seeds from std::random_device, i.e. it is not
    reproducible -- here it is)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.lognormal(mean, stddev, n)


def gen_dot(n: int, c: float, seed: int = 1):
    """Ogita, Rump, Oishi, "Accurate sum and dot product", SIAM J. Sci. Comput. 26(6), 2005,
    Algorithm 6.1 (GenDot): vectors x, y of length n whose dot product has condition number ~ c.
    Host only (uses exact rational arithmetic for the second half).  Returns (x, y, exact dot as a
    Fraction, achieved condition number)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n2 = (n + 1) // 2
    b = math.log2(c)
    x = np.zeros(n)
    y = np.zeros(n)
    e = np.rint(rng.random(n2) * b / 2)
    e[0] = round(b / 2) + 1
    e[-1] = 0
    x[:n2] = (2 * rng.random(n2) - 1) * np.exp2(e)
    y[:n2] = (2 * rng.random(n2) - 1) * np.exp2(e)
    acc = sum(Fraction(float(x[i])) * Fraction(float(y[i])) for i in range(n2))
    e2 = np.rint(np.linspace(b / 2, 0, n - n2))
    for k, i in enumerate(range(n2, n)):
        x[i] = (2 * rng.random() - 1) * 2.0 ** e2[k]
        y[i] = ((2 * rng.random() - 1) * 2.0 ** e2[k] - float(acc)) / x[i]
        acc += Fraction(float(x[i])) * Fraction(float(y[i]))
    perm = rng.permutation(n)
    x, y = x[perm], y[perm]
    absdot = sum(abs(Fraction(float(a)) * Fraction(float(bb))) for a, bb in zip(x, y))
    cond = float(2 * absdot / abs(acc)) if acc != 0 else math.inf
    return x, y, acc, cond


def cancelling_pair(n: int, kind: str = "dot", c: float = 1e32, seed: int = 1, residual: float = 1.5, device=None):
    """Full-size ill-conditioned inputs with a KNOWN exact answer (size-independent parity property).

    First half: init_ill_cond values; second half: the same magnitudes at permuted positions with the
    sign of one factor flipped, so every term cancels exactly; one term is replaced by `residual`.
    kind == "sum": returns a with exsum(a) == residual exactly.
    kind == "dot": returns (a, b) with exdot(a, b) == residual exactly and condition number
    sum|a_i b_i| / |residual| >> c.
    Works for numpy (device None) and torch; n must be even."""
    assert n % 2 == 0 and n >= 4
    h = n // 2
    if device is None:
        cat, arange = np.concatenate, (lambda k: np.arange(k, dtype=np.int64))
    else:
        import torch
        cat = torch.cat
        arange = lambda k: torch.arange(k, dtype=torch.int64, device=device)  # noqa: E731
    perm = (arange(h) * 0x9E3779B1 + 12345) % h if (h & (h - 1)) == 0 else (h - 1 - arange(h))
    a1 = init_ill_cond(h, c, seed=seed, device=device)
    if kind == "sum":
        a = cat([a1, -a1[perm]])
        a[0] = residual
        inv0 = int((perm == 0).nonzero()[0][0]) if device is None else int((perm == 0).nonzero()[0, 0])
        a[h + inv0] = 0.0
        return a
    b1 = init_ill_cond(h, c, seed=seed + 1000, device=device)
    a = cat([a1, a1[perm]])
    b = cat([b1, -b1[perm]])
    inv0 = int((perm == 0).nonzero()[0][0]) if device is None else int((perm == 0).nonzero()[0, 0])
    a[0] = 1.0
    b[0] = residual
    a[h + inv0] = 0.0
    return a, b
