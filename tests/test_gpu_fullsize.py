"""GPU tests at BASELINE.json's full sizes (2^30 elements): size-independent properties, because
the scalar oracle cannot run 2^30 in seconds.

 * closed form: exsum of n copies of 1.1 == correctly rounded n * double(1.1) (exact rational);
 * known answer: cancelling_pair() vectors sum / dot to exactly 1.5 however large the partial sums;
 * partition additivity: limbs(whole) == integer sum of limbs(shards) -- the multi-GPU merge;
 * every FPE / early-exit variant returns the same limbs;
 * an oracle check on a slice regenerated from the counter-based generator.
"""
import math
from fractions import Fraction

import numpy as np
import pytest

from helpers import limbs_from_fraction, same_double
from exblas_b200 import common as cm

pytestmark = pytest.mark.gpu

N = 1 << 30


def _sum(gpu, t, fpe, ee, rm=0, n=None, off=0):
    gpu.exsum_async(t.numel() - off if n is None else n, t, 1, off, fpe, ee, rm)
    return gpu.fetch()


def test_naive_closed_form_2p30(gpu):
    import torch
    a = torch.full((N,), 1.1, dtype=torch.float64, device="cuda")
    exact = Fraction(1.1) * N
    want_limbs = limbs_from_fraction(exact)
    for fpe, ee in [(0, False), (3, False), (4, False), (8, False), (8, True)]:
        v, l, st = _sum(gpu, a, fpe, ee, rm=1)
        assert st == 0 and (l == want_limbs).all(), (fpe, ee)
        assert same_double(v, float(exact))
    v, _, _ = _sum(gpu, a, 8, True, rm=0)
    assert v == float(exact)          # reference Round() is exact here (top limb well filled)


def test_loguniform_2p30_variants_and_partition(gpu, oracle):
    import exblas_b200 as xb
    a = cm.init_fpuniform(N, 664, 332, seed=1, neg_ratio=2, device="cuda")
    v0, l0, st = _sum(gpu, a, 0, False)
    assert st == 0
    for fpe, ee in [(3, False), (4, False), (8, False), (4, True), (8, True)]:
        v, l, st = _sum(gpu, a, fpe, ee)
        assert same_double(v, v0) and (l == l0).all() and st == 0, (fpe, ee)
    # 8-way partition (what 8 GPUs would each reduce), merged as integers on the host
    acc = np.zeros(39, dtype=np.int64)
    per = N // 8
    for g in range(8):
        _, l, _ = _sum(gpu, a, 4, False, n=per, off=g * per)
        acc = xb.merge_limbs(acc, l)
    assert (acc == l0).all()
    assert same_double(xb.round_limbs(acc, 0), v0)
    # oracle on a slice regenerated independently on the host
    lo, hi = 123_456_789, 123_456_789 + 200_000
    host = cm.init_fpuniform(N, 664, 332, seed=1, neg_ratio=2, lo=lo, hi=hi)
    w, wl = oracle.exsum(host, fpe=0)
    v, l, _ = _sum(gpu, a, 8, True, n=hi - lo, off=lo)
    assert same_double(v, w) and (l == wl).all()


def test_cancelling_sum_2p30_known_answer(gpu):
    a = cm.cancelling_pair(N, "sum", device="cuda")
    for fpe, ee in [(0, False), (4, False), (8, True)]:
        for rm in (0, 1):
            v, l, st = _sum(gpu, a, fpe, ee, rm)
            assert v == 1.5 and st == 0, (fpe, ee, rm)


def test_cancelling_dot_2p30_known_answer(gpu):
    """BASELINE config 3: 2^30-element ill-conditioned dot product (cond > 1e32), exact answer 1.5"""
    a, b = cm.cancelling_pair(N, "dot", device="cuda")
    for fpe, ee in [(0, False), (3, False), (8, False), (8, True)]:
        gpu.exdot_async(N, a, 1, 0, b, 1, 0, fpe, ee, 1)
        v, l, st = gpu.fetch()
        assert v == 1.5 and st == 0, (fpe, ee)


def test_naive_closed_form_2p32(gpu):
    """BASELINE config 4's largest size: 2^32 doubles (32 GiB) -- 64-bit lengths end to end"""
    import torch
    n = 1 << 32
    free, _ = torch.cuda.mem_get_info()
    if free < (n * 8) + (2 << 30):
        pytest.skip("not enough free device memory for a 32 GiB vector")
    a = torch.full((n,), 1.1, dtype=torch.float64, device="cuda")
    a[n - 1] = -3.25                                         # the very last element must be seen
    exact = Fraction(1.1) * (n - 1) + Fraction(-3.25)
    want = limbs_from_fraction(exact)
    for fpe, ee in [(0, False), (3, False), (8, True)]:
        v, l, st = _sum(gpu, a, fpe, ee, rm=1)
        assert st == 0 and (l == want).all() and v == float(exact), (fpe, ee)
    del a
    torch.cuda.empty_cache()
