"""CPU tests: pin the oracle (oracle/exblas_oracle.c).

 * against the golden vectors generated from the unmodified reference (tests/golden/make_golden.py);
 * against the reference library itself on fresh random inputs, when oracle/_ref is present;
 * against independent exact arithmetic (fractions / math.fsum) for the correctly rounded finaliser
   and for the executable model of the reference Round().
"""
import math

import numpy as np
import pytest

from helpers import (cpu41_to_gpu39, exact_dot_fraction, exact_sum_fraction, limbs_from_fraction, ref_round_model,
                     round_fraction, same_double)
from exblas_b200 import common as cm


def test_golden_exsum(oracle, golden):
    variants = [tuple(v) for v in golden["variants"]]
    for name in golden["sum_cases"]:
        a = golden[f"{name}/a"]
        ref_vals = golden[f"{name}/ref_exsum"]
        l41 = golden[f"{name}/ref_limbs41"]
        ref_round = float(golden[f"{name}/ref_round"][0])
        mpfr = float(golden[f"{name}/mpfr"][0])
        # every FPE variant of the reference agrees with its own superaccumulator (test.exsum.cpu.cpp:138-146)
        for v in ref_vals:
            assert math.isnan(v) or same_double(v, ref_round), name
        r41, o41 = oracle.superacc_limbs(a, layout=0)
        assert (o41 == l41).all(), name
        assert same_double(r41, ref_round), name
        l39 = cpu41_to_gpu39(l41)
        for (fpe, ee) in variants:
            v, l = oracle.exsum(a, fpe=int(fpe), early_exit=bool(ee), round_mode=0)
            assert same_double(v, ref_round), (name, fpe, ee)
            assert (l == l39).all(), (name, fpe, ee)
        vx, _ = oracle.exsum(a, fpe=4, round_mode=1)
        assert same_double(vx, mpfr), name
        assert same_double(vx, math.fsum(a)), name


def test_golden_exdot(oracle, golden):
    for name in golden["dot_cases"]:
        a, b = golden[f"{name}/a"], golden[f"{name}/b"]
        l39 = cpu41_to_gpu39(golden[f"{name}/ref_limbs41"])
        ref_round = float(golden[f"{name}/ref_round"][0])
        mpfr = float(golden[f"{name}/mpfr"][0])
        for fpe, ee in [(0, 0), (3, 0), (4, 0), (8, 0), (4, 1), (6, 1), (8, 1)]:
            v, l = oracle.exdot(a, b, fpe=fpe, early_exit=bool(ee))
            assert same_double(v, ref_round), (name, fpe, ee)
            assert (l == l39).all(), (name, fpe, ee)
        vx, _ = oracle.exdot(a, b, fpe=0, round_mode=1)
        assert same_double(vx, mpfr), name
        assert same_double(vx, round_fraction(exact_dot_fraction(a, b))), name


def test_golden_covers_known_round_defect(golden):
    """The reference Round() is not always correctly rounded; the golden set
    must contain such cases so that reference-parity and exact rounding are really told apart."""
    diff = sum(1 for n in golden["sum_cases"]
               if float(golden[f"{n}/ref_round"][0]) != float(golden[f"{n}/mpfr"][0]))
    assert diff >= 3


def test_oracle_vs_reference_random(oracle, reference):
    rng = np.random.default_rng(11)
    for trial in range(60):
        n = int(rng.integers(8, 3000))
        kind = trial % 4
        if kind == 0:
            a = np.ldexp(rng.uniform(1, 2, n), rng.integers(-900, 900, n)) * rng.choice([-1.0, 1.0], n)
        elif kind == 1:
            a = cm.init_fpuniform(n, 664, 332, seed=trial)
        elif kind == 2:
            a = cm.init_ill_cond(n, 1e32, seed=trial)
        else:
            a = cm.cancelling_pair(2 * (n // 2) + 2, "sum", seed=trial)
        r, l41 = reference.superacc_limbs(a)
        ro, lo = oracle.superacc_limbs(a, layout=0)
        assert (l41 == lo).all() and same_double(r, ro)
        for fpe, ee in [(0, 0), (2, 0), (4, 0), (8, 0), (4, 1), (8, 1)]:
            assert same_double(reference.exsum(a, fpe=fpe, early_exit=bool(ee)), r)
            v, l = oracle.exsum(a, fpe=fpe, early_exit=bool(ee))
            assert same_double(v, r) and (l == cpu41_to_gpu39(l41)).all()
        b = np.ldexp(rng.uniform(1, 2, a.size), rng.integers(-300, 300, a.size)) * rng.choice([-1.0, 1.0], a.size)
        a2 = np.ldexp(rng.uniform(1, 2, a.size), rng.integers(-300, 300, a.size)) * rng.choice([-1.0, 1.0], a.size)
        rd, ld41 = reference.exdot_superacc(a2, b)
        vd, ld = oracle.exdot(a2, b, fpe=4)
        assert same_double(vd, rd) and (ld == cpu41_to_gpu39(ld41)).all()
        assert same_double(oracle.exdot(a2, b, fpe=0, round_mode=1)[0], reference.exdot_mpfr(a2, b))


def test_oracle_round_model_and_exact(oracle):
    """Round(): oracle C == Appendix-A2 python model (driven from the exact rational sum);
    round_exact == correctly rounded rational; normal-form limbs == limbs_from_fraction."""
    rng = np.random.default_rng(5)
    mism = 0
    for trial in range(400):
        n = int(rng.integers(1, 60))
        top = int(rng.integers(-750, 900))
        a = np.ldexp(rng.uniform(1, 2, n), rng.integers(top - 150, top + 1, n)) * rng.choice([-1.0, 1.0], n)
        if trial % 5 == 0:
            a = np.concatenate([a, -a[: n // 2]])
        fr = exact_sum_fraction(a)
        limbs = limbs_from_fraction(fr)
        v0, l = oracle.exsum(a, fpe=0, round_mode=0)
        v1, _ = oracle.exsum(a, fpe=3, round_mode=1)
        assert (l == limbs).all()
        assert same_double(v0, ref_round_model(limbs))
        assert same_double(v1, round_fraction(fr))
        mism += v0 != v1
    assert mism > 0       # the defect is exercised


def test_oracle_strides_offsets_and_empty(oracle):
    a = cm.init_fpuniform(1000, 100, 50, seed=9, neg_ratio=2)
    for off, inc in [(0, 1), (1, 1), (3, 2), (7, 5)]:
        m = (a.size - off + inc - 1) // inc
        v, l = oracle.exsum(a, inca=inc, offset=off, n=m, fpe=4)
        assert same_double(oracle.exsum(a[off::inc], fpe=0)[0], v)
    assert oracle.exsum(np.zeros(0), fpe=4)[0] == 0.0
    assert oracle.exdot(np.zeros(0), np.zeros(0), fpe=4)[0] == 0.0


def test_oracle_merge_is_exact(oracle):
    """per-rank normalised limbs summed as integers == limbs of the whole (cpu ExSUM.cpp:266-273)"""
    a = cm.init_fpuniform(4000, 664, 332, seed=4, neg_ratio=2)
    whole_v, whole_l = oracle.exsum(a, fpe=0)
    for parts in (2, 3, 8):
        shards = np.array_split(a, parts)
        per = np.stack([oracle.exsum(s, fpe=4)[1] for s in shards])
        v, merged = oracle.merge_round(per)
        assert (merged == whole_l).all() and same_double(v, whole_v)
