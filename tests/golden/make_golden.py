"""Generate tests/golden/golden.npz from the UNMODIFIED reference code (oracle/_ref/libexblas_ref.so,
built by oracle/Makefile from /root/reference) -- run in the build container only:

    python tests/golden/make_golden.py

The reference ships no golden vectors; these are outputs of the reference
itself on fixed inputs, so the parity pin travels to machines without /root/reference.

Per case the file stores the input vector(s) and, computed by the reference:
  ref_exsum[v]   exsum(N, a, 1, 0, fpe, early_exit) for the variants of tests/test.exsum.cpu.cpp:107-112
  ref_limbs41    normalised Superaccumulator limbs (CPU layout, 41 limbs)
  ref_round      Superaccumulator::Round() of those limbs
  mpfr           the reference tests' MPFR checker (correctly rounded)
and for dot cases the restated ExDOT.Superacc.cl on the reference Superaccumulator class + MPFR.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from exblas_b200 import common as cm          # noqa: E402
from oracle.oracle import Reference           # noqa: E402

VARIANTS = [(0, 0), (2, 0), (3, 0), (4, 0), (8, 0), (4, 1), (6, 1), (8, 1)]


def cases():
    rng = np.random.Generator(np.random.PCG64(2026))
    out = {}
    for n in (8, 64, 1000, 4096):
        out[f"naive_{n}"] = cm.init_naive(n)
        out[f"loguniform_{n}"] = cm.init_fpuniform(n, 664, 332, seed=n)
        out[f"loguniform_signed_{n}"] = cm.init_fpuniform(n, 664, 332, seed=n + 1, neg_ratio=2)
        out[f"illcond_{n}"] = cm.init_ill_cond(n, 1e32, seed=n)
        out[f"lognormal_{n}"] = cm.init_lognormal(n, 0.0, 2.0, seed=n)
    out["cancel_4096"] = cm.cancelling_pair(4096, "sum")
    out["negative_1000"] = -np.abs(cm.init_fpuniform(1000, 200, 100, seed=5))
    # adversarial rounding cases: sums whose top limb holds 1..4 bits, ties, exact zero
    for k in range(24):
        e = int(rng.integers(-750, 900))
        top = 52 * (e // 52) + int(rng.integers(0, 4))
        a = np.ldexp(rng.uniform(1, 2, 33), rng.integers(top - 120, top, 33)) * rng.choice([-1.0, 1.0], 33)
        a[0] = np.ldexp(1.0, top) * rng.choice([-1.0, 1.0])
        out[f"adversarial_{k}"] = a
    out["tie_even"] = np.array([1.0, 2.0 ** -53, 0, 0, 0, 0, 0, 0])
    out["tie_odd"] = np.array([1.0 + 2.0 ** -52, 2.0 ** -53, 0, 0, 0, 0, 0, 0])
    out["zero_sum"] = np.array([3.5, -1.25, -2.25, 1e-30, -1e-30, 0, 0, 0])
    return out


def main():
    ref = Reference()
    data = {}
    names = []
    for name, a in cases().items():
        a = np.ascontiguousarray(a, dtype=np.float64)
        names.append(name)
        data[f"{name}/a"] = a
        vals = []
        for fpe, ee in VARIANTS:
            if a.size < 8 and fpe >= 2:
                vals.append(np.nan)
            else:
                vals.append(ref.exsum(a, fpe=fpe, early_exit=bool(ee)))
        data[f"{name}/ref_exsum"] = np.array(vals)
        r, l41 = ref.superacc_limbs(a)
        data[f"{name}/ref_limbs41"] = l41
        data[f"{name}/ref_round"] = np.array([r])
        data[f"{name}/mpfr"] = np.array([ref.exsum_mpfr(a)])
    dnames = []
    for n in (8, 100, 1000, 4096):
        for kind in ("loguniform", "illcond", "naive"):
            if kind == "loguniform":
                a = cm.init_fpuniform(n, 664, 332, seed=n + 7, neg_ratio=2)
                b = cm.init_fpuniform(n, 664, 332, seed=n + 8, neg_ratio=3)
            elif kind == "illcond":
                a = cm.init_ill_cond(n, 1e32, seed=n + 7)
                b = cm.init_ill_cond(n, 1e32, seed=n + 8)
            else:
                a = cm.init_naive(n)
                b = cm.init_naive(n)
            name = f"dot_{kind}_{n}"
            dnames.append(name)
            data[f"{name}/a"] = a
            data[f"{name}/b"] = b
            r, l41 = ref.exdot_superacc(a, b)
            data[f"{name}/ref_limbs41"] = l41
            data[f"{name}/ref_round"] = np.array([r])
            data[f"{name}/mpfr"] = np.array([ref.exdot_mpfr(a, b)])
    x, y, exact, cond = cm.gen_dot(2000, 1e32, seed=3)
    name = "dot_gendot_2000"
    dnames.append(name)
    data[f"{name}/a"] = x
    data[f"{name}/b"] = y
    r, l41 = ref.exdot_superacc(x, y)
    data[f"{name}/ref_limbs41"] = l41
    data[f"{name}/ref_round"] = np.array([r])
    data[f"{name}/mpfr"] = np.array([ref.exdot_mpfr(x, y)])
    data["sum_cases"] = np.array(names)
    data["dot_cases"] = np.array(dnames)
    data["variants"] = np.array(VARIANTS)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes;", len(names), "sum cases,", len(dnames), "dot cases")


if __name__ == "__main__":
    main()
