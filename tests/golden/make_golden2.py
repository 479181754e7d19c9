"""Generate tests/golden/golden2.npz from the UNMODIFIED reference code (oracle/_ref/libexblas_ref.so) -- run in the
build container only:

    python tests/golden/make_golden2.py

Golden vectors for the rows of SURVEY section 8f that golden.npz does not cover:
  gemv_<k>/...     ExGEMV cases ('N' and 'T'): inputs + the reference tests' MPFR checker
                   (tests/test.exgemv.gpu.cpp:35-78) = the correctly rounded y
  seg/...          a batched reduction: data, offsets, and per segment the reference's own CPU exsum()
                   (Superaccumulator path, reference Round()) and the restated ExDOT.Superacc.cl on the reference
                   Superaccumulator class, plus the MPFR checkers
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from exblas_b200 import common as cm          # noqa: E402
from oracle.oracle import Reference           # noqa: E402

GEMV = [("N", 40, 33, 43, "loguniform", 1.0, 1.0), ("N", 64, 100, 64, "narrow", 1.0, 0.0), ("N", 17, 300, 20, "illcond", 1.0, -2.5),
        ("T", 300, 20, 300, "narrow", 1.0, 1.0), ("T", 513, 9, 520, "illcond", 1.0, 0.0), ("T", 280, 12, 281, "loguniform", 1.0, 0.5),
        ("N", 30, 50, 30, "narrow", -0.75, 2.5)]


def gemv_inputs(m, n, lda, kind, seed):
    if kind == "loguniform":
        A = cm.init_fpuniform(m * n, 300, 150, seed=seed, neg_ratio=2)
        x = cm.init_fpuniform(max(m, n), 300, 150, seed=seed + 1, neg_ratio=3)
    elif kind == "illcond":
        A = cm.init_ill_cond(max(m * n, 2), 1e32, seed=seed)[: m * n]
        x = cm.init_ill_cond(max(m, n, 2), 1e32, seed=seed + 1)[: max(m, n)]
    else:
        A = cm.init_fpuniform(m * n, 10, 5, seed=seed, neg_ratio=2)
        x = cm.init_fpuniform(max(m, n), 10, 5, seed=seed + 1, neg_ratio=2)
    a = np.zeros(lda * n)
    a.reshape(n, lda)[:, :m] = A.reshape(n, m)
    y = cm.init_fpuniform(max(m, n), 100, 50, seed=seed + 2, neg_ratio=2)
    return a, x, y


def main():
    ref = Reference()
    data = {}
    for k, (trans, m, n, lda, kind, alpha, beta) in enumerate(GEMV):
        a, x, y = gemv_inputs(m, n, lda, kind, 100 + k)
        nin, nout = (m, n) if trans == "T" else (n, m)
        data[f"gemv_{k}/a"], data[f"gemv_{k}/x"], data[f"gemv_{k}/y"] = a, x[:nin].copy(), y[:nout].copy()
        data[f"gemv_{k}/mpfr"] = ref.exgemv_mpfr(trans, m, n, alpha, a, lda, x[:nin], 1, beta, y[:nout], 1)
    # batched reduction: 260 segments of 0 .. 400 elements, three data kinds interleaved
    rng = np.random.Generator(np.random.PCG64(77))
    lengths = np.concatenate([[0, 1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 400], rng.integers(0, 120, size=247)])
    seg = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    total = int(seg[-1])
    a = np.empty(total)
    b = np.empty(total)
    for i in range(lengths.size):
        lo, hi = int(seg[i]), int(seg[i + 1])
        if hi == lo:
            continue
        kind = i % 3
        if kind == 0:
            a[lo:hi] = cm.init_fpuniform(hi - lo, 10, 5, seed=1000 + i, neg_ratio=2)
            b[lo:hi] = cm.init_fpuniform(hi - lo, 10, 5, seed=2000 + i, neg_ratio=2)
        elif kind == 1:
            a[lo:hi] = cm.init_fpuniform(hi - lo, 664, 332, seed=1000 + i, neg_ratio=2)
            b[lo:hi] = cm.init_fpuniform(hi - lo, 300, 150, seed=2000 + i, neg_ratio=2)
        else:
            a[lo:hi] = cm.init_ill_cond(max(hi - lo, 2), 1e32, seed=1000 + i)[: hi - lo]
            b[lo:hi] = cm.init_ill_cond(max(hi - lo, 2), 1e32, seed=2000 + i)[: hi - lo]
    data["seg/a"], data["seg/b"], data["seg/offsets"] = a, b, seg
    rs, ms, rd, md = [], [], [], []
    for i in range(lengths.size):
        lo, hi = int(seg[i]), int(seg[i + 1])
        if hi == lo:
            rs.append(0.0); ms.append(0.0); rd.append(0.0); md.append(0.0)
            continue
        rs.append(ref.exsum(a[lo:hi], fpe=0))                   # the reference's CPU exsum(): superaccumulator + Round()
        ms.append(ref.exsum_mpfr(a[lo:hi]))
        rd.append(ref.exdot_superacc(a[lo:hi], b[lo:hi])[0])
        md.append(ref.exdot_mpfr(a[lo:hi], b[lo:hi]))
    data["seg/ref_exsum"], data["seg/mpfr_sum"] = np.array(rs), np.array(ms)
    data["seg/ref_exdot"], data["seg/mpfr_dot"] = np.array(rd), np.array(md)
    data["gemv_cases"] = np.array([f"{t},{m},{n},{lda},{kind},{alpha},{beta}" for (t, m, n, lda, kind, alpha, beta) in GEMV])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden2.npz")
    np.savez_compressed(out, **data)
    print("wrote", out, os.path.getsize(out), "bytes;", len(GEMV), "gemv cases,", lengths.size, "segments,", total, "elements")


if __name__ == "__main__":
    main()
