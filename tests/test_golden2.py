"""tests/golden/golden2.npz: outputs of the UNMODIFIED reference (its CPU exsum(), its Superaccumulator class fed
as ExDOT.Superacc.cl does, and its tests' MPFR checkers for ExSUM / ExDOT / ExGEMV) for the batched reductions
and ExGEMV (made by tests/golden/make_golden2.py in the build container).  CPU: the oracle must reproduce them;
GPU: the CUDA path must, through the C ABI."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g2():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden2.npz"))


def gemv_cases(g2):
    for k, line in enumerate(g2["gemv_cases"]):
        t, m, n, lda, kind, alpha, beta = str(line).split(",")
        yield k, t, int(m), int(n), int(lda), kind, float(alpha), float(beta)


def same_bits(x, y):
    return (np.asarray(x, dtype=np.float64).view(np.uint64) == np.asarray(y, dtype=np.float64).view(np.uint64)) | \
           ((np.asarray(x) == 0) & (np.asarray(y) == 0))


def test_oracle_matches_reference_gemv_and_segments(oracle, g2):
    for k, t, m, n, lda, kind, alpha, beta in gemv_cases(g2):
        got = oracle.exgemv(t, m, n, alpha, g2[f"gemv_{k}/a"], lda, g2[f"gemv_{k}/x"], 1, beta, g2[f"gemv_{k}/y"], 1, 0, False, 1)
        assert same_bits(got, g2[f"gemv_{k}/mpfr"]).all(), (k, t, m, n, kind)
    a, b, seg = g2["seg/a"], g2["seg/b"], g2["seg/offsets"]
    for i in range(seg.size - 1):
        lo, hi = int(seg[i]), int(seg[i + 1])
        if hi == lo:
            continue
        assert same_bits(oracle.exsum(a[lo:hi], fpe=0, round_mode=0)[0], g2["seg/ref_exsum"][i]), i
        assert same_bits(oracle.exsum(a[lo:hi], fpe=0, round_mode=1)[0], g2["seg/mpfr_sum"][i]), i
        assert same_bits(oracle.exdot(a[lo:hi], b[lo:hi], fpe=0, round_mode=0)[0], g2["seg/ref_exdot"][i]), i
        assert same_bits(oracle.exdot(a[lo:hi], b[lo:hi], fpe=0, round_mode=1)[0], g2["seg/mpfr_dot"][i]), i


@pytest.mark.gpu
def test_gpu_matches_reference_gemv(gpu, g2):
    import torch
    import exblas_b200 as xb
    for k, t, m, n, lda, kind, alpha, beta in gemv_cases(g2):
        a, x, y = g2[f"gemv_{k}/a"], g2[f"gemv_{k}/x"], g2[f"gemv_{k}/y"]
        for fpe, ee in ((0, False), (3, False), (8, True)):
            dy = torch.from_numpy(y.copy()).cuda()
            xb.exgemv(t, m, n, alpha, torch.from_numpy(a).cuda(), lda, 0, torch.from_numpy(x).cuda(), 1, 0, beta, dy, 1, 0, fpe, ee,
                      round_mode=xb.ROUND_EXACT, handle=gpu)
            assert same_bits(dy.cpu().numpy(), g2[f"gemv_{k}/mpfr"]).all(), (k, t, m, n, kind, fpe, ee)
        hy = y.copy()                                            # host operands, as the reference's exgemv takes them
        xb.exgemv(t, m, n, alpha, a, lda, 0, x, 1, 0, beta, hy, 1, 0, 0, False, round_mode=xb.ROUND_EXACT, handle=gpu)
        assert same_bits(hy, g2[f"gemv_{k}/mpfr"]).all(), (k, "host")


@pytest.mark.gpu
def test_gpu_matches_reference_segments(gpu, g2):
    import torch
    import exblas_b200 as xb
    a, b, seg = g2["seg/a"], g2["seg/b"], g2["seg/offsets"]
    da, db, ds = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(seg).cuda()
    for rm, ws, wd in ((xb.ROUND_REFERENCE, g2["seg/ref_exsum"], g2["seg/ref_exdot"]), (xb.ROUND_EXACT, g2["seg/mpfr_sum"], g2["seg/mpfr_dot"])):
        assert same_bits(gpu.exsum_segments(da, ds, round_mode=rm).cpu().numpy(), ws).all()
        assert same_bits(gpu.exdot_segments(da, db, ds, round_mode=rm).cpu().numpy(), wd).all()
        assert same_bits(gpu.exsum_segments(a, seg, round_mode=rm), ws).all()               # host operands
        assert same_bits(gpu.exdot_segments(a, b, seg, round_mode=rm), wd).all()
    # each segment on its own through the plain entry points (what the reference's callers do)
    for i in (3, 7, 12, 50, 200):
        lo, hi = int(seg[i]), int(seg[i + 1])
        if hi > lo:
            assert same_bits(gpu.exsum(hi - lo, a[lo:hi].copy(), 1, 0, 0), g2["seg/ref_exsum"][i])
