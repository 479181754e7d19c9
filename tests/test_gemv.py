"""ExGEMV 'N' (BASELINE config 5).

CPU: the oracle's exgemv against exact rationals and against the reference tests' MPFR checker
(tests/test.exgemv.gpu.cpp:35-78, compiled into oracle/_ref).  GPU: the CUDA path through the C ABI
against the oracle, bit for bit, including alpha / beta / lda / inc / offsets / column splits."""
from fractions import Fraction

import numpy as np
import pytest

from exblas_b200 import common as cm

VARIANTS = [(0, False), (1, False), (2, False), (3, False), (4, False), (8, False), (4, True), (6, True), (8, True)]


def make(m, n, lda, kind, seed):
    a = np.zeros(lda * n)
    if kind == "loguniform":
        A = cm.init_fpuniform(m * n, 300, 150, seed=seed, neg_ratio=2)
        x = cm.init_fpuniform(n, 300, 150, seed=seed + 1, neg_ratio=3)
    elif kind == "illcond":
        A = cm.init_ill_cond(max(m * n, 2), 1e32, seed=seed)[: m * n]
        x = cm.init_ill_cond(max(n, 2), 1e32, seed=seed + 1)[:n]
    else:
        A = np.full(m * n, 1.1)
        x = np.full(n, 1.1)
    a.reshape(n, lda)[:, :m] = A.reshape(n, m)
    y = cm.init_fpuniform(m, 100, 50, seed=seed + 2, neg_ratio=2)
    return a, x, y


def exact(m, n, alpha, a, lda, x, beta, y):
    A = a.reshape(n, lda)
    return np.array([float(sum(Fraction(alpha) * Fraction(float(A[j, i])) * Fraction(float(x[j])) for j in range(n))
                           + Fraction(beta) * Fraction(float(y[i]))) for i in range(m)])


def test_oracle_exgemv_exact_and_mpfr(oracle):
    for (m, n, lda) in [(1, 1, 1), (5, 7, 5), (40, 33, 43), (64, 64, 64)]:
        for kind in ("loguniform", "illcond", "naive"):
            a, x, y = make(m, n, lda, kind, seed=m + n)
            for alpha, beta in [(1.0, 1.0), (1.0, 0.0), (1.0, -2.5), (3.7, 0.3)]:
                want = exact(m, n, alpha, a, lda, x, beta, y)
                for fpe, ee in VARIANTS:
                    got = oracle.exgemv("N", m, n, alpha, a, lda, x, 1, beta, y, 1, fpe, ee, 1)
                    assert (got == want).all(), (m, n, kind, alpha, beta, fpe, ee)
                # reference-compatible rounding differs from exact only through Round()'s known defect
                g0 = oracle.exgemv("N", m, n, alpha, a, lda, x, 1, beta, y, 1, 4, False, 0)
                assert np.allclose(g0, want, rtol=4e-16, atol=0)


def test_oracle_exgemv_vs_reference_mpfr_checker(oracle, reference):
    for (m, n, lda) in [(40, 33, 43), (64, 100, 64)]:
        a, x, y = make(m, n, lda, "loguniform", seed=7)
        for beta in (1.0, 0.0, 0.5):
            mp = reference.exgemv_mpfr("N", m, n, 1.0, a, lda, x, 1, beta, y, 1)
            got = oracle.exgemv("N", m, n, 1.0, a, lda, x, 1, beta, y, 1, 8, True, 1)
            assert (got == mp).all()


@pytest.mark.gpu
def test_exgemv_gpu_vs_oracle(gpu, oracle):
    import torch
    import exblas_b200 as xb
    shapes = [(1, 1, 1), (5, 7, 5), (33, 100, 40), (512, 64, 512), (513, 1030, 520), (1500, 257, 1500), (100, 5000, 128)]
    for (m, n, lda) in shapes:
        for kind in ("loguniform", "illcond", "naive"):
            a, x, y = make(m, n, lda, kind, seed=m * 3 + n)
            for alpha, beta in [(1.0, 1.0), (1.0, 0.0), (-0.75, 2.5)]:
                w0 = oracle.exgemv("N", m, n, alpha, a, lda, x, 1, beta, y, 1, 0, False, 0)
                w1 = oracle.exgemv("N", m, n, alpha, a, lda, x, 1, beta, y, 1, 0, False, 1)
                da, dx = torch.from_numpy(a).cuda(), torch.from_numpy(x).cuda()
                for fpe, ee in VARIANTS:
                    for rm, want in ((0, w0), (1, w1)):
                        dy = torch.from_numpy(y).cuda()
                        xb.exgemv("N", m, n, alpha, da, lda, 0, dx, 1, 0, beta, dy, 1, 0, fpe, ee, round_mode=rm, handle=gpu)
                        got = dy.cpu().numpy()
                        assert (got.view(np.uint64) == want.view(np.uint64)).all(), (m, n, kind, alpha, beta, fpe, ee, rm)
                # host pointers (what the reference's exgemv takes)
                hy = y.copy()
                xb.exgemv("N", m, n, alpha, a, lda, 0, x, 1, 0, beta, hy, 1, 0, 4, False, handle=gpu)
                assert (hy.view(np.uint64) == w0.view(np.uint64)).all()
    assert gpu.last_status() == 0


@pytest.mark.gpu
def test_exgemv_offsets_strides_and_splits(gpu, oracle):
    import torch
    import exblas_b200 as xb
    m, n, lda = 300, 777, 320
    a, x, y = make(m, n, lda, "loguniform", seed=11)
    offa, offx, offy, incx, incy = 5, 3, 2, 2, 3
    abuf = np.concatenate([np.full(offa, 9.0), a])
    xbuf = np.full(offx + (n - 1) * incx + 1, 7.0)
    xbuf[offx::incx][:n] = x
    ybuf = np.full(offy + (m - 1) * incy + 1, 5.0)
    ybuf[offy::incy][:m] = y
    want = oracle.exgemv("N", m, n, 1.0, a, lda, x, 1, 1.0, y, 1, 0, False, 0)
    try:
        for parts in (0, 1, 2, 7, 13):
            gpu.set_option("gemv_parts", parts)
            for fpe, ee in [(0, False), (3, False), (8, True)]:
                dy = torch.from_numpy(ybuf.copy()).cuda()
                xb.exgemv("N", m, n, 1.0, torch.from_numpy(abuf).cuda(), lda, offa, torch.from_numpy(xbuf).cuda(), incx, offx,
                          1.0, dy, incy, offy, fpe, ee, handle=gpu)
                out = dy.cpu().numpy()
                assert (out[offy::incy][:m].view(np.uint64) == want.view(np.uint64)).all(), (parts, fpe, ee)
                mask = np.ones(out.size, dtype=bool)
                mask[offy::incy] = False
                assert (out[mask] == 5.0).all()                      # nothing else touched
    finally:
        gpu.set_option("gemv_parts", 0)


@pytest.mark.gpu
def test_exgemv_large_rows_match_exdot(gpu):
    """size-independent property at a BASELINE-like shape: every row of exgemv equals exdot of that row"""
    import torch
    import exblas_b200 as xb
    m, n = 4096, 8192
    A = cm.init_fpuniform(m * n, 664, 332, seed=3, neg_ratio=2, device="cuda")      # column-major m x n
    x = cm.init_fpuniform(n, 100, 50, seed=4, neg_ratio=2, device="cuda")
    y = torch.zeros(m, dtype=torch.float64, device="cuda")
    xb.exgemv("N", m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, 8, True, handle=gpu)
    rows = [0, 1, 511, 512, 2047, 4095]
    Am = A.view(n, m)
    for r in rows:
        row = Am[:, r].contiguous()
        gpu.exdot_async(n, row, 1, 0, x, 1, 0, 3, False, 0)
        v, _, st = gpu.fetch()
        assert v == float(y[r]) and st == 0, r
    # naive closed form: every entry 1.1 * 1.1 summed n times
    A.fill_(1.1)
    x.fill_(1.1)
    xb.exgemv("N", m, n, 1.0, A, m, 0, x, 1, 0, 0.0, y, 1, 0, 4, False, round_mode=xb.ROUND_EXACT, handle=gpu)
    want = float(Fraction(1.1) * Fraction(1.1) * n)
    assert bool((y == want).all())


# ---- register-window kernels (exblas_b200/csrc/window.cuh): 'N' with fpe < 2, and 'T' ----

def make_kind(m, n, lda, kind, seed):
    """column-major m x n matrix (leading dimension lda), x of length max(m, n), y likewise"""
    rng = np.random.default_rng(seed)
    A = cm.init_fpuniform(m * n, 10, 5, seed=seed, neg_ratio=2).reshape(n, m)          # narrow: 10 binades
    x = cm.init_fpuniform(max(m, n), 10, 5, seed=seed + 1, neg_ratio=2)
    if kind == "rowscale":         # every row at its own magnitude: per-thread windows ('N'), drifting windows ('T')
        A = A * np.exp2(rng.integers(-150, 150, size=m)).reshape(1, m)
    elif kind == "colscale":
        A = A * np.exp2(rng.integers(-150, 150, size=n)).reshape(n, 1)
    elif kind == "drift":          # the scale moves slowly along both directions: windows must follow
        A = A * np.exp2((np.arange(n) // 37).reshape(n, 1) * 3.0 - (np.arange(m) // 50).reshape(1, m) * 2.0)
    elif kind == "sparse":         # exact zeros, -0.0 and a few far outliers
        A = A.copy()
        A[rng.random((n, m)) < 0.3] = 0.0
        A[rng.random((n, m)) < 0.01] *= 2.0 ** 200
        x = x.copy()
        x[::7] = -0.0
    elif kind == "wide":
        A = cm.init_fpuniform(m * n, 600, 300, seed=seed, neg_ratio=2).reshape(n, m)
    a = np.zeros(lda * n)
    a.reshape(n, lda)[:, :m] = A
    y = cm.init_fpuniform(max(m, n), 100, 50, seed=seed + 2, neg_ratio=2)
    return a, x, y


KINDS = ("narrow", "rowscale", "colscale", "drift", "sparse", "wide")


@pytest.mark.gpu
def test_exgemv_n_window_kernel(gpu, oracle):
    import torch
    import exblas_b200 as xb
    try:
        for (m, n, lda) in [(700, 3000, 704), (384, 70, 384), (1000, 1037, 1003), (33, 9000, 40)]:
            for kind in KINDS:
                a, x, y = make_kind(m, n, lda, kind, seed=m + n)
                da, dx = torch.from_numpy(a).cuda(), torch.from_numpy(x[:n].copy()).cuda()
                for beta in (0.0, 1.0):
                    w0 = oracle.exgemv("N", m, n, 1.0, a, lda, x[:n], 1, beta, y[:m], 1, 0, False, 0)
                    w1 = oracle.exgemv("N", m, n, 1.0, a, lda, x[:n], 1, beta, y[:m], 1, 0, False, 1)
                    for window, nshape in ((1, 0), (1, 1), (1, 2), (0, 0)):
                        gpu.set_option("window", window)
                        gpu.set_option("gemv_n_shape", nshape)
                        # fpe >= 2 with alpha == 1 takes the window kernel too; window = 0 keeps the expansion kernels covered
                        for fpe in ((0, 1, 3, 8) if window else (0, 1, 2, 4, 8)):
                            for rm, want in ((0, w0), (1, w1)):
                                dy = torch.from_numpy(y[:m].copy()).cuda()
                                xb.exgemv("N", m, n, 1.0, da, lda, 0, dx, 1, 0, beta, dy, 1, 0, fpe, False, round_mode=rm, handle=gpu)
                                got = dy.cpu().numpy()
                                assert (got.view(np.uint64) == want.view(np.uint64)).all(), (m, n, kind, beta, window, nshape, fpe, rm)
                assert gpu.last_status() == 0
        # column splits of the window kernel (x slices staged per part) and a strided x
        m, n, lda = 500, 2500, 512
        a, x, y = make_kind(m, n, lda, "drift", seed=5)
        xbuf = np.full(3 * n, 9.0)
        xbuf[1::3][:n] = x[:n]
        want = oracle.exgemv("N", m, n, 1.0, a, lda, x[:n], 1, 1.0, y[:m], 1, 0, False, 0)
        gpu.set_option("window", 2)
        for parts in (1, 2, 7, 40):
            gpu.set_option("gemv_parts", parts)
            dy = torch.from_numpy(y[:m].copy()).cuda()
            xb.exgemv("N", m, n, 1.0, torch.from_numpy(a).cuda(), lda, 0, torch.from_numpy(xbuf).cuda(), 3, 1, 1.0, dy, 1, 0, 0, False, handle=gpu)
            assert (dy.cpu().numpy().view(np.uint64) == want.view(np.uint64)).all(), parts
    finally:
        gpu.set_option("window", 2)
        gpu.set_option("gemv_parts", 0)
        gpu.set_option("gemv_n_shape", 1)


@pytest.mark.gpu
def test_exgemv_t_kernels(gpu, oracle):
    """'T': y_j = sum_i A[i, j] x[i].  Columns >= 256 rows run the warp-per-output window kernel (both
    launch shapes), shorter ones the strided thread-per-output kernel; all must equal the oracle."""
    import torch
    import exblas_b200 as xb
    try:
        for (m, n, lda) in [(300, 40, 300), (2048, 24, 2048), (5000, 17, 5003), (255, 10, 256), (4224, 30, 4224), (3200, 9, 3201),
                            (1151, 13, 1152)]:
            for kind in KINDS:
                a, x, y = make_kind(m, n, lda, kind, seed=2 * m + n)
                da, dx = torch.from_numpy(a).cuda(), torch.from_numpy(x[:m].copy()).cuda()
                for alpha, beta in ((1.0, 0.0), (1.0, 1.0), (1.0, -2.5), (0.3, 1.0)):
                    w0 = oracle.exgemv("T", m, n, alpha, a, lda, x[:m], 1, beta, y[:n], 1, 0, False, 0)
                    w1 = oracle.exgemv("T", m, n, alpha, a, lda, x[:m], 1, beta, y[:n], 1, 0, False, 1)
                    for shape in (0, 1, 2, 3, 4, 5):
                        gpu.set_option("gemv_t_shape", shape)
                        for fpe, ee in ((0, False), (8, True)):
                            for rm, want in ((0, w0), (1, w1)):
                                dy = torch.from_numpy(y[:n].copy()).cuda()
                                xb.exgemv("T", m, n, alpha, da, lda, 0, dx, 1, 0, beta, dy, 1, 0, fpe, ee, round_mode=rm, handle=gpu)
                                got = dy.cpu().numpy()
                                assert (got.view(np.uint64) == want.view(np.uint64)).all(), (m, n, kind, alpha, beta, shape, fpe, ee, rm)
                assert gpu.last_status() == 0
        # offsets and strides through the window kernel
        m, n, lda = 2100, 19, 2104
        a, x, y = make_kind(m, n, lda, "rowscale", seed=9)
        offa, offx, offy, incx, incy = 3, 2, 1, 2, 3
        abuf = np.concatenate([np.full(offa, 9.0), a])
        xbuf = np.full(offx + (m - 1) * incx + 1, 7.0)
        xbuf[offx::incx][:m] = x[:m]
        ybuf = np.full(offy + (n - 1) * incy + 1, 5.0)
        ybuf[offy::incy][:n] = y[:n]
        want = oracle.exgemv("T", m, n, 1.0, a, lda, x[:m], 1, 1.0, y[:n], 1, 0, False, 0)
        for shape in (0, 1, 2, 3, 4, 5):
            gpu.set_option("gemv_t_shape", shape)
            dy = torch.from_numpy(ybuf.copy()).cuda()
            xb.exgemv("T", m, n, 1.0, torch.from_numpy(abuf).cuda(), lda, offa, torch.from_numpy(xbuf).cuda(), incx, offx, 1.0, dy,
                      incy, offy, 0, False, handle=gpu)
            out = dy.cpu().numpy()
            assert (out[offy::incy][:n].view(np.uint64) == want.view(np.uint64)).all(), shape
            mask = np.ones(out.size, dtype=bool)
            mask[offy::incy] = False
            assert (out[mask] == 5.0).all()
        # specials keep their IEEE meaning through the window kernels
        m, n = 1024, 8
        a, x, y = make_kind(m, n, m, "narrow", seed=4)
        a2 = a.copy()
        a2[5] = np.inf
        a2[m + 7] = np.nan
        dy = torch.zeros(n, dtype=torch.float64, device="cuda")
        xb.exgemv("T", m, n, 1.0, torch.from_numpy(a2).cuda(), m, 0, torch.from_numpy(x[:m].copy()).cuda(), 1, 0, 0.0, dy, 1, 0, 0, False, handle=gpu)
        out = dy.cpu().numpy()
        assert np.isinf(out[0]) and np.isnan(out[1]) and np.isfinite(out[2:]).all()
    finally:
        gpu.set_option("gemv_t_shape", 2)


@pytest.mark.gpu
def test_exgemv_t_x_pipeline_many_chunks_and_sets(gpu):
    """'T' window kernel on a matrix whose columns span several x chunks (with a remainder) and whose outputs need several
    sets per CTA, so that the TMA / mbarrier pipeline of x wraps around many times: every launch shape, x staged by
    cp.async.bulk and by the plain-copy fallback (option gemv_tma = 0, and an x that is not 16-byte aligned), must give
    the bits of the general kernel (option window = 0, checked against the oracle elsewhere)."""
    import torch
    import exblas_b200 as xb
    from exblas_b200 import common as cm
    rows, cols = 9000 + 77, 148 * 16 * 2 + 5
    dev = torch.device("cuda:0")
    for kind in ("narrow", "wide"):
        er, eo = (10, 5) if kind == "narrow" else (600, 300)
        A = cm.init_fpuniform(rows * cols, er, eo, seed=11, neg_ratio=2, device=dev)
        x0 = cm.init_fpuniform(rows, 10, 5, seed=12, neg_ratio=2, device=dev)
        x1 = torch.zeros(rows + 1, dtype=torch.float64, device=dev)
        x1[1:] = x0                                         # the same x, one element (8 bytes) further on
        y0 = cm.init_fpuniform(cols, 100, 50, seed=13, neg_ratio=2, device=dev)
        try:
            gpu.set_option("window", 0)
            want = y0.clone()
            xb.exgemv("T", rows, cols, 1.0, A, rows, 0, x0, 1, 0, 1.0, want, 1, 0, 0, False, handle=gpu)
            gpu.set_option("window", 2)
            for shape in (0, 1, 2, 3, 4, 5):
                gpu.set_option("gemv_t_shape", shape)
                for tma, offx in ((1, 0), (0, 0), (1, 1)):
                    gpu.set_option("gemv_tma", tma)
                    got = y0.clone()
                    xb.exgemv("T", rows, cols, 1.0, A, rows, 0, x1 if offx else x0, 1, offx, 1.0, got, 1, 0, 3, False, handle=gpu)
                    assert bool((got.view(torch.int64) == want.view(torch.int64)).all()), (kind, shape, tma, offx)
                    assert gpu.last_status() == 0
        finally:
            gpu.set_option("window", 2)
            gpu.set_option("gemv_t_shape", 2)
            gpu.set_option("gemv_tma", 1)


@pytest.mark.gpu
def test_exgemv_t_random_shapes_against_general_kernel(gpu):
    """'T' window kernel against the general kernel (option window = 0) on random shapes around the chunk and round
    boundaries of every launch shape: odd row counts, single columns, padded leading dimensions, beta != 0."""
    import torch
    import exblas_b200 as xb
    from exblas_b200 import common as cm
    rng = np.random.default_rng(2024)
    dev = torch.device("cuda:0")
    edges = [256, 257, 383, 384, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 3071, 3072, 3073, 4095, 4096, 4097, 6143, 6144, 6145,
             8191, 8192, 8193, 12287, 12289, 16385]
    try:
        for t in range(24):
            rows = int(edges[t % len(edges)] if t < 16 else rng.integers(256, 20000))
            cols = int(rng.choice([1, 2, 15, 16, 17, 33, 150]))
            lda = rows + int(rng.integers(0, 5))
            A = cm.init_fpuniform(lda * cols, 40, 20, seed=100 + t, neg_ratio=2, device=dev)
            x = cm.init_fpuniform(rows, 10, 5, seed=200 + t, neg_ratio=2, device=dev)
            y0 = cm.init_fpuniform(cols, 30, 15, seed=300 + t, neg_ratio=2, device=dev)
            beta = float(rng.choice([0.0, 1.0, -0.75]))
            gpu.set_option("window", 0)
            want = y0.clone()
            xb.exgemv("T", rows, cols, 1.0, A, lda, 0, x, 1, 0, beta, want, 1, 0, 0, False, handle=gpu)
            gpu.set_option("window", 2)
            for shape in (0, 1, 2, 3, 4, 5):
                gpu.set_option("gemv_t_shape", shape)
                for tma in (1, 0):
                    gpu.set_option("gemv_tma", tma)
                    got = y0.clone()
                    xb.exgemv("T", rows, cols, 1.0, A, lda, 0, x, 1, 0, beta, got, 1, 0, 0, False, handle=gpu)
                    assert bool((got.view(torch.int64) == want.view(torch.int64)).all()), (rows, cols, lda, beta, shape, tma)
            assert gpu.last_status() == 0
    finally:
        gpu.set_option("window", 2)
        gpu.set_option("gemv_t_shape", 2)
        gpu.set_option("gemv_tma", 1)


@pytest.mark.gpu
def test_exgemv_exact_scaling_domain_is_flagged(gpu):
    """alpha * a[i, j] / beta * y[i] that cannot be split exactly into two doubles are dropped and flagged, never summed
    inexactly (include/exblas_b200.h: domain of the exact scaling)."""
    import torch
    import exblas_b200 as xb
    m, n = 40, 64
    a = np.full(m * n, 3.0)
    x = np.full(n, 1.0)
    for trans in ("N", "T"):
        nout = m if trans == "N" else n
        # ordinary alpha: exact, no flag
        dy = torch.zeros(nout, dtype=torch.float64).cuda()
        xb.exgemv(trans, m, n, 0.5, torch.from_numpy(a).cuda(), m, 0, torch.from_numpy(np.full(max(m, n), 1.0)).cuda(), 1, 0, 0.0, dy, 1, 0, 3, False, handle=gpu)
        assert gpu.last_status() == 0 and (dy.cpu().numpy() == 1.5 * (n if trans == "N" else m)).all()
        # alpha * a underflows: TwoProd would be inexact -> TOOSMALL
        a2 = a.copy()
        a2[5] = 2.0 ** -500
        dy = torch.zeros(nout, dtype=torch.float64).cuda()
        xb.exgemv(trans, m, n, 2.0 ** -600, torch.from_numpy(a2).cuda(), m, 0, torch.from_numpy(np.full(max(m, n), 1.0)).cuda(), 1, 0, 0.0, dy, 1, 0, 0, False, handle=gpu)
        assert gpu.last_status() & xb.ST_TOOSMALL, trans
        # alpha * a overflows with finite operands -> TOOLARGE, not Inf
        a3 = a.copy()
        a3[7] = 2.0 ** 600
        dy = torch.zeros(nout, dtype=torch.float64).cuda()
        xb.exgemv(trans, m, n, 2.0 ** 600, torch.from_numpy(a3).cuda(), m, 0, torch.from_numpy(np.full(max(m, n), 1.0)).cuda(), 1, 0, 0.0, dy, 1, 0, 4, False, handle=gpu)
        st = gpu.last_status()
        assert (st & xb.ST_TOOLARGE) and not (st & (xb.ST_POSINF | xb.ST_NAN)), (trans, st)
    # beta * y underflow
    dy = torch.full((m,), 2.0 ** -700, dtype=torch.float64).cuda()
    xb.exgemv("N", m, n, 1.0, torch.from_numpy(a).cuda(), m, 0, torch.from_numpy(x).cuda(), 1, 0, 2.0 ** -400, dy, 1, 0, 0, False, handle=gpu)
    assert gpu.last_status() & xb.ST_TOOSMALL
    # a clean call afterwards is clean
    dy = torch.zeros(m, dtype=torch.float64).cuda()
    xb.exgemv("N", m, n, 1.0, torch.from_numpy(a).cuda(), m, 0, torch.from_numpy(x).cuda(), 1, 0, 0.0, dy, 1, 0, 0, False, handle=gpu)
    assert gpu.last_status() == 0 and (dy.cpu().numpy() == 3.0 * n).all()
