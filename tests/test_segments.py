"""Batched ("segmented") ExSUM / ExDOT and the CSR sparse matrix-vector product built on it
(SURVEY section 8f rank 2): one launch, one warp per segment.  Every segment must equal the oracle's
exsum / exdot of that slice bit for bit, for both finalisers, whatever its length or data."""
import numpy as np
import pytest

from exblas_b200 import common as cm

LENGTHS = [0, 1, 2, 3, 31, 32, 33, 127, 128, 129, 255, 256, 257, 1000, 4099, 0, 70001, 5, 16390]


def make_data(total, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "narrow":
        return cm.init_fpuniform(total, 10, 5, seed=seed, neg_ratio=2)
    if kind == "wide":
        return cm.init_fpuniform(total, 664, 332, seed=seed, neg_ratio=2)
    if kind == "illcond":
        return cm.init_ill_cond(total, 1e32, seed=seed)
    if kind == "steps":            # narrow inside a segment-sized block, very different from block to block
        a = cm.init_fpuniform(total, 6, 3, seed=seed, neg_ratio=2)
        return a * np.exp2(rng.integers(-200, 200, size=(total + 499) // 500).repeat(500)[:total])
    raise ValueError(kind)


def offsets(lengths):
    return np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)


@pytest.mark.gpu
def test_segments_match_oracle(gpu, oracle):
    import torch
    seg = offsets(LENGTHS)
    total = int(seg[-1])
    for kind in ("narrow", "wide", "illcond", "steps"):
        a = make_data(total, kind, 3)
        b = make_data(total, kind if kind != "steps" else "narrow", 4)
        for rm in (0, 1):
            want_s = np.array([oracle.exsum(a[seg[i]:seg[i + 1]], fpe=0, round_mode=rm)[0] if seg[i + 1] > seg[i] else 0.0
                               for i in range(len(LENGTHS))])
            want_d = np.array([oracle.exdot(a[seg[i]:seg[i + 1]], b[seg[i]:seg[i + 1]], fpe=0, round_mode=rm)[0]
                               if seg[i + 1] > seg[i] else 0.0 for i in range(len(LENGTHS))])
            # device pointers
            da, db, ds = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(seg).cuda()
            got, st = gpu.exsum_segments(da, ds, round_mode=rm, want_status=True)
            assert (got.cpu().numpy().view(np.uint64) == want_s.view(np.uint64)).all(), (kind, rm, "sum dev")
            assert int(st.abs().sum()) == 0
            got = gpu.exdot_segments(da, db, ds, round_mode=rm)
            assert (got.cpu().numpy().view(np.uint64) == want_d.view(np.uint64)).all(), (kind, rm, "dot dev")
            # host pointers
            got = gpu.exsum_segments(a, seg, fpe=8, early_exit=True, round_mode=rm)
            assert (got.view(np.uint64) == want_s.view(np.uint64)).all(), (kind, rm, "sum host")
            got = gpu.exdot_segments(a, b, seg, fpe=3, round_mode=rm)
            assert (got.view(np.uint64) == want_d.view(np.uint64)).all(), (kind, rm, "dot host")
    assert gpu.last_status() == 0


@pytest.mark.gpu
def test_segments_many_short_and_gaps(gpu, oracle):
    """20 000 short segments (the k-means / SpMV-row regime), a third of them empty or a single element"""
    import torch
    rng = np.random.default_rng(5)
    lengths = rng.integers(0, 70, size=20000)
    short = rng.random(lengths.size) < 0.33
    lengths[short] = rng.integers(0, 2, size=int(short.sum()))
    seg = offsets(lengths)
    total = int(seg[-1]) + 5                                   # a few elements after the last segment are never read
    a = make_data(total, "wide", 6)
    got = gpu.exsum_segments(torch.from_numpy(a).cuda(), torch.from_numpy(seg).cuda()).cpu().numpy()
    idx = rng.choice(lengths.size, size=400, replace=False)
    for i in idx:
        want = oracle.exsum(a[seg[i]:seg[i + 1]], fpe=0)[0] if lengths[i] else 0.0
        assert np.float64(got[i]).view(np.uint64) == np.float64(want).view(np.uint64), (i, lengths[i])
    # the whole vector as ONE segment equals the plain exsum kernel
    one = gpu.exsum_segments(torch.from_numpy(a).cuda(), torch.tensor([0, total], dtype=torch.int64, device="cuda")).cpu().numpy()
    assert one[0] == gpu.exsum(total, torch.from_numpy(a).cuda(), 1, 0, 0)


@pytest.mark.gpu
def test_segments_csr_spmv(gpu, oracle):
    """CSR sparse matrix-vector product = exdot_segments with a gather index (what the reference's spmv
    example does with one exsum per row, src/cpu/examples/spmv/main.cpp:85)"""
    import torch
    rng = np.random.default_rng(8)
    nrows, ncols = 3000, 2500
    row_nnz = rng.integers(0, 60, size=nrows)
    row_nnz[7] = 900
    row_nnz[100] = 0
    rowptr = offsets(row_nnz)
    nnz = int(rowptr[-1])
    colidx = rng.integers(0, ncols, size=nnz).astype(np.int32)
    vals = make_data(nnz, "wide", 9)
    x = make_data(ncols, "narrow", 10)
    want = np.array([oracle.exdot(vals[rowptr[i]:rowptr[i + 1]], x[colidx[rowptr[i]:rowptr[i + 1]]], fpe=0)[0]
                     if row_nnz[i] else 0.0 for i in range(nrows)])
    got = gpu.exdot_segments(torch.from_numpy(vals).cuda(), torch.from_numpy(x).cuda(), torch.from_numpy(rowptr).cuda(),
                             gather=torch.from_numpy(colidx).cuda()).cpu().numpy()
    assert (got.view(np.uint64) == want.view(np.uint64)).all()
    got = gpu.exdot_segments(vals, x, rowptr, gather=colidx)                    # host operands
    assert (got.view(np.uint64) == want.view(np.uint64)).all()
    with pytest.raises(Exception):
        bad = colidx.copy()
        bad[3] = ncols
        gpu.exdot_segments(vals, x, rowptr, gather=bad)


@pytest.mark.gpu
def test_segments_status_per_segment(gpu):
    import exblas_b200 as xb
    a = np.ones(300)
    a[10] = np.inf
    a[150] = np.nan
    a[290] = 1e300
    seg = np.array([0, 100, 200, 280, 300], dtype=np.int64)
    got, st = gpu.exsum_segments(a, seg, want_status=True)
    assert np.isposinf(got[0]) and np.isnan(got[1]) and got[2] == 80.0
    assert st[0] == xb.ST_POSINF and st[1] == xb.ST_NAN and st[2] == 0 and st[3] == xb.ST_TOOLARGE
    assert gpu.last_status() == (xb.ST_POSINF | xb.ST_NAN | xb.ST_TOOLARGE)


@pytest.mark.gpu
def test_segments_any_decomposition(gpu, oracle):
    """The element range is cut into one piece per warp of the grid: segments far longer than a piece are reduced piecewise
    (scratch accumulators + last-arriver finish), runs of short segments take the lane-per-segment path, and everything in
    between is mixed.  Every result must still be the oracle's bits; leading / trailing / interior empty segments included."""
    import torch
    rng = np.random.default_rng(11)
    layouts = {
        "two_huge": [1_500_000, 0, 2_000_003],
        "huge_then_short": [3_000_000] + [7] * 500 + [0] * 40,
        "short_runs_and_long": [0, 0] + list(rng.integers(1, 65, size=3000)) + [100_000] + list(rng.integers(60, 70, size=400))
                               + [65, 64, 63, 1, 0, 0, 0],
        "all_16": [16] * 50_000,
        "thresholds": [63, 64, 65, 127, 128, 129, 3, 3, 3, 3, 64, 64, 64, 64, 65, 2047, 2048, 2049, 1, 1, 1],
        "few_elements": [1, 0, 2, 0, 0, 3],
    }
    for name, lengths in layouts.items():
        seg = offsets(np.array(lengths, dtype=np.int64))
        total = int(seg[-1])
        for kind in ("wide", "narrow"):
            a = make_data(max(total, 1), kind, 21)
            b = make_data(max(total, 1), "illcond", 22)
            da, db, ds = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(seg).cuda()
            gs, st = gpu.exsum_segments(da, ds, want_status=True)
            gd = gpu.exdot_segments(da, db, ds, round_mode=1)
            gs, gd = gs.cpu().numpy(), gd.cpu().numpy()
            assert int(st.abs().sum()) == 0
            check = range(len(lengths)) if len(lengths) <= 600 else list(rng.choice(len(lengths), size=300, replace=False)) + [0, len(lengths) - 1]
            for i in check:
                lo, hi = int(seg[i]), int(seg[i + 1])
                ws = oracle.exsum(a[lo:hi], fpe=0)[0] if hi > lo else 0.0
                wd = oracle.exdot(a[lo:hi], b[lo:hi], fpe=0, round_mode=1)[0] if hi > lo else 0.0
                assert np.float64(gs[i]).view(np.uint64) == np.float64(ws).view(np.uint64), (name, kind, i, lengths[i], "sum")
                assert np.float64(gd[i]).view(np.uint64) == np.float64(wd).view(np.uint64), (name, kind, i, lengths[i], "dot")
    # a run of launches leaves the scratch accumulators clean: same answer every time
    seg = offsets(np.array(layouts["two_huge"], dtype=np.int64))
    a = make_data(int(seg[-1]), "wide", 23)
    da, ds = torch.from_numpy(a).cuda(), torch.from_numpy(seg).cuda()
    first = gpu.exsum_segments(da, ds).cpu().numpy()
    for _ in range(5):
        assert (gpu.exsum_segments(da, ds).cpu().numpy().view(np.uint64) == first.view(np.uint64)).all()
    # offsets that do not start at zero
    seg2 = seg + 0
    seg2[0] = 12345
    got = gpu.exsum_segments(da, torch.from_numpy(seg2).cuda()).cpu().numpy()
    assert got[0] == oracle.exsum(a[12345:int(seg2[1])], fpe=0)[0] and got[2] == first[2]
