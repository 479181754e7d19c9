"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against the oracle.

Bit-exact everywhere: the value (both round modes) AND the 39 normalised limbs.  Sizes here are
those the scalar oracle finishes in seconds; full-size (2^30) checks are in test_gpu_fullsize.py.
"""
import math
import os
import subprocess

import numpy as np
import pytest

from helpers import VARIANTS_DOT, VARIANTS_SUM, cpu41_to_gpu39, same_double
from exblas_b200 import common as cm

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_golden_exsum_through_abi(gpu, golden):
    for name in golden["sum_cases"]:
        a = golden[f"{name}/a"]
        l39 = cpu41_to_gpu39(golden[f"{name}/ref_limbs41"])
        ref_round = float(golden[f"{name}/ref_round"][0])
        mpfr = float(golden[f"{name}/mpfr"][0])
        d = dev(a)
        for fpe, ee in VARIANTS_SUM:
            for src in (d, a):                                   # device pointer and host pointer
                v, l = gpu.exsum(a.size, src, 1, 0, fpe, ee, want_limbs=True)
                assert same_double(v, ref_round), (name, fpe, ee)
                assert (l == l39).all(), (name, fpe, ee)
            assert same_double(gpu.exsum(a.size, d, 1, 0, fpe, ee, round_mode=1), mpfr), (name, fpe, ee)
        assert gpu.last_status() == 0


def test_golden_exdot_through_abi(gpu, golden):
    for name in golden["dot_cases"]:
        a, b = golden[f"{name}/a"], golden[f"{name}/b"]
        l39 = cpu41_to_gpu39(golden[f"{name}/ref_limbs41"])
        ref_round = float(golden[f"{name}/ref_round"][0])
        mpfr = float(golden[f"{name}/mpfr"][0])
        da, db = dev(a), dev(b)
        for fpe, ee in VARIANTS_DOT:
            for sa, sb in ((da, db), (a, b)):
                v, l = gpu.exdot(a.size, sa, 1, 0, sb, 1, 0, fpe, ee, want_limbs=True)
                assert same_double(v, ref_round), (name, fpe, ee)
                assert (l == l39).all(), (name, fpe, ee)
            assert same_double(gpu.exdot(a.size, da, 1, 0, db, 1, 0, fpe, ee, round_mode=1), mpfr), (name, fpe, ee)


SIZES = [0, 1, 2, 3, 4, 5, 7, 8, 31, 33, 127, 2047, 2048, 2049, 8191, 8192, 8193, 65535, 65536, 65537, 100003,
         1 << 20, (1 << 20) + 1, 148 * 8192, 148 * 8192 + 3 * 8192 + 17]


@pytest.mark.parametrize("kind", ["naive", "loguniform", "loguniform_signed", "illcond", "cancel"])
def test_exsum_vs_oracle_sizes(gpu, oracle, kind):
    for n in SIZES:
        if kind == "naive":
            a = cm.init_naive(n)
        elif kind == "loguniform":
            a = cm.init_fpuniform(n, 664, 332, seed=n + 1)
        elif kind == "loguniform_signed":
            a = cm.init_fpuniform(n, 664, 332, seed=n + 2, neg_ratio=2)
        elif kind == "illcond":
            a = cm.init_ill_cond(n, 1e32, seed=n + 3) if n >= 2 else cm.init_naive(n)
        else:
            if n < 4:
                continue
            a = cm.cancelling_pair(2 * (n // 2), "sum", seed=n)
        v0, l0 = oracle.exsum(a, fpe=0, round_mode=0)
        v1, _ = oracle.exsum(a, fpe=0, round_mode=1)
        d = dev(a)
        for fpe, ee in VARIANTS_SUM:
            v, l = gpu.exsum(a.size, d, 1, 0, fpe, ee, want_limbs=True)
            assert same_double(v, v0) and (l == l0).all(), (kind, n, fpe, ee)
        assert same_double(gpu.exsum(a.size, d, 1, 0, 4, False, round_mode=1), v1), (kind, n)
        if kind == "cancel":
            assert v1 == 1.5


@pytest.mark.parametrize("kind", ["loguniform_signed", "illcond", "cancel"])
def test_exdot_vs_oracle_sizes(gpu, oracle, kind):
    for n in [0, 1, 3, 4, 9, 1023, 4096, 4097, 65536, 65539, 300001, 148 * 4096 + 5]:
        if kind == "loguniform_signed":
            a = cm.init_fpuniform(n, 664, 332, seed=n + 5, neg_ratio=2)
            b = cm.init_fpuniform(n, 664, 332, seed=n + 6, neg_ratio=3)
        elif kind == "illcond":
            if n < 2:
                continue
            a = cm.init_ill_cond(n, 1e32, seed=n + 5)
            b = cm.init_ill_cond(n, 1e32, seed=n + 6)
        else:
            if n < 4:
                continue
            a, b = cm.cancelling_pair(2 * (n // 2), "dot", seed=n)
        v0, l0 = oracle.exdot(a, b, fpe=0, round_mode=0)
        v1, _ = oracle.exdot(a, b, fpe=0, round_mode=1)
        da, db = dev(a), dev(b)
        for fpe, ee in VARIANTS_DOT:
            v, l = gpu.exdot(a.size, da, 1, 0, db, 1, 0, fpe, ee, want_limbs=True)
            if n == 0:
                assert v == 0.0
                continue
            assert same_double(v, v0) and (l == l0).all(), (kind, n, fpe, ee)
        if n:
            assert same_double(gpu.exdot(a.size, da, 1, 0, db, 1, 0, 3, False, round_mode=1), v1)
            if kind == "cancel":
                assert v1 == 1.5


def test_offsets_strides_alignment(gpu, oracle):
    """inca / offset follow the reference GPU kernels: a[offset + i*inca], Ng = element count
    (ExSUM.FPE.cl:298-299); also every 8-byte misalignment of the base pointers."""
    n = 70001
    a = cm.init_fpuniform(n, 664, 332, seed=77, neg_ratio=2)
    b = cm.init_ill_cond(n, 1e32, seed=78)
    da, db = dev(a), dev(b)
    for off in (0, 1, 2, 3, 5):
        for inc in (1, 2, 3, 7):
            m = (n - off + inc - 1) // inc
            v0, l0 = oracle.exsum(a, inca=inc, offset=off, n=m, fpe=0)
            for fpe, ee in [(0, False), (4, False), (8, True)]:
                for src in (da, a):
                    v, l = gpu.exsum(m, src, inc, off, fpe, ee, want_limbs=True)
                    assert same_double(v, v0) and (l == l0).all(), (off, inc, fpe, ee)
    for offa, offb, inca, incb in [(0, 0, 1, 1), (1, 1, 1, 1), (1, 2, 1, 1), (3, 0, 1, 1), (0, 0, 2, 1), (5, 1, 3, 2)]:
        m = min((n - offa + inca - 1) // inca, (n - offb + incb - 1) // incb)
        v0, l0 = oracle.exdot(a, b, inca=inca, offa=offa, incb=incb, offb=offb, fpe=0, n=m)
        for fpe, ee in [(0, False), (3, False), (8, True)]:
            for sa, sb in ((da, db), (a, b)):
                v, l = gpu.exdot(m, sa, inca, offa, sb, incb, offb, fpe, ee, want_limbs=True)
                assert same_double(v, v0) and (l == l0).all(), (offa, offb, inca, incb, fpe, ee)


def test_host_streaming_chunks(gpu, oracle):
    """host inputs are streamed in chunks that all add into one device accumulator"""
    n = 1_000_003
    a = cm.init_fpuniform(n, 664, 332, seed=5, neg_ratio=2)
    b = cm.init_fpuniform(n, 300, 150, seed=6, neg_ratio=2)
    v0, l0 = oracle.exsum(a, fpe=0)
    d0, dl0 = oracle.exdot(a, b, fpe=0)
    try:
        for chunk in (4096, 100_000, 1 << 23):
            gpu.set_option("host_chunk_elems", chunk)
            v, l = gpu.exsum(n, a, 1, 0, 4, False, want_limbs=True)
            assert same_double(v, v0) and (l == l0).all(), chunk
            v, l = gpu.exdot(n, a, 1, 0, b, 1, 0, 4, True, want_limbs=True)
            assert same_double(v, d0) and (l == dl0).all(), chunk
            v, l = gpu.exsum(n // 3, a, 3, 1, 0, False, want_limbs=True)
            w, wl = oracle.exsum(a, inca=3, offset=1, n=n // 3, fpe=0)
            assert same_double(v, w) and (l == wl).all(), chunk
        # pageable memory (numpy) takes the pinned bounce ring filled by the copy threads; pinned memory and
        # host_threads = 1 (the driver's own pageable staging) take the direct H2D path: all three must agree
        import torch
        pin_a = torch.from_numpy(a).pin_memory()
        pin_b = torch.from_numpy(b).pin_memory()
        for threads, pchunk in ((0, 1 << 21), (3, 7001), (8, 65536), (1, 1 << 21)):
            gpu.set_option("host_threads", threads)
            gpu.set_option("pageable_chunk_elems", pchunk)
            for src_a, src_b in ((a, b), (pin_a, pin_b)):
                v, l = gpu.exsum(n, src_a, 1, 0, 3, False, want_limbs=True)
                assert same_double(v, v0) and (l == l0).all(), (threads, pchunk)
                v, l = gpu.exdot(n, src_a, 1, 0, src_b, 1, 0, 0, False, want_limbs=True)
                assert same_double(v, d0) and (l == dl0).all(), (threads, pchunk)
            v, l = gpu.exsum(n // 3, a, 3, 1, 8, True, want_limbs=True)
            w, wl = oracle.exsum(a, inca=3, offset=1, n=n // 3, fpe=0)
            assert same_double(v, w) and (l == wl).all(), (threads, pchunk)
    finally:
        gpu.set_option("host_chunk_elems", 1 << 23)
        gpu.set_option("host_threads", 0)
        gpu.set_option("pageable_chunk_elems", 1 << 21)


def test_result_independent_of_launch_shape(gpu, oracle):
    """grid / block size are performance knobs only"""
    n = 3_000_017
    a = cm.init_fpuniform(n, 664, 332, seed=15, neg_ratio=2)
    v0, l0 = oracle.exsum(a, fpe=0)
    d = dev(a)
    try:
        for T in (32, 64, 128, 256, 384, 512):
            for blocks in (0, 1, 7, 148, 296, 1000):
                gpu.set_option("block_threads", T)
                gpu.set_option("blocks", blocks)
                for fpe, ee in [(0, False), (3, False), (8, True)]:
                    v, l = gpu.exsum(n, d, 1, 0, fpe, ee, want_limbs=True)
                    assert same_double(v, v0) and (l == l0).all(), (T, blocks, fpe, ee)
        # hand-set small CTAs on SHORT vectors: the touched-row tracking path with fewer threads than limbs
        for n2 in (100, 5000, 70001):
            w0, wl0 = oracle.exsum(a[:n2], fpe=0)
            for T in (32, 64, 96):
                for blocks in (1, 3):
                    gpu.set_option("block_threads", T)
                    gpu.set_option("blocks", blocks)
                    for fpe, ee in [(0, False), (4, False), (8, True)]:
                        v, l = gpu.exsum(n2, d, 1, 0, fpe, ee, want_limbs=True)
                        assert same_double(v, w0) and (l == wl0).all(), (n2, T, blocks, fpe, ee)
    finally:
        gpu.set_option("auto_shape", 1)


def test_small_and_mid_sizes_every_shape_regime(gpu, oracle):
    """The size-dependent launch shapes (one CTA published from shared memory, 256-thread CTAs, equal tiles +
    evenly spread remainder) around each of their thresholds, ExSUM and ExDOT, positive and signed data."""
    nmax = (1 << 22) + 4099
    a = cm.init_fpuniform(nmax, 664, 332, seed=21, neg_ratio=2)
    b = cm.init_ill_cond(nmax, 1e32, seed=22)
    da, db = dev(a), dev(b)
    sizes = [1, 5, 511, 2048, 4096, 8191, 8192, 8193, 8197, 1 << 14, (1 << 16) + 1, 148 * 1024 * 4, 148 * 1024 * 4 + 5,
             (1 << 20) - 3, 1 << 21, (1 << 22) - 1, 1 << 22, (1 << 22) + 1, nmax]
    for n in sizes:
        v0, l0 = oracle.exsum(a[:n], fpe=0)
        d0, dl0 = oracle.exdot(a[:n], b[:n], fpe=0)
        for fpe, ee in [(0, False), (3, False), (8, True)]:
            for track in (0, 1 << 20):          # the kernel fpe selects / the library's default routing of mid sizes
                gpu.set_option("track_max_elems", track)
                v, l = gpu.exsum(n, da, 1, 0, fpe, ee, want_limbs=True)
                assert same_double(v, v0) and (l == l0).all(), ("sum", n, fpe, ee, track)
            gpu.set_option("track_max_elems", 0)
            v, l = gpu.exdot(n, da, 1, 0, db, 1, 0, fpe, ee, want_limbs=True)
            assert same_double(v, d0) and (l == dl0).all(), ("dot", n, fpe, ee)
        # misaligned start: alignment head + vector region + tail
        v, l = gpu.exsum(n - 1, da, 1, 1, 4, False, want_limbs=True) if n > 1 else (0.0, None)
        if n > 1:
            w, wl = oracle.exsum(a[1:n], fpe=0)
            assert same_double(v, w) and (l == wl).all(), ("sum+1", n)


def test_exdot_handoff_to_the_window_loop(gpu, oracle):
    """ExDOT with fpe >= 3: a warp whose first tile thrashes the expansion hands the rest of its rows to the 5-digit window
    loop (option dot_handoff_tiles; default 16 tiles per CTA, here 1 so that vectors of a few million elements take it).
    Ill-conditioned products (the case it is for), wide-range products (too wide for any window: no hand-off) and narrow
    ones (no thrashing: no hand-off) must all give the oracle's value and limbs, with and without the hand-off."""
    nmax = 3 * (1 << 21) + 4111
    pairs = {"illcond": (cm.init_ill_cond(nmax, 1e32, seed=31), cm.init_ill_cond(nmax, 1e32, seed=32)),
             "wide": (cm.init_fpuniform(nmax, 664, 332, seed=33, neg_ratio=2), cm.init_fpuniform(nmax, 200, 100, seed=34, neg_ratio=2)),
             "narrow": (cm.init_fpuniform(nmax, 10, 5, seed=35, neg_ratio=2), cm.init_fpuniform(nmax, 10, 5, seed=36, neg_ratio=2))}
    try:
        for kind, (a, b) in pairs.items():
            da, db = dev(a), dev(b)
            for n in ((1 << 21) + 5, 1 << 22, nmax):
                d0, dl0 = oracle.exdot(a[:n], b[:n], fpe=0)
                for handoff in (1, 0):
                    gpu.set_option("dot_handoff_tiles", handoff)
                    for fpe, ee in [(3, False), (4, False), (8, False), (6, True), (8, True)]:
                        v, l = gpu.exdot(n, da, 1, 0, db, 1, 0, fpe, ee, want_limbs=True)
                        assert same_double(v, d0) and (l == dl0).all(), (kind, n, handoff, fpe, ee)
    finally:
        gpu.set_option("dot_handoff_tiles", 16)


def test_partial_sums_near_the_top_of_the_layout(gpu):
    """Inputs below 2^988 whose partial sums inside one expansion exceed it: every fpe must give the exact sum
    (the expansion flush adds such a partial sum to limb 38 as an integer)."""
    from fractions import Fraction
    from helpers import limbs_from_fraction
    for n, sign_every in [(1024, 0), (1024, 3), (300, 0), (4096 + 64, 2)]:
        k = np.arange(n, dtype=np.float64)
        a = (2.0 ** 987) * (1.0 + k * 2.0 ** -30)
        if sign_every:
            a[::sign_every] *= -1.0
        exact = sum((Fraction(float(x)) for x in a), Fraction(0))
        want = limbs_from_fraction(exact)
        d = dev(a)
        for fpe, ee in [(0, False), (2, False), (4, False), (8, False), (4, True), (8, True)]:
            gpu.exsum_async(n, d, 1, 0, fpe, ee, 1)
            v, l, st = gpu.fetch()
            assert st == 0 and (l == want).all() and v == float(exact), (n, sign_every, fpe, ee, st)


def test_permutation_invariance_and_rerun(gpu):
    """RNGExample's strong-reproducibility pattern (RNGExample.cpp:300-333): shuffle between runs,
    every bit of the result must stay the same."""
    import torch
    n = 2_000_000
    a = cm.init_fpuniform(n, 664, 332, seed=31, neg_ratio=2, device="cuda")
    v0, l0, _ = _sum(gpu, a, 4, False)
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    for it in range(6):
        p = a[torch.randperm(n, device="cuda", generator=g)]
        for fpe, ee in [(0, False), (4, False), (8, True)]:
            v, l, _ = _sum(gpu, p, fpe, ee)
            assert same_double(v, v0) and (l == l0).all()
    for it in range(50):
        v, l, _ = _sum(gpu, a, 8, True)
        assert same_double(v, v0) and (l == l0).all()


def _sum(gpu, t, fpe, ee, rm=0):
    gpu.exsum_async(t.numel(), t, 1, 0, fpe, ee, rm)
    return gpu.fetch()


def test_specials_and_domain(gpu, oracle):
    import exblas_b200 as xb
    base = cm.init_fpuniform(5000, 100, 50, seed=1, neg_ratio=2)

    def run(extra, fpe=4, ee=False):
        a = np.concatenate([base[:2500], np.array(extra, dtype=np.float64), base[2500:]])
        gpu.exsum_async(a.size, dev(a), 1, 0, fpe, ee, 1)
        return gpu.fetch()

    for fpe, ee in [(0, False), (4, False), (8, True)]:
        v, _, st = run([np.nan], fpe, ee)
        assert math.isnan(v) and st & xb.ST_NAN
        v, _, st = run([np.inf], fpe, ee)
        assert v == math.inf and st == xb.ST_POSINF
        v, _, st = run([-np.inf, -np.inf], fpe, ee)
        assert v == -math.inf and st == xb.ST_NEGINF
        v, _, st = run([np.inf, -np.inf], fpe, ee)
        assert math.isnan(v)
        v, _, st = run([1e300], fpe, ee)                     # >= 2^988: outside the 39-limb layout
        assert st & xb.ST_TOOLARGE
        v, _, st = run([0.0, -0.0, 0.0], fpe, ee)            # zeros are ordinary
        assert st == 0 and v == math.fsum(base)
        v, _, st = run([2.0 ** -1040, -(2.0 ** -1030), 2.0 ** -1000], fpe, ee)   # tiny but representable
        assert st == 0 and v == math.fsum(np.concatenate([base, [2.0 ** -1040, -(2.0 ** -1030), 2.0 ** -1000]]))
        v, _, st = run([5e-324], fpe, ee)                    # below 2^-1040: truncated and flagged
        assert st & xb.ST_TOOSMALL
    # after a flagged call the handle is clean again
    v, _, st = run([])
    assert st == 0 and v == math.fsum(base)
    # ExDOT specials
    a = np.array([1.0, 2.0, np.inf, 3.0]); b = np.array([1.0, 1.0, 0.0, 1.0])
    gpu.exdot_async(4, dev(a), 1, 0, dev(b), 1, 0, 3, False, 1)
    v, _, st = gpu.fetch()
    assert math.isnan(v)                                      # inf * 0
    a = np.array([1e200, 2.0]); b = np.array([1e200, 1.0])
    gpu.exdot_async(2, dev(a), 1, 0, dev(b), 1, 0, 0, False, 1)
    v, _, st = gpu.fetch()
    assert st & xb.ST_TOOLARGE
    a = np.array([2.0 ** -500, 3.0, 0.0]); b = np.array([2.0 ** -520, 5.0, 1e300])
    gpu.exdot_async(3, dev(a), 1, 0, dev(b), 1, 0, 4, True, 1)
    v, _, st = gpu.fetch()
    assert st == 0 and v == 15.0 + 2.0 ** -1020               # tiny exact product, zero times huge


def test_argument_errors(gpu):
    import exblas_b200 as xb
    a = np.ones(16)
    with pytest.raises(ValueError):
        gpu.exsum(16, a, 0, 0, 4)                            # inc < 1: rejected by the host mirror ...
    import ctypes as C
    res = C.c_double()
    rc = gpu.lib.exblas_b200_exsum(gpu._h, a.ctypes.data, 16, 0, 0, 4, 0, 0, C.byref(res))
    assert rc == -1                                          # ... and EINVAL at the C ABI
    rc = gpu.lib.exblas_b200_exsum(gpu._h, a.ctypes.data, 16, 1, 0, -2, 0, 0, C.byref(res))
    assert rc == -1                                          # fpe < 0
    with pytest.raises(ValueError):
        gpu.exsum(17, a, 1, 0, 4)                            # reads past the end: caught by the host mirror
    with pytest.raises(xb.ExblasB200Error):
        gpu.exsum_async(16, a, 1, 0, 4)                      # async entry points need device pointers


def test_blas1_cpp_dropin(gpu, oracle, tmp_path):
    """A C++ program written against the reference's blas1.hpp, linked to libexblas_b200.so."""
    src = tmp_path / "main.cpp"
    src.write_text(r'''
#include "blas1.hpp"
#include <cstdio>
#include <vector>
int main() {
    const int n = 1 << 16;
    std::vector<double> a(n), b(n);
    for (int i = 0; i < n; ++i) { a[i] = 1.1; b[i] = (i % 7) - 3.25; }
    printf("%a %a %a %a %a\n", exsum(n, a.data(), 1, 0, 0), exsum(n, a.data(), 1, 0, 4), exsum(n, a.data(), 1, 0, 8, true),
           exdot(n, a.data(), 1, 0, b.data(), 1, 0, 3), exdot(n, a.data(), 1, 0, b.data(), 1, 0, 8, true));
    return 0;
}''')
    exe = tmp_path / "main"
    libdir = os.path.join(ROOT, "exblas_b200")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lexblas_b200", f"-Wl,-rpath,{libdir}",
                           "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    out = subprocess.check_output([str(exe)], text=True).split()
    a = np.full(1 << 16, 1.1)
    b = (np.arange(1 << 16) % 7) - 3.25
    s = oracle.exsum(a, fpe=0)[0]
    d = oracle.exdot(a, b, fpe=0)[0]
    assert [float.fromhex(x) for x in out] == [s, s, s, d, d]


def test_superacc_only_window_variants(gpu, oracle):
    """fpe 0 (superaccumulator-only kernel): the register windows -- off, narrow only, narrow then wide, wide only --
    must all give the oracle's value AND limbs, on data that stays in the narrow window, needs the wide one,
    drifts so that the windows must follow, or fits neither."""
    n = (1 << 24) + 37                      # ~55 rows of 2048 per CTA: every loop of the kernel runs, plus a ragged tail
    rng = np.random.default_rng(3)
    narrow = cm.init_fpuniform(n, 10, 5, seed=1, neg_ratio=2)
    ill = cm.init_ill_cond(n, 1e32, seed=2)
    ill2 = cm.init_ill_cond(n, 1e32, seed=3)
    drift = narrow * np.exp2((np.arange(n) // (1 << 19)).astype(np.float64) * 9.0 - 100.0)   # moves 9 binades every 2^19 elements
    wide = cm.init_fpuniform(n, 664, 332, seed=4, neg_ratio=2)
    spiky = narrow.copy()
    spiky[rng.integers(0, n, size=200)] *= 2.0 ** 300                                        # rare far outliers
    spiky[rng.integers(0, n, size=200)] = 0.0
    try:
        for name, a, b in (("narrow", narrow, narrow[::-1].copy()), ("ill", ill, narrow), ("ill x ill", ill, ill2), ("drift", drift, narrow),
                           ("wide", wide, narrow), ("spiky", spiky, ill)):
            vs, ls = oracle.exsum(a, fpe=0, round_mode=0)
            vd, ld = oracle.exdot(a, b, fpe=0, round_mode=0)
            da, db = dev(a), dev(b)
            for window in (0, 1, 2, 3):
                gpu.set_option("window", window)
                v, l = gpu.exsum(n, da, 1, 0, 0, False, want_limbs=True)
                assert same_double(v, vs) and (l == ls).all(), ("exsum", name, window)
                v, l = gpu.exdot(n, da, 1, 0, db, 1, 0, 0, False, want_limbs=True)
                assert same_double(v, vd) and (l == ld).all(), ("exdot", name, window)
                assert gpu.last_status() == 0
    finally:
        gpu.set_option("window", 2)
