"""ctypes loader for the CHECKERS under oracle/ (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (exblas_b200) never does.

* ``Oracle``    -- oracle/liboracle.so, our C restatement (oracle/exblas_oracle.c).
* ``Reference`` -- oracle/_ref/libexblas_ref.so, the unmodified reference CPU code compiled by
                   oracle/Makefile from /root/reference (absent on machines without the prebuilt .so).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libexblas_ref.so")

_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> None:
    """Compile liboracle.so (always) and _ref/libexblas_ref.so (when /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or \
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "exblas_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, os.path.join(HERE, "liboracle.so")])
    if os.path.isdir("/root/reference/src/cpu/blas/blas1"):
        if force or not os.path.exists(REF_SO):
            subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
        # the reference's own GPU test mains, linked against the product library (needs it built first)
        prod = os.path.join(HERE, "..", "exblas_b200", "libexblas_b200.so")
        if os.path.exists(prod) and (force or not os.path.exists(os.path.join(HERE, "_ref", "test.exsum.gpu"))):
            subprocess.check_call(["make", "-s", "-C", HERE, "ref_tests"])


def _as_f64(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class Oracle:
    """C restatement of the reference algorithm (39-limb GPU layout unless stated)."""

    def __init__(self) -> None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.oracle_exsum.restype = C.c_double
        L.oracle_exsum.argtypes = [C.c_int64, _dp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, _lp]
        L.oracle_exdot.restype = C.c_double
        L.oracle_exdot.argtypes = [C.c_int64, _dp, C.c_int64, C.c_int64, _dp, C.c_int64, C.c_int64,
                                   C.c_int, C.c_int, C.c_int, _lp]
        L.oracle_superacc_limbs.restype = C.c_double
        L.oracle_superacc_limbs.argtypes = [C.c_int64, _dp, C.c_int, _lp, C.c_int]
        L.oracle_round_limbs.restype = C.c_double
        L.oracle_round_limbs.argtypes = [_lp, C.c_int, C.c_int]
        L.oracle_merge_round.restype = C.c_double
        L.oracle_merge_round.argtypes = [_lp, C.c_int, C.c_int, C.c_int, _lp]
        L.oracle_exsum_parallel.restype = C.c_double
        L.oracle_exsum_parallel.argtypes = [C.c_int64, _dp, C.c_int, C.c_int, C.c_int]
        L.oracle_exgemv.restype = C.c_int
        L.oracle_exgemv.argtypes = [C.c_char, C.c_int64, C.c_int64, C.c_double, _dp, C.c_int64, _dp, C.c_int64,
                                    C.c_double, _dp, C.c_int64, C.c_int, C.c_int, C.c_int]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        self.L = L

    def exsum(self, a, inca=1, offset=0, fpe=0, early_exit=False, round_mode=0, n=None):
        a = _as_f64(a)
        if n is None:
            n = 0 if a.size <= offset else (a.size - offset + inca - 1) // inca
        limbs = np.zeros(39, dtype=np.int64)
        r = self.L.oracle_exsum(n, _ptr(a), inca, offset, fpe, int(early_exit), round_mode,
                                limbs.ctypes.data_as(_lp))
        return r, limbs

    def exdot(self, a, b, inca=1, offa=0, incb=1, offb=0, fpe=0, early_exit=False, round_mode=0, n=None):
        a = _as_f64(a)
        b = _as_f64(b)
        if n is None:
            n = 0 if a.size <= offa else (a.size - offa + inca - 1) // inca
        limbs = np.zeros(39, dtype=np.int64)
        r = self.L.oracle_exdot(n, _ptr(a), inca, offa, _ptr(b), incb, offb, fpe, int(early_exit),
                                round_mode, limbs.ctypes.data_as(_lp))
        return r, limbs

    def superacc_limbs(self, a, layout=1, round_mode=0):
        a = _as_f64(a)
        limbs = np.zeros(41 if layout == 0 else 39, dtype=np.int64)
        r = self.L.oracle_superacc_limbs(a.size, _ptr(a), layout, limbs.ctypes.data_as(_lp), round_mode)
        return r, limbs

    def round_limbs(self, limbs, round_mode=0):
        limbs = np.ascontiguousarray(limbs, dtype=np.int64)
        layout = 0 if limbs.size == 41 else 1
        return self.L.oracle_round_limbs(limbs.ctypes.data_as(_lp), layout, round_mode)

    def merge_round(self, limbs_per_rank, round_mode=0):
        arr = np.ascontiguousarray(limbs_per_rank, dtype=np.int64)
        nr, nl = arr.shape
        out = np.zeros(nl, dtype=np.int64)
        r = self.L.oracle_merge_round(arr.ctypes.data_as(_lp), nr, 0 if nl == 41 else 1, round_mode,
                                      out.ctypes.data_as(_lp))
        return r, out

    def exgemv(self, trans, m, n, alpha, a, lda, x, incx, beta, y, incy, fpe=0, early_exit=False, round_mode=0):
        """returns the new y (input y is not modified)"""
        a = _as_f64(a)
        x = _as_f64(x)
        out = np.array(y, dtype=np.float64, copy=True)
        self.L.oracle_exgemv(trans.encode(), m, n, alpha, _ptr(a), lda, _ptr(x), incx, beta, _ptr(out), incy, fpe,
                             int(early_exit), round_mode)
        return out

    def exsum_parallel(self, a, fpe=8, early_exit=True, round_mode=0):
        a = _as_f64(a)
        return self.L.oracle_exsum_parallel(a.size, _ptr(a), fpe, int(early_exit), round_mode)

    def max_threads(self) -> int:
        return int(self.L.oracle_max_threads())

    def use_all_cores(self) -> int:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.L.oracle_set_threads(n)
        return self.max_threads()


class Reference:
    """The unmodified reference CPU library (41-limb CPU layout)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self) -> None:
        if not os.path.exists(REF_SO):
            build()
        L = C.CDLL(REF_SO)
        L.ref_exsum.restype = C.c_double
        L.ref_exsum.argtypes = [C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_superacc_limbs.restype = C.c_double
        L.ref_superacc_limbs.argtypes = [_dp, C.c_long, _lp]
        L.ref_round_limbs.restype = C.c_double
        L.ref_round_limbs.argtypes = [_lp, C.c_int]
        L.ref_exdot_superacc.restype = C.c_double
        L.ref_exdot_superacc.argtypes = [C.c_long, _dp, _dp, _lp]
        L.ref_exsum_mpfr.restype = C.c_double
        L.ref_exsum_mpfr.argtypes = [C.c_long, _dp]
        L.ref_exdot_mpfr.restype = C.c_double
        L.ref_exdot_mpfr.argtypes = [C.c_long, _dp, _dp]
        L.ref_exgemv_mpfr.argtypes = [C.c_char, C.c_int, C.c_int, C.c_double, _dp, C.c_int, _dp, C.c_int, C.c_double,
                                      _dp, C.c_int, _dp]
        L.ref_omp_max_threads.restype = C.c_int
        L.ref_omp_set_threads.argtypes = [C.c_int]
        L.ref_limb_count.restype = C.c_int
        L.ref_f_words.restype = C.c_int
        L.ref_srand.argtypes = [C.c_uint]
        L.ref_init_naive.argtypes = [C.c_int, _dp]
        L.ref_init_fpuniform.argtypes = [C.c_int, _dp, C.c_int, C.c_int]
        L.ref_init_ill_cond.argtypes = [C.c_int, _dp, C.c_double]
        self.L = L

    def exsum(self, a, fpe=0, early_exit=False, parallel=True, inca=1, offset=0, n=None):
        a = _as_f64(a)
        n = a.size if n is None else n
        return self.L.ref_exsum(n, _ptr(a), inca, offset, fpe, int(early_exit), int(parallel))

    def superacc_limbs(self, a):
        a = _as_f64(a)
        limbs = np.zeros(41, dtype=np.int64)
        r = self.L.ref_superacc_limbs(_ptr(a), a.size, limbs.ctypes.data_as(_lp))
        return r, limbs

    def round_limbs(self, limbs41):
        limbs = np.ascontiguousarray(limbs41, dtype=np.int64)
        assert limbs.size == 41
        return self.L.ref_round_limbs(limbs.ctypes.data_as(_lp), 41)

    def exdot_superacc(self, a, b):
        a = _as_f64(a)
        b = _as_f64(b)
        limbs = np.zeros(41, dtype=np.int64)
        r = self.L.ref_exdot_superacc(a.size, _ptr(a), _ptr(b), limbs.ctypes.data_as(_lp))
        return r, limbs

    def exgemv_mpfr(self, trans, m, n, alpha, a, lda, x, incx, beta, y, incy):
        a = _as_f64(a)
        x = _as_f64(x)
        y = _as_f64(y)
        nout = n if trans == "T" else m
        out = np.zeros(nout, dtype=np.float64)
        self.L.ref_exgemv_mpfr(trans.encode(), m, n, alpha, _ptr(a), lda, _ptr(x), incx, beta, _ptr(y), incy, _ptr(out))
        return out

    def exsum_mpfr(self, a):
        a = _as_f64(a)
        return self.L.ref_exsum_mpfr(a.size, _ptr(a))

    def exdot_mpfr(self, a, b):
        a = _as_f64(a)
        b = _as_f64(b)
        return self.L.ref_exdot_mpfr(a.size, _ptr(a), _ptr(b))

    def max_threads(self) -> int:
        return int(self.L.ref_omp_max_threads())

    def use_all_cores(self) -> int:
        """torchrun exports OMP_NUM_THREADS=1; the CPU baseline must use every host core."""
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.L.ref_omp_set_threads(n)
        return self.max_threads()

    def init_naive(self, n):
        a = np.empty(n, dtype=np.float64)
        self.L.ref_init_naive(n, _ptr(a))
        return a

    def init_fpuniform(self, n, rng, emax, seed=1):
        a = np.empty(n, dtype=np.float64)
        self.L.ref_srand(seed)
        self.L.ref_init_fpuniform(n, _ptr(a), rng, emax)
        return a

    def init_ill_cond(self, n, c, seed=1):
        a = np.empty(n, dtype=np.float64)
        self.L.ref_srand(seed)
        self.L.ref_init_ill_cond(n, _ptr(a), float(c))
        return a


def cpu41_to_gpu39(limbs41: np.ndarray) -> np.ndarray:
    """cpu limb j+1 == gpu limb j; requires cpu limbs 0 and 40 to be empty."""
    limbs41 = np.asarray(limbs41, dtype=np.int64)
    assert limbs41[0] == 0, "value has bits below 2^-1040: outside the 39-limb layout"
    top = int(limbs41[40])
    assert top in (0, -1), "value beyond the 39-limb layout"
    out = limbs41[1:40].copy()
    if top == -1:          # negative: the cpu layout carried the sign one limb further up
        out[38] -= 1 << 52
    return out
