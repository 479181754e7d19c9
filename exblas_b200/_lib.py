"""ctypes binding of include/exblas_b200.h.  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EXBLAS_B200_LIB") or os.path.join(HERE, "libexblas_b200.so")   # env override: tuning builds only

LIMBS = 39
ROUND_REFERENCE = 0
ROUND_EXACT = 1

ST_NAN, ST_POSINF, ST_NEGINF, ST_TOOLARGE, ST_TOOSMALL, ST_PEERTIMEOUT = 1, 2, 4, 8, 16, 32

_dp = C.c_void_p           # data pointers are passed as raw addresses (host or device)
_i64 = C.c_int64
_h = C.c_void_p

# every symbol include/exblas_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "exblas_b200_version": (C.c_int, []),
    "exblas_b200_strerror": (C.c_char_p, [C.c_int]),
    "exblas_b200_create": (C.c_int, [C.POINTER(_h), C.c_int]),
    "exblas_b200_destroy": (C.c_int, [_h]),
    "exblas_b200_set_stream": (C.c_int, [_h, C.c_void_p]),
    "exblas_b200_set_option": (C.c_int, [_h, C.c_char_p, _i64]),
    "exblas_b200_exsum": (C.c_int, [_h, _dp, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "exblas_b200_exdot": (C.c_int, [_h, _dp, _i64, _i64, _dp, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_double)]),
    "exblas_b200_exsum_limbs": (C.c_int, [_h, _dp, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(_i64), C.POINTER(C.c_double)]),
    "exblas_b200_exdot_limbs": (C.c_int, [_h, _dp, _i64, _i64, _dp, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(_i64), C.POINTER(C.c_double)]),
    "exblas_b200_exsum_async": (C.c_int, [_h, _dp, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int]),
    "exblas_b200_exdot_async": (C.c_int, [_h, _dp, _i64, _i64, _dp, _i64, _i64, _i64, C.c_int, C.c_int, C.c_int]),
    "exblas_b200_fetch": (C.c_int, [_h, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(C.c_uint32)]),
    "exblas_b200_result_ptr": (C.c_int, [_h, C.POINTER(C.c_void_p)]),
    "exblas_b200_exgemv": (C.c_int, [_h, C.c_char, _i64, _i64, C.c_double, _dp, _i64, _i64, _dp, _i64, _i64, C.c_double,
                                     _dp, _i64, _i64, C.c_int, C.c_int, C.c_int]),
    "exblas_b200_exsum_segments": (C.c_int, [_h, _dp, C.c_void_p, _i64, C.c_int, C.c_int, C.c_int, _dp, C.c_void_p]),
    "exblas_b200_exdot_segments": (C.c_int, [_h, _dp, _dp, C.c_void_p, _i64, C.c_void_p, _i64, C.c_int, C.c_int, C.c_int, _dp,
                                             C.c_void_p]),
    "exblas_b200_sync": (C.c_int, [_h]),
    "exblas_b200_round": (C.c_int, [C.POINTER(_i64), C.c_int, C.POINTER(C.c_double)]),
    "exblas_b200_merge_limbs": (C.c_int, [C.POINTER(_i64), C.POINTER(_i64)]),
    "exblas_b200_normalize": (C.c_int, [C.POINTER(_i64), C.POINTER(C.c_int)]),
    "exblas_b200_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "exblas_b200_comm_init": (C.c_int, [_h, C.c_int, C.c_int, C.c_void_p]),
    "exblas_b200_allreduce_async": (C.c_int, [_h, C.c_int]),
    "exblas_b200_peer_export": (C.c_int, [_h, C.c_void_p]),
    "exblas_b200_peer_attach": (C.c_int, [_h, C.c_int, C.c_int, C.c_void_p]),
    "exblas_b200_last_status": (C.c_int, [_h, C.POINTER(C.c_uint32)]),
    "exblas_b200_last_error": (C.c_char_p, [_h]),
    "exblas_b200_launch_count": (_i64, [_h]),
    "exblas_b200_phase_times": (_i64, [_h, C.POINTER(C.c_uint64), _i64]),
    "exblas_b200_microbench": (C.c_int, [_h, C.c_int, _dp, _i64, C.POINTER(C.c_double)]),
    "exblas_b200_last_kernel": (C.c_char_p, [_h]),
}

_lib = None


class ExblasB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libexblas_b200.so (built by `python -m exblas_b200.build`).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ExblasB200Error(
            f"{LIB_PATH} is missing: build it with `python -m exblas_b200.build` "
            "(nvcc, sm_100a).  exblas_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)          # RTLD_LOCAL: exsum/exdot must not interpose other libraries
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI / header mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, handle=None) -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.exblas_b200_strerror(rc).decode()
    if handle:
        detail = lib.exblas_b200_last_error(handle).decode()
        if detail:
            msg += f": {detail}"
    raise ExblasB200Error(f"exblas_b200 error {rc}: {msg}")
