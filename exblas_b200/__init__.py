"""exblas_b200 -- B200-native ExSUM / ExDOT (exact, reproducible sum and dot product).

Hand-written CUDA for sm_100a behind a C ABI (include/exblas_b200.h); this package is the thin
Python host layer that mirrors the reference's blas1.hpp interface.  No CPU fallback.
"""
from ._lib import (LIMBS, ROUND_EXACT, ROUND_REFERENCE, ST_NAN, ST_NEGINF, ST_POSINF, ST_TOOLARGE, ST_TOOSMALL,
                   ExblasB200Error)
from .blas2 import exgemv
from .blas1 import (Handle, default_handle, exdot, exsum, merge_limbs, nccl_unique_id, normalize_limbs,
                    round_limbs)

__all__ = [
    "LIMBS", "ROUND_EXACT", "ROUND_REFERENCE", "ST_NAN", "ST_NEGINF", "ST_POSINF", "ST_TOOLARGE", "ST_TOOSMALL",
    "ExblasB200Error", "Handle", "default_handle", "exdot", "exgemv", "exsum", "merge_limbs", "nccl_unique_id",
    "normalize_limbs", "round_limbs",
]
