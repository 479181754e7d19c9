"""Host-side mirror of the reference's BLAS-2 interface, ExGEMV only (reference include/blas2.hpp:95):

    exgemv(transa, m, n, alpha, a, lda, offseta, x, incx, offsetx, beta, y, incy, offsety, fpe,
           early_exit=False) -> 0,  y overwritten with alpha*A*x + beta*y, every element the rounded
                                    exact value

transa == 'N' (BASELINE config 5: thread per row) and 'T' (warp per output) both run register-window kernels for
alpha == 1 and the general expansion kernel otherwise.  a, x, y: numpy float64 arrays (host, y is
updated in place) or torch CUDA float64 tensors (device pointers, asynchronous on the handle's
stream; the wrapper synchronises before returning, like the reference).
"""
from __future__ import annotations

from typing import Optional

from . import blas1
from ._lib import ROUND_REFERENCE, check


def exgemv(transa, m, n, alpha, a, lda, offseta, x, incx, offsetx, beta, y, incy, offsety, fpe, early_exit=False, *,
           round_mode=ROUND_REFERENCE, handle: Optional[blas1.Handle] = None, sync: bool = True) -> int:
    h = handle or blas1.default_handle()
    if isinstance(transa, str):
        transa = transa.encode()
    aa, sa, ka = blas1._address(a)
    ax, sx, kx = blas1._address(x)
    ay, sy, ky = blas1._address(y)
    if m < 0 or n < 0 or lda < max(1, m) or incx < 1 or incy < 1 or min(offseta, offsetx, offsety) < 0:
        raise ValueError("invalid m / n / lda / inc / offset")
    if m and n and offseta + lda * (n - 1) + m > sa:
        raise ValueError("a is too small for m, n, lda, offseta")
    tr = transa in (b"T", b"t")
    nin, nout = (m, n) if tr else (n, m)
    if nin and offsetx + (nin - 1) * incx >= sx:
        raise ValueError("x is too small")
    if nout and offsety + (nout - 1) * incy >= sy:
        raise ValueError("y is too small")
    check(h.lib.exblas_b200_exgemv(h._h, transa, m, n, float(alpha), aa, lda, offseta, ax, incx, offsetx, float(beta),
                                   ay, incy, offsety, fpe, int(early_exit), round_mode), h._h)
    if sync:
        check(h.lib.exblas_b200_sync(h._h), h._h)
    del ka, kx, ky
    return 0
