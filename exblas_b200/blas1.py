"""Host-side mirror of the reference's BLAS-1 interface for the ExSUM / ExDOT path.

Same names, argument order and meaning as the reference's include/blas1.hpp:48,74
(`exsum(Ng, ag, inca, offset, fpe, early_exit=False, parallel=True)`,
 `exdot(Ng, ag, inca, offseta, bg, incb, offsetb, fpe, early_exit=False)`), calling the
C ABI (include/exblas_b200.h) through ctypes.  `ag` / `bg` may be

* a numpy float64 array (host memory; streamed to the GPU in chunks by the library), or
* a torch float64 tensor on the GPU (device pointer; no copy), or on the CPU (host pointer).

PyTorch is only used to obtain device pointers; no torch op is on the compute path, and there is
no CPU fallback: without the CUDA library / a B200 the calls raise.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import LIMBS, ROUND_EXACT, ROUND_REFERENCE, ExblasB200Error, check


def _address(x) -> Tuple[int, int, object]:
    """-> (address, number of float64 elements, keep-alive object)."""
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64:
            raise TypeError("exblas_b200 works on float64 data")
        if not x.flags.c_contiguous:
            raise ValueError("pass a contiguous array and express strides with inca / incb")
        return x.ctypes.data, x.size, x
    if hasattr(x, "data_ptr") and hasattr(x, "dtype"):      # torch tensor, without importing torch here
        import torch
        if x.dtype != torch.float64:
            raise TypeError("exblas_b200 works on float64 data")
        if not x.is_contiguous():
            raise ValueError("pass a contiguous tensor and express strides with inca / incb")
        return x.data_ptr(), x.numel(), x
    arr = np.ascontiguousarray(x, dtype=np.float64)
    return arr.ctypes.data, arr.size, arr


class Handle:
    """Owns the device workspace, stream and (optionally) the NCCL communicator of one GPU.
    Replaces the per-call OpenCL set-up of the reference (src/gpu/blas/blas1/ExSUM.cpp:86-209)."""

    def __init__(self, device: int = -1):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.exblas_b200_create(C.byref(h), device))
        self._h = h
        self.nranks = 1

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.exblas_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- options -----------------------------------------------------------------------------
    def set_option(self, name: str, value: int) -> None:
        check(self.lib.exblas_b200_set_option(self._h, name.encode(), int(value)), self._h)

    def set_stream(self, cuda_stream_ptr: int) -> None:
        check(self.lib.exblas_b200_set_stream(self._h, C.c_void_p(cuda_stream_ptr)), self._h)

    # -- synchronous -------------------------------------------------------------------------
    def exsum(self, Ng, ag, inca=1, offset=0, fpe=0, early_exit=False, round_mode=ROUND_REFERENCE,
              want_limbs=False):
        addr, size, keep = _address(ag)
        _check_extent(Ng, inca, offset, size, "ag")
        res = C.c_double()
        limbs = (C.c_int64 * LIMBS)()
        check(self.lib.exblas_b200_exsum_limbs(self._h, addr, Ng, inca, offset, fpe, int(early_exit), round_mode,
                                               limbs, C.byref(res)), self._h)
        del keep
        return (res.value, np.array(limbs, dtype=np.int64)) if want_limbs else res.value

    def exdot(self, Ng, ag, inca, offseta, bg, incb, offsetb, fpe=0, early_exit=False,
              round_mode=ROUND_REFERENCE, want_limbs=False):
        aa, sa, ka = _address(ag)
        ab, sb, kb = _address(bg)
        _check_extent(Ng, inca, offseta, sa, "ag")
        _check_extent(Ng, incb, offsetb, sb, "bg")
        res = C.c_double()
        limbs = (C.c_int64 * LIMBS)()
        check(self.lib.exblas_b200_exdot_limbs(self._h, aa, inca, offseta, ab, incb, offsetb, Ng, fpe,
                                               int(early_exit), round_mode, limbs, C.byref(res)), self._h)
        del ka, kb
        return (res.value, np.array(limbs, dtype=np.int64)) if want_limbs else res.value

    # -- batched ("segmented") reductions: one launch, one warp per segment -------------------
    def exsum_segments(self, a, seg, fpe=0, early_exit=False, round_mode=ROUND_REFERENCE, out=None, want_status=False):
        """results[s] = exact sum of a[seg[s]:seg[s+1]], rounded.  a / seg / out: all numpy (host) or all
        torch CUDA tensors (seg int64).  Returns out (allocated when None) [, per-segment status flags]."""
        return self._segments(a, None, None, seg, fpe, early_exit, round_mode, out, want_status)

    def exdot_segments(self, a, b, seg, gather=None, fpe=0, early_exit=False, round_mode=ROUND_REFERENCE, out=None,
                       want_status=False):
        """results[s] = exact sum of a[i] * b[gather[i] if gather is not None else i] over seg[s] <= i < seg[s+1].
        With gather (int32 column indices) this is a CSR sparse matrix-vector product."""
        return self._segments(a, b, gather, seg, fpe, early_exit, round_mode, out, want_status)

    def _segments(self, a, b, gather, seg, fpe, early_exit, round_mode, out, want_status):
        on_gpu = hasattr(a, "data_ptr")
        if on_gpu:
            import torch
            if seg.dtype != torch.int64 or (gather is not None and gather.dtype != torch.int32):
                raise TypeError("seg must be int64, gather int32")
            nseg = seg.numel() - 1
            if out is None:
                out = torch.empty(max(nseg, 0), dtype=torch.float64, device=a.device)
            st = torch.zeros(max(nseg, 0), dtype=torch.int32, device=a.device) if want_status else None
            ptr = lambda t: t.data_ptr() if t is not None else None
            nb = b.numel() if b is not None else 0
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
            b = np.ascontiguousarray(b, dtype=np.float64) if b is not None else None
            seg = np.ascontiguousarray(seg, dtype=np.int64)
            gather = np.ascontiguousarray(gather, dtype=np.int32) if gather is not None else None
            nseg = seg.size - 1
            if nseg >= 0 and seg.size and int(seg[-1]) > a.size:
                raise ValueError("seg reads past the end of a")
            if b is not None and gather is None and nseg >= 0 and seg.size and int(seg[-1]) > b.size:
                raise ValueError("seg reads past the end of b")
            if out is None:
                out = np.empty(max(nseg, 0), dtype=np.float64)
            st = np.zeros(max(nseg, 0), dtype=np.uint32) if want_status else None
            ptr = lambda t: t.ctypes.data if t is not None else None
            nb = b.size if b is not None else 0
        if nseg < 0:
            raise ValueError("seg needs at least one offset")
        if b is None:
            check(self.lib.exblas_b200_exsum_segments(self._h, ptr(a), ptr(seg), nseg, fpe, int(early_exit), round_mode,
                                                      ptr(out), ptr(st)), self._h)
        else:
            check(self.lib.exblas_b200_exdot_segments(self._h, ptr(a), ptr(b), ptr(gather), nb, ptr(seg), nseg, fpe,
                                                      int(early_exit), round_mode, ptr(out), ptr(st)), self._h)
        if on_gpu:
            check(self.lib.exblas_b200_sync(self._h), self._h)
        return (out, st) if want_status else out

    # -- asynchronous, device-resident ---------------------------------------------------------
    def exsum_async(self, Ng, d_ag, inca=1, offset=0, fpe=0, early_exit=False, round_mode=ROUND_REFERENCE):
        addr, size, _ = _address(d_ag)
        _check_extent(Ng, inca, offset, size, "ag")
        check(self.lib.exblas_b200_exsum_async(self._h, addr, Ng, inca, offset, fpe, int(early_exit), round_mode),
              self._h)

    def exdot_async(self, Ng, d_ag, inca, offseta, d_bg, incb, offsetb, fpe=0, early_exit=False,
                    round_mode=ROUND_REFERENCE):
        aa, sa, _ = _address(d_ag)
        ab, sb, _ = _address(d_bg)
        _check_extent(Ng, inca, offseta, sa, "ag")
        _check_extent(Ng, incb, offsetb, sb, "bg")
        check(self.lib.exblas_b200_exdot_async(self._h, aa, inca, offseta, ab, incb, offsetb, Ng, fpe,
                                               int(early_exit), round_mode), self._h)

    def fetch(self):
        """-> (value, limbs[39] int64, status flags) of the last asynchronous call."""
        res = C.c_double()
        limbs = (C.c_int64 * LIMBS)()
        st = C.c_uint32()
        check(self.lib.exblas_b200_fetch(self._h, C.byref(res), limbs, C.byref(st)), self._h)
        return res.value, np.array(limbs, dtype=np.int64), int(st.value)

    def result_ptr(self) -> int:
        p = C.c_void_p()
        check(self.lib.exblas_b200_result_ptr(self._h, C.byref(p)), self._h)
        return int(p.value)

    # -- multi-GPU -----------------------------------------------------------------------------
    def comm_init(self, nranks: int, rank: int, unique_id: bytes) -> None:
        buf = C.create_string_buffer(unique_id, 128)
        check(self.lib.exblas_b200_comm_init(self._h, nranks, rank, buf), self._h)
        self.nranks = nranks

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(self.lib.exblas_b200_peer_export(self._h, buf), self._h)
        return buf.raw

    def peer_attach(self, nranks: int, rank: int, handles) -> None:
        blob = b"".join(handles)
        assert len(blob) == 64 * nranks
        buf = C.create_string_buffer(blob, len(blob))
        check(self.lib.exblas_b200_peer_attach(self._h, nranks, rank, buf), self._h)

    def allreduce_async(self, round_mode=ROUND_REFERENCE) -> None:
        check(self.lib.exblas_b200_allreduce_async(self._h, round_mode), self._h)

    # -- diagnostics ---------------------------------------------------------------------------
    def last_status(self) -> int:
        st = C.c_uint32()
        check(self.lib.exblas_b200_last_status(self._h, C.byref(st)), self._h)
        return int(st.value)

    def launch_count(self) -> int:
        return int(self.lib.exblas_b200_launch_count(self._h))

    def last_kernel(self) -> str:
        return self.lib.exblas_b200_last_kernel(self._h).decode()

    def microbench(self, what: int, d_buf=None) -> float:
        """0: FP64 DADD lane-instructions/s; 1: read-only stream GB/s over the device tensor d_buf."""
        res = C.c_double()
        addr, n = (0, 0)
        if d_buf is not None:
            addr, n, _ = _address(d_buf)
        check(self.lib.exblas_b200_microbench(self._h, what, addr, n, C.byref(res)), self._h)
        return res.value

    def phase_times(self) -> np.ndarray:
        """With option "phase_timing" = 1: [CTAs][16] globaltimer stamps (ns) of the last reduction kernel."""
        buf = (C.c_uint64 * (2048 * 16))()
        nb = int(self.lib.exblas_b200_phase_times(self._h, buf, 2048 * 16))
        return np.frombuffer(buf, dtype=np.uint64)[:nb * 16].reshape(nb, 16).copy()


def _check_extent(n, inc, off, size, name):
    if n < 0 or inc < 1 or off < 0:
        raise ValueError(f"invalid Ng / inc / offset for {name}")
    if n > 0 and off + (n - 1) * inc >= size:
        raise ValueError(f"{name} has {size} elements; Ng={n}, inc={inc}, offset={off} reads past its end")


def nccl_unique_id() -> bytes:
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    check(lib.exblas_b200_nccl_unique_id(buf))
    return buf.raw


# ---- limb arithmetic on the host (pure C, no GPU needed) ---------------------------------------
def round_limbs(limbs, round_mode=ROUND_REFERENCE) -> float:
    lib = _lib.load()
    arr = (C.c_int64 * LIMBS)(*[int(v) for v in limbs])
    res = C.c_double()
    check(lib.exblas_b200_round(arr, round_mode, C.byref(res)))
    return res.value


def merge_limbs(dst, src) -> np.ndarray:
    lib = _lib.load()
    a = (C.c_int64 * LIMBS)(*[int(v) for v in dst])
    b = (C.c_int64 * LIMBS)(*[int(v) for v in src])
    check(lib.exblas_b200_merge_limbs(a, b))
    return np.array(a, dtype=np.int64)


def normalize_limbs(limbs) -> Tuple[np.ndarray, bool]:
    lib = _lib.load()
    a = (C.c_int64 * LIMBS)(*[int(v) for v in limbs])
    neg = C.c_int()
    check(lib.exblas_b200_normalize(a, C.byref(neg)))
    return np.array(a, dtype=np.int64), bool(neg.value)


# ---- the reference's free functions --------------------------------------------------------------
_default: Optional[Handle] = None


def default_handle() -> Handle:
    global _default
    if _default is None:
        _default = Handle(-1)
    return _default


def exsum(Ng, ag, inca, offset, fpe, early_exit=False, parallel=True, *, round_mode=ROUND_REFERENCE,
          handle: Optional[Handle] = None) -> float:
    """reference include/blas1.hpp:48.  `parallel` is accepted and ignored, as in the reference's
    GPU implementation (src/gpu/blas/blas1/ExSUM.cpp:61)."""
    if fpe < 0:   # cpu ExSUM.cpp:25-28 prints this and exits
        sys.stderr.write("Size of floating-point expansion should be a positive number. "
                         "Preferably, it should be in the interval [2, 8]\n")
        raise SystemExit(1)
    return (handle or default_handle()).exsum(max(int(Ng), 0), ag, inca, offset, fpe, early_exit, round_mode)


def exdot(Ng, ag, inca, offseta, bg, incb, offsetb, fpe, early_exit=False, *, round_mode=ROUND_REFERENCE,
          handle: Optional[Handle] = None) -> float:
    """reference include/blas1.hpp:74; Ng <= 0 returns 0.0 (src/gpu/blas/blas1/ExDOT.cpp:70-71)."""
    if Ng <= 0:
        return 0.0
    if fpe < 0:
        sys.stderr.write("Size of floating-point expansion should be a positive number. "
                         "Preferably, it should be in the interval [3, 8]\n")
        raise SystemExit(1)
    return (handle or default_handle()).exdot(int(Ng), ag, inca, offseta, bg, incb, offsetb, fpe, early_exit,
                                              round_mode)
