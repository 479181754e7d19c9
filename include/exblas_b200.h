/* exblas_b200.h -- C ABI of the B200-native ExSUM / ExDOT path.
 *
 * This is the drop-in boundary for the reference's BLAS-1 entry points
 *     double exsum(int Ng, double* ag, int inca, int offset, int fpe, bool early_exit, bool parallel)
 *                                                      (reference include/blas1.hpp:48)
 *     double exdot(int Ng, double* ag, int inca, int offseta, double* bg, int incb, int offsetb,
 *                  int fpe, bool early_exit)           (reference include/blas1.hpp:74)
 * implemented for the GPU in src/gpu/blas/blas1/ExSUM.cpp:64-209 and ExDOT.cpp:69-223.
 * include/blas1.hpp in this repository keeps those two C++ signatures on top of this ABI.
 *
 * Plain pointers and sizes only.  Every data pointer may be a host pointer (pageable or pinned)
 * or a device / managed pointer on the handle's GPU; the library finds out with
 * cudaPointerGetAttributes.  Host inputs are streamed through device staging buffers in chunks,
 * copy overlapped with the reduction (the reference copies the whole vector on every call,
 * ExSUM.cpp:126); pageable memory -- what new[] / malloc / _mm_malloc return -- is first gathered into
 * pinned bounce buffers by a few library threads, chunk k + 1 while chunk k is on the bus.
 *
 * Argument meaning follows the reference's GPU kernels (ExSUM.FPE.cl:252,298-299): `n` elements
 * a[offset + i*inca], i = 0..n-1.  fpe < 2 (exsum) / fpe < 3 (exdot) selects the
 * superaccumulator-only kernel; with early_exit the expansion size is bucketed to 4 / 6 / 8 exactly
 * as ExSUM.cpp:73-79 does; fpe > 8 is treated as 8 (the reference's CPU path returns 0.0 there).
 * The result does not depend on fpe / early_exit (they are performance knobs).
 *
 * All functions return 0 on success or a negative EXBLAS_B200_E* code; they never call exit().
 * Status flags about the DATA (NaN, Inf, out-of-range) are reported separately, see below.
 */
#ifndef EXBLAS_B200_H_
#define EXBLAS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EXBLAS_B200_LIMBS 39          /* reference bin_count, include/common.hpp:43 */

/* round_mode */
#define EXBLAS_B200_ROUND_REFERENCE 0 /* bit-for-bit the reference's Superaccumulator::Round()
                                         (superaccumulator.cpp:80-134; not always correctly rounded) */
#define EXBLAS_B200_ROUND_EXACT 1     /* correctly rounded, nearest-even */

/* error codes */
#define EXBLAS_B200_OK 0
#define EXBLAS_B200_EINVAL (-1)       /* bad argument (negative n, fpe < 0, null pointer, inc == 0 ...) */
#define EXBLAS_B200_ECUDA (-2)        /* a CUDA runtime call failed; see exblas_b200_last_error() */
#define EXBLAS_B200_ENOGPU (-3)       /* no usable sm_100 device */
#define EXBLAS_B200_ENCCL (-4)        /* NCCL missing or a NCCL call failed */
#define EXBLAS_B200_ENOMEM (-5)
#define EXBLAS_B200_EPEER (-6)        /* fused multi-GPU exchange: a peer did not arrive in time (see peer_attach) */

/* data status flags (bitwise OR), returned by exblas_b200_last_status() */
#define EXBLAS_B200_ST_NAN 1u         /* NaN met: value is NaN */
#define EXBLAS_B200_ST_POSINF 2u      /* +Inf met */
#define EXBLAS_B200_ST_NEGINF 4u      /* -Inf met */
#define EXBLAS_B200_ST_TOOLARGE 8u    /* finite |x| (or product) >= 2^988: outside the 39-limb layout, dropped */
#define EXBLAS_B200_ST_TOOSMALL 16u   /* bits below 2^-1040 truncated: result not exact */
#define EXBLAS_B200_ST_PEERTIMEOUT 32u /* fused multi-GPU exchange: a peer never arrived; the value is NaN, the limbs are
                                          partial and exblas_b200_fetch / the synchronous calls return EXBLAS_B200_EPEER */

typedef struct exblas_b200_handle_s* exblas_b200_handle_t;

/* Handle: owns the device workspace (global accumulator, result slot, staging buffers), a stream
 * and, optionally, a NCCL communicator.  Replaces the per-call OpenCL context / queue / JIT of
 * ExSUM.cpp:86-209 and ExSUM.Launcher.cpp:47-133.  device < 0 means the current device. */
int exblas_b200_create(exblas_b200_handle_t* handle, int device);
int exblas_b200_destroy(exblas_b200_handle_t handle);

/* Use `stream` (a cudaStream_t) for all subsequent work of this handle.  A new handle uses the
 * legacy default stream (NULL), like cuBLAS: ordered after prior default-stream work of the process. */
int exblas_b200_set_stream(exblas_b200_handle_t handle, void* stream);

/* Tuning knobs (performance only, never the result): "block_threads" (multiple of 32, <= 512) and
 * "blocks" (0 = automatic) fix the launch shape by hand ("auto_shape" = 1 returns to the size-dependent
 * one: a single CTA up to "solo_max_elems" elements, 256-thread CTAs up to "small_max_elems", else one
 * 512-thread CTA per SM), "track_max_elems" (ExSUM vectors up to this length, default 2^20, take the superaccumulator-only
 * kernel without its unrolled body whatever fpe says: it merges only the limb rows it touched, the cheapest launch at
 * those sizes; 0 = the kernel fpe selects), "host_chunk_elems" (elements per H2D chunk of pinned host input), "host_threads" /
 * "pageable_chunk_elems" (pageable host input goes through pinned bounce buffers filled by that many copy
 * threads, 0 = automatic, 1 = none: the driver's own pageable staging), "gemv_parts" (0 = automatic column split, <= 1024),
 * "adaptive" (1 = a warp bypasses the expansion and deposits straight into its superaccumulators while
 * the expansion overflows on most elements; 0 = always walk all fpe levels, as the reference kernels do),
 * "dot_handoff_tiles" (ExDOT with fpe >= 3: a warp whose first tile thrashes the expansion hands the rest of its rows to
 * the 5-digit register window when the vector gives every CTA at least this many tiles, default 16, 0 = never),
 * "window" (register window of the superaccumulator-only kernels: 0 off, 1 narrow windows, 2 narrow then
 * wide windows (default), 3 wide only), "gemv_n_shape" (0..2) / "gemv_t_shape" (0..3): launch shapes of
 * the ExGEMV window kernels, "gemv_prefetch" (their TMA-engine L2 prefetch distance in rounds, 0 = off, default 2),
 * "gemv_tma" (ExGEMV 'T': 1 (default) = x is staged by cp.async.bulk whenever it is contiguous and 16-byte aligned,
 * 0 = always by the copying warp), "reduce_prefetch" (L2 prefetch distance of the expansion kernel in tiles, default 0),
 * "fused_allreduce" (1 = exchange limbs inside the kernel once peers are attached),
 * "world_size" (declares a multi-rank job: exblas_b200_allreduce_async then fails with EXBLAS_B200_ENCCL
 * instead of finishing locally when no transport covers that many ranks),
 * "peer_timeout_ms" (fused exchange: how long a rank waits for its peers, default 600000, 0 = for ever),
 * "phase_timing" (diagnostics: per-CTA globaltimer stamps, read with exblas_b200_phase_times). */
int exblas_b200_set_option(exblas_b200_handle_t handle, const char* name, int64_t value);

/* ---- synchronous entry points: the ones a reference binding calls --------------------------- */

/* replaces exsum()  (reference include/blas1.hpp:48; src/gpu/blas/blas1/ExSUM.cpp:64-84) */
int exblas_b200_exsum(exblas_b200_handle_t handle, const double* a, int64_t n, int64_t inca, int64_t offset,
                      int fpe, int early_exit, int round_mode, double* result);

/* replaces exdot()  (reference include/blas1.hpp:74; src/gpu/blas/blas1/ExDOT.cpp:69-92) */
int exblas_b200_exdot(exblas_b200_handle_t handle, const double* a, int64_t inca, int64_t offseta,
                      const double* b, int64_t incb, int64_t offsetb, int64_t n,
                      int fpe, int early_exit, int round_mode, double* result);

/* Same reductions, but return the exact sum as normalised limbs (limb j weighs 2^(52*(j-20)); limbs
 * 0..37 in [0,2^52), limb 38 signed) instead of / in addition to the rounded value.  This is what the
 * reference's MPI path exchanges (cpu ExSUM.cpp:266-273).  result may be NULL. */
int exblas_b200_exsum_limbs(exblas_b200_handle_t handle, const double* a, int64_t n, int64_t inca, int64_t offset,
                            int fpe, int early_exit, int round_mode, int64_t limbs[EXBLAS_B200_LIMBS], double* result);
int exblas_b200_exdot_limbs(exblas_b200_handle_t handle, const double* a, int64_t inca, int64_t offseta,
                            const double* b, int64_t incb, int64_t offsetb, int64_t n,
                            int fpe, int early_exit, int round_mode, int64_t limbs[EXBLAS_B200_LIMBS], double* result);

/* ---- asynchronous, device-resident entry points (what bench.py times as `value`) ------------- */

/* Enqueue the reduction of DEVICE data on the handle's stream and return at once.  The result
 * (double value, uint32 status, int64 limbs[39]) is left in the handle's device result slot;
 * fetch it with exblas_b200_fetch() or read it on the device through exblas_b200_result_ptr(). */
int exblas_b200_exsum_async(exblas_b200_handle_t handle, const double* d_a, int64_t n, int64_t inca, int64_t offset,
                            int fpe, int early_exit, int round_mode);
int exblas_b200_exdot_async(exblas_b200_handle_t handle, const double* d_a, int64_t inca, int64_t offseta,
                            const double* d_b, int64_t incb, int64_t offsetb, int64_t n,
                            int fpe, int early_exit, int round_mode);
/* Wait for the stream and copy the result slot to the host.  Any of the outputs may be NULL. */
int exblas_b200_fetch(exblas_b200_handle_t handle, double* result, int64_t limbs[EXBLAS_B200_LIMBS], uint32_t* status);
/* Device address of the result slot: { double value; uint32 status; uint32 pad; int64 limbs[39]; } */
int exblas_b200_result_ptr(exblas_b200_handle_t handle, void** d_result);

/* ---- ExGEMV 'N' (BASELINE config 5) --------------------------------- */

/* replaces exgemv()  (reference include/blas2.hpp:95; src/gpu/blas/blas2/ExGEMV.cpp:81-234) for
 * transa == 'N': y := alpha*A*x + beta*y, A column-major m x n with leading dimension lda, every
 * y[i] the rounded EXACT value (per-row ExDOT).  Operands may all be host or all be device
 * pointers; with device pointers the call is asynchronous on the handle's stream (follow with
 * exblas_b200_sync to read the status flags).  fpe: 0 / 1 superaccumulators only (the reference's
 * fpe == 1 is a plain DGEMV comparator), early_exit buckets 4 / 6 / 8, else the expansion size.
 * Deviation from the reference kernels: alpha is applied (exactly); ExGEMV.FPE.cl:246 ignores it.
 * Domain of the exact scaling: alpha * a[i,j] and beta * y[i] are split into two doubles (TwoProductFMA), which is
 * exact for 2^-959 <= |product| < 2^1024.  A non-zero product outside that range is dropped and reported
 * (EXBLAS_B200_ST_TOOSMALL: the affected y[i] is not exact; EXBLAS_B200_ST_TOOLARGE), never summed inexactly. */
int exblas_b200_exgemv(exblas_b200_handle_t handle, char trans, int64_t m, int64_t n, double alpha,
                       const double* a, int64_t lda, int64_t offseta, const double* x, int64_t incx,
                       int64_t offsetx, double beta, double* y, int64_t incy, int64_t offsety,
                       int fpe, int early_exit, int round_mode);
/* ---- batched ("segmented") ExSUM / ExDOT ---------------------------- */

/* nseg independent exact reductions in ONE launch, one warp per segment:
 *     results[s] = round( sum_{i = seg[s]}^{seg[s+1]-1} a[i] )                    (exsum_segments)
 *     results[s] = round( sum_{i = seg[s]}^{seg[s+1]-1} a[i] * b[ gather ? gather[i] : i ] )   (exdot_segments)
 * seg holds nseg + 1 non-decreasing offsets.  This is what the reference's application callers do with
 * one exsum() per short host vector (src/cpu/examples/kmeans/kmeans_clustering.cpp:213,
 * spmv/main.cpp:85, mri-gridding/CPU_kernels.cpp:293); with a gather index it is a CSR sparse
 * matrix-vector product (a = values, gather = column indices, seg = row pointers, b = x).
 * statuses (optional) receives each segment's EXBLAS_B200_ST_* flags.  All pointers host, or all
 * device (then asynchronous on the handle's stream; nb is only needed for host operands with gather:
 * the length of b).  fpe / early_exit are accepted for symmetry and never change a result. */
int exblas_b200_exsum_segments(exblas_b200_handle_t handle, const double* a, const int64_t* seg, int64_t nseg,
                               int fpe, int early_exit, int round_mode, double* results, uint32_t* statuses);
int exblas_b200_exdot_segments(exblas_b200_handle_t handle, const double* a, const double* b, const int32_t* gather,
                               int64_t nb, const int64_t* seg, int64_t nseg, int fpe, int early_exit, int round_mode,
                               double* results, uint32_t* statuses);

/* Wait for the handle's stream and latch the status flags of the last asynchronous call. */
int exblas_b200_sync(exblas_b200_handle_t handle);

/* ---- limb arithmetic on the host (no GPU needed) ------------------------------------------- */

/* Round normalised-or-not limbs to a double (restates Superaccumulator::Round / exact RN-even). */
int exblas_b200_round(const int64_t limbs[EXBLAS_B200_LIMBS], int round_mode, double* result);
/* dst += src limb-wise, then normalise dst (Superaccumulator::Accumulate(Superaccumulator&),
 * superaccumulator.cpp:68-78). */
int exblas_b200_merge_limbs(int64_t dst[EXBLAS_B200_LIMBS], const int64_t src[EXBLAS_B200_LIMBS]);
/* Carry-normalise in place; *negative receives the sign (Superaccumulator::Normalize). */
int exblas_b200_normalize(int64_t limbs[EXBLAS_B200_LIMBS], int* negative);

/* ---- multi-GPU: exact combination of per-rank limbs over NCCL ------------------------------- */

/* Size of a NCCL unique id in bytes and creation of one (rank 0 calls this and broadcasts it). */
int exblas_b200_nccl_unique_id(void* id128);
/* Join a communicator of `nranks` ranks (one process per GPU). */
int exblas_b200_comm_init(exblas_b200_handle_t handle, int nranks, int rank, const void* id128);
/* All-reduce (integer sum, exact) of the limbs + status left in the result slot by the last
 * *_async call, then normalise and round on every rank: every rank ends with identical bits.
 * Replaces MPI_Reduce(MPI_LONG, MPI_SUM) + Round() of cpu ExSUM.cpp:266-273. */
int exblas_b200_allreduce_async(exblas_b200_handle_t handle, int round_mode);

/* Fused alternative to the NCCL path: the limb exchange happens INSIDE the closing reduction kernel
 * over peer-mapped memory (NVLink / NVSwitch), no NCCL call and no extra launch.  Every rank calls
 * peer_export (a 64-byte CUDA IPC handle of its mailbox), the 64-byte handles of all ranks are
 * gathered in rank order by any means, and every rank calls peer_attach.  From then on every
 * *_async reduction on this handle is a COLLECTIVE: all ranks must issue the same sequence of
 * reductions; each ends with identical value / limbs / status on every rank, and
 * exblas_b200_allreduce_async becomes a no-op.  Option "fused_allreduce" = 0 switches back to NCCL.
 * One process per GPU, at most 8 ranks on one NVLink domain.
 * A rank waits up to "peer_timeout_ms" (default 10 minutes; 0 = for ever, like a NCCL collective) for its
 * peers' limbs.  If they do not arrive, that rank's result carries EXBLAS_B200_ST_PEERTIMEOUT, its value is
 * NaN and exblas_b200_fetch / the synchronous entry points return EXBLAS_B200_EPEER.  The ranks then no longer
 * agree on which reductions have completed: results are undefined until every rank has called peer_export
 * and peer_attach again. */
int exblas_b200_peer_export(exblas_b200_handle_t handle, void* handle64);
int exblas_b200_peer_attach(exblas_b200_handle_t handle, int nranks, int rank, const void* handles);

/* ---- diagnostics ----------------------------------------------------------------------------- */
int exblas_b200_last_status(exblas_b200_handle_t handle, uint32_t* status_flags);
const char* exblas_b200_last_error(exblas_b200_handle_t handle);
const char* exblas_b200_strerror(int code);
/* With option "phase_timing" = 1: copies the globaltimer stamps (ns) of the last reduction kernel, 16 per CTA
 * (0 kernel start, 1 unrolled body done, 2..4 remainder / scalar part / flush, 5 block merge, 6 global merge,
 * 7 normalised, 8 peers merged, 9 published), into out[capacity]; returns the number of CTAs written. */
int64_t exblas_b200_phase_times(exblas_b200_handle_t handle, uint64_t* out, int64_t capacity);
/* Measured ceilings for the roofline, on this GPU at this moment (diagnostics; what bench.py divides by):
 * what = 0: FP64 pipe, DADD lane-instructions per second (d_buf / n unused);
 * what = 1: HBM read-only stream in GB/s over the DEVICE buffer d_buf[0, n) (32-byte aligned), read with the
 *           reduction kernels' own 256-bit L1-bypassing loads. */
int exblas_b200_microbench(exblas_b200_handle_t handle, int what, const double* d_buf, int64_t n, double* result);
/* Template instance and launch shape of the reduction kernel this handle launched last. */
const char* exblas_b200_last_kernel(exblas_b200_handle_t handle);
/* Kernels launched by this handle since creation (bench.py reports it as gpu_launches). */
int64_t exblas_b200_launch_count(exblas_b200_handle_t handle);
int exblas_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EXBLAS_B200_H_ */
