// reduce_kernel.cuh -- the ExSUM / ExDOT streaming kernel for sm_100a.
//
// Replaces the reference's OpenCL kernels ExSUM / ExDOT (+ ExSUMComplete / ExDOTComplete):
//   src/gpu/blas/blas1/ExSUM.FPE.cl:230-453, ExSUM.FPE.EX.{4,6,8}.cl, ExSUM.Superacc.cl:212-294,
//   src/gpu/blas/blas1/ExDOT.FPE.cl:201-388, ExDOT.FPE.EX.{4,6,8}.cl, ExDOT.Superacc.cl:218-320
// and the CPU driver ExSUMFPE (src/cpu/blas/blas1/ExSUM.cpp:219-311).
//
// Design (B200-first, not a translation):
//   * one persistent CTA per SM; every thread streams 256-bit coalesced, L1-bypassing loads
//     (ld.global.nc.L1::no_allocate.v4.f64) with U vectors in flight, rolling;
//   * each thread keeps a floating-point expansion a[F] in registers (Knuth TwoSum, and
//     TwoProd via FMA for ExDOT); early exit is a warp-uniform vote on the integer bit
//     pattern of the residual, so it costs no FP64-pipe slot and no divergence;
//   * a non-zero residual is deposited into the thread's PRIVATE 39-limb superaccumulator
//     column in shared memory (plain LDS/STS, conflict-free, no atomics) -- see superacc.cuh;
//     F == 0 deposits every element directly (the reference's "superacc only" mode);
//   * columns are carry-normalised on a fixed schedule (every <= 2046 deposits) so no limb can
//     overflow at any N;
//   * block merge: column sums by warp shuffles -> normalise -> 39 native 64-bit REDs into one
//     global accumulator; the last CTA to arrive (ticket counter) normalises that accumulator,
//     optionally rounds, publishes, and resets the workspace.  Integer adds commute, so the
//     result is bit-identical for any grid, block size, FPE size or schedule.
#pragma once
#include <cuda_runtime.h>
#include "superacc.cuh"
#include "window.cuh"

namespace exb {

struct alignas(32) Vec4 { double x, y, z, w; };

struct Workspace {                       // lives in device memory, zero between calls
    unsigned long long gacc[kLimbs];     // global accumulator (always left normalised)
    unsigned counter;                    // CTA arrival ticket
    unsigned status;                     // OR of status flags
};

constexpr int kFlagSlots = 5;            // one counter per status flag, so that flags survive an integer all-reduce

struct Result {                          // published by the last CTA
    double value;
    unsigned status;
    unsigned pad;
    long long limbs[kLimbs];             // normalised
    long long flagcnt[kFlagSlots];       // flagcnt[k] = 1 when status bit k is set (contiguous with limbs)
};

// Peer-memory mailbox for the fused limb exchange over NVLink / NVSwitch.  Rank r's last CTA stores
// its 44 int64 (normalised limbs + status-flag counters) into slot [epoch & 1][r] of EVERY rank's
// mailbox with plain peer stores, then publishes the epoch with a release store; each rank then
// waits for all its slots of this epoch, sums them as integers, normalises and rounds -- inside the
// reduction kernel, with no NCCL call and no extra launch.  Two slot sets suffice: a rank cannot
// start epoch e+2 before every peer has finished reading epoch e (it needs their epoch e+1 data).
constexpr int kMaxPeers = 8;
constexpr int kMsgWords = kLimbs + kFlagSlots;   // 44
struct MailSlot {
    unsigned long long data[kMsgWords];
    unsigned long long seq;
    unsigned long long pad[3];                   // 48 words = 384 B per slot
};
struct Mailbox {
    MailSlot slot[2][kMaxPeers];
};

struct ReduceParams {
    const double* a;                     // already offset by `offset`
    const double* b;                     // ExDOT only
    long long n;                         // number of elements
    long long inca, incb;                // element strides
    long long head;                      // scalar elements before the vector body
    long long ntiles;                    // full tiles in the vector body (0 => all scalar)
    Workspace* ws;
    Result* out;
    int finalize;                        // 1: publish value/limbs/status and reset workspace
    int round_mode;                      // 0 reference Round(), 1 exact RN-even
    int keep;                            // 1: do not reset the accumulator after publishing
    int adaptive;                        // 1: bypass the expansion while it thrashes (performance only)
    int fresh;                           // 1: the workspace accumulator is known to be zero (no pending chunks)
    int window;                          // 1: F == 0 kernels keep a register window of the superaccumulator (performance only)
    // fused multi-GPU exchange (0 ranks = off): every rank's mailbox, mapped into this process
    Mailbox* peers[kMaxPeers];
    int nranks, rank;
    unsigned long long epoch;            // identifies this collective reduction (same on every rank, >= 1)
};

EXB_D Vec4 ldg256(const double* p) {
    Vec4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
}

EXB_D double ldg64(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

EXB_D void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
EXB_D void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
EXB_D unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
EXB_D unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

EXB_D bool nonzero_bits(double x) {
    return (((unsigned)__double2hiint(x) & 0x7fffffffu) | (unsigned)__double2loint(x)) != 0u;
}

// ---- per-thread state --------------------------------------------------------------------------
// (col, stride) of the private accumulator column, status flags, and M independent expansions
// a[level][m] in registers.  Two expansions per thread (F <= 4) give the FP64 pipe independent
// TwoSum chains to overlap (elements 0,2 of a vector feed expansion 0, elements 1,3 expansion 1);
// the exact sum does not care how summands are distributed over expansions.
__host__ __device__ constexpr int expansions(int f) { return f <= 4 ? 2 : 1; }   // large F: registers go to the load window instead

// Knuth TwoSum, un-contracted (ExSUM.FPE.cl:27-32): a + x = r + s exactly; a <- r, x <- s.
EXB_D void two_sum(double& a, double& x) {
    const double r = __dadd_rn(a, x);
    const double z = __dsub_rn(r, a);
    const double s = __dadd_rn(__dsub_rn(a, __dsub_rn(r, z)), __dsub_rn(x, z));
    a = r;
    x = s;
}

// Four summands through the expansion levels [first, F), level by level so that four chains are in
// flight.  With EE the walk stops at the first level after which no lane of the warp (UNIFORM) /
// this thread (!UNIFORM) holds a non-zero residual: one vote per level per four elements, taken
// on the integer bit patterns (no FP64-pipe compare, no divergence).  Residuals come back in x[].
template <int F, bool EE, bool UNIFORM>
EXB_D void fpe_push4(double (&a)[F > 0 ? F : 1][expansions(F)], double (&x)[4], int first) {
    constexpr int M1 = expansions(F) - 1;
#pragma unroll
    for (int i = 0; i < F; ++i) {
        if (i < first) continue;
        two_sum(a[i][0], x[0]);
        two_sum(a[i][M1], x[1]);
        two_sum(a[i][0], x[2]);
        two_sum(a[i][M1], x[3]);
        if (EE && i + 1 < F) {
            const unsigned any = ((unsigned)__double2hiint(x[0]) | (unsigned)__double2hiint(x[1]) |
                                  (unsigned)__double2hiint(x[2]) | (unsigned)__double2hiint(x[3])) << 1 |
                                 ((unsigned)__double2loint(x[0]) | (unsigned)__double2loint(x[1]) |
                                  (unsigned)__double2loint(x[2]) | (unsigned)__double2loint(x[3]));
            if (UNIFORM) {
                if (!__any_sync(0xffffffffu, any != 0u)) break;
            } else {
                if (any == 0u) break;
            }
        }
    }
}

// One summand through the levels [first, F) of expansion 0 (alignment heads, tails and strided
// vectors; lanes are not converged there, so the early exit is per thread).
template <int F, bool EE>
EXB_D double fpe_push1(double (&a)[F > 0 ? F : 1][expansions(F)], double x, int first) {
#pragma unroll
    for (int i = 0; i < F; ++i) {
        if (i < first) continue;
        two_sum(a[i][0], x);
        if (EE && i + 1 < F && !nonzero_bits(x)) break;
    }
    return x;
}

// Residuals that fell off the last level go to the superaccumulator column.  Returns how many.
EXB_D int deposit_residuals(unsigned col, unsigned stride, const double (&x)[4], unsigned& status) {
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (nonzero_bits(x[k])) {
            deposit(col, stride, x[k], status);
            ++cnt;
        }
    }
    return cnt;
}

// Four ordinary inputs (any doubles) through the expansion.  No lane leaves early (the votes in
// fpe_push4 must be reached by all lanes): Inf / NaN / |x| >= 2^988 are diverted and replaced by 0.
template <int F, bool EE, bool UNIFORM>
EXB_D int add4(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1][expansions(F)], unsigned& status, double (&x)[4]) {
    const unsigned h0 = (unsigned)__double2hiint(x[0]) & 0x7fffffffu, h1 = (unsigned)__double2hiint(x[1]) & 0x7fffffffu;
    const unsigned h2 = (unsigned)__double2hiint(x[2]) & 0x7fffffffu, h3 = (unsigned)__double2hiint(x[3]) & 0x7fffffffu;
    if (max(max(h0, h1), max(h2, h3)) >= (kELim << 20)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned hi = (unsigned)__double2hiint(x[k]);
            if ((hi & 0x7fffffffu) >= (kELim << 20)) {
                status |= deposit_slow(col, stride, (unsigned)__double2loint(x[k]), hi);
                x[k] = 0.0;
            }
        }
    }
    fpe_push4<F, EE, UNIFORM>(a, x, 0);
    const unsigned any = (unsigned)nonzero_bits(x[0]) | (unsigned)nonzero_bits(x[1]) | (unsigned)nonzero_bits(x[2]) |
                         (unsigned)nonzero_bits(x[3]);
    return any ? deposit_residuals(col, stride, x, status) : 0;
}

// Products that are zero, special, too large, or so small that TwoProd may be inexact.
__device__ __noinline__ unsigned product_slow(unsigned col, unsigned stride, double x, double y, double p, double e) {
    const unsigned xh = (unsigned)__double2hiint(x) & 0x7fffffffu, yh = (unsigned)__double2hiint(y) & 0x7fffffffu;
    const bool xz = (xh | (unsigned)__double2loint(x)) == 0u, yz = (yh | (unsigned)__double2loint(y)) == 0u;
    if (xh >= 0x7ff00000u || yh >= 0x7ff00000u)                    // Inf or NaN operand: p is NaN / +-Inf (IEEE)
        return deposit_slow(col, stride, (unsigned)__double2loint(p), (unsigned)__double2hiint(p));
    if (xz || yz) return 0u;                                       // exact zero product
    const unsigned ph = (unsigned)__double2hiint(p) & 0x7fffffffu;
    if (ph >= (kELim << 20)) return kStTooLarge;                   // finite operands, product >= 2^988
    // Tiny product (|p| < 2^-935, or underflowed): TwoProd may have lost bits, so redo it in exact
    // integer arithmetic.  x = mx * 2^(Ex-1075), y likewise; P = mx*my (<= 106 bits) sits at bit
    // position pos = Ex + Ey - 2150 + 1040 <= 0 relative to the accumulator LSB (2^-1040).
    unsigned long long mx = ((unsigned long long)(xh & 0xfffffu) << 32) | (unsigned)__double2loint(x);
    unsigned long long my = ((unsigned long long)(yh & 0xfffffu) << 32) | (unsigned)__double2loint(y);
    int ex = (int)(xh >> 20), ey = (int)(yh >> 20);
    if (ex == 0) ex = 1; else mx |= 1ull << 52;
    if (ey == 0) ey = 1; else my |= 1ull << 52;
    unsigned __int128 P = (unsigned __int128)mx * my;
    const int pos = ex + ey - 2150 + 1040;
    unsigned st = 0u;
    if (pos < 0) {
        const int sh = -pos;
        if (sh >= 128) {
            if (P != 0) st |= kStTooSmall;
            P = 0;
        } else {
            if (P & ((((unsigned __int128)1) << sh) - 1)) st |= kStTooSmall;   // truncated toward zero
            P >>= sh;
        }
    } else {
        P <<= pos;                                                             // pos is 0 here at most a few bits
    }
    const bool neg = ((unsigned)__double2hiint(x) ^ (unsigned)__double2hiint(y)) >> 31;
    unsigned a0 = col;
#pragma unroll 1
    for (int j = 0; j < 3; ++j, a0 += stride) {
        const unsigned long long d = (unsigned long long)(P & (unsigned __int128)kLimbMask);
        P >>= kDigits;
        if (d) sts64(a0, neg ? lds64(a0) - d : lds64(a0) + d);
    }
    return st;
}

// Four products: TwoProductFMA (ExDOT.FPE.cl:25-29), then p through all levels and the error terms
// through the lower levels (ExDOT.FPE.cl:254-258: level F-3; ExDOT.FPE.EX.4.cl: level 1 with early
// exit).  F == 0: both parts are deposited directly (ExDOT.Superacc.cl:244-253).
template <int F, bool EE, bool UNIFORM>
EXB_D int mul_add4(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1][expansions(F)], unsigned& status,
                   const double (&x)[4], const double (&y)[4]) {
    double p[4], e[4];
    unsigned worst = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        p[k] = __dmul_rn(x[k], y[k]);
        e[k] = __fma_rn(x[k], y[k], -p[k]);
        // TwoProd is exact and both parts lie inside the layout when 2^-935 <= |p| < 2^988
        worst = max(worst, ((((unsigned)__double2hiint(p[k]) & 0x7fffffffu) >> 20) - 88u));
    }
    if (worst >= (kELim - 88u)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if ((((((unsigned)__double2hiint(p[k]) & 0x7fffffffu) >> 20) - 88u)) >= (kELim - 88u)) {
                status |= product_slow(col, stride, x[k], y[k], p[k], e[k]);
                p[k] = 0.0;
                e[k] = 0.0;
            }
        }
    }
    if (F == 0) {
        deposit4<false>(col, stride, p[0], p[1], p[2], p[3], status);
        deposit4<false>(col, stride, e[0], e[1], e[2], e[3], status);     // zeros take the (cheap) slow path
        return 0;
    }
    int cnt = 0;
    fpe_push4<F, EE, UNIFORM>(a, p, 0);
    unsigned any = (unsigned)nonzero_bits(p[0]) | (unsigned)nonzero_bits(p[1]) | (unsigned)nonzero_bits(p[2]) |
                   (unsigned)nonzero_bits(p[3]);
    if (any) cnt += deposit_residuals(col, stride, p, status);
    constexpr int first = EE ? (F > 1 ? 1 : 0) : (F > 3 ? F - 3 : 0);
    fpe_push4<F, EE, UNIFORM>(a, e, first);
    any = (unsigned)nonzero_bits(e[0]) | (unsigned)nonzero_bits(e[1]) | (unsigned)nonzero_bits(e[2]) |
          (unsigned)nonzero_bits(e[3]);
    if (any) cnt += deposit_residuals(col, stride, e, status);
    return cnt;
}

// Scalar versions for the non-vector part (one element per thread and iteration).
template <int F, bool EE>
EXB_D void add1(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1][expansions(F)], unsigned& status, double x) {
    if (F == 0) {
        deposit(col, stride, x, status);
        return;
    }
    const unsigned hi = (unsigned)__double2hiint(x);
    if ((hi & 0x7fffffffu) >= (kELim << 20)) {                   // Inf / NaN / too large: keep out of the expansion
        status |= deposit_slow(col, stride, (unsigned)__double2loint(x), hi);
        return;
    }
    const double r = fpe_push1<F, EE>(a, x, 0);
    if (nonzero_bits(r)) deposit(col, stride, r, status);
}

template <int F, bool EE>
EXB_D void mul_add1(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1][expansions(F)], unsigned& status, double x,
                    double y) {
    const double p = __dmul_rn(x, y);
    const double e = __fma_rn(x, y, -p);
    if ((((((unsigned)__double2hiint(p) & 0x7fffffffu) >> 20) - 88u)) >= (kELim - 88u)) {
        status |= product_slow(col, stride, x, y, p, e);
        return;
    }
    if (F == 0) {
        deposit(col, stride, p, status);
        deposit(col, stride, e, status);
        return;
    }
    double r = fpe_push1<F, EE>(a, p, 0);
    if (nonzero_bits(r)) deposit(col, stride, r, status);
    constexpr int first = EE ? (F > 1 ? 1 : 0) : (F > 3 ? F - 3 : 0);
    r = fpe_push1<F, EE>(a, e, first);
    if (nonzero_bits(r)) deposit(col, stride, r, status);
}

// Carry-normalise a 39-limb array that lives in shared memory: pull it into registers first so the
// carry chain runs at ALU latency instead of shared-memory round trips.
EXB_D bool normalize_shared(long long* limbs) {
    long long r[kLimbs];
#pragma unroll
    for (int j = 0; j < kLimbs; ++j) r[j] = limbs[j];
    long long carry = 0;
#pragma unroll
    for (int j = 0; j < kLimbs - 1; ++j) {
        const long long v = r[j] + carry;
        carry = v >> kDigits;
        r[j] = v & kLimbMask;
    }
    r[kLimbs - 1] += carry;
#pragma unroll
    for (int j = 0; j < kLimbs; ++j) limbs[j] = r[j];
    return r[kLimbs - 1] < 0;
}

// Sum one limb row over all T columns; result valid in lane 0.
EXB_D long long row_sum(unsigned row_addr, unsigned T, unsigned lane) {
    long long s = 0;
    for (unsigned t = lane; t < T; t += 32) s += (long long)lds64(row_addr + 8u * t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    return s;
}

// ---- register window (window.cuh): the rare, out-of-line halves --------------------------------
// Window state is passed and returned BY VALUE so that it stays in registers in the hot loops.
// Ordinary path for one group of four columns + window bookkeeping (rare: out of line, by value so
// that the window stays in registers in the hot loop).
__device__ __noinline__ Window prod_slow_group(Window w, unsigned col, unsigned stride, double a0, double a1, double a2,
                                               double a3, double x0, double x1, double x2, double x3, bool mine,
                                               bool track) {
    const double xa[4] = {a0, a1, a2, a3}, xb[4] = {x0, x1, x2, x3};
    double none[1][expansions(0)];
    unsigned status = w.st;
    mul_add4<0, false, false>(col, stride, none, status, xa, xb);
    if (track) {
        unsigned hi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) hi[k] = (unsigned)__double2hiint(__dmul_rn(xa[k], xb[k]));
        win_after_slow_group<4>(w, mine, hi, false, [&](double v) { deposit(col, stride, v, status); });
    }
    w.st = status;
    return w;
}

__device__ __noinline__ Window win_flush_products(Window w, unsigned col, unsigned stride) {
    double out[4];
    win_drain(w, out);
    unsigned status = w.st;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (out[k] != 0.0) deposit(col, stride, out[k], status);
    w.st = status;
    return w;
}

// ordinary path for four single summands + window bookkeeping (rare: out of line, window by value)
__device__ __noinline__ Window sum_slow_group(Window w, unsigned col, unsigned stride, double x0, double x1, double x2,
                                              double x3, bool mine) {
    unsigned status = w.st;
    deposit(col, stride, x0, status);
    deposit(col, stride, x1, status);
    deposit(col, stride, x2, status);
    deposit(col, stride, x3, status);
    const unsigned hi[4] = {(unsigned)__double2hiint(x0), (unsigned)__double2hiint(x1), (unsigned)__double2hiint(x2),
                            (unsigned)__double2hiint(x3)};
    win_after_slow_group<4>(w, mine, hi, true, [&](double v) { deposit(col, stride, v, status); });
    w.st = status;
    return w;
}

__device__ __noinline__ Window win_flush_singles(Window w, unsigned col, unsigned stride) {
    double out[4];
    win_drain_single(w, out);
    unsigned status = w.st;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (out[k] != 0.0) deposit(col, stride, out[k], status);
    w.st = status;
    return w;
}

// Thrash control for the expansion (performance only; the sum is exact either way).  When most
// elements fall off the last level -- data whose dynamic range exceeds what F doubles can hold,
// e.g. log-uniform over 2^+-332 -- the TwoSum walk is pure overhead, so the warp bypasses the
// expansion and deposits directly for `kBypassTiles` tiles, then probes the expansion again.
constexpr int kBypassTiles = 32;
constexpr int kBypassMax = 4096;

// Everything after the vector body, shared by the streaming kernels: the scalar part (alignment head, tail, or
// the whole strided vector), the flush of the expansions, the block merge, and the last CTA's global merge,
// (optional) peer exchange, rounding and publication.  `body` = elements the vector body has consumed.
template <int F, bool EE, bool DOT>
EXB_D void reduce_finish(const ReduceParams& prm, const unsigned col, const unsigned stride, const unsigned smem_base,
                         const unsigned T, const unsigned tid, const long long body,
                         double (&a)[F > 0 ? F : 1][expansions(F)], unsigned status) {
    constexpr int kM = expansions(F);
    constexpr int kDepPerElem = DOT ? 2 : 1;               // at most one deposit per summand
    __shared__ long long block_limbs[kLimbs];
    __shared__ unsigned is_last;
    __shared__ unsigned block_status;
    // Latency regime: a single CTA reducing into an empty workspace publishes straight from shared
    // memory -- no global atomics, fence or ticket.
    const bool solo = (gridDim.x == 1) && prm.fresh && prm.finalize;
    if (tid == 0) block_status = 0;

    // ---------------- scalar part: alignment head, tail, or the whole strided vector ----------
    {
        const long long nscalar = prm.n - body;            // head + tail (or everything)
        const long long gthreads = (long long)gridDim.x * T;
        int since_norm = 0;
        long long k = (long long)blockIdx.x * T + tid;
        // Strided vectors (inca / incb != 1: no vector body at all) and long tails: eight independent
        // loads per thread in flight, consumed four at a time through the same four-summand code as
        // the vector body (per-thread early exit: lanes need not be converged here).  A warp load of
        // 32 consecutive elements touches 32 * inc * 8 bytes, so the useful fraction of every sector
        // is 1 / inc whatever is done here; what this loop buys is memory-level parallelism.
        for (; k + 7 * gthreads < nscalar; k += 8 * gthreads) {
            double xa[8], xb[DOT ? 8 : 1];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long kk = k + j * gthreads;
                const long long idx = kk < prm.head ? kk : kk + body;
                xa[j] = ldg64(prm.a + idx * prm.inca);
                if (DOT) xb[j] = ldg64(prm.b + idx * prm.incb);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double x4[4] = {xa[4 * h], xa[4 * h + 1], xa[4 * h + 2], xa[4 * h + 3]};
                if (DOT) {
                    const double y4[4] = {xb[4 * h], xb[4 * h + 1], xb[4 * h + 2], xb[4 * h + 3]};
                    mul_add4<F, EE, false>(col, stride, a, status, x4, y4);
                } else if (F == 0) {
                    deposit4<false>(col, stride, x4[0], x4[1], x4[2], x4[3], status);
                } else {
                    add4<F, EE, false>(col, stride, a, status, x4);
                }
            }
            since_norm += 8 * kDepPerElem;
            if (since_norm > kMaxDepositsPerNormalize - 8 * kDepPerElem - 2 * kM * (F + 2)) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
        for (; k < nscalar; k += gthreads) {
            const long long idx = k < prm.head ? k : k + body;
            if (DOT) mul_add1<F, EE>(col, stride, a, status, prm.a[idx * prm.inca], prm.b[idx * prm.incb]);
            else add1<F, EE>(col, stride, a, status, prm.a[idx * prm.inca]);
            since_norm += kDepPerElem;
            if (since_norm > kMaxDepositsPerNormalize - kDepPerElem - 2 * kM * (F + 2)) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
    }
    if (F > 0) {
#pragma unroll
        for (int i = 0; i < F; ++i)
#pragma unroll
            for (int m = 0; m < kM; ++m) deposit(col, stride, a[i][m], status);
    }
    bound_column(col, stride);
    __syncthreads();                                       // block_status initialised; all columns final
    if (status) {
        if (solo) atomicOr(&block_status, status);
        else atomicOr(&prm.ws->status, status);
    }

    // ---------------- block merge: 39 row sums -> normalise -> global accumulator --------------
    {
        const unsigned warp = tid >> 5, ln = tid & 31u, nwarps = T >> 5;
        for (unsigned j = warp; j < (unsigned)kLimbs; j += nwarps) {
            const long long s = row_sum(smem_base + j * stride, T, ln);   // |s| < T * (2^52 + 2^11) < 2^62
            if (ln == 0) block_limbs[j] = s;
        }
        __syncthreads();
        if (!solo) {
            if (tid == 0) normalize_shared(block_limbs);
            __syncthreads();
            for (unsigned j = tid; j < (unsigned)kLimbs; j += T) {
                const long long v = block_limbs[j];
                if (v != 0) atomicAdd(&prm.ws->gacc[j], (unsigned long long)v);
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned ticket = atomicAdd(&prm.ws->counter, 1u);
                is_last = (ticket == gridDim.x - 1);
            }
            __syncthreads();
        }
    }

    // ---------------- last CTA: normalise the global accumulator, publish ----------------------
    if (solo || is_last) {
        if (!solo) {
            __threadfence();
            for (unsigned j = tid; j < (unsigned)kLimbs; j += T)
                block_limbs[j] = (long long)atomicExch(&prm.ws->gacc[j], 0ull);
            __syncthreads();
        }
        __shared__ unsigned final_status;
        if (tid == 0) {
            unsigned st;
            if (solo) st = block_status;
            else st = prm.finalize && !prm.keep ? atomicExch(&prm.ws->status, 0u) : atomicOr(&prm.ws->status, 0u);
            normalize_shared(block_limbs);
            final_status = st;
        }
        __syncthreads();
        // ---- fused multi-GPU exchange over peer memory (only the closing launch of a reduction) ----
        if (prm.finalize && prm.nranks > 1) {
            const unsigned set = (unsigned)(prm.epoch & 1ull);
            const unsigned st = final_status;
            for (unsigned j = tid; j < (unsigned)kMsgWords * (unsigned)prm.nranks; j += T) {
                const unsigned r = j / kMsgWords, w = j % kMsgWords;
                const unsigned long long v = w < (unsigned)kLimbs ? (unsigned long long)block_limbs[w]
                                                                  : (unsigned long long)((st >> (w - kLimbs)) & 1u);
                st_relaxed_sys(&prm.peers[r]->slot[set][prm.rank].data[w], v);
            }
            __threadfence_system();
            __syncthreads();
            if (tid < (unsigned)prm.nranks) st_release_sys(&prm.peers[tid]->slot[set][prm.rank].seq, prm.epoch);
            // wait for every rank's contribution of this epoch in MY mailbox (bounded spin)
            __shared__ unsigned timed_out;
            if (tid == 0) timed_out = 0;
            __syncthreads();
            if (tid < (unsigned)prm.nranks) {
                const unsigned long long* seq = &prm.peers[prm.rank]->slot[set][tid].seq;
                const long long t0 = clock64();
                while (ld_acquire_sys(seq) != prm.epoch) {
                    if (clock64() - t0 > 20000000000ll) {        // ~10 s: a peer never arrived
                        atomicOr(&timed_out, 1u);
                        break;
                    }
                    __nanosleep(200);
                }
            }
            __syncthreads();
            __shared__ unsigned long long merged[kMsgWords];
            for (unsigned w = tid; w < (unsigned)kMsgWords; w += T) {
                unsigned long long sum = 0;
                for (int r = 0; r < prm.nranks; ++r) sum += ld_relaxed_sys(&prm.peers[prm.rank]->slot[set][r].data[w]);
                merged[w] = sum;                                 // <= 8 normalised limbs: no overflow
            }
            __syncthreads();
            if (tid == 0) {
                unsigned stm = timed_out ? kStPeerTimeout : 0u;
                for (int k = 0; k < kFlagSlots; ++k)
                    if (merged[kLimbs + k] != 0) stm |= 1u << k;
                for (int j = 0; j < kLimbs; ++j) block_limbs[j] = (long long)merged[j];
                normalize_shared(block_limbs);
                final_status = stm;
            }
            __syncthreads();
        }
        if (tid == 0) {
            const unsigned st = final_status;
            const bool neg = block_limbs[kLimbs - 1] < 0;
            if (prm.finalize) {
                Result* out = prm.out;
                for (int j = 0; j < kLimbs; ++j) out->limbs[j] = block_limbs[j];
                double v;
                if ((st & kStNaN) || ((st & kStPosInf) && (st & kStNegInf))) v = __longlong_as_double(0x7ff8000000000000ll);
                else if (st & kStPosInf) v = __longlong_as_double(0x7ff0000000000000ll);
                else if (st & kStNegInf) v = __longlong_as_double(0xfff0000000000000ll);
                else v = prm.round_mode ? round_exact(block_limbs, neg) : round_ref_compat(block_limbs, neg);
                out->value = v;
                out->status = st;
                for (int k = 0; k < kFlagSlots; ++k) out->flagcnt[k] = (st >> k) & 1u;
            }
            if (!solo) prm.ws->counter = 0;
        }
        __syncthreads();
        // leave the (normalised) partial sum in the workspace unless this call closes the reduction
        if (!solo && (!prm.finalize || prm.keep))
            for (unsigned j = tid; j < (unsigned)kLimbs; j += T) prm.ws->gacc[j] = (unsigned long long)block_limbs[j];
    }
}

// (Measured and rejected, round 1: moving the thrash bypass or the expansion walk of THIS kernel into out-of-line
// functions, as exblas_reduce0_kernel does with its window loop.  The expansion walk contains calls (deposits of
// residuals), and a function that is itself called and calls on spills its prefetch slots: 1.9 instead of
// 6.0 TB/s on narrow data; a call to an out-of-line bypass from inside this loop cost the walk 40 %.)
template <int F, bool EE, bool DOT, int U, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) exblas_reduce_kernel(const ReduceParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;
    const unsigned tid = threadIdx.x;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned stride = 8u * T;
    const unsigned col = smem_base + 8u * tid;

    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
    // columns are thread-private: no barrier needed before use

    unsigned status = 0;
    constexpr int kM = expansions(F);
    double a[F > 0 ? F : 1][kM];
#pragma unroll
    for (int i = 0; i < (F > 0 ? F : 1); ++i)
#pragma unroll
        for (int m = 0; m < kM; ++m) a[i][m] = 0.0;

    constexpr int kDepPerElem = DOT ? 2 : 1;               // at most one deposit per summand
    constexpr int kDepPerTile = 4 * U * kDepPerElem;
    const long long TILE = (long long)T * 4 * U;

    // ---------------- vector body: full tiles, 256-bit loads, rolling prefetch ----------------
    if (prm.ntiles > 0 && (long long)blockIdx.x < prm.ntiles) {
        // this CTA owns tiles blockIdx.x, blockIdx.x + grid, ...: `iters` of them
        const unsigned iters = (unsigned)((prm.ntiles - 1 - blockIdx.x) / gridDim.x) + 1u;
        const long long tile_step = (long long)gridDim.x * TILE;               // elements between my tiles
        const double* pa = prm.a + prm.head + (long long)blockIdx.x * TILE + (long long)tid * 4;
        const double* pb = DOT ? prm.b + prm.head + (long long)blockIdx.x * TILE + (long long)tid * 4 : nullptr;
        const long long vstep = (long long)T * 4;                               // elements between my vectors
        Vec4 va[U];
        Vec4 vb[DOT ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            va[u] = ldg256(pa + u * vstep);
            if (DOT) vb[u] = ldg256(pb + u * vstep);
        }
        int since_norm = 0;
        int bypass = 0, backoff = kBypassTiles;
        for (unsigned it = 0; it < iters; ++it) {
            pa += tile_step;
            if (DOT) pb += tile_step;
            const bool has_next = it + 1 < iters;
            const bool direct = (F == 0) || (prm.adaptive && bypass > 0);
            int deposits = 0;
            // Two separately unrolled tile bodies (only one is hot at a time, so each fits the
            // instruction cache): direct deposits, or the expansion walk.  Each vector slot is
            // refilled for the next tile right after it is consumed.
            if (direct) {
                // ExSUM: one vote per tile on the signs of everything in the register window; an
                // all-positive tile (the reference generator's data, norms, energies ...) takes the
                // sign-free deposit, 7 integer instructions per element cheaper.
                bool all_pos = false;
#ifndef EXB_NO_POS
                if (!DOT) {
                    unsigned hs = 0u;
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        hs |= (unsigned)__double2hiint(va[u].x) | (unsigned)__double2hiint(va[u].y) |
                              (unsigned)__double2hiint(va[u].z) | (unsigned)__double2hiint(va[u].w);
                    all_pos = !__any_sync(0xffffffffu, (int)hs < 0);
                }
#endif
                if (!DOT && all_pos) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        deposit4<true>(col, stride, va[u].x, va[u].y, va[u].z, va[u].w, status);
                        if (has_next) va[u] = ldg256(pa + u * vstep);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (DOT) {
                            const double x[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
                            const double y[4] = {vb[u].x, vb[u].y, vb[u].z, vb[u].w};
                            double none[1][expansions(0)];
                            mul_add4<0, false, true>(col, stride, none, status, x, y);
                        } else {
                            deposit4<false>(col, stride, va[u].x, va[u].y, va[u].z, va[u].w, status);
                        }
                        if (has_next) {
                            va[u] = ldg256(pa + u * vstep);
                            if (DOT) vb[u] = ldg256(pb + u * vstep);
                        }
                    }
                }
            } else if (F > 0) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double x[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
                    if (DOT) {
                        const double y[4] = {vb[u].x, vb[u].y, vb[u].z, vb[u].w};
                        deposits += mul_add4<F, EE, true>(col, stride, a, status, x, y);
                    } else {
                        deposits += add4<F, EE, true>(col, stride, a, status, x);
                    }
                    if (has_next) {
                        va[u] = ldg256(pa + u * vstep);
                        if (DOT) vb[u] = ldg256(pb + u * vstep);
                    }
                }
            }
            if (F > 0 && prm.adaptive) {
                if (bypass > 0) {
                    --bypass;
                } else {
                    // Warp-uniform decision.  A deposit is an out-of-line, divergent call: once more than
                    // ~1.5 % of the warp's summands need one, nearly every vector step pays for it and
                    // depositing everything directly is cheaper.  Back off exponentially while it lasts.
                    const int total = __reduce_add_sync(0xffffffffu, deposits);
                    if (total * 64 >= 32 * kDepPerTile) {
                        bypass = backoff;
                        backoff = min(backoff * 16, kBypassMax);   // a second thrashing probe in a row: stay away for long
                    } else {
                        backoff = kBypassTiles;
                    }
                }
            }
            since_norm += kDepPerTile;
            if (since_norm > kMaxDepositsPerNormalize - kDepPerTile - 2 * kM * (F + 2)) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
        bound_column(col, stride);
    }

    reduce_finish<F, EE, DOT>(prm, col, stride, smem_base, T, tid, prm.ntiles * TILE, a, status);
}

// ------------------------------------------------------------------------------------------------
// Superaccumulator-only streaming kernel (fpe < 2 for ExSUM, fpe < 3 for ExDOT): the reference's
// ExSUM.Superacc.cl:212-294 / ExDOT.Superacc.cl:218-320 mode, with a register window (window.cuh) in
// front of the shared-memory superaccumulator.
//
// Geometry: a ROW is T * 4 consecutive elements (one 256-bit vector per thread); this CTA owns rows
// blockIdx, blockIdx + grid, ...  Two loops over its rows, each with its own prefetch depth, so that
// each gets the registers it needs (a prefetch slot that spills to local memory stalls on its own
// load and serialises the whole window -- measured: 4.3 instead of 7.2 TB/s):
//   1. window loop, DW rows in flight: a vector whose four summands (products) lie inside the window of
//      EVERY lane of the warp (one vote) is accumulated in registers -- 4 FP64 + 4 integer instructions
//      per summand, 2 + 8 FP64 + 8 integer per product, no shared-memory traffic.  Any other vector
//      takes the ordinary deposits out of line.  Two blocks of DW rows in a row that mostly miss
//      (wide-range data such as the log-uniform benchmark vector) end this loop for good;
//   2. direct loop, DD >= DW rows in flight: every summand is deposited (sign-free when a whole block of
//      rows is positive, as in exblas_reduce_kernel).
// Tail, merge and publication are reduce_finish, the same code as the expansion kernels.
// ------------------------------------------------------------------------------------------------
// Loop 1 of exblas_reduce0_kernel, out of line ON PURPOSE: as separate functions the window loops get their own
// register allocation, so that their window state and temporaries cannot push the direct loop's prefetch slots or
// pointers into local memory (inlined, the direct loop lost 8-12 %).  Each consumes rows 0, DW, 2 DW, ... of this
// CTA while its window holds and returns the number of rows consumed (a multiple of DW).  Rows it had in flight
// but not consumed when it stops are simply loaded again by the caller (they come from L2).
// The windows are the W-digit ones of window.cuh with the WARP-UNIFORM, range-covering anchoring: a vector that
// misses moves every lane's window so that it admits all exponents that have missed so far; [emin, emax] travels
// from one attempt to the next, so that a warp whose data cannot fit (log-uniform 2^+-332: the first vector
// already spans more than any window) leaves after ONE row and skips the wider attempt altogether.

// ---- ExDOT: W-digit product window (WindowP<W>): W = 3 (50 binades, 10 FP64 instructions per product) first,
// then W = 5 (154 binades, 22) for products further apart, e.g. ill-conditioned dot products ----
template <int W>
__device__ __noinline__ WindowP<W> prodw_slow_group(WindowP<W> w, unsigned col, unsigned stride, double a0, double a1, double a2,
                                                    double a3, double x0, double x1, double x2, double x3, bool mine) {
    const double xa[4] = {a0, a1, a2, a3}, xb[4] = {x0, x1, x2, x3};
    double none[1][expansions(0)];
    unsigned status = w.st;
    mul_add4<0, false, false>(col, stride, none, status, xa, xb);
    // exponent range of this group's products over the WHOLE warp (zeros and specials do not count): every lane
    // re-anchors to the same window, the one that admits everything that has missed so far
    int gmin = 4096, gmax = -4096;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int E = (int)(((unsigned)__double2hiint(__dmul_rn(xa[k], xb[k])) >> 20) & 0x7ffu);
        if (E != 0 && E != 0x7ff) {
            gmin = min(gmin, E - 1023);
            gmax = max(gmax, E - 1023);
        }
    }
    gmin = __reduce_min_sync(0xffffffffu, gmin);
    gmax = __reduce_max_sync(0xffffffffu, gmax);
    winp_cover<W>(w, gmin, gmax, [&](double v) { deposit(col, stride, v, status); });
    (void)mine;
    w.st = status;
    return w;
}

template <int W>
__device__ __noinline__ WindowP<W> winp_flush(WindowP<W> w, unsigned col, unsigned stride) {
    double out[W + 1];
    winp_drain(w, out);
    unsigned status = w.st;
#pragma unroll
    for (int k = 0; k <= W; ++k)
        if (out[k] != 0.0) deposit(col, stride, out[k], status);
    w.st = status;
    return w;
}

template <int DW, int W, bool EARLY = false>
__device__ __noinline__ unsigned reduce0_window_rows_wide(const double* pa, const double* pb, const long long row_step,
                                                          const unsigned iters, const unsigned col, const unsigned stride,
                                                          unsigned* status_io, int* range_io) {
    if (range_io[1] - range_io[0] + 1 > 50 + 52 * (W - 3)) return 0u;     // what has missed so far cannot fit this window
    Vec4 va[DW];
    Vec4 vb[DW];
    unsigned loaded = 0;
    auto load_row = [&](int u) {
        va[u] = ldg256(pa);
        vb[u] = ldg256(pb);
        pa += row_step;
        pb += row_step;
        ++loaded;
    };
#pragma unroll
    for (int u = 0; u < DW; ++u)
        if (loaded < iters) load_row(u);
    WindowP<W> w;
    winp_reset(w);
    w.emin = range_io[0];
    w.emax = range_io[1];
    unsigned k = 0;
    int since_norm = 0;
    for (int bad = 0; k + DW <= iters && bad < 2; k += DW) {
        int missed = 0;
#pragma unroll
        for (int u = 0; u < DW; ++u) {
            const double a0 = va[u].x, a1 = va[u].y, a2 = va[u].z, a3 = va[u].w;
            const double b0 = vb[u].x, b1 = vb[u].y, b2 = vb[u].z, b3 = vb[u].w;
            if (EARLY && loaded < iters) load_row(u);
            const double p0 = __dmul_rn(a0, b0), p1 = __dmul_rn(a1, b1), p2 = __dmul_rn(a2, b2), p3 = __dmul_rn(a3, b3);
            const unsigned k0 = ((unsigned)__double2hiint(p0) & 0x7fffffffu) - w.key0;
            const unsigned k1 = ((unsigned)__double2hiint(p1) & 0x7fffffffu) - w.key0;
            const unsigned k2 = ((unsigned)__double2hiint(p2) & 0x7fffffffu) - w.key0;
            const unsigned k3 = ((unsigned)__double2hiint(p3) & 0x7fffffffu) - w.key0;
            if constexpr (W >= 5) {
                // two products per vote: the temporaries of four interleaved 22-instruction splits + the 5-digit state do
                // not fit 128 registers, and what spills is a prefetch slot (which then stalls on its own load)
                const bool mine01 = max(k0, k1) < w.span, mine23 = max(k2, k3) < w.span;
                if (__all_sync(0xffffffffu, mine01)) {
                    winp_add_product(w, p0, __fma_rn(a0, b0, -p0));
                    winp_add_product(w, p1, __fma_rn(a1, b1, -p1));
                    w.cnt += 2u;
                } else {
                    w = prodw_slow_group<W>(w, col, stride, a0, a1, 0.0, 0.0, b0, b1, 0.0, 0.0, mine01);
                    ++missed;
                }
                if (__all_sync(0xffffffffu, mine23)) {
                    winp_add_product(w, p2, __fma_rn(a2, b2, -p2));
                    winp_add_product(w, p3, __fma_rn(a3, b3, -p3));
                    w.cnt += 2u;
                } else {
                    w = prodw_slow_group<W>(w, col, stride, a2, a3, 0.0, 0.0, b2, b3, 0.0, 0.0, mine23);
                    ++missed;
                }
            } else {
                const bool mine = max(max(k0, k1), max(k2, k3)) < w.span;
                if (__all_sync(0xffffffffu, mine)) {
                    winp_add_product(w, p0, __fma_rn(a0, b0, -p0));
                    winp_add_product(w, p1, __fma_rn(a1, b1, -p1));
                    winp_add_product(w, p2, __fma_rn(a2, b2, -p2));
                    winp_add_product(w, p3, __fma_rn(a3, b3, -p3));
                    w.cnt += 4u;
                } else {
                    w = prodw_slow_group<W>(w, col, stride, a0, a1, a2, a3, b0, b1, b2, b3, mine);
                    missed += 2;
                }
            }
            // the slot is refilled AFTER its row has been consumed: its registers are free by then, which keeps the
            // W = 5 window (27 registers of state) from pushing a load in flight into local memory
            if (!EARLY && loaded < iters) load_row(u);          // row k + DW + u
        }
        bad = (missed > DW) ? bad + 1 : 0;                      // `missed` counts half rows here
        if (w.span == 0u) bad = 2;                              // what has missed no longer fits the window (warp-uniform): leave
        if (w.cnt > (unsigned)(kWinFlushEvery - 4 * DW)) {
            w = winp_flush<W>(w, col, stride);
            since_norm += W + 1;
        }
        since_norm += missed * (8 + W + 1);                     // ordinary deposits + a drain when the window moves
        if (since_norm > kMaxDepositsPerNormalize - 2 * DW * (8 + W + 1) - (W + 1) - 16) {   // room for one more block of half rows + a flush
            bound_column(col, stride);
            since_norm = 0;
        }
    }
    w = winp_flush<W>(w, col, stride);
    *status_io |= w.st;
    range_io[0] = w.emin;
    range_io[1] = w.emax;
    bound_column(col, stride);
    return k;
}

// ---- ExSUM: W-digit single-summand window: W = 2 (51 binades, 4 FP64 instructions per summand) first, then
// W = 3 (103 binades, 7), e.g. for the reference's ill-conditioned generator (init_ill_cond: ~65-85 binades) ----
template <int W>
__device__ __noinline__ WindowP<W> sumw_slow_group(WindowP<W> w, unsigned col, unsigned stride, double x0, double x1, double x2,
                                                   double x3) {
    unsigned status = w.st;
    deposit(col, stride, x0, status);
    deposit(col, stride, x1, status);
    deposit(col, stride, x2, status);
    deposit(col, stride, x3, status);
    const double xs[4] = {x0, x1, x2, x3};
    int gmin = 4096, gmax = -4096;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int E = (int)(((unsigned)__double2hiint(xs[k]) >> 20) & 0x7ffu);
        if (E != 0 && E != 0x7ff) {
            gmin = min(gmin, E - 1023);
            gmax = max(gmax, E - 1023);
        }
    }
    gmin = __reduce_min_sync(0xffffffffu, gmin);
    gmax = __reduce_max_sync(0xffffffffu, gmax);
    wins_cover<W>(w, gmin, gmax, [&](double v) { deposit(col, stride, v, status); });
    w.st = status;
    return w;
}

template <int W>
__device__ __noinline__ WindowP<W> wins_flush(WindowP<W> w, unsigned col, unsigned stride) {
    double out[W + 1];
    wins_drain(w, out);
    unsigned status = w.st;
#pragma unroll
    for (int k = 0; k <= W; ++k)
        if (out[k] != 0.0) deposit(col, stride, out[k], status);
    w.st = status;
    return w;
}

template <int DW, int W>
__device__ __noinline__ unsigned reduce0_window_rows_wide_sum(const double* pa, const long long row_step, const unsigned iters,
                                                              const unsigned col, const unsigned stride, unsigned* status_io,
                                                              int* range_io) {
    if (range_io[1] - range_io[0] + 1 > 51 + 52 * (W - 2)) return 0u;
    Vec4 va[DW];
    unsigned loaded = 0;
    auto load_row = [&](int u) {
        va[u] = ldg256(pa);
        pa += row_step;
        ++loaded;
    };
#pragma unroll
    for (int u = 0; u < DW; ++u)
        if (loaded < iters) load_row(u);
    WindowP<W> w;
    winp_reset(w);
    w.emin = range_io[0];
    w.emax = range_io[1];
    unsigned k = 0;
    int since_norm = 0;
    for (int bad = 0; k + DW <= iters && bad < 2; k += DW) {
        int missed = 0;
#pragma unroll
        for (int u = 0; u < DW; ++u) {
            const double a0 = va[u].x, a1 = va[u].y, a2 = va[u].z, a3 = va[u].w;
            const unsigned k0 = ((unsigned)__double2hiint(a0) & 0x7fffffffu) - w.key0;
            const unsigned k1 = ((unsigned)__double2hiint(a1) & 0x7fffffffu) - w.key0;
            const unsigned k2 = ((unsigned)__double2hiint(a2) & 0x7fffffffu) - w.key0;
            const unsigned k3 = ((unsigned)__double2hiint(a3) & 0x7fffffffu) - w.key0;
            const bool mine = max(max(k0, k1), max(k2, k3)) < w.span;
            if (__all_sync(0xffffffffu, mine)) {
                wins_add(w, a0);
                wins_add(w, a1);
                wins_add(w, a2);
                wins_add(w, a3);
                w.cnt += 4u;
            } else {
                w = sumw_slow_group<W>(w, col, stride, a0, a1, a2, a3);
                ++missed;
            }
            if (loaded < iters) load_row(u);                    // row k + DW + u (after the slot has been consumed)
        }
        bad = (2 * missed > DW) ? bad + 1 : 0;
        if (w.span == 0u) bad = 2;
        if (w.cnt > (unsigned)(kWinFlushEvery - 4 * DW)) {
            w = wins_flush<W>(w, col, stride);
            since_norm += W + 1;
        }
        since_norm += missed * (4 + W + 1);
        if (since_norm > kMaxDepositsPerNormalize - DW * (4 + W + 1) - (W + 1) - 16) {       // room for one more block + a flush
            bound_column(col, stride);
            since_norm = 0;
        }
    }
    w = wins_flush<W>(w, col, stride);
    *status_io |= w.st;
    range_io[0] = w.emin;
    range_io[1] = w.emax;
    bound_column(col, stride);
    return k;
}

template <bool DOT, int DW, int DD, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) exblas_reduce0_kernel(const ReduceParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;
    const unsigned tid = threadIdx.x;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned stride = 8u * T;
    const unsigned col = smem_base + 8u * tid;
    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);

    unsigned status = 0;
    constexpr int kDepPerElem = DOT ? 2 : 1;
    const long long ROW = (long long)T * 4;

    if (prm.ntiles > 0 && (long long)blockIdx.x < prm.ntiles) {
        const unsigned iters = (unsigned)((prm.ntiles - 1 - blockIdx.x) / gridDim.x) + 1u;     // my rows
        const long long row_step = (long long)gridDim.x * ROW;
        const double* pa = prm.a + prm.head + (long long)blockIdx.x * ROW + (long long)tid * 4;  // my first row
        const double* pb = DOT ? prm.b + prm.head + (long long)blockIdx.x * ROW + (long long)tid * 4 : nullptr;
        // ---------------- loop 1: register window (out of line) ----------------
        unsigned k = 0;                                        // rows consumed so far
        // (vectors too short for the windows to pay -- fewer than 32 rows per CTA, n < ~2^23 -- go straight to loop 2)
        if constexpr (DW > 0) if (prm.window && iters >= 32u) {
            unsigned st1 = 0;
            int range[2] = {4096, -4096};                      // exponents that have missed so far, warp-uniform
            if constexpr (DOT) {
                if (prm.window != 3)
                    k = reduce0_window_rows_wide<DW, 3>(pa, pb, row_step, iters, col, stride, &st1, range);
                if (k < iters && prm.window > 1)
                    // (two rows in flight, refilled BEFORE the row is consumed: measured best of 2 / 3 / 4 rows, early / late
                    // refill: 6.7 against 5.0-6.4 TB/s on ill-conditioned data; the wide loop spends ~45 instructions per product)
                    k += reduce0_window_rows_wide<DW, 5, true>(pa + (long long)k * row_step, pb + (long long)k * row_step, row_step,
                                                               iters - k, col, stride, &st1, range);
            } else {
                if (prm.window != 3)
                    k = reduce0_window_rows_wide_sum<DW, 2>(pa, row_step, iters, col, stride, &st1, range);
                if (k < iters && prm.window > 1)
                    k += reduce0_window_rows_wide_sum<DW + 2, 3>(pa + (long long)k * row_step, row_step, iters - k, col, stride, &st1, range);
            }
            status |= st1;
            pa += (long long)k * row_step;
            if (DOT) pb += (long long)k * row_step;
        }
        // ---------------- loop 2: direct deposits, DD rows in flight; slot u holds row k + u ----------------
        Vec4 va[DD];
        Vec4 vb[DOT ? DD : 1];
        unsigned loaded = k;                                   // rows loaded (or consumed by loop 1) so far
        auto load_row = [&](int u) {
            va[u] = ldg256(pa);
            if (DOT) vb[DOT ? u : 0] = ldg256(pb);
            pa += row_step;
            if (DOT) pb += row_step;
            ++loaded;
        };
        int since_norm = 0;
#pragma unroll
        for (int u = 0; u < DD; ++u)
            if (loaded < iters) load_row(u);
        auto consume = [&](int u, bool all_pos) {
            if (DOT) {
                const double x[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
                const double y[4] = {vb[DOT ? u : 0].x, vb[DOT ? u : 0].y, vb[DOT ? u : 0].z, vb[DOT ? u : 0].w};
                double none[1][expansions(0)];
                mul_add4<0, false, true>(col, stride, none, status, x, y);
            } else if (all_pos) {
                deposit4<true>(col, stride, va[u].x, va[u].y, va[u].z, va[u].w, status);
            } else {
                deposit4<false>(col, stride, va[u].x, va[u].y, va[u].z, va[u].w, status);
            }
        };
        for (; k + DD <= iters; k += DD) {
            // ExSUM: one vote per block of rows on the signs of everything in the register window; an all-positive
            // block (the reference generator's data, norms, energies ...) takes the sign-free deposit
            bool all_pos = false;
            if (!DOT) {
                unsigned hs = 0u;
#pragma unroll
                for (int u = 0; u < DD; ++u)
                    hs |= (unsigned)__double2hiint(va[u].x) | (unsigned)__double2hiint(va[u].y) |
                          (unsigned)__double2hiint(va[u].z) | (unsigned)__double2hiint(va[u].w);
                all_pos = !__any_sync(0xffffffffu, (int)hs < 0);
            }
            if (all_pos) {
#pragma unroll
                for (int u = 0; u < DD; ++u) {
                    consume(u, true);
                    if (loaded < iters) load_row(u);
                }
            } else {
#pragma unroll
                for (int u = 0; u < DD; ++u) {
                    consume(u, false);
                    if (loaded < iters) load_row(u);
                }
            }
            since_norm += 4 * DD * kDepPerElem;
            if (since_norm > kMaxDepositsPerNormalize - 8 * DD * kDepPerElem - 16) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
        // rows left in the slots (< DD)
#pragma unroll
        for (int u = 0; u < DD; ++u)
            if (k + u < iters) consume(u, false);
        bound_column(col, stride);
    }
    double none[1][expansions(0)];
    reduce_finish<0, false, DOT>(prm, col, stride, smem_base, T, tid, prm.ntiles * ROW, none, status);
}

// Multi-GPU epilogue: the result slot's limbs and flag counters have been summed over ranks by an
// integer all-reduce; normalise, rebuild the status word and round.  Every rank runs this on the
// same integers, so every rank gets the same bits.
__global__ void exblas_finalize_kernel(Result* res, int round_mode) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long acc[kLimbs];
        for (int j = 0; j < kLimbs; ++j) acc[j] = res->limbs[j];
        unsigned st = 0;
        for (int k = 0; k < kFlagSlots; ++k)
            if (res->flagcnt[k] != 0) st |= (1u << k);
        res->value = finalize_value(acc, st, round_mode);
        res->status = st;
        for (int j = 0; j < kLimbs; ++j) res->limbs[j] = acc[j];
        for (int k = 0; k < kFlagSlots; ++k) res->flagcnt[k] = (st >> k) & 1u;
    }
}

}  // namespace exb
