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

// Peer-memory mailbox for the fused limb exchange over NVLink / NVSwitch.  Rank r's closing warp stores its
// message into slot [epoch & 1][r] of EVERY rank's mailbox with plain 8-byte peer stores; every word carries its
// own 12-bit sequence tag next to a 52-bit payload (the "LL" idea: data and flag travel in ONE store, so there
// is no fence, no separate flag store and no second NVLink traversal).  Each rank polls the words of all its
// slots until they carry this epoch's tag, sums them as integers, normalises and rounds -- inside the reduction
// kernel, with no NCCL call and no extra launch.  Two slot sets suffice: a rank cannot start epoch e+2 before
// every peer has finished reading epoch e (it needs their epoch e+1 data first).
//   word w < 38 : limb w (normalised: in [0, 2^52))
//   word 38, 39 : limb 38 (signed, 64 bits) as low 52 bits / arithmetic high 12 bits
//   word 40     : status flags
constexpr int kMaxPeers = 8;
constexpr int kMsgWords = kLimbs + 2;            // 41
struct MailSlot {
    unsigned long long data[48];                 // 41 used; 384 B per slot
};
struct Mailbox {
    MailSlot slot[2][kMaxPeers];
};
EXB_HD unsigned long long peer_tag(unsigned long long epoch) { return (epoch % 4095ull) + 1ull; }   // never 0 (= empty mailbox)

constexpr int kPhaseSlots = 16;          // per-CTA globaltimer stamps (diagnostics: option "phase_timing")

struct ReduceParams {
    const double* a;                     // already offset by `offset`
    const double* b;                     // ExDOT only
    long long n;                         // number of elements
    long long inca, incb;                // element strides
    long long head;                      // scalar elements before the 32-byte aligned vector region (0 when nvec == 0)
    long long nvec;                      // full 256-bit vectors after the head (0 => everything is scalar)
    long long iters;                     // tiles (of T * U vectors) every CTA streams in the unrolled vector body
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
    unsigned long long peer_timeout_ns;  // give up waiting for a peer after this long (0 = wait for ever)
    int l2_prefetch;                     // expansion kernel: bulk L2 prefetch distance in tiles (0 = off; performance only)
    int handoff_tiles;                   // ExDOT expansion kernels: hand a thrashing vector of at least this many tiles per CTA to the 5-digit window loop (0 = never)
    unsigned long long* phase;           // optional [gridDim.x][kPhaseSlots] globaltimer stamps (nullptr = off)
};

EXB_D Vec4 ldg256(const double* p) {
    Vec4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
}

// TMA-engine prefetch of a contiguous run of global memory into L2 (no destination in shared memory, no completion to
// wait for): SASS UBLKPF.  `bytes` must be a multiple of 16 and `p` 16-byte aligned.
EXB_D void bulk_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
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

EXB_D unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define EXB_PHASE(k)                                                                                         \
    do {                                                                                                     \
        if (prm.phase != nullptr && threadIdx.x == 0) prm.phase[(size_t)blockIdx.x * kPhaseSlots + (k)] = globaltimer_ns(); \
    } while (0)

EXB_D bool nonzero_bits(double x) {
    return (((unsigned)__double2hiint(x) & 0x7fffffffu) | (unsigned)__double2loint(x)) != 0u;
}

// ---- per-thread state --------------------------------------------------------------------------
// (col, stride) of the private accumulator column, status flags, and M independent expansions
// a[level][m] in registers.  Two expansions per thread (F <= 4) give the FP64 pipe independent
// TwoSum chains to overlap (elements 0,2 of a vector feed expansion 0, elements 1,3 expansion 1);
// the exact sum does not care how summands are distributed over expansions.
// (Measured, round 2: two expansions for F > 4 as well -- profiles/ab_sustained_r02.jsonl, EXB_EXP2_MAXF = 6 / 8 -- LOSE
// 17-26 % on the large early-exit ExDOT kernels and spill in the F = 8 ones; the serial chain is not their limiter.)
#ifndef EXB_EXP2_MAXF
#define EXB_EXP2_MAXF 4
#endif
__host__ __device__ constexpr int expansions(int f) { return f <= EXB_EXP2_MAXF ? 2 : 1; }   // large F: registers go to the load window instead

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
EXB_D int fpe_push4(double (&a)[F > 0 ? F : 1][expansions(F)], double (&x)[4], int first) {
    constexpr int M1 = expansions(F) - 1;
    int walked = 0;                                        // levels visited (warp-uniform when UNIFORM)
#pragma unroll
    for (int i = 0; i < F; ++i) {
        if (i < first) continue;
        ++walked;
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
    return walked;
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
EXB_D int add4(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1][expansions(F)], unsigned& status, double (&x)[4],
               int* walked = nullptr) {
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
    const int lv = fpe_push4<F, EE, UNIFORM>(a, x, 0);
    if (walked) *walked += lv;
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
template <int F, bool EE, bool UNIFORM, bool P3 = UNIFORM>
EXB_D int mul_add4(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1][expansions(F)], unsigned& status,
                   const double (&x)[4], const double (&y)[4], int* walked = nullptr) {
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
#ifndef EXB_NO_PRODUCT3
        // p and its error term share their middle limb: three read-modify-writes per product instead of four.  (In the
        // streaming loops of the ExSUM / ExDOT / batched kernels only -- P3: in the out-of-line miss paths of the window
        // kernels and in the ExGEMV window kernels' fallback bodies the extra code cost the surrounding window loops
        // registers: -6 % on ill-conditioned ExDOT, -15 % on ExGEMV 'T'.)
        // (Exact zero products -- zeros in the data, products diverted above -- pass the test: they add zeros.)
        if (P3)
        if (product3_ok((unsigned)__double2hiint(p[0]), (unsigned)__double2loint(p[0])) & product3_ok((unsigned)__double2hiint(p[1]), (unsigned)__double2loint(p[1])) &
            product3_ok((unsigned)__double2hiint(p[2]), (unsigned)__double2loint(p[2])) & product3_ok((unsigned)__double2hiint(p[3]), (unsigned)__double2loint(p[3]))) {
#pragma unroll
            for (int k = 0; k < 4; ++k) deposit_product3(col, stride, p[k], e[k]);
            return 0;
        }
#endif
        deposit4<false>(col, stride, p[0], p[1], p[2], p[3], status);
        deposit4<false>(col, stride, e[0], e[1], e[2], e[3], status);     // zeros take the (cheap) slow path
        return 0;
    }
    int cnt = 0;
    int lv = fpe_push4<F, EE, UNIFORM>(a, p, 0);
    unsigned any = (unsigned)nonzero_bits(p[0]) | (unsigned)nonzero_bits(p[1]) | (unsigned)nonzero_bits(p[2]) |
                   (unsigned)nonzero_bits(p[3]);
    if (any) cnt += deposit_residuals(col, stride, p, status);
    constexpr int first = EE ? (F > 1 ? 1 : 0) : (F > 3 ? F - 3 : 0);
    lv += fpe_push4<F, EE, UNIFORM>(a, e, first);
    if (walked) *walked += lv;
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

// ------------------------------------------------------------------------------------------------
// Touched-row tracking for the latency regime (the reference keeps imin / imax in its Superaccumulator for the same
// reason, superaccumulator.hpp:125, superaccumulator.cpp:138-162): a launch over a few thousand elements touches 2-15 of
// the 39 limb rows, and summing all 39 x T columns was the largest single cost of such a launch.  The kernels of
// this regime record the exponent range of what they deposit (two integer min / max per summand, off the hot
// streaming path) and the block merge only sums the rows that range can have touched.
// ------------------------------------------------------------------------------------------------
struct RowRange {
    unsigned emin, emax;     // biased exponents of the deposited values (emin > emax: nothing deposited)
};
EXB_D RowRange rr_empty() { return RowRange{0x7ffu, 0u}; }
EXB_D RowRange rr_full() { return RowRange{35u, 2010u}; }
// a value with high / low words (hi, lo) is about to be deposited (or fed to an expansion)
EXB_D void rr_note(RowRange& r, unsigned hi, unsigned lo) {
    const unsigned ahi = hi & 0x7fffffffu;
    if ((ahi | lo) == 0u) return;                              // exact zero: no deposit
    unsigned E = ahi >> 20;
    E = E < 35u ? 35u : (E > 2010u ? 2010u : E);               // tiny values land in limb 0; specials / too large are not deposited
    r.emin = min(r.emin, E);
    r.emax = max(r.emax, E);
}
EXB_D void rr_note(RowRange& r, double x) { rr_note(r, (unsigned)__double2hiint(x), (unsigned)__double2loint(x)); }
// a product (its TwoProd error term reaches 105 bits below its exponent; products below 2^-935 take the integer path
// that touches limbs 0..2)
EXB_D void rr_note_product(RowRange& r, double x, double y) {
    const double p = __dmul_rn(x, y);
    const unsigned ahi = (unsigned)__double2hiint(p) & 0x7fffffffu;
    const unsigned xz = ((unsigned)__double2hiint(x) & 0x7fffffffu) | (unsigned)__double2loint(x);
    const unsigned yz = ((unsigned)__double2hiint(y) & 0x7fffffffu) | (unsigned)__double2loint(y);
    if (xz == 0u || yz == 0u) return;
    unsigned E = ahi >> 20;
    E = E < 88u + 105u ? 35u + 105u : (E > 2010u ? 2010u : E);   // (E - 105 below is then >= 35: row 0 after the margin)
    r.emin = min(r.emin, E - 105u);
    r.emax = max(r.emax, E);
}
// limb rows [lo, hi] that deposits of values in the range can have touched; `grow` = binades by which partial sums of the
// deposited values (expansion levels) may exceed them
EXB_D void rr_rows(const RowRange& r, unsigned grow, unsigned& lo, unsigned& hi) {
    if (r.emin > r.emax) {
        lo = (unsigned)kLimbs - 1u;
        hi = 0u;
        return;
    }
    const unsigned jlo = __umulhi(r.emin + 17u, 82595525u);    // J1 of the smallest value: it touches rows J1 - 1, J1
    const unsigned jhi = __umulhi(min(r.emax + grow, 2010u) + 17u, 82595525u);
    lo = jlo > 2u ? jlo - 2u : 0u;                             // (one more row: residuals of an expansion reach 52 bits lower)
    hi = jhi < (unsigned)kLimbs - 1u ? jhi : (unsigned)kLimbs - 1u;
}

// ------------------------------------------------------------------------------------------------
// Warp-parallel limb arithmetic for the epilogue.  A 39-limb array is spread over one warp: lane l holds limb l
// in `.a` (l = 0..31) and limb 32 + l in `.b` (l = 0..6; `.b` is zero in the other lanes).  Normalisation and
// both roundings then cost a few dozen instructions and a handful of shuffles / votes instead of a 39-step
// serial carry chain and serial limb scans by one thread (those were ~1.5 us of every launch).
// ------------------------------------------------------------------------------------------------
struct WarpLimbs { long long a, b; };
constexpr unsigned kFullWarp = 0xffffffffu;

EXB_D unsigned long long wl_ballot(bool pa, bool pb) {          // bit j = predicate of limb j
    return (unsigned long long)__ballot_sync(kFullWarp, pa) | ((unsigned long long)(__ballot_sync(kFullWarp, pb) & 0x7fu) << 32);
}
EXB_D long long wl_get(const WarpLimbs& x, int k) {             // limb k (k warp-uniform, 0..38)
    const long long va = __shfl_sync(kFullWarp, x.a, k & 31);
    const long long vb = __shfl_sync(kFullWarp, x.b, k & 31);
    return k < 32 ? va : vb;
}
EXB_D WarpLimbs wl_shift_up(const WarpLimbs& c, unsigned lane) {    // out.limb[j] = c.limb[j-1], out.limb[0] = 0
    long long ua = __shfl_up_sync(kFullWarp, c.a, 1);
    long long ub = __shfl_up_sync(kFullWarp, c.b, 1);
    const long long a31 = __shfl_sync(kFullWarp, c.a, 31);
    if (lane == 0) {
        ua = 0;
        ub = a31;
    }
    WarpLimbs r;
    r.a = ua;
    r.b = ub;
    return r;
}
EXB_D unsigned long long low_bits(int n) { return n >= 64 ? ~0ull : (n <= 0 ? 0ull : ((1ull << n) - 1ull)); }   // bits 0..n-1

// Normal form (limbs 0..37 in [0, 2^52), limb 38 keeps the signed remainder: normalize(), superacc.cuh) of ANY
// int64 limbs.  Two local splits bring every carry down to one bit, then the ripple is resolved with the
// classic generate / propagate trick on vote masks: carries = ((g | p) + g) ^ p.  To keep every carry
// non-negative, limbs 1..37 first borrow one unit of the limb above (+2^52 here, -1 there: value unchanged).
// Returns true when the value is negative.
EXB_D bool warp_normalize(WarpLimbs& x, unsigned lane) {
    const bool has_b = lane < 7u, low_b = lane < 6u;             // .b: limbs 32..38; limb 38 (lane 6) is never reduced
    const long long R = 1ll << kDigits;
    WarpLimbs c, u;
    // split 1: carries in [-2^11, 2^11)
    c.a = x.a >> kDigits;
    x.a &= kLimbMask;
    c.b = low_b ? (x.b >> kDigits) : 0;
    if (low_b) x.b &= kLimbMask;
    u = wl_shift_up(c, lane);
    x.a += u.a;
    if (has_b) x.b += u.b;
    // pre-borrow: limbs 1..37 += 2^52, limbs 2..38 -= 1
    if (lane >= 1u) x.a += R;
    if (lane >= 2u) x.a -= 1;
    if (low_b) x.b += R;
    if (has_b) x.b -= 1;
    // split 2: limbs 0..37 are non-negative and below 2^54 now; carries in {0..3}
    c.a = x.a >> kDigits;
    x.a &= kLimbMask;
    c.b = low_b ? (x.b >> kDigits) : 0;
    if (low_b) x.b &= kLimbMask;
    u = wl_shift_up(c, lane);
    x.a += u.a;
    if (has_b) x.b += u.b;
    // limbs 0..37 in [0, 2^52 + 3]: one-bit carries, resolved by an integer add of the vote masks
    const unsigned long long g = wl_ballot(x.a >= R, low_b && x.b >= R);
    const unsigned long long pm = wl_ballot(x.a == R - 1, low_b && x.b == R - 1);
    const unsigned long long carry = ((g | pm) + g) ^ pm;          // bit j = carry INTO limb j
    x.a = (x.a + (long long)((carry >> lane) & 1ull)) & kLimbMask;
    if (has_b) {
        x.b += (long long)((carry >> (32u + lane)) & 1ull);
        if (low_b) x.b &= kLimbMask;
    }
    return __shfl_sync(kFullWarp, x.b, 6) < 0;
}

// Superaccumulator::Round() (superaccumulator.cpp:80-134) on normalised limbs held by the warp: the scans of
// round_ref_compat() as votes, then the shared arithmetic (round_ref_parts).  Every lane returns the value.
EXB_D double warp_round_ref(const WarpLimbs& x, bool negative) {
    const unsigned long long nz = wl_ballot(x.a != 0, x.b != 0);
    int i = nz ? 63 - __clzll((long long)nz) : -1;                                       // :91-94
    if (negative && i >= 0) {                                                            // :95-101
        const unsigned long long ones = wl_ballot((x.a & kLimbMask) == kLimbMask, (x.b & kLimbMask) == kLimbMask);
        const unsigned long long cand = ~ones & low_bits(i + 1);
        i = cand ? 63 - __clzll((long long)cand) : -1;
    }
    const long long acc_i = wl_get(x, i < 0 ? 0 : i);
    const long long acc_im1 = wl_get(x, i < 1 ? 0 : i - 1);
    // :116-119 -- limbs 0 .. i-2; for a negative value every term (2^52 - acc[j]) is non-zero
    const bool sticky = i >= 2 && (negative || (nz & low_bits(i - 1)) != 0ull);
    return round_ref_parts(i, acc_i, acc_im1, sticky, negative);
}

// Correct rounding (round_exact(), superacc.cuh) on normalised limbs held by the warp.
__device__ __noinline__ double warp_round_exact(const WarpLimbs& x, bool negative, unsigned lane) {
    unsigned long long ma = (unsigned long long)x.a, mb = (unsigned long long)x.b;
    if (negative) {      // magnitude = two's complement negation, digit by digit: the +1 ripples through the zero limbs
        const unsigned long long nz = wl_ballot(x.a != 0, x.b != 0);
        const int tz = __ffsll((long long)nz) - 1;               // lowest non-zero limb (the value is not zero)
        const long long R = 1ll << kDigits;
        const int ja = (int)lane, jb = 32 + (int)lane;
        ma = ja < tz ? 0ull : (unsigned long long)((ja == tz ? R : R - 1) - x.a);
        if (lane < 6u) mb = jb < tz ? 0ull : (unsigned long long)((jb == tz ? R : R - 1) - x.b);
        else if (lane == 6u) mb = (unsigned long long)((tz == 38 ? 0ll : -1ll) - x.b);
        else mb = 0ull;
    }
    const unsigned long long mz = wl_ballot(ma != 0ull, mb != 0ull);
    const int top = mz ? 63 - __clzll((long long)mz) : -1;
    WarpLimbs m;
    m.a = (long long)ma;
    m.b = (long long)mb;
    const unsigned long long m_top = (unsigned long long)wl_get(m, top < 0 ? 0 : top);
    const unsigned long long m1 = top >= 1 ? (unsigned long long)wl_get(m, top - 1) : 0ull;
    const unsigned long long m2 = top >= 2 ? (unsigned long long)wl_get(m, top - 2) : 0ull;
    const bool sticky = top >= 3 && (mz & low_bits(top - 2)) != 0ull;       // limbs 0 .. top-3
    return round_exact_parts(top, m_top, m1, m2, sticky, negative);
}

// value from limbs + status flags (finalize_value(), superacc.cuh); x must be normalised
EXB_D double warp_value(const WarpLimbs& x, bool negative, unsigned st, int round_mode, unsigned lane) {
    if ((st & kStNaN) || ((st & kStPosInf) && (st & kStNegInf)) || (st & kStPeerTimeout)) return __longlong_as_double(0x7ff8000000000000ll);
    if (st & kStPosInf) return __longlong_as_double(0x7ff0000000000000ll);
    if (st & kStNegInf) return __longlong_as_double(0xfff0000000000000ll);
    return round_mode ? warp_round_exact(x, negative, lane) : warp_round_ref(x, negative);
}

// Sum one limb row over all T columns WITHOUT requiring the columns to be bounded first: the low 52 bits and the
// carry-save bits of every limb are summed separately (the carry part belongs to the row above), so any int64
// limb is fine and the separate bound_column pass before the merge is gone.  The cross-lane step is four
// independent 32-bit REDUX (three 19-bit chunks of the low part + the carry part) instead of a 15-deep chain of
// dependent shuffles, which used to dominate the epilogue (2-3 us of every launch).  Results valid in all lanes.
EXB_D void row_sum_split(unsigned row_addr, unsigned T, unsigned lane, long long& lo, int& hi) {
    unsigned long long slo = 0;
    int shi = 0;
#pragma unroll 4
    for (unsigned t = lane; t < T; t += 32) {
        const long long v = (long long)lds64(row_addr + 8u * t);
        slo += (unsigned long long)(v & kLimbMask);
        shi += (int)(v >> kDigits);
    }
    // slo < (T / 32) * 2^52 <= 2^57: chunks of 19 bits, each lane sum < 2^24
    const unsigned r0 = __reduce_add_sync(kFullWarp, (unsigned)slo & 0x7ffffu);
    const unsigned r1 = __reduce_add_sync(kFullWarp, (unsigned)(slo >> 19) & 0x7ffffu);
    const unsigned r2 = __reduce_add_sync(kFullWarp, (unsigned)(slo >> 38));
    hi = __reduce_add_sync(kFullWarp, shi);                    // |.| <= 1024 * 2^11
    lo = (long long)((unsigned long long)r0 + ((unsigned long long)r1 << 19) + ((unsigned long long)r2 << 38));   // < 1024 * 2^52
}

// ticket counter with acquire + release semantics at GPU scope: orders this warp's REDs (made visible to lane 0 by
// __syncwarp) before the ticket, and the last CTA's reads after it -- one instruction instead of two full fences
EXB_D unsigned ticket_acq_rel(unsigned* counter) {
    unsigned old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
    return old;
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
        win_after_slow_group<4>(w, mine, hi, false, [&](double v) { deposit_sum(col, stride, v, status); });
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
        if (out[k] != 0.0) deposit_sum(col, stride, out[k], status);
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
    win_after_slow_group<4>(w, mine, hi, true, [&](double v) { deposit_sum(col, stride, v, status); });
    w.st = status;
    return w;
}

__device__ __noinline__ Window win_flush_singles(Window w, unsigned col, unsigned stride) {
    double out[4];
    win_drain_single(w, out);
    unsigned status = w.st;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (out[k] != 0.0) deposit_sum(col, stride, out[k], status);
    w.st = status;
    return w;
}

// Thrash control for the expansion (performance only; the sum is exact either way).  When most
// elements fall off the last level -- data whose dynamic range exceeds what F doubles can hold,
// e.g. log-uniform over 2^+-332 -- the TwoSum walk is pure overhead, so the warp bypasses the
// expansion and deposits directly for `kBypassTiles` tiles, then probes the expansion again.
constexpr int kBypassTiles = 32;
constexpr int kBypassMax = 4096;

// The fused limb exchange (see Mailbox): send this rank's normalised limbs + status to every rank's mailbox, collect
// every rank's words of this epoch from MY mailbox, sum, normalise.  Out of line: single-GPU launches never fetch it.
__device__ __noinline__ bool peer_exchange(const ReduceParams& prm, WarpLimbs& x, unsigned& final_status, const unsigned ln) {
    const unsigned set = (unsigned)(prm.epoch & 1ull);
    const unsigned long long tag = peer_tag(prm.epoch) << kDigits;
    // my words: lane l owns word l (limb l) and, for l < 9, word 32 + l (limbs 32..37, limb 38 low / high, status)
    const long long top = __shfl_sync(kFullWarp, x.b, 6);
    unsigned long long w0 = (unsigned long long)x.a, w1 = 0ull;
    if (ln < 6u) w1 = (unsigned long long)x.b;
    else if (ln == 6u) w1 = (unsigned long long)top & (unsigned long long)kLimbMask;
    else if (ln == 7u) w1 = (unsigned long long)(top >> kDigits) & (unsigned long long)kLimbMask;
    else if (ln == 8u) w1 = (unsigned long long)final_status;
    for (int r = 0; r < prm.nranks; ++r) {
        unsigned long long* dst = prm.peers[r]->slot[set][prm.rank].data;
        st_relaxed_sys(dst + ln, tag | w0);
        if (ln < 9u) st_relaxed_sys(dst + 32 + ln, tag | w1);
    }
    // collect every rank's words of this epoch from MY mailbox
    const unsigned long long t0 = globaltimer_ns();
    bool timed_out = false;
    unsigned long long s0 = 0ull, s1 = 0ull;          // sums over ranks of word ln / word 32 + ln
    long long s1_signed = 0;
    for (int r = 0; r < prm.nranks; ++r) {
        const unsigned long long* src = prm.peers[prm.rank]->slot[set][r].data;
        unsigned long long v0, v1 = tag;
        unsigned spins = 0;
        for (;;) {
            v0 = ld_relaxed_sys(src + ln);
            if (ln < 9u) v1 = ld_relaxed_sys(src + 32 + ln);
            if ((v0 >> kDigits) == (tag >> kDigits) && (v1 >> kDigits) == (tag >> kDigits)) break;
            if (((++spins) & 63u) == 0u && prm.peer_timeout_ns != 0ull && globaltimer_ns() - t0 > prm.peer_timeout_ns) {
                timed_out = true;
                break;
            }
        }
        if (timed_out) break;
        s0 += v0 & (unsigned long long)kLimbMask;
        const unsigned long long p1 = v1 & (unsigned long long)kLimbMask;
        if (ln == 7u) s1_signed += ((long long)(p1 << 12)) >> 12;      // sign-extend the 52-bit field
        else if (ln == 8u) s1 |= p1;                                    // status flags: OR
        else s1 += p1;
    }
    timed_out = __any_sync(kFullWarp, timed_out);
    const long long top_lo = (long long)__shfl_sync(kFullWarp, s1, 6);
    const long long top_hi = __shfl_sync(kFullWarp, s1_signed, 7);
    unsigned stm = (unsigned)__shfl_sync(kFullWarp, s1, 8);
    if (timed_out) stm |= kStPeerTimeout;
    x.a = (long long)s0;                               // <= 8 normalised limbs: no overflow
    x.b = ln < 6u ? (long long)s1 : (ln == 6u ? top_lo + (top_hi << kDigits) : 0ll);
    final_status = stm;
    return warp_normalize(x, ln);
}

// Steps 4 and 5 of the epilogue (see reduce_finish), shared by every reduction kernel of this file:
//   4. block merge: split row sums -> one local carry split -> native 64-bit REDs into the global accumulator ->
//      ticket (skipped by a `solo` CTA, which publishes straight from shared memory);
//   5. the last CTA's closing WARP: gather, warp-parallel normalise, (optional) peer exchange, round, publish.
// Not inlined: one copy of this (cold, run-once) code serves all kernels, which keeps them small.
__device__ __noinline__ void block_merge_and_close(const ReduceParams& prm, const unsigned stride, const unsigned smem_base,
                                                   const unsigned T, const unsigned tid, unsigned status, const bool solo,
                                                   long long* block_lo, int* block_hi, unsigned* block_status,
                                                   unsigned row_lo, unsigned row_hi) {
    // block_status[0] = status flags, [1] / [2] = first / last limb row any thread of the CTA has touched
    const unsigned warp = tid >> 5, ln = tid & 31u, nwarps = T >> 5;
    if (tid == 0) {
        block_status[0] = 0u;
        block_status[1] = (unsigned)kLimbs - 1u;
        block_status[2] = 0u;
    }
    __syncthreads();
    row_lo = __reduce_min_sync(kFullWarp, row_lo);
    row_hi = __reduce_max_sync(kFullWarp, row_hi);
    if (ln == 0) {
        atomicMin(&block_status[1], row_lo);
        atomicMax(&block_status[2], row_hi);
    }
    for (unsigned j = tid; j < (unsigned)kLimbs; j += T) {       // (T may be as small as 32)
        block_lo[j] = 0;
        block_hi[j] = 0;
    }
    __syncthreads();                                       // row range and block sums initialised; all columns final
    if (status) {
        if (solo) atomicOr(block_status, status);
        else atomicOr(&prm.ws->status, status);
    }
    row_lo = block_status[1];
    row_hi = block_status[2];
#pragma unroll 1
    for (unsigned j = row_lo + warp; j <= row_hi; j += nwarps) {      // (rolled on purpose: this code runs once, from a cold instruction cache)
        long long lo;
        int hi;
        row_sum_split(smem_base + j * stride, T, ln, lo, hi);
        if (ln == 0) {
            block_lo[j] = lo;
            block_hi[j] = hi;
        }
    }
    __syncthreads();
    EXB_PHASE(5);
    if (warp != 0) return;                                 // the rest is one warp's work

    // limb j of this CTA's sum = lo[j] + hi[j-1]  (|.| < 2^62), spread over the warp
    WarpLimbs x;
    x.a = block_lo[ln] + (ln > 0 ? (long long)block_hi[ln - 1] : 0ll);
    x.b = ln < 7u ? block_lo[32 + ln] + (long long)block_hi[31 + ln] : 0ll;
    if (ln == 6u) x.b += ((long long)block_hi[38]) << kDigits;     // carry-save bits of the top limb stay in the top limb
    unsigned final_status = 0;
    bool last = solo;
    if (!solo) {
        // one local split keeps every contribution below 2^52 + 2^10 in magnitude, so that <= 2048 CTAs (and a
        // pending normalised partial sum) cannot overflow a 64-bit global limb; no carry chain needed here
        WarpLimbs c, u;
        c.a = x.a >> kDigits;
        x.a &= kLimbMask;
        c.b = ln < 6u ? (x.b >> kDigits) : 0;
        if (ln < 6u) x.b &= kLimbMask;
        u = wl_shift_up(c, ln);
        x.a += u.a;
        if (ln < 7u) x.b += u.b;
        if (x.a != 0) atomicAdd(&prm.ws->gacc[ln], (unsigned long long)x.a);
        if (ln < 7u && x.b != 0) atomicAdd(&prm.ws->gacc[32 + ln], (unsigned long long)x.b);
        __syncwarp();
        unsigned ticket = 0;
        if (ln == 0) ticket = ticket_acq_rel(&prm.ws->counter);
        ticket = __shfl_sync(kFullWarp, ticket, 0);
        last = (ticket == gridDim.x - 1);
    }
    EXB_PHASE(6);
    if (!last) return;

    // ---------------- 5. closing warp: normalise the global accumulator, exchange, round, publish ----
    if (!solo) {
        x.a = (long long)atomicExch(&prm.ws->gacc[ln], 0ull);
        x.b = ln < 7u ? (long long)atomicExch(&prm.ws->gacc[32 + ln], 0ull) : 0ll;
        if (ln == 0) {
            final_status = prm.finalize && !prm.keep ? atomicExch(&prm.ws->status, 0u) : atomicOr(&prm.ws->status, 0u);
            prm.ws->counter = 0;
        }
    } else {
        if (ln == 0) final_status = *block_status;
    }
    final_status = __shfl_sync(kFullWarp, final_status, 0);
    bool neg = warp_normalize(x, ln);
    EXB_PHASE(7);
    // ---- fused multi-GPU exchange over peer memory (only the closing launch of a reduction) ----
    if (prm.finalize && prm.nranks > 1) neg = peer_exchange(prm, x, final_status, ln);
    EXB_PHASE(8);
    if (prm.finalize) {
        Result* out = prm.out;
        const double v = warp_value(x, neg, final_status, prm.round_mode, ln);
        out->limbs[ln] = x.a;
        if (ln < 7u) out->limbs[32 + ln] = x.b;
        if (ln < (unsigned)kFlagSlots) out->flagcnt[ln] = (final_status >> ln) & 1u;
        if (ln == 0) {
            out->value = v;
            out->status = final_status;
        }
    }
    // leave the (normalised) partial sum in the workspace unless this call closes the reduction
    if (!solo && (!prm.finalize || prm.keep)) {
        prm.ws->gacc[ln] = (unsigned long long)x.a;
        if (ln < 7u) prm.ws->gacc[32 + ln] = (unsigned long long)x.b;
    }
    EXB_PHASE(9);
}

// Everything after the unrolled vector body, shared by the streaming kernels:
//   1. the REMAINDER of the vector region (whatever the equal-sized tiles of the body left over -- for vectors
//      below ~2^21 elements that is everything): guarded 256-bit loads, four in flight per thread, spread evenly
//      over all threads of the grid;
//   2. the scalar part: alignment head, the last n mod 4 elements, or the whole vector when it is strided /
//      misaligned;
//   3. the flush of the expansions;
//   4. the block merge: split row sums (no bound_column pass) -> one local carry split -> native 64-bit REDs into
//      the global accumulator -> ticket;
//   5. the last CTA's closing WARP: gather, warp-parallel normalise, (optional) peer exchange, round, publish.
// `body_vecs` = vectors the body has consumed; `since_norm` = deposits into this thread's column since it was
// last bounded; `bypass_hint` = the body ended in thrash-bypass mode (its data overflows the expansion).
template <int F, bool EE, bool DOT>
EXB_D void reduce_finish(const ReduceParams& prm, const unsigned col, const unsigned stride, const unsigned smem_base,
                         const unsigned T, const unsigned tid, const long long body_vecs,
                         double (&a)[F > 0 ? F : 1][expansions(F)], unsigned status, int since_norm, bool bypass_hint) {
    constexpr int kM = expansions(F);
    constexpr int kDepPerElem = DOT ? 2 : 1;               // at most one deposit per summand
    constexpr int kSlack = 2 * kM * (F + 2) + 16;          // room for the flush of the expansions
    __shared__ long long block_lo[kLimbs];
    __shared__ int block_hi[kLimbs];
    __shared__ unsigned block_status[3];
    // Latency regime: a single CTA reducing into an empty workspace publishes straight from shared
    // memory -- no global atomics, fence or ticket.
    const bool solo = (gridDim.x == 1) && prm.fresh && prm.finalize;
    // Without an unrolled body (vectors below ~2^20 elements) everything this CTA deposits passes through the loops
    // below, which then record its exponent range: the merge only sums the limb rows that range can have touched.
    // (ExSUM only: in the ExDOT instantiations the extra live state spilled registers of the streaming loops, -8 %.)
    const bool track = !DOT && prm.iters == 0;
    RowRange rr = track ? rr_empty() : rr_full();
    EXB_PHASE(2);

    // ---------------- 1. remainder of the vector region ----------------
    {
        const long long nrem = prm.nvec - body_vecs;
        const long long gthreads = (long long)gridDim.x * T;
        long long r = (long long)blockIdx.x * T + tid;
        if (r < nrem) {
            constexpr int UR = DOT ? 2 : 4;               // vectors in flight per thread and stream (same bytes for both)
            const double* pa = prm.a + prm.head + 4 * (body_vecs + r);
            const double* pb = DOT ? prm.b + prm.head + 4 * (body_vecs + r) : nullptr;
            const long long vstep = 4 * gthreads;
            Vec4 va[UR];
            Vec4 vb[DOT ? UR : 1];
            bool direct = (F > 0) && prm.adaptive && bypass_hint;
            int fell = 0;
#pragma unroll
            for (int u = 0; u < UR; ++u)
                if (r + u * gthreads < nrem) {
                    va[u] = ldg256(pa + u * vstep);
                    if (DOT) vb[DOT ? u : 0] = ldg256(pb + u * vstep);
                }
            for (; r < nrem; r += UR * gthreads) {
                pa += UR * vstep;
                if (DOT) pb += UR * vstep;
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    if (r + u * gthreads < nrem) {
                        double x4[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
                        if (!DOT && track) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) rr_note(rr, x4[q]);
                        }
                        if (DOT) {
                            const double y4[4] = {vb[DOT ? u : 0].x, vb[DOT ? u : 0].y, vb[DOT ? u : 0].z, vb[DOT ? u : 0].w};
                            if (F == 0 || direct) {
                                double none[1][expansions(0)];
                                mul_add4<0, false, false>(col, stride, none, status, x4, y4);
                            } else {
                                fell += mul_add4<F, EE, false>(col, stride, a, status, x4, y4);
                            }
                        } else if (F == 0 || direct) {
                            deposit4<false>(col, stride, x4[0], x4[1], x4[2], x4[3], status);
                        } else {
                            fell += add4<F, EE, false>(col, stride, a, status, x4);
                        }
                        if (r + (u + UR) * gthreads < nrem) {
                            va[u] = ldg256(pa + u * vstep);
                            if (DOT) vb[DOT ? u : 0] = ldg256(pb + u * vstep);
                        }
                    }
                }
                // thrash control as in the body, but per thread (lanes are not converged here): once a quarter of this
                // thread's summands fall off the last level, it deposits the rest of its (short) remainder directly
                if (F > 0 && prm.adaptive && 4 * fell >= 4 * UR * kDepPerElem) direct = true;
                fell = 0;
                since_norm += 4 * UR * kDepPerElem;
                if (since_norm > kMaxDepositsPerNormalize - 4 * UR * kDepPerElem - kSlack) {
                    bound_column(col, stride);
                    since_norm = 0;
                }
            }
        }
    }
    EXB_PHASE(3);

    // ---------------- 2. scalar part: alignment head, tail, or the whole strided vector ----------
    {
        const long long body = 4 * prm.nvec;               // elements of the vector region
        const long long nscalar = prm.n - body;            // head + tail (or everything)
        const long long gthreads = (long long)gridDim.x * T;
        long long k = (long long)blockIdx.x * T + tid;
        // Strided vectors (inca / incb != 1: no vector region at all): eight independent loads per thread in
        // flight, consumed four at a time through the same four-summand code as the vector body (per-thread early
        // exit: lanes need not be converged here).  A warp load of 32 consecutive elements touches 32 * inc * 8
        // bytes, so the useful fraction of every sector is 1 / inc whatever is done here; what this loop buys is
        // memory-level parallelism.
        for (; k + 7 * gthreads < nscalar; k += 8 * gthreads) {
            double xa[8], xb[DOT ? 8 : 1];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long kk = k + j * gthreads;
                const long long idx = kk < prm.head ? kk : kk + body;
                xa[j] = ldg64(prm.a + idx * prm.inca);
                if (DOT) xb[DOT ? j : 0] = ldg64(prm.b + idx * prm.incb);
            }
            if (!DOT && track) {
#pragma unroll
                for (int j = 0; j < 8; ++j) rr_note(rr, xa[j]);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double x4[4] = {xa[4 * h], xa[4 * h + 1], xa[4 * h + 2], xa[4 * h + 3]};
                if (DOT) {
                    const double y4[4] = {xb[DOT ? 4 * h : 0], xb[DOT ? 4 * h + 1 : 0], xb[DOT ? 4 * h + 2 : 0], xb[DOT ? 4 * h + 3 : 0]};
                    mul_add4<F, EE, false>(col, stride, a, status, x4, y4);
                } else if (F == 0) {
                    deposit4<false>(col, stride, x4[0], x4[1], x4[2], x4[3], status);
                } else {
                    add4<F, EE, false>(col, stride, a, status, x4);
                }
            }
            since_norm += 8 * kDepPerElem;
            if (since_norm > kMaxDepositsPerNormalize - 8 * kDepPerElem - kSlack) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
        for (; k < nscalar; k += gthreads) {
            const long long idx = k < prm.head ? k : k + body;
            if (!DOT && track) rr_note(rr, prm.a[idx * prm.inca]);
            if (DOT) mul_add1<F, EE>(col, stride, a, status, prm.a[idx * prm.inca], prm.b[idx * prm.incb]);
            else add1<F, EE>(col, stride, a, status, prm.a[idx * prm.inca]);
            since_norm += kDepPerElem;
            if (since_norm > kMaxDepositsPerNormalize - kDepPerElem - kSlack) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
    }
    // ---------------- 3. flush the expansions (exact sums of in-range inputs: deposit_sum) ----------
    if (F > 0) {
#pragma unroll
        for (int i = 0; i < F; ++i)
#pragma unroll
            for (int m = 0; m < kM; ++m) deposit_sum(col, stride, a[i][m], status);
    }
    EXB_PHASE(4);
    unsigned row_lo, row_hi;
    rr_rows(rr, F > 0 ? 64u : 0u, row_lo, row_hi);        // partial sums in the expansions may grow beyond the largest summand
    block_merge_and_close(prm, stride, smem_base, T, tid, status, solo, block_lo, block_hi, block_status, row_lo, row_hi);
}

// (defined further down, with the superaccumulator-only kernel)
template <int DW, int W, bool EARLY>
__device__ __noinline__ unsigned reduce0_window_rows_wide(const double* pa, const double* pb, const long long row_step,
                                                          const unsigned iters, const unsigned col, const unsigned stride, unsigned* status_io, int* range_io);

// Hand-off of the ExDOT expansion kernels [r2]: products that thrash an expansion but fit the 5-digit register window
// (ill-conditioned dot products: ~125 binades) stream faster through that window than through direct deposits (6.1 against
// 5.2-5.5 TB/s), and fpe never changes the result.  A warp whose first tile thrashed hands the REST of its rows (the ExDOT
// tiles are rows of the superaccumulator-only kernel's geometry, so a thread's vectors form one arithmetic progression) to
// that kernel's window loop; rows the loop leaves (the window no longer fits) are deposited directly.  Out of line: the
// streaming loops of the kernel must not pay for this code with registers.  Returns status flags; the column is bounded.
__device__ __noinline__ unsigned dot_handoff(const double* qa, const double* qb, const long long row_step, const unsigned rows,
                                             const unsigned col, const unsigned stride) {
    unsigned status = 0;
    bound_column(col, stride);
    int range[2] = {4096, -4096};
    unsigned k = reduce0_window_rows_wide<3, 5, false>(qa, qb, row_step, rows, col, stride, &status, range);
    int since_norm = 0;
    for (; k < rows; ++k) {                                                      // (rare: the window gave up on the way)
        const Vec4 xa = ldg256(qa + (long long)k * row_step), xb = ldg256(qb + (long long)k * row_step);
        const double x[4] = {xa.x, xa.y, xa.z, xa.w}, y[4] = {xb.x, xb.y, xb.z, xb.w};
        double none[1][expansions(0)];
        mul_add4<0, false, false>(col, stride, none, status, x, y);
        since_norm += 8;
        if (since_norm > kMaxDepositsPerNormalize - 64) {
            bound_column(col, stride);
            since_norm = 0;
        }
    }
    bound_column(col, stride);
    return status;
}

// True when the exponents of the tile a warp holds in registers span more than an F-level expansion can hold (53 F bits
// plus slack): such a tile is not worth walking.  Warp-uniform.
template <int F, bool DOT, int U>
EXB_D bool tile_too_wide(const Vec4 (&va)[U], const Vec4 (&vb)[DOT ? U : 1]) {
    unsigned emax = 0u, emin = 0xfffu;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const double xs[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
        const double ys[4] = {vb[DOT ? u : 0].x, vb[DOT ? u : 0].y, vb[DOT ? u : 0].z, vb[DOT ? u : 0].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned e = ((unsigned)__double2hiint(xs[k]) >> 20) & 0x7ffu;
            if (DOT) {
                const unsigned eb = ((unsigned)__double2hiint(ys[k]) >> 20) & 0x7ffu;
                e = (e && eb) ? e + eb : 0u;                 // exponent of the product (+ 1023), 0 for zeros
            }
            emax = max(emax, e);
            emin = min(emin, e ? e : 0xfffu);
        }
    }
    emax = __reduce_max_sync(kFullWarp, emax);
    emin = __reduce_min_sync(kFullWarp, emin);
    return emax > emin && emax - emin > 53u * F + 64u;
}

// (Measured and rejected, round 1: moving the thrash bypass or the expansion walk of THIS kernel into out-of-line
// functions, as exblas_reduce0_kernel does with its window loop.  The expansion walk contains calls (deposits of
// residuals), and a function that is itself called and calls on spills its prefetch slots: 1.9 instead of
// 6.0 TB/s on narrow data; a call to an out-of-line bypass from inside this loop cost the walk 40 %.)
template <int F, bool EE, bool DOT, int U, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) exblas_reduce_kernel(const __grid_constant__ ReduceParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;
    const unsigned tid = threadIdx.x;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned stride = 8u * T;
    const unsigned col = smem_base + 8u * tid;

    EXB_PHASE(0);
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
    int since_norm = 0;
    bool bypass_hint = false;
    bool handed_off = false;

    // ---------------- vector body: prm.iters full tiles per CTA, 256-bit loads, rolling prefetch ----------------
    // (every CTA streams the same number of tiles: tiles blockIdx.x, blockIdx.x + grid, ...; what that leaves over
    // is spread evenly over all threads by reduce_finish)
    if (prm.iters > 0) {
        const unsigned iters = (unsigned)prm.iters;
        // ExSUM: a tile is T * U consecutive vectors.  ExDOT: a tile is U ROWS of T vectors, grid * T vectors apart (the
        // geometry of the superaccumulator-only kernel), so that the vectors a thread owns form ONE arithmetic progression
        // and a warp can hand its rest to that kernel's window loop on its own (see the hand-off below).
        const long long vstep = DOT ? (long long)gridDim.x * T * 4 : (long long)T * 4;      // elements between my vectors
        const long long tile_step = DOT ? (long long)U * vstep : (long long)gridDim.x * TILE;   // elements between my tiles
        const long long first = DOT ? (long long)blockIdx.x * T * 4 : (long long)blockIdx.x * TILE;
        const double* pa = prm.a + prm.head + first + (long long)tid * 4;
        const double* pb = DOT ? prm.b + prm.head + first + (long long)tid * 4 : nullptr;
        Vec4 va[U];
        Vec4 vb[DOT ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            va[u] = ldg256(pa + u * vstep);
            if (DOT) vb[u] = ldg256(pb + u * vstep);
        }
        // clear the (thread-private: no barrier needed) column while the first loads are in flight
        for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
        // TMA-engine L2 prefetch of whole tiles (one UBLKPF per tile and stream, thread 0): D tiles ahead of the loads
        const int pfd = DOT ? 0 : prm.l2_prefetch;                              // (contiguous tiles only)
        const double* pfa = prm.a + prm.head + (long long)blockIdx.x * TILE;
        const double* pfb = DOT ? prm.b + prm.head + (long long)blockIdx.x * TILE : nullptr;
        if (pfd > 0 && tid == 0) {
            for (int d = 1; d <= pfd && (unsigned)d < iters; ++d) {
                bulk_prefetch_l2(pfa + (long long)d * tile_step, (unsigned)(TILE * 8));
                if (DOT) bulk_prefetch_l2(pfb + (long long)d * tile_step, (unsigned)(TILE * 8));
            }
        }
        int bypass = 0, backoff = kBypassTiles;
        // The first probe of the expansion is skipped when the exponents of the first tile alone span more than an
        // F-level expansion can hold (log-uniform 2^+-332 against 53 F bits): walking a thrashing tile costs ~10 direct
        // tiles, which vectors of 2^22..2^26 elements never amortise.  Later probes (after a bypass period) are real.
#ifdef EXB_NO_HANDOFF
        bool handoff_ok = false;
#else
        bool handoff_ok = DOT && F > 0 && prm.adaptive && prm.window > 1 && prm.handoff_tiles > 0 && iters >= (unsigned)prm.handoff_tiles;
#endif
        if (F > 0 && prm.adaptive && tile_too_wide<F, DOT, U>(va, vb)) {
            bypass = backoff;
            backoff = min(backoff * 16, kBypassMax);
            handoff_ok = false;                                // wider than any window: the direct tiles below are the right path
        }
        for (unsigned it = 0; it < iters; ++it) {
            pa += tile_step;
            if (DOT) pb += tile_step;
            const bool has_next = it + 1 < iters;
            if (pfd > 0 && tid == 0 && it + 1u + (unsigned)pfd < iters) {
                bulk_prefetch_l2(pfa + (long long)(it + 1u + (unsigned)pfd) * tile_step, (unsigned)(TILE * 8));
                if (DOT) bulk_prefetch_l2(pfb + (long long)(it + 1u + (unsigned)pfd) * tile_step, (unsigned)(TILE * 8));
            }
#ifndef EXB_NO_REPROBE_CHECK
            // A probe that follows a bypass period looks at the exponents first, like the first one: on data that stays
            // wide (the log-uniform benchmark vector) no tile is ever walked.
            if (F > 0 && prm.adaptive && bypass == 0 && backoff > kBypassTiles && tile_too_wide<F, DOT, U>(va, vb)) {
                bypass = backoff;
                backoff = min(backoff * 16, kBypassMax);
            }
#endif
            const bool direct = (F == 0) || (prm.adaptive && bypass > 0);
            int deposits = 0;
            int walked = 0;                                 // expansion levels visited in this tile (warp-uniform)
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
                        deposits += mul_add4<F, EE, true>(col, stride, a, status, x, y, &walked);
                    } else {
                        deposits += add4<F, EE, true>(col, stride, a, status, x, &walked);
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
                    // Early-exit kernels [r2]: a walk that keeps going deep is slower than depositing directly although
                    // nothing falls off the end -- every level is a vote and six dependent FP64 adds for four summands
                    // (ExDOT fpe 8 ee on ill-conditioned products: ~8 levels per product, 3.3 TB/s, FP64 pipe 56 % busy
                    // on latency; direct three-limb deposits run the same data at 5.3+).  More than 4 levels per ExSUM
                    // vector / 6 per ExDOT vector (p and e together) on average over a tile counts as thrashing too.
                    constexpr int kDeepWalk = (DOT ? 6 : 4) * U;
                    if (total * 64 >= 32 * kDepPerTile || (EE && walked > kDeepWalk)) {
                        bypass = backoff;
                        backoff = min(backoff * 16, kBypassMax);   // a second thrashing probe in a row: stay away for long
                    } else {
                        backoff = kBypassTiles;
                    }
                }
            }
            since_norm += kDepPerTile;
            if (since_norm > kMaxDepositsPerNormalize - kDepPerTile - 2 * kM * (F + 2) - 16) {
                bound_column(col, stride);
                since_norm = 0;
            }
            if (DOT && F > 0 && it == 0u && handoff_ok && bypass > 0) {   // (never with EXB_NO_HANDOFF)
                handed_off = true;                             // the first tile thrashed the expansion (warp-uniform)
                break;
            }
        }
        bypass_hint = bypass > 0;
#ifndef EXB_NO_HANDOFF
        if constexpr (DOT && F > 0) {
            if (handed_off) {
                status |= dot_handoff(prm.a + prm.head + (long long)blockIdx.x * T * 4 + (long long)tid * 4 + (long long)U * vstep,
                                      prm.b + prm.head + (long long)blockIdx.x * T * 4 + (long long)tid * 4 + (long long)U * vstep, vstep,
                                      (iters - 1u) * (unsigned)U, col, stride);
                since_norm = 0;
            }
        }
#endif
    } else {
        for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
    }
    EXB_PHASE(1);
    reduce_finish<F, EE, DOT>(prm, col, stride, smem_base, T, tid, prm.iters * (long long)gridDim.x * T * U, a, status, since_norm, bypass_hint);
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
    winp_cover<W>(w, gmin, gmax, [&](double v) { deposit_sum(col, stride, v, status); });
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
        if (out[k] != 0.0) deposit_sum(col, stride, out[k], status);
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
    wins_cover<W>(w, gmin, gmax, [&](double v) { deposit_sum(col, stride, v, status); });
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
        if (out[k] != 0.0) deposit_sum(col, stride, out[k], status);
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
    bool abandoned = false;
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
                if (w.span == 0u) {          // what has missed cannot fit this window (warp-uniform): leave NOW, not a block later
                    k += (unsigned)u + 1u;   // rows consumed so far, this one included
                    abandoned = true;
                    break;
                }
            }
            if (loaded < iters) load_row(u);                    // row k + DW + u (after the slot has been consumed)
        }
        if (abandoned) break;
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
__global__ void __launch_bounds__(MAXT, 1) exblas_reduce0_kernel(const __grid_constant__ ReduceParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;
    const unsigned tid = threadIdx.x;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned stride = 8u * T;
    const unsigned col = smem_base + 8u * tid;
    EXB_PHASE(0);

    unsigned status = 0;
    constexpr int kDepPerElem = DOT ? 2 : 1;
    const long long ROW = (long long)T * 4;
    bool zeroed = false;                                       // (the column is cleared behind the first loads where possible)

    if (prm.iters > 0) {
        const unsigned iters = (unsigned)prm.iters;            // my rows: blockIdx.x, blockIdx.x + grid, ... (same count in every CTA)
        const long long row_step = (long long)gridDim.x * ROW;
        const double* pa = prm.a + prm.head + (long long)blockIdx.x * ROW + (long long)tid * 4;  // my first row
        const double* pb = DOT ? prm.b + prm.head + (long long)blockIdx.x * ROW + (long long)tid * 4 : nullptr;
        unsigned k = 0;                                        // rows consumed so far
        // The first DD rows are loaded at once (slot u holds row k + u), the column is cleared behind those loads, and
        // the exponents of those rows decide -- for free, they are in registers -- whether the register windows are
        // worth trying at all: data wider than the widest window (log-uniform 2^+-332) goes straight to the direct
        // loop without paying for a window attempt (which cost ~5 us per launch before this check).
        Vec4 va[DD];
        Vec4 vb[DOT ? DD : 1];
        unsigned loaded = 0;                                   // rows loaded (or consumed by loop 1) so far
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
        for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
        zeroed = true;
        // ---------------- loop 1: register window (out of line) ----------------
        // (vectors too short for the windows to pay -- fewer than 32 rows per CTA, n < ~2^23 -- go straight to loop 2)
        if constexpr (DW > 0) if (prm.window && iters >= 32u) {
            unsigned emax = 0u, emin = 0xfffu;
#pragma unroll
            for (int u = 0; u < DD; ++u) {
                const double xs[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
                const double ys[4] = {vb[DOT ? u : 0].x, vb[DOT ? u : 0].y, vb[DOT ? u : 0].z, vb[DOT ? u : 0].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    unsigned e = ((unsigned)__double2hiint(xs[q]) >> 20) & 0x7ffu;
                    if (DOT) {
                        const unsigned eb = ((unsigned)__double2hiint(ys[q]) >> 20) & 0x7ffu;
                        e = (e && eb) ? e + eb : 0u;           // exponent of the product (+ 1023), 0 for zeros
                    }
                    emax = max(emax, e);
                    emin = min(emin, e ? e : 0xfffu);
                }
            }
            emax = __reduce_max_sync(kFullWarp, emax);
            emin = __reduce_min_sync(kFullWarp, emin);
            constexpr unsigned kWidest = DOT ? 50u + 52u * 2u : 51u + 52u;      // W = 5 products / W = 3 summands
            if (!(emax > emin && emax - emin + 1u > kWidest)) {
                const double* pa0 = pa - (long long)loaded * row_step;           // back to my first row
                const double* pb0 = DOT ? pb - (long long)loaded * row_step : nullptr;
                unsigned st1 = 0;
                int range[2] = {4096, -4096};                  // exponents that have missed so far, warp-uniform
                if constexpr (DOT) {
                    if (prm.window != 3)
                        k = reduce0_window_rows_wide<DW, 3>(pa0, pb0, row_step, iters, col, stride, &st1, range);
                    if (k < iters && prm.window > 1)
                        // (three rows in flight, each slot refilled AFTER its row is consumed: 5.3 TB/s sustained on the ill-conditioned
                        // pair against 5.2 / 5.1 / 4.8 for 2 / 4 / 1 rows and 4.75 for two rows refilled BEFORE the row is consumed --
                        // profiles/ab_w5_r02.jsonl.  Called from this kernel the function has ~30 registers less than on its own:
                        // the early refill, once the best, now spills its load slots; the wide loop spends ~45 instructions per product)
                        k += reduce0_window_rows_wide<3, 5, false>(pa0 + (long long)k * row_step, pb0 + (long long)k * row_step, row_step,
                                                                   iters - k, col, stride, &st1, range);
                } else {
                    if (prm.window != 3)
                        k = reduce0_window_rows_wide_sum<DW, 2>(pa0, row_step, iters, col, stride, &st1, range);
                    if (k < iters && prm.window > 1)
                        k += reduce0_window_rows_wide_sum<DW + 2, 3>(pa0 + (long long)k * row_step, row_step, iters - k, col, stride, &st1, range);
                }
                status |= st1;
                // ---------------- loop 2 restarts at row k (rows that were in flight come from L2 now) ----------------
                pa = pa0 + (long long)k * row_step;
                if (DOT) pb = pb0 + (long long)k * row_step;
                loaded = k;
#pragma unroll
                for (int u = 0; u < DD; ++u) {
                    // (every slot is redefined here, so that none of the first prefetch stays live across the calls above)
                    if (loaded < iters) {
                        load_row(u);
                    } else {
                        va[u] = Vec4{0.0, 0.0, 0.0, 0.0};
                        if (DOT) vb[DOT ? u : 0] = Vec4{0.0, 0.0, 0.0, 0.0};
                    }
                }
            }
        }
        // ---------------- loop 2: direct deposits, DD rows in flight; slot u holds row k + u ----------------
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
    if (!zeroed)
        for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
    EXB_PHASE(1);
    double none[1][expansions(0)];
    reduce_finish<0, false, DOT>(prm, col, stride, smem_base, T, tid, prm.iters * (long long)gridDim.x * T, none, status, 0, false);
}

// scalar part of exblas_small_kernel: alignment head, tail, or the whole strided / misaligned vector (four loads in
// flight).  Out of line so that the usual aligned unit-stride call does not even fetch it.
template <bool DOT>
__device__ __noinline__ unsigned small_scalar_part(const ReduceParams& prm, const unsigned col, const unsigned stride,
                                                   const unsigned tid, const unsigned T) {
    unsigned status = 0;
    double none[1][expansions(0)];
    const long long body = 4 * prm.nvec;
    const long long nscalar = prm.n - body;
    for (long long k = tid; k < nscalar; k += 4ll * T) {
        double xa[4], xb[DOT ? 4 : 1];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long kk = k + (long long)j * T;
            if (kk < nscalar) {
                const long long idx = kk < prm.head ? kk : kk + body;
                xa[j] = ldg64(prm.a + idx * prm.inca);
                if (DOT) xb[DOT ? j : 0] = ldg64(prm.b + idx * prm.incb);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (k + (long long)j * T < nscalar) {
                if (DOT) mul_add1<0, false>(col, stride, none, status, xa[j], xb[DOT ? j : 0]);
                else deposit(col, stride, xa[j], status);
            }
        }
    }
    return status;
}

// ------------------------------------------------------------------------------------------------
// Latency regime (n <= a few thousand elements): ONE CTA, and a kernel that is SMALL.  A reduction of 8 KB is over
// in the time a streaming kernel spends fetching its own (unrolled, run-once, instruction-cache-cold) code, so this
// kernel trades every throughput device for short code: up to four 256-bit loads per thread issued first, direct
// deposits through the out-of-line deposit (one hot copy of the code), no expansions -- fpe never changes the
// result, and with a handful of summands per thread there is nothing for an expansion to amortise -- no bounding
// (<= 64 deposits per column), and the shared epilogue published straight from shared memory.
// The host guarantees nvec <= 4 * T and a fresh, closing launch (solo).
// ------------------------------------------------------------------------------------------------
template <bool DOT>
__global__ void __launch_bounds__(512, 1) exblas_small_kernel(const __grid_constant__ ReduceParams prm) {
    extern __shared__ long long smem[];
    __shared__ long long block_lo[kLimbs];
    __shared__ int block_hi[kLimbs];
    __shared__ unsigned block_status[3];
    const unsigned T = blockDim.x;
    const unsigned tid = threadIdx.x;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned stride = 8u * T;
    const unsigned col = smem_base + 8u * tid;
    EXB_PHASE(0);
    RowRange rr = rr_empty();
    Vec4 va[4];
    Vec4 vb[DOT ? 4 : 1];
    const long long nvec = prm.nvec;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long long r = (long long)tid + (long long)u * T;
        if (r < nvec) {
            va[u] = ldg256(prm.a + prm.head + 4 * r);
            if (DOT) vb[DOT ? u : 0] = ldg256(prm.b + prm.head + 4 * r);
        }
    }
#pragma unroll 13
    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
    unsigned status = 0;
    double none[1][expansions(0)];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if ((long long)tid + (long long)u * T < nvec) {
            const double x4[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (DOT) {
                    const double y4[4] = {vb[DOT ? u : 0].x, vb[DOT ? u : 0].y, vb[DOT ? u : 0].z, vb[DOT ? u : 0].w};
                    rr_note_product(rr, x4[k], y4[k]);
                    mul_add1<0, false>(col, stride, none, status, x4[k], y4[k]);
                } else {
                    rr_note(rr, x4[k]);
                    deposit(col, stride, x4[k], status);
                }
            }
        }
    }
    EXB_PHASE(1);
    // scalar part: alignment head, tail, or the whole strided / misaligned vector (out of line: usually absent)
    if (prm.n != 4 * nvec) {
        status |= small_scalar_part<DOT>(prm, col, stride, tid, T);
        rr = rr_full();                                    // (rare path: not tracked)
    }
    EXB_PHASE(4);
    unsigned row_lo, row_hi;
    rr_rows(rr, 0u, row_lo, row_hi);
    block_merge_and_close(prm, stride, smem_base, T, tid, status, true, block_lo, block_hi, block_status, row_lo, row_hi);
}

// Multi-GPU epilogue: the result slot's limbs and flag counters have been summed over ranks by an
// integer all-reduce; normalise, rebuild the status word and round.  Every rank runs this on the
// same integers, so every rank gets the same bits.
__global__ void exblas_finalize_kernel(Result* res, int round_mode) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const unsigned ln = threadIdx.x;
    WarpLimbs x;
    x.a = res->limbs[ln];
    x.b = ln < 7u ? res->limbs[32 + ln] : 0ll;
    unsigned st = 0;
    for (int k = 0; k < kFlagSlots; ++k)
        if (res->flagcnt[k] != 0) st |= (1u << k);
    const bool neg = warp_normalize(x, ln);
    const double v = warp_value(x, neg, st, round_mode, ln);
    __syncwarp();
    res->limbs[ln] = x.a;
    if (ln < 7u) res->limbs[32 + ln] = x.b;
    if (ln < (unsigned)kFlagSlots) res->flagcnt[ln] = (st >> ln) & 1u;
    if (ln == 0) {
        res->value = v;
        res->status = st;
    }
}

}  // namespace exb
