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
};

EXB_D Vec4 ldg256(const double* p) {
    Vec4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
}

EXB_D bool nonzero_bits(double x) {
    return (((unsigned)__double2hiint(x) & 0x7fffffffu) | (unsigned)__double2loint(x)) != 0u;
}

// One element through the expansion, starting at level `first`.  Returns the residual.
// Knuth TwoSum, un-contracted (ExSUM.FPE.cl:27-32).  With EE the walk stops as soon as no lane
// of the warp (UNIFORM) or this thread (!UNIFORM) has a non-zero residual.
template <int F, bool EE, bool UNIFORM>
EXB_D double fpe_push(double (&a)[F > 0 ? F : 1], double x, int first = 0) {
#pragma unroll
    for (int i = 0; i < F; ++i) {
        if (i < first) continue;
        const double r = __dadd_rn(a[i], x);
        const double z = __dsub_rn(r, a[i]);
        const double s = __dadd_rn(__dsub_rn(a[i], __dsub_rn(r, z)), __dsub_rn(x, z));
        a[i] = r;
        x = s;
        if (EE && i + 1 < F) {
            if (UNIFORM) {
                if (!__any_sync(0xffffffffu, nonzero_bits(x))) break;
            } else {
                if (!nonzero_bits(x)) break;
            }
        }
    }
    return x;
}

// ---- per-thread state: (col, stride) of the private accumulator column, status flags, and the
// expansion a[F].  Everything is passed by reference into force-inlined helpers so that it stays
// in registers (a struct with out-of-line members would be spilled to local memory).

// One summand (any double).  No lane leaves early: the warp-uniform early-exit votes inside
// fpe_push must be reached by every lane, so special values are diverted and replaced by 0.
template <int F, bool EE, bool UNIFORM>
EXB_D void add_value(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1], unsigned& status, double x) {
    if (F == 0) {
        deposit(col, stride, x, status);
    } else {
        const unsigned hi = (unsigned)__double2hiint(x);
        if ((hi & 0x7fffffffu) >= (kELim << 20)) {             // Inf / NaN / too large: keep out of the FPE
            status |= deposit_slow(col, stride, (unsigned)__double2loint(x), hi);
            x = 0.0;
        }
        const double r = fpe_push<F, EE, UNIFORM>(a, x, 0);
        if (nonzero_bits(r)) deposit(col, stride, r, status);
    }
}

// The error term of a product enters the expansion lower down (ExDOT.FPE.cl:254-258,
// ExDOT.FPE.EX.4.cl: level 1 with early exit).
template <int F, bool EE, bool UNIFORM>
EXB_D void add_error(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1], unsigned& status, double e) {
    if (F == 0) {
        deposit(col, stride, e, status);
    } else {
        constexpr int first = EE ? (F > 1 ? 1 : 0) : (F > 3 ? F - 3 : 0);
        const double r = fpe_push<F, EE, UNIFORM>(a, e, first);
        if (nonzero_bits(r)) deposit(col, stride, r, status);
    }
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
    // tiny product (|p| < 2^-935).  The FMA error term is exact iff the true error's LSB,
    // 2^(Ex-1075 + Ey-1075), is representable (>= 2^-1074); otherwise bits were lost.
    unsigned st = 0u;
    const unsigned ex = (xh >> 20) ? (xh >> 20) : 1u, ey = (yh >> 20) ? (yh >> 20) : 1u;
    if (ex + ey < 1076u) st |= kStTooSmall;
    deposit(col, stride, p, st);                                   // flags kStTooSmall itself if it truncates
    deposit(col, stride, e, st);
    return st;
}

template <int F, bool EE, bool UNIFORM>
EXB_D void add_product(unsigned col, unsigned stride, double (&a)[F > 0 ? F : 1], unsigned& status, double x, double y) {
    double p = __dmul_rn(x, y);
    double e = __fma_rn(x, y, -p);                                 // TwoProductFMA, ExDOT.FPE.cl:25-29
    const unsigned ph = (unsigned)__double2hiint(p) & 0x7fffffffu;
    // TwoProd is exact and both parts lie inside the layout when 2^-935 <= |p| < 2^988
    if (((ph >> 20) - 88u) >= (kELim - 88u)) {
        status |= product_slow(col, stride, x, y, p, e);
        p = 0.0;
        e = 0.0;
    }
    add_value<F, EE, UNIFORM>(col, stride, a, status, p);
    add_error<F, EE, UNIFORM>(col, stride, a, status, e);
}

// Sum one limb row over all T columns; result valid in lane 0.
EXB_D long long row_sum(unsigned row_addr, unsigned T, unsigned lane) {
    long long s = 0;
    for (unsigned t = lane; t < T; t += 32) s += (long long)lds64(row_addr + 8u * t);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    return s;
}

template <int F, bool EE, bool DOT, int U, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) exblas_reduce_kernel(const ReduceParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;
    const unsigned tid = threadIdx.x;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned stride = 8u * T;
    const unsigned col = smem_base + 8u * tid;
    __shared__ long long block_limbs[kLimbs];
    __shared__ unsigned is_last;

    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
    // columns are thread-private: no barrier needed before use

    unsigned status = 0;
    double a[F > 0 ? F : 1];
#pragma unroll
    for (int i = 0; i < (F > 0 ? F : 1); ++i) a[i] = 0.0;

    constexpr int kDepPerElem = DOT ? 2 : 1;               // at most one deposit per summand
    constexpr int kDepPerTile = 4 * U * kDepPerElem;
    const long long TILE = (long long)T * 4 * U;

    // ---------------- vector body: full tiles, 256-bit loads, rolling prefetch ----------------
    if (prm.ntiles > 0) {
        const double* pa = prm.a + prm.head + (long long)tid * 4;
        const double* pb = DOT ? prm.b + prm.head + (long long)tid * 4 : nullptr;
        Vec4 va[U];
        Vec4 vb[DOT ? U : 1];
        long long tile = blockIdx.x;
        if (tile < prm.ntiles) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                va[u] = ldg256(pa + tile * TILE + (long long)u * T * 4);
                if (DOT) vb[u] = ldg256(pb + tile * TILE + (long long)u * T * 4);
            }
        }
        int since_norm = 0;
        while (tile < prm.ntiles) {
            const long long next = tile + gridDim.x;
            const bool has_next = next < prm.ntiles;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const Vec4 x = va[u];
                Vec4 y;
                if (DOT) y = vb[u];
                if (has_next) {
                    va[u] = ldg256(pa + next * TILE + (long long)u * T * 4);
                    if (DOT) vb[u] = ldg256(pb + next * TILE + (long long)u * T * 4);
                }
                if (DOT) {
                    add_product<F, EE, true>(col, stride, a, status, x.x, y.x);
                    add_product<F, EE, true>(col, stride, a, status, x.y, y.y);
                    add_product<F, EE, true>(col, stride, a, status, x.z, y.z);
                    add_product<F, EE, true>(col, stride, a, status, x.w, y.w);
                } else if (F == 0) {
                    deposit4(col, stride, x.x, x.y, x.z, x.w, status);
                } else {
                    add_value<F, EE, true>(col, stride, a, status, x.x);
                    add_value<F, EE, true>(col, stride, a, status, x.y);
                    add_value<F, EE, true>(col, stride, a, status, x.z);
                    add_value<F, EE, true>(col, stride, a, status, x.w);
                }
            }
            tile = next;
            since_norm += kDepPerTile;
            if (since_norm > kMaxDepositsPerNormalize - kDepPerTile - 2 * (F + 2)) {
                normalize_column(col, stride);
                since_norm = 0;
            }
        }
        normalize_column(col, stride);
    }

    // ---------------- scalar part: alignment head, tail, or the whole strided vector ----------
    {
        const long long body = prm.ntiles * TILE;
        const long long nscalar = prm.n - body;            // head + tail (or everything)
        const long long gthreads = (long long)gridDim.x * T;
        int since_norm = 0;
        for (long long k = (long long)blockIdx.x * T + tid; k < nscalar; k += gthreads) {
            const long long idx = k < prm.head ? k : k + body;
            if (DOT) add_product<F, EE, false>(col, stride, a, status, prm.a[idx * prm.inca], prm.b[idx * prm.incb]);
            else add_value<F, EE, false>(col, stride, a, status, prm.a[idx * prm.inca]);
            since_norm += kDepPerElem;
            if (since_norm > kMaxDepositsPerNormalize - kDepPerElem - 2 * (F + 2)) {
                normalize_column(col, stride);
                since_norm = 0;
            }
        }
    }
    if (F > 0) {
#pragma unroll
        for (int i = 0; i < F; ++i) deposit(col, stride, a[i], status);
    }
    normalize_column(col, stride);
    if (status) atomicOr(&prm.ws->status, status);
    __syncthreads();

    // ---------------- block merge: 39 row sums -> normalise -> global accumulator --------------
    {
        const unsigned warp = tid >> 5, ln = tid & 31u, nwarps = T >> 5;
        for (unsigned j = warp; j < (unsigned)kLimbs; j += nwarps) {
            const long long s = row_sum(smem_base + j * stride, T, ln);   // |s| < T * 2^52 <= 2^62
            if (ln == 0) block_limbs[j] = s;
        }
        __syncthreads();
        if (tid == 0) normalize(block_limbs);
        __syncthreads();
        if (tid < (unsigned)kLimbs) {
            const long long v = block_limbs[tid];
            if (v != 0) atomicAdd(&prm.ws->gacc[tid], (unsigned long long)v);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            const unsigned ticket = atomicAdd(&prm.ws->counter, 1u);
            is_last = (ticket == gridDim.x - 1);
        }
        __syncthreads();
    }

    // ---------------- last CTA: normalise the global accumulator, publish ----------------------
    if (is_last) {
        __threadfence();
        if (tid < (unsigned)kLimbs)
            block_limbs[tid] = (long long)atomicExch(&prm.ws->gacc[tid], 0ull);
        __syncthreads();
        if (tid == 0) {
            unsigned st = prm.finalize && !prm.keep ? atomicExch(&prm.ws->status, 0u) : atomicOr(&prm.ws->status, 0u);
            const bool neg = normalize(block_limbs);
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
            prm.ws->counter = 0;
        }
        __syncthreads();
        // leave the (normalised) partial sum in the workspace unless this call closes the reduction
        if ((!prm.finalize || prm.keep) && tid < (unsigned)kLimbs)
            prm.ws->gacc[tid] = (unsigned long long)block_limbs[tid];
    }
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
