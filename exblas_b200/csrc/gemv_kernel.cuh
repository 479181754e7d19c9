// gemv_kernel.cuh -- ExGEMV (y := alpha*op(A)*x + beta*y, column-major A) for sm_100a.
// 'N' is the tuned case (coalesced); 'T' runs the same kernel with the strides swapped (correct,
// uncoalesced: SURVEY section 8f rank 3 leaves its fast path for later).
//
// SURVEY.md section 8f rank 1 / BASELINE config 5.  Replaces the reference's OpenCL kernels
//   gemv / gemv_reduce   src/gpu/blas/blas2/ExGEMV.FPE.cl:199-379, 561-580, ExGEMV.FPE.EX.{4,6,8}.cl,
//                        ExGEMV.Superacc.cl:192-290, 397
// by per-row reuse of the ExDOT device code (reduce_kernel.cuh): a thread owns one row, streams it
// with coalesced column loads (a warp reads 32 consecutive rows of one column = 256 B), multiplies
// by x with TwoProductFMA and feeds the two parts to its register expansion / private
// shared-memory superaccumulator column exactly as ExDOT does.
//
// What is different from the reference kernels:
//   * columns are split over `parts` CTAs per row block so that 148 SMs are busy even when m is
//     small (the reference hard-codes p = 1, ExGEMV.cpp:165); the per-part limbs go to a scratch
//     array laid out [part][limb][row] (coalesced), and exgemv_finish_kernel sums them as integers;
//   * x is read from global memory through L1 (the reference stages all n values in local memory,
//     ExGEMV.FPE.cl:216-232, which cannot work for n = 32768);
//   * alpha is honoured (the reference's non-transpose FPE kernel ignores it, ExGEMV.FPE.cl:246):
//     alpha == 1 multiplies directly; any other alpha is applied exactly, alpha*a = p1 + e1
//     (TwoProd), then (p1 + e1)*x by two more TwoProds, so the row sum stays exact;
//   * beta*y is added exactly (TwoProd) as ExGEMV.FPE.cl:346-377 does, for any beta.
#pragma once
#include "reduce_kernel.cuh"

namespace exb {

struct GemvParams {
    const double* a;        // already offset by offseta; output r, summand c reads a[r * rs + c * cs]
    const double* x;        // already offset by offsetx
    double* y;              // already offset by offsety
    long long m, n, rs, cs, incx, incy;   // m outputs, n summands each ('N': rs = 1, cs = lda; 'T': rs = lda, cs = 1)
    double alpha, beta;
    long long cols_per_part;      // multiple of 4
    int parts;
    long long* scratch;     // [parts][kLimbs][m]
    unsigned* row_status;   // [parts][m]
    Workspace* ws;
    int round_mode;
    int adaptive;
    int x_vec_ok;           // x contiguous and 32-byte aligned at every part start
};

EXB_D Vec4 ldg256_cached(const double* p) {     // through L1: x is re-read by every warp of the CTA
    Vec4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}

EXB_D double ldg64(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

template <int F, bool EE, bool ALPHA1, int U, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) exgemv_n_kernel(const GemvParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;
    const unsigned tid = threadIdx.x;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    const unsigned stride = 8u * T;
    const unsigned col = smem_base + 8u * tid;
    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);

    unsigned status = 0;
    constexpr int kM = expansions(F);
    double a[F > 0 ? F : 1][kM];
#pragma unroll
    for (int i = 0; i < (F > 0 ? F : 1); ++i)
#pragma unroll
        for (int m = 0; m < kM; ++m) a[i][m] = 0.0;

    const long long row_raw = (long long)blockIdx.x * T + tid;
    const bool valid = row_raw < prm.m;
    const long long row = valid ? row_raw : prm.m - 1;       // idle lanes redo the last row (keeps votes uniform)
    const long long c0 = (long long)blockIdx.y * prm.cols_per_part;
    long long c1 = c0 + prm.cols_per_part;
    if (c1 > prm.n) c1 = prm.n;
    const long long ncols = c1 > c0 ? c1 - c0 : 0;
    const long long ngroups = ncols / 4;                     // full groups of 4 columns
    constexpr bool unit_alpha = ALPHA1;                      // alpha == 1 (the only case the reference tests) is its own instantiation
    constexpr int kDepPerGroup = 4 * 2 * 2;                  // 4 columns, <= 2 products each when alpha != 1, 2 parts
    const double* pa = prm.a + row * prm.rs + prm.cs * c0;
    const double* px = prm.x + prm.incx * c0;
    const long long astep = 4 * prm.cs, xstep = 4 * prm.incx;

    // Rolling window of U column groups (4 columns each) of A per thread, refilled right after use.
    // All addressing is by running pointers (no 64-bit multiplies in the loop).  x[k..k+3] is one
    // 256-bit broadcast load when x is contiguous and 32-byte aligned (every lane reads the same
    // address: one L1 wavefront), else four scalar broadcast loads.
    double va[U][4];
    Vec4 vx[U];
    const long long cs = prm.cs, incx = prm.incx;
    const bool xvec = prm.x_vec_ok != 0;
    const double* qa = pa;                                   // next group of A to load
    const double* qx = px;                                   // next group of x to load
    auto load_group = [&](int u) {
        va[u][0] = ldg64(qa);
        va[u][1] = ldg64(qa + cs);
        va[u][2] = ldg64(qa + 2 * cs);
        va[u][3] = ldg64(qa + 3 * cs);
        if (xvec) {
            vx[u] = ldg256_cached(qx);
        } else {
            vx[u].x = __ldg(qx);
            vx[u].y = __ldg(qx + incx);
            vx[u].z = __ldg(qx + 2 * incx);
            vx[u].w = __ldg(qx + 3 * incx);
        }
        qa += astep;
        qx += xstep;
    };
    const long long rounds = ngroups / U;                    // full rounds of U groups
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (rounds > 0) load_group(u);

    int since_norm = 0;
    int bypass = 0, backoff = kBypassTiles;
    auto consume = [&](const double (&xa_in)[4], const Vec4& xv, bool direct, int& deposits) {
        double xa[4] = {xa_in[0], xa_in[1], xa_in[2], xa_in[3]};
        const double xb[4] = {xv.x, xv.y, xv.z, xv.w};
        if constexpr (unit_alpha) {
            if (direct) {
                double none[1][expansions(0)];
                mul_add4<0, false, true>(col, stride, none, status, xa, xb);
            } else {
                deposits += mul_add4<F, EE, true>(col, stride, a, status, xa, xb);
            }
        } else {
            // alpha * a = p1 + e1 exactly; then p1 * x and e1 * x
            double p1[4], e1[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                p1[k] = __dmul_rn(prm.alpha, xa[k]);
                e1[k] = __fma_rn(prm.alpha, xa[k], -p1[k]);
            }
            if (direct) {
                double none[1][expansions(0)];
                mul_add4<0, false, true>(col, stride, none, status, p1, xb);
                mul_add4<0, false, true>(col, stride, none, status, e1, xb);
            } else {
                deposits += mul_add4<F, EE, true>(col, stride, a, status, p1, xb);
                deposits += mul_add4<F, EE, true>(col, stride, a, status, e1, xb);
            }
        }
    };
    for (long long r = 0; r < rounds; ++r) {
        const bool direct = (F == 0) || (prm.adaptive && bypass > 0);
        const bool has_next = r + 1 < rounds;
        int deposits = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double xa[4] = {va[u][0], va[u][1], va[u][2], va[u][3]};
            const Vec4 xv = vx[u];
            if (has_next) load_group(u);
            consume(xa, xv, direct, deposits);
        }
        if (F > 0 && prm.adaptive) {
            if (bypass > 0) {
                --bypass;
            } else {
                const int total = __reduce_add_sync(0xffffffffu, deposits);
                if (total * 64 >= 32 * U * kDepPerGroup) {
                    bypass = backoff;
                    backoff = min(backoff * 16, kBypassMax);   // a second thrashing probe in a row: stay away for long
                } else {
                    backoff = kBypassTiles;
                }
            }
        }
        since_norm += U * kDepPerGroup;
        if (since_norm > kMaxDepositsPerNormalize - U * kDepPerGroup - 2 * kM * (F + 2) - 64) {
            bound_column(col, stride);
            since_norm = 0;
        }
    }
    // groups left over after the full rounds (< U of them), then columns left over (< 4)
    for (long long g = rounds * U; g < ngroups; ++g) {
        const double* ra = pa + g * astep;
        const double* rx = px + g * xstep;
        double xa[4] = {ra[0], ra[cs], ra[2 * cs], ra[3 * cs]};
        Vec4 xv;
        xv.x = rx[0]; xv.y = rx[incx]; xv.z = rx[2 * incx]; xv.w = rx[3 * incx];
        int deposits = 0;
        consume(xa, xv, F == 0, deposits);
    }
    bound_column(col, stride);
    // leftover columns (< 4)
    for (long long c = c0 + ngroups * 4; c < c1; ++c) {
        const double av = prm.a[row * prm.rs + prm.cs * c], xv = prm.x[prm.incx * c];
        if constexpr (unit_alpha) {
            mul_add1<F, EE>(col, stride, a, status, av, xv);
        } else {
            const double p1 = __dmul_rn(prm.alpha, av);
            const double e1 = __fma_rn(prm.alpha, av, -p1);
            mul_add1<F, EE>(col, stride, a, status, p1, xv);
            mul_add1<F, EE>(col, stride, a, status, e1, xv);
        }
    }
    if (F > 0) {
#pragma unroll
        for (int i = 0; i < F; ++i)
#pragma unroll
            for (int m = 0; m < kM; ++m) deposit(col, stride, a[i][m], status);
    }
    bound_column(col, stride);
    if (status && valid) atomicOr(&prm.ws->status, status);
    if (valid) prm.row_status[(long long)blockIdx.y * prm.m + row_raw] = status;
    // this thread's limbs -> scratch[part][limb][row]
    if (valid) {
        long long* out = prm.scratch + (long long)blockIdx.y * kLimbs * prm.m + row_raw;
        for (int j = 0; j < kLimbs; ++j) out[(long long)j * prm.m] = (long long)lds64(col + j * stride);
    }
}

// One thread per row: integer sum of the per-part limbs, + beta*y exactly, round, store.
__global__ void exgemv_finish_kernel(const GemvParams prm) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= prm.m) return;
    long long acc[kLimbs];
    for (int j = 0; j < kLimbs; ++j) {
        long long s = 0;
        for (int p = 0; p < prm.parts; ++p) s += prm.scratch[((long long)p * kLimbs + j) * prm.m + row];
        acc[j] = s;                                          // parts <= 2048 normalised limbs: no overflow
    }
    unsigned st = 0;
    for (int p = 0; p < prm.parts; ++p) st |= prm.row_status[(long long)p * prm.m + row];
    double* yp = prm.y + row * prm.incy;
    if (prm.beta != 0.0) {
        const double yv = *yp;
        if (prm.beta == 1.0) {
            st |= accumulate_double(acc, yv);
        } else {
            const double p = __dmul_rn(prm.beta, yv);
            const double e = __fma_rn(prm.beta, yv, -p);
            st |= accumulate_double(acc, p);
            if (!(e != e)) st |= accumulate_double(acc, e);
        }
    }
    *yp = finalize_value(acc, st, prm.round_mode);
    if (st) atomicOr(&prm.ws->status, st);
}

}  // namespace exb
