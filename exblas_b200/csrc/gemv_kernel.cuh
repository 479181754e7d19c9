// gemv_kernel.cuh -- ExGEMV (y := alpha*op(A)*x + beta*y, column-major A) for sm_100a.
// 'N' (thread per row, coalesced column loads) and 'T' (warp per output, coalesced column streaming) each have a
// register-window kernel for alpha == 1 (every fpe value) and share exgemv_n_kernel -- expansions, any alpha, any
// strides -- as the general path.
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
    int l2_prefetch;        // window kernels: bulk L2 prefetch distance in rounds (0 = off; needs A 16-byte aligned, lda even)
};


EXB_D Vec4 ldg256_cached(const double* p) {     // through L1: x is re-read by every warp of the CTA
    Vec4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}


// alpha * a = p + e exactly (TwoProductFMA) -- provided nothing underflows or overflows on the way.  The exact range is
// 2^-959 <= |alpha * a| < 2^1024; outside it (and for non-zero operands) the scaled value cannot be represented by two
// doubles, so it is dropped and FLAGGED instead of being summed inexactly: kStTooSmall (bits were lost: the row sum is
// no longer exact) / kStTooLarge.  Inf / NaN operands pass through with their IEEE meaning.
EXB_D void scale_exact(double alpha, double a, double& p, double& e, unsigned& status) {
    p = __dmul_rn(alpha, a);
    e = __fma_rn(alpha, a, -p);
    const unsigned ph = (unsigned)__double2hiint(p) & 0x7fffffffu;
    if (ph - (64u << 20) >= ((0x7ffu - 64u) << 20)) {                     // |p| < 2^-959, or Inf / NaN
        const unsigned ah = (unsigned)__double2hiint(alpha) & 0x7fffffffu, xh = (unsigned)__double2hiint(a) & 0x7fffffffu;
        const bool finite = ah < 0x7ff00000u && xh < 0x7ff00000u;
        const bool nonzero = ((ah | (unsigned)__double2loint(alpha)) != 0u) && ((xh | (unsigned)__double2loint(a)) != 0u);
        if (finite && nonzero) {
            status |= ph >= 0x7ff00000u ? kStTooLarge : kStTooSmall;
            p = 0.0;
        }
        if (finite) e = 0.0;                                              // (e of an exact zero / a dropped value)
    }
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
                mul_add4<0, false, true, true>(col, stride, none, status, xa, xb);
            } else {
                deposits += mul_add4<F, EE, true>(col, stride, a, status, xa, xb);
            }
        } else {
            // alpha * a = p1 + e1 exactly; then p1 * x and e1 * x
            double p1[4], e1[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) scale_exact(prm.alpha, xa[k], p1[k], e1[k], status);
            if (direct) {
                double none[1][expansions(0)];
                mul_add4<0, false, true, false>(col, stride, none, status, p1, xb);
                mul_add4<0, false, true, false>(col, stride, none, status, e1, xb);
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
            double p1, e1;
            scale_exact(prm.alpha, av, p1, e1, status);
            mul_add1<F, EE>(col, stride, a, status, p1, xv);
            mul_add1<F, EE>(col, stride, a, status, e1, xv);
        }
    }
    if (F > 0) {
#pragma unroll
        for (int i = 0; i < F; ++i)
#pragma unroll
            for (int m = 0; m < kM; ++m) deposit_sum(col, stride, a[i][m], status);
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

// ------------------------------------------------------------------------------------------------
// ExGEMV 'N', superaccumulator-only mode (fpe < 2, alpha == 1) with a register window (window.cuh).
//
// Same decomposition as exgemv_n_kernel (thread per row, column split over blockIdx.y, per-part
// limbs to scratch), but:
//   * the CTA's slice of x is staged ONCE in shared memory behind the accumulator columns and read
//     back with broadcast LDS.128 (no per-thread x registers in the load window, no L1 traffic);
//   * A runs U column groups (4 columns each) ahead: 32 8-byte loads in flight per thread,
//     ~96 KB per SM;
//   * a group whose four products lie inside the window of EVERY lane of the warp (one vote) is
//     accumulated in registers: 2 + 8 FP64 and 8 integer instructions per element, no shared-memory
//     traffic.  Any other group takes the ordinary path (mul_add4<0>: TwoProd + two deposits per
//     element) out of line, and lanes that keep missing re-anchor their window there;
//   * a warp whose groups mostly miss (wide-range rows) stops voting for a while (exponential
//     back-off), so such data costs what it cost before.
// The result is the exact row sum whichever path each element took.
// ------------------------------------------------------------------------------------------------
EXB_D void lds128(unsigned addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}

constexpr int kGemvXsMax = 8192;          // doubles of x staged per CTA (64 KB)

template <int U, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) exgemv_n_win_kernel(const GemvParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;                           // <= MAXT
    const unsigned tid = threadIdx.x;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    constexpr unsigned stride = 8u * MAXT;                   // compile-time limb stride: column addresses are immediates
    const unsigned col = smem_base + 8u * tid;
#pragma unroll
    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);

    const long long row_raw = (long long)blockIdx.x * T + tid;
    const bool valid = row_raw < prm.m;
    const long long row = valid ? row_raw : prm.m - 1;       // idle lanes redo the last row (keeps votes uniform)
    const long long c0 = (long long)blockIdx.y * prm.cols_per_part;
    long long c1 = c0 + prm.cols_per_part;
    if (c1 > prm.n) c1 = prm.n;
    const int ncols = (int)(c1 > c0 ? c1 - c0 : 0);          // <= kGemvXsMax
    const int ngroups = ncols / 4;

    // stage x[c0 .. c1) behind the columns
    double* xs = reinterpret_cast<double*>(smem + (size_t)kLimbs * MAXT);
    const unsigned xs_base = smem_base + stride * (unsigned)kLimbs;
    for (int k = (int)tid; k < ncols; k += (int)T) xs[k] = prm.x[prm.incx * (c0 + k)];
    __syncthreads();

    Window w;
    win_reset(w);
    const long long cs = prm.cs;
    const double* qa = prm.a + row * prm.rs + cs * c0;       // next group of A to load
    const long long astep = 4 * cs;
    double va[U][4];
    // (Refills are predicated on "there is a next round".  Making them unconditional by parking the pointer in
    // the last round was measured 30 % slower: the compiler hoists the loads and spills the window.)
    auto load_group = [&](int u) {
        va[u][0] = ldg64(qa);
        va[u][1] = ldg64(qa + cs);
        va[u][2] = ldg64(qa + 2 * cs);
        va[u][3] = ldg64(qa + 3 * cs);
        qa += astep;
    };
    const int rounds = ngroups / U;
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (rounds > 0) load_group(u);

    unsigned xaddr = xs_base;                                // x of the group being consumed
    int since_norm = 0;
    int r = 0;
    // ---- loop 1: register window.  Two rounds in a row that mostly miss (wide-range rows) end it. ----
    // L2 prefetch (TMA engine, UBLKPF): the kernel is bound by the latency of its 8-byte global loads (ncu, round 1: a
    // quarter of all stall samples on the first use of a loaded value; 64 KB in flight per SM is all the registers
    // allow).  Every round, warp w asks for column 4 U (r + D) + w of this CTA's row block -- one contiguous run of
    // T * 8 bytes -- so that D rounds later the loads hit L2 instead of DRAM.
    const long long rows_here = prm.m - (long long)blockIdx.x * T < (long long)T ? prm.m - (long long)blockIdx.x * T : (long long)T;
    int pf_cidx = prm.l2_prefetch * 4 * U + (int)(tid >> 5);                      // this warp's column (of the part) to ask for next
    const double* pf_ptr = prm.a + (long long)blockIdx.x * T * prm.rs + cs * (c0 + pf_cidx);
    const unsigned pf_bytes = (prm.l2_prefetch > 0 && (tid & 31u) == 0u && (tid >> 5) < 4u * U) ? ((unsigned)(rows_here * 8) & ~15u) : 0u;
    auto prefetch_round = [&]() {
        if (pf_bytes && pf_cidx < ncols) bulk_prefetch_l2(pf_ptr, pf_bytes);
        pf_ptr += astep * U;
        pf_cidx += 4 * U;
    };
    for (int bad = 0; r < rounds && bad < 2; ++r) {
        const bool has_next = r + 1 < rounds;
        int missed = 0;
        prefetch_round();
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double a0 = va[u][0], a1 = va[u][1], a2 = va[u][2], a3 = va[u][3];
            double x0, x1, x2, x3;
            lds128(xaddr, x0, x1);
            lds128(xaddr + 16u, x2, x3);
            xaddr += 32u;
            const double p0 = __dmul_rn(a0, x0), p1 = __dmul_rn(a1, x1), p2 = __dmul_rn(a2, x2), p3 = __dmul_rn(a3, x3);
            const unsigned k0 = ((unsigned)__double2hiint(p0) & 0x7fffffffu) - w.key0;
            const unsigned k1 = ((unsigned)__double2hiint(p1) & 0x7fffffffu) - w.key0;
            const unsigned k2 = ((unsigned)__double2hiint(p2) & 0x7fffffffu) - w.key0;
            const unsigned k3 = ((unsigned)__double2hiint(p3) & 0x7fffffffu) - w.key0;
            const bool mine = max(max(k0, k1), max(k2, k3)) < w.span;
            if (__all_sync(0xffffffffu, mine)) {
                win_add_product(w, p0, __fma_rn(a0, x0, -p0));
                win_add_product(w, p1, __fma_rn(a1, x1, -p1));
                win_add_product(w, p2, __fma_rn(a2, x2, -p2));
                win_add_product(w, p3, __fma_rn(a3, x3, -p3));
                w.cnt += 4u;
            } else {
                w = prod_slow_group(w, col, stride, a0, a1, a2, a3, x0, x1, x2, x3, mine, true);
                ++missed;
            }
            // the slot is refilled AFTER its values are dead: a predicated load issued while they were still needed made
            // the compiler copy all eight registers first (8 moves per group, ~8 % of the loop's instructions)
            if (has_next) load_group(u);
        }
        bad = (2 * missed > U) ? bad + 1 : 0;                // warp-uniform (the votes are)
        if (w.cnt > (unsigned)(kWinFlushEvery - 4 * U)) {
            w = win_flush_products(w, col, stride);
            since_norm += 4;
        }
        since_norm += missed * 12;                           // <= 8 deposits + a 4-deposit drain per slow group
        if (since_norm > kMaxDepositsPerNormalize - 12 * U - 8) {
            bound_column(col, stride);
            since_norm = 0;
        }
    }
    // ---- loop 2 (only after loop 1 gave up): every product takes the ordinary path, inlined ----
    if (r < rounds) {
        w = win_flush_products(w, col, stride);
        since_norm += 4;
        unsigned status = w.st;
        for (; r < rounds; ++r) {
            const bool has_next = r + 1 < rounds;
            prefetch_round();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double xa[4] = {va[u][0], va[u][1], va[u][2], va[u][3]};
                double xb[4];
                lds128(xaddr, xb[0], xb[1]);
                lds128(xaddr + 16u, xb[2], xb[3]);
                xaddr += 32u;
                double none[1][expansions(0)];
                mul_add4<0, false, true, true>(col, stride, none, status, xa, xb);
                if (has_next) load_group(u);
            }
            since_norm += 8 * U;
            if (since_norm > kMaxDepositsPerNormalize - 12 * U - 8) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
        w.st = status;
    }
    // groups left over after the full rounds (< U of them), then columns left over (< 4): ordinary path
    {
        const double* ra = prm.a + row * prm.rs + cs * (c0 + (long long)rounds * U * 4);
        for (int g = rounds * U; g < ngroups; ++g, ra += astep) {
            const double* xg = xs + 4 * g;
            w = prod_slow_group(w, col, stride, ra[0], ra[cs], ra[2 * cs], ra[3 * cs], xg[0], xg[1], xg[2], xg[3], true,
                                false);
        }
        unsigned status = w.st;
        double none[1][expansions(0)];
        for (int c = ngroups * 4; c < ncols; ++c, ra += cs) mul_add1<0, false>(col, stride, none, status, ra[0], xs[c]);
        w.st = status;
    }
    w = win_flush_products(w, col, stride);
    bound_column(col, stride);
    const unsigned status = w.st;
    if (status && valid) atomicOr(&prm.ws->status, status);
    if (valid) prm.row_status[(long long)blockIdx.y * prm.m + row_raw] = status;
    if (valid) {
        long long* out = prm.scratch + (long long)blockIdx.y * kLimbs * prm.m + row_raw;
        for (int j = 0; j < kLimbs; ++j) out[(long long)j * prm.m] = (long long)lds64(col + j * stride);
    }
}

// ------------------------------------------------------------------------------------------------
// ExGEMV 'T' (y_j = alpha * sum_i A[i + j*lda] * x[i] + beta * y_j), alpha == 1, any fpe: one WARP per
// output.  Column j of A is contiguous, so lane l streams rows l, l+32, l+64, ... with coalesced
// 8-byte loads (256 B per warp instruction, no alignment requirement), U groups of 4 rows ahead.
// All warps of the CTA walk the same rows of their respective columns, so x is staged once per CTA
// in chunks (TMA bulk copies completing on mbarriers: see the x pipeline below) and read back with
// conflict-free LDS.64.  Each lane accumulates in its register window (window.cuh) and, for what
// falls outside, its private shared-memory column; at the end of the column lane 0 deposits beta*y
// exactly into its column, the warp drains the windows and sums the 32 columns limb by limb with
// shuffles (limbs that are zero in every lane are skipped after one vote) so that limb j ends up
// in lane j, then normalises and rounds AS A WARP (warp_normalize / warp_value) and stores y_j --
// no scratch, no second kernel, no local memory.  Replaces the reference's gemvT kernels
// (ExGEMV.FPE.cl:382-557, ExGEMV.Superacc.cl:295-395), whose threads walk a row of the transposed
// matrix with stride lda.
// ------------------------------------------------------------------------------------------------
EXB_D double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// The out-of-line halves of exgemv_t_win_kernel's window loop.  Besides the work itself they keep the count of ordinary
// deposits since the column was last bounded -- in bits 16.. of w.st (the status flags live in bits 0..5 and are only
// ever OR-ed in) -- and bound the column when it is due, so that the streaming loop carries no counter for it.
constexpr unsigned kTDepUnit = 1u << 16;
constexpr unsigned kTDepLimit = (unsigned)(kMaxDepositsPerNormalize - 160);   // room for a slow group, a flush and the <= 36-row remainder
EXB_D Window t_dep_note(Window w, unsigned col, unsigned stride, unsigned deposits) {
    w.st += deposits * kTDepUnit;
    if ((w.st >> 16) > kTDepLimit) {
        bound_column(col, stride);
        w.st &= 0xffffu;
    }
    return w;
}
__device__ __noinline__ Window t_slow_group(Window w, unsigned col, unsigned stride, double a0, double a1, double a2, double a3,
                                            double x0, double x1, double x2, double x3, bool mine) {
    w = prod_slow_group(w, col, stride, a0, a1, a2, a3, x0, x1, x2, x3, mine, true);
    return t_dep_note(w, col, stride, 12u);                  // <= 8 deposits + a 4-deposit drain
}
__device__ __noinline__ Window t_flush(Window w, unsigned col, unsigned stride) {
    w = win_flush_products(w, col, stride);
    return t_dep_note(w, col, stride, 4u);
}

// ---- x pipeline of exgemv_t_win_kernel: TMA bulk copies + mbarriers, no CTA-wide barrier ----------------------------
// All warps of a CTA walk the same rows of their columns, so x is staged per CTA, kGemvTChunk rows at a time, in NB
// buffers.  Chunks are numbered g = 0, 1, 2, ... through the whole life of the CTA (x is the same for every set of
// columns, so chunk g is rows (g mod nchunks) * CH ... of x) and live in buffer g mod NB (NB buffers):
//   * a warp about to read chunk g waits on the buffer's "full" mbarrier (phase (g / NB) & 1);
//   * a warp that has finished chunk g counts itself out on the buffer's counter; the LAST warp out resets the counter and
//     issues chunk g + NB into the buffer: one cp.async.bulk (TMA engine, SASS UBLKCP) that completes on the mbarrier --
//     or, when x is strided or not 16-byte aligned, a copy by that warp followed by an ordinary arrive.
// Nobody waits for a slower warp unless it runs a whole chunk ahead of it (round 1 / early round 2: cp.async by all
// threads + one bar.sync per chunk, 10 % of all warp stall cycles).
EXB_D void mbar_init(unsigned addr, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory"); }
EXB_D void mbar_arrive(unsigned addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
EXB_D void mbar_expect_tx(unsigned addr, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
EXB_D void mbar_wait(unsigned addr, unsigned parity) {
    unsigned done;
    do {                                                                 // try_wait sleeps in hardware for a while before it gives up
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
EXB_D void bulk_copy_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(mbar) : "memory");
}

struct TPipe {                 // addresses in the shared window
    unsigned xs_base;          // 2 x CH doubles
    unsigned ctl;              // full[0], full[1] (mbarriers, 8 B each), out[0], out[1] (counters, 4 B each)
};

// Whole warp: bring chunk g of the CTA's sequence into buffer g & 1 (out of line: once per chunk and CTA).
template <int CH, int NB>
__device__ __noinline__ void t_issue_chunk(const TPipe tp, const double* x, long long incx, int nrows, int nchunks, unsigned g,
                                           int bulk_ok) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned b = g % (unsigned)NB;
    const int chunk = (int)(g % (unsigned)nchunks);
    const int r0 = chunk * CH;
    const int cnt = nrows - r0 < CH ? nrows - r0 : CH;
    const unsigned dst = tp.xs_base + b * (CH * 8u);
    const unsigned full = tp.ctl + 8u * b;
    const int nb = bulk_ok ? (cnt & ~1) : 0;                           // 16-byte granules through the TMA engine
    for (int k0 = nb + (int)lane; k0 < cnt; k0 += 256) {               // (strided / unaligned x, or the odd last row) eight loads in flight per lane
        double v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = k0 + 32 * i < cnt ? ldg64(x + incx * (long long)(r0 + k0 + 32 * i)) : 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (k0 + 32 * i < cnt) sts64(dst + 8u * (unsigned)(k0 + 32 * i), (unsigned long long)__double_as_longlong(v[i]));
    }
    __syncwarp();
    if (lane == 0u) {
        if (nb > 0) {
            mbar_expect_tx(full, (unsigned)nb * 8u);                   // the arrival + the bytes the copy will deliver
            bulk_copy_g2s(dst, x + r0, (unsigned)nb * 8u, full);
        } else {
            mbar_arrive(full);                                         // release: the warp's stores above are ordered before it
        }
    }
}

// One column (output) per warp: everything exgemv_t_win_kernel does for one set of columns.  Out of line ON PURPOSE [r2]:
// as a function of its own the column code gets its own register allocation -- the kernel's per-set bookkeeping stays
// out of it, and the streaming loop runs without a single local-memory access (inlined into the loop over the sets,
// ptxas spilled the round counter, the miss counters and the prefetch base of that loop).  The window travels by value.
// (One set per CTA with grid = number of sets, i.e. no loop at all, measured the same on narrow data.)
template <int U, int MAXT, int kGemvTChunk, int NB>
__device__ __noinline__ Window t_column(Window w, const GemvParams& prm, const int set, const unsigned gbase, const unsigned gtotal) {
    // (A variant in which lane l reads rows 4l .. 4l+3 with one 256-bit load was measured slower: its
    // x reads, 32 bytes apart per lane, conflict in shared memory, and the wider requests bought nothing.)
    // here prm.m = number of outputs (columns of A), prm.n = rows summed per output (< 2^31, the host checks),
    // prm.rs = lda, prm.cs = 1
    //
    // Register discipline [r2]: at 512 threads a thread has 128 registers, 32 of them hold A values in flight and ~45 the
    // window and its split.  Everything the streaming loop does not need every round lives elsewhere -- the deposit count
    // in w.st (above), the "skip the window" count in the warp's shared-memory slot, 32-bit round counters, the prefetch
    // address derived from the load pointer.
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;                                                   // <= MAXT
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    constexpr unsigned stride = 8u * MAXT;                                           // compile-time limb stride
    const unsigned col = smem_base + 8u * tid;
    constexpr unsigned kXsOff = stride * (unsigned)kLimbs;                           // NB x kGemvTChunk doubles
    constexpr unsigned kWlOff = kXsOff + 8u * NB * kGemvTChunk;                         // per warp: a 320-byte slot; its last word is the skip count
    constexpr unsigned kCtlOff = kWlOff + (MAXT / 32u) * 320u;                       // the pipeline's mbarriers and counters
    const unsigned wl_addr = smem_base + kWlOff + (tid >> 5) * 320u;
    TPipe tp;
    tp.xs_base = smem_base + kXsOff;
    tp.ctl = smem_base + kCtlOff;
    const int nrows = (int)prm.n;
    const int nwarps = (int)(T >> 5);
    const int nchunks = (nrows + kGemvTChunk - 1) / kGemvTChunk;
    const int rounds = nrows / (128 * U);                                            // rounds of U groups of 4 rows per lane, all rows valid
    constexpr int RPC = kGemvTChunk / (128 * U);                                     // rounds per chunk of x
    static_assert(RPC * 128 * U == kGemvTChunk && (RPC & (RPC - 1)) == 0, "a chunk of x holds a power of two of whole rounds");
    // chunk g: wait until it is resident; returns the shared address of this lane's x for the first round of the chunk
    auto acquire = [&](unsigned g) -> unsigned {
        mbar_wait(tp.ctl + 8u * (g % (unsigned)NB), (g / (unsigned)NB) & 1u);
        return tp.xs_base + (g % (unsigned)NB) * (kGemvTChunk * 8u) + 8u * lane;
    };
    // chunk g: this warp is done with it; the last warp out refills the buffer with chunk g + 2
    auto release = [&](unsigned g) {
        __syncwarp();
        unsigned old = 0u;
        asm volatile("fence.acq_rel.cta;" ::: "memory");
        if (lane == 0u) old = atomicAdd((unsigned*)((char*)smem + kCtlOff + 8u * NB + 4u * (g % (unsigned)NB)), 1u);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == (unsigned)nwarps - 1u) {
            if (lane == 0u) *(volatile unsigned*)((char*)smem + kCtlOff + 8u * NB + 4u * (g % (unsigned)NB)) = 0u;
            if (g + (unsigned)NB < gtotal) t_issue_chunk<kGemvTChunk, NB>(tp, prm.x, prm.incx, nrows, nchunks, g + (unsigned)NB, prm.x_vec_ok);
        }
    };

    const double* qa;                                                            // this lane's next group of 4 rows
    {
        const long long jraw = (long long)set * nwarps + (tid >> 5);
        const long long j = jraw < prm.m ? jraw : prm.m - 1;                     // spare warps redo the last column (they take part in the x pipeline)
        qa = prm.a + j * prm.rs + lane;
    }
    double va[U][4];
    auto load_group = [&](int u) {
        va[u][0] = ldg64(qa);
        va[u][1] = ldg64(qa + 32);
        va[u][2] = ldg64(qa + 64);
        va[u][3] = ldg64(qa + 96);
        qa += 128;
    };
    auto load_x = [&](unsigned xaddr, double& x0, double& x1, double& x2, double& x3) {
        x0 = lds_f64(xaddr); x1 = lds_f64(xaddr + 256u); x2 = lds_f64(xaddr + 512u); x3 = lds_f64(xaddr + 768u);
    };
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (rounds > 0) load_group(u);
    // L2 prefetch (TMA engine, UBLKPF): at the top of round r the load pointer of lane 0 stands at the first row of round
    // r + 1, so the 128 U rows that round r + D will read start (D - 1) rounds further on
    const int pf_dist = prm.l2_prefetch;
    auto prefetch_round = [&](int r) {
        if (pf_dist > 0 && lane == 0u && r + pf_dist < rounds) bulk_prefetch_l2(qa + (pf_dist - 1) * (128 * U), 128u * U * 8u);
    };
    // top of round r: every RPC rounds hand the finished chunk back and pick up the next one
    auto chunk_gate = [&](int r) -> unsigned {
        const unsigned g = gbase + (unsigned)(r / RPC);
        if (r > 0) release(g - 1u);
        return acquire(g);
    };
    int r = 0;
    unsigned xaddr = 0u;
    // ---- loop 1: register window; a warp whose groups keep missing leaves it (for this and the next 7 columns) ----
    const unsigned skip_window = (unsigned)lds64(wl_addr + 312u);
    if (__any_sync(0xffffffffu, skip_window > 0u)) {                             // (a vote: the branch is warp-uniform and the compiler knows it)
        __syncwarp();
        if (lane == 0u) sts64(wl_addr + 312u, (unsigned long long)(skip_window - 1u));
    } else {
        // leaky bucket, warp-uniform (the votes are): +2 per group that missed, -U per round; two rounds in a row
        // that miss throughout end the loop
        for (int score = 0; r < rounds && score <= U; ++r) {
            if ((r & (RPC - 1)) == 0) xaddr = chunk_gate(r);
            const bool has_next = r + 1 < rounds;
            prefetch_round(r);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double a0 = va[u][0], a1 = va[u][1], a2 = va[u][2], a3 = va[u][3];
                double x0, x1, x2, x3;
                load_x(xaddr, x0, x1, x2, x3);
                xaddr += 1024u;
                const double p0 = __dmul_rn(a0, x0), p1 = __dmul_rn(a1, x1), p2 = __dmul_rn(a2, x2), p3 = __dmul_rn(a3, x3);
                const unsigned k0 = ((unsigned)__double2hiint(p0) & 0x7fffffffu) - w.key0;
                const unsigned k1 = ((unsigned)__double2hiint(p1) & 0x7fffffffu) - w.key0;
                const unsigned k2 = ((unsigned)__double2hiint(p2) & 0x7fffffffu) - w.key0;
                const unsigned k3 = ((unsigned)__double2hiint(p3) & 0x7fffffffu) - w.key0;
                const bool mine = max(max(k0, k1), max(k2, k3)) < w.span;
                if (__all_sync(0xffffffffu, mine)) {
                    win_add_product(w, p0, __fma_rn(a0, x0, -p0));
                    win_add_product(w, p1, __fma_rn(a1, x1, -p1));
                    win_add_product(w, p2, __fma_rn(a2, x2, -p2));
                    win_add_product(w, p3, __fma_rn(a3, x3, -p3));
                    w.cnt += 4u;
                } else {
                    w = t_slow_group(w, col, stride, a0, a1, a2, a3, x0, x1, x2, x3, mine);
                    score += 2;
                }
                if (has_next) load_group(u);                                     // after the slot's values are dead (see exgemv_n_win_kernel)
            }
            score = max(score - U, 0);
            if (w.cnt > (unsigned)(kWinFlushEvery - 4 * U)) w = t_flush(w, col, stride);
        }
        if (r < rounds) {
            __syncwarp();
            if (lane == 0u) sts64(wl_addr + 312u, 7ull);
        }
    }
    // ---- loop 2: wide-range column, every product takes the ordinary path, inlined ----
    if (r < rounds) {
        unsigned status = w.st & 0xffffu;
        int since_norm = (int)(w.st >> 16);
        for (; r < rounds; ++r) {
            if ((r & (RPC - 1)) == 0) xaddr = chunk_gate(r);
            const bool has_next = r + 1 < rounds;
            prefetch_round(r);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double xa[4] = {va[u][0], va[u][1], va[u][2], va[u][3]};
                double xb[4];
                load_x(xaddr, xb[0], xb[1], xb[2], xb[3]);
                xaddr += 1024u;
                double none[1][expansions(0)];
                mul_add4<0, false, true, true>(col, stride, none, status, xa, xb);
                if (has_next) load_group(u);
            }
            since_norm += 8 * U;
            if (since_norm > (int)kTDepLimit - 8 * U) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
        w.st = status | ((unsigned)since_norm << 16);
    }
    // rows the full rounds do not cover (fewer than 128 U): they lie in ONE chunk, the last -- which is either the
    // chunk the last round was in or the one after it
    {
        const int r0 = rounds * U * 128;
        int cur = rounds > 0 ? (rounds - 1) / RPC : -1;                          // chunk this warp holds
        if (r0 < nrows) {
            const int chunk = r0 / kGemvTChunk;
            if (chunk > cur) {
                if (cur >= 0) release(gbase + (unsigned)cur);
                (void)acquire(gbase + (unsigned)chunk);
                cur = chunk;
            }
            unsigned status = w.st;                                              // (the count in bits 16.. rides along untouched)
            double none[1][expansions(0)];
            const unsigned xb = tp.xs_base + ((gbase + (unsigned)chunk) % (unsigned)NB) * (kGemvTChunk * 8u);
            const double* ra = qa;                                               // U * rounds groups have been loaded: this lane's row r0 + lane
            for (int rr = r0 + (int)lane; rr < nrows; rr += 32, ra += 32) {      // < 4 U rows per lane
                const double xv = lds_f64(xb + 8u * (unsigned)(rr - chunk * kGemvTChunk));
                mul_add1<0, false>(col, stride, none, status, ra[0], xv);
            }
            w.st = status;
        }
        release(gbase + (unsigned)cur);                                          // nrows >= 1: cur is the last chunk, nchunks - 1
    }
    // ---- end of the column: beta * y, drain, warp-sum the 32 private columns, normalise + round as a warp, store ----
    // (Round 1 / early round 2 handed the 39 summed limbs to lane 0, which added beta * y, normalised and rounded them
    // serially in LOCAL memory: ~15 % of all warp stall samples of the kernel sat in that code, waiting on local loads.
    // Now limb j of the column's sum lives in lane j -- WarpLimbs, as in the closing warp of the reduction kernels.)
    const long long jraw = (long long)set * nwarps + (tid >> 5);
    const bool valid = jraw < prm.m;
    double* const yp = prm.y + (valid ? jraw : 0) * prm.incy;
    if (lane == 0u && valid && prm.beta != 0.0) {                                    // exact beta * y_j joins lane 0's column
        unsigned status = w.st;
        const double yv = *yp;
        if (prm.beta == 1.0) {
            deposit(col, stride, yv, status);
        } else {
            double p, e;
            scale_exact(prm.beta, yv, p, e, status);
            deposit(col, stride, p, status);
            if (!(e != e)) deposit(col, stride, e, status);
        }
        w.st = status;
    }
    w = win_flush_products(w, col, stride);
    bound_column(col, stride);
    const unsigned st_all = __reduce_or_sync(0xffffffffu, w.st & 0xffffu);
    w.st = 0u;
    WarpLimbs xl;
    xl.a = 0;
    xl.b = 0;
#pragma unroll 1
    for (int jl = 0; jl < kLimbs; ++jl) {
        const unsigned addr = col + jl * stride;
        long long v = (long long)lds64(addr);
        if (__any_sync(0xffffffffu, v != 0)) {
            sts64(addr, 0ull);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // 32 x (2^52 + 2^11) < 2^58
            if ((int)lane == (jl & 31)) {
                if (jl < 32) xl.a = v; else xl.b = v;
            }
        }
    }
    const bool neg = warp_normalize(xl, lane);
    const double yv = warp_value(xl, neg, st_all, prm.round_mode, lane);
    if (lane == 0u && valid) {
        *yp = yv;
        if (st_all) atomicOr(&prm.ws->status, st_all);
    }
    __syncwarp();
    return w;
}

template <int U, int MAXT, int kGemvTChunk, int NB>
__global__ void __launch_bounds__(MAXT, 1) exgemv_t_win_kernel(const __grid_constant__ GemvParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;                                                   // <= MAXT
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    constexpr unsigned stride = 8u * MAXT;                                           // compile-time limb stride
    const unsigned col = smem_base + 8u * tid;
#pragma unroll
    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
    constexpr unsigned kXsOff = stride * (unsigned)kLimbs;                           // NB x kGemvTChunk doubles
    constexpr unsigned kWlOff = kXsOff + 8u * NB * kGemvTChunk;                         // per warp: a 320-byte slot; its last word is the skip count
    constexpr unsigned kCtlOff = kWlOff + (MAXT / 32u) * 320u;                       // the pipeline's mbarriers and counters
    const unsigned wl_addr = smem_base + kWlOff + (tid >> 5) * 320u;
    TPipe tp;
    tp.xs_base = smem_base + kXsOff;
    tp.ctl = smem_base + kCtlOff;

    const int nrows = (int)prm.n;
    const int nwarps = (int)(T >> 5);
    const int nsets = (int)((prm.m + nwarps - 1) / nwarps);
    const int nchunks = (nrows + kGemvTChunk - 1) / kGemvTChunk;
    const int rounds = nrows / (128 * U);                                            // rounds of U groups of 4 rows per lane, all rows valid
    constexpr int RPC = kGemvTChunk / (128 * U);                                     // rounds per chunk of x
    static_assert(RPC * 128 * U == kGemvTChunk && (RPC & (RPC - 1)) == 0, "a chunk of x holds a power of two of whole rounds");
    const int my_sets = (int)blockIdx.x < nsets ? (nsets - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const unsigned gtotal = (unsigned)my_sets * (unsigned)nchunks;                   // chunks this CTA will consume

    if (lane == 0u) sts64(wl_addr + 312u, 0ull);
    if (tid == 0u) {
        for (unsigned b = 0; b < (unsigned)NB; ++b) {
            mbar_init(tp.ctl + 8u * b, 1u);
            *(volatile unsigned*)((char*)smem + kCtlOff + 8u * NB + 4u * b) = 0u;    // the buffer's "warps out" counter
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32u) {                                                                 // warp 0 starts the pipeline
        for (unsigned g = 0; g < (unsigned)NB && g < gtotal; ++g)
            t_issue_chunk<kGemvTChunk, NB>(tp, prm.x, prm.incx, nrows, nchunks, g, prm.x_vec_ok);
    }
    Window w;
    win_reset(w);
    unsigned gbase = 0u;                                                             // sequence number of chunk 0 of the current set
    for (int set = (int)blockIdx.x; set < nsets; set += (int)gridDim.x, gbase += (unsigned)nchunks)
        w = t_column<U, MAXT, kGemvTChunk, NB>(w, prm, set, gbase, gtotal);
}

// Closing kernel of the 'N' paths: y_r = round(sum over the parts of row r's limbs + beta * y_r), 32 rows per CTA.
// Warp w sums the per-part limbs 10 w .. 10 w + 9 of the CTA's rows (lane = row: coalesced, scratch is
// [part][limb][row]; four parts in flight) into a row of shared memory, warp 0 then adds beta * y exactly, and each warp
// normalises and rounds eight of the rows AS A WARP (limb j in lane j: warp_normalize / warp_value).  Round 1 did all
// of it per thread on a 39-limb array in local memory: 56 us for 32768 rows, 4 % of the whole ExGEMV.
constexpr int kFinishT = 128;
__global__ void __launch_bounds__(kFinishT) exgemv_finish_kernel(const GemvParams prm) {
    __shared__ long long rows[32][kLimbs + 2];                 // 41 words per row: conflict-free 64-bit accesses both ways
    __shared__ unsigned row_st[32];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const long long row = (long long)blockIdx.x * 32 + lane;
    const bool valid = row < prm.m;
    const int j0 = 10 * (int)warp, j1 = j0 + 10 < kLimbs ? j0 + 10 : kLimbs;
    for (int j = j0; j < j1; ++j) {
        long long s = 0;
        if (valid) {
            const long long* q = prm.scratch + (long long)j * prm.m + row;
            const long long pstep = (long long)kLimbs * prm.m;
            int p = 0;
            for (; p + 4 <= prm.parts; p += 4, q += 4 * pstep) {
                const long long v0 = q[0], v1 = q[pstep], v2 = q[2 * pstep], v3 = q[3 * pstep];
                s += (v0 + v1) + (v2 + v3);                     // parts <= 1024 bounded limbs (< 2^52 + 2^11 each): no overflow
            }
            for (; p < prm.parts; ++p, q += pstep) s += q[0];
        }
        rows[lane][j] = s;
    }
    __syncthreads();
    if (warp == 0u) {
        unsigned st = 0;
        if (valid) {
            for (int p = 0; p < prm.parts; ++p) st |= prm.row_status[(long long)p * prm.m + row];
            if (prm.beta != 0.0) {
                const double yv = prm.y[row * prm.incy];
                if (prm.beta == 1.0) {
                    st |= accumulate_double(rows[lane], yv);
                } else {
                    double p, e;
                    scale_exact(prm.beta, yv, p, e, st);
                    st |= accumulate_double(rows[lane], p);
                    if (!(e != e)) st |= accumulate_double(rows[lane], e);
                }
            }
            if (st) atomicOr(&prm.ws->status, st);
        }
        row_st[lane] = st;
    }
    __syncthreads();
#pragma unroll 1
    for (unsigned k = 8u * warp; k < 8u * warp + 8u; ++k) {
        WarpLimbs x;
        x.a = rows[k][lane];
        x.b = lane < 7u ? rows[k][32 + lane] : 0ll;
        const bool neg = warp_normalize(x, lane);
        const double v = warp_value(x, neg, row_st[k], prm.round_mode, lane);
        const long long r = (long long)blockIdx.x * 32 + k;
        if (lane == 0u && r < prm.m) prm.y[r * prm.incy] = v;
    }
}

}  // namespace exb
