// segments_kernel.cuh -- batched ("segmented") ExSUM / ExDOT: many independent exact reductions in ONE
// launch.
//
// The reference's callers outside its tests are applications that reduce MANY SHORT vectors: k-means
// distance sums, SpMV rows, MRI gridding bins.  They gather each vector on the host and call exsum()
// once per vector (src/cpu/examples/kmeans/kmeans_clustering.cpp:213, spmv/main.cpp:85,
// mri-gridding/CPU_kernels.cpp:293) -- one OpenCL context, JIT build and two launches per call in the
// GPU build (src/gpu/blas/blas1/ExSUM.cpp:86-209).  Here segment s = [seg[s], seg[s+1]) of the
// input is reduced by one WARP:
//   * lanes stride over the segment with coalesced 8-byte loads (no alignment requirement), two groups
//     of 4 elements per lane in flight;
//   * summands inside the lane's register window (window.cuh) are accumulated there, everything else
//     is deposited into the lane's private shared-memory superaccumulator column;
//   * at the end of the segment the warp drains the windows, sums its 32 columns limb by limb with
//     shuffles (limbs zero in every lane are skipped after one vote), and lane 0 rounds and stores
//     results[s] (+ the segment's status flags).
// ExDOT segments optionally read their second operand through an index array, b[gather[i]], which
// makes a CSR sparse matrix-vector product y = A x one call: a = values, gather = column indices,
// seg = row pointers, b = x.
#pragma once
#include "gemv_kernel.cuh"

namespace exb {

struct SegParams {
    const double* a;
    const double* b;              // nullptr for sums
    const int* gather;            // optional: second operand is b[gather[i]]
    const long long* seg;         // nseg + 1 non-decreasing offsets
    long long nseg;
    double* results;
    unsigned* statuses;           // per segment; may be nullptr
    Workspace* ws;
    int round_mode;
};

// lane 0 at the end of a segment: summed limbs -> rounded result (+ status)
__device__ __noinline__ void segment_store(const long long* wl, unsigned st, double* out, unsigned* st_out, int round_mode,
                                           unsigned* ws_status) {
    long long acc[kLimbs];
    for (int j = 0; j < kLimbs; ++j) acc[j] = wl[j];
    *out = finalize_value(acc, st, round_mode);
    if (st_out) *st_out = st;
    if (st) atomicOr(ws_status, st);
}

template <bool DOT, bool GATHER, int MAXT>
__global__ void __launch_bounds__(MAXT, 2) exblas_segments_kernel(const SegParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;                                    // <= MAXT
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarps = T >> 5;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    constexpr unsigned stride = 8u * MAXT;
    const unsigned col = smem_base + 8u * tid;
#pragma unroll
    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);
    long long* wl = smem + (size_t)kLimbs * MAXT + (size_t)warp * 40;

    Window w;
    win_reset(w);
    const long long wstep = (long long)gridDim.x * nwarps;
    for (long long s = (long long)blockIdx.x * nwarps + warp; s < prm.nseg; s += wstep) {
        const long long lo = prm.seg[s], hi = prm.seg[s + 1];
        const long long n = hi > lo ? hi - lo : 0;
        const long long full = n / 128;                               // groups in which every lane has 4 elements
        const double* pa = prm.a + lo + lane;
        const double* pb = DOT && !GATHER ? prm.b + lo + lane : nullptr;
        const int* pg = GATHER ? prm.gather + lo + lane : nullptr;
        double na[4], nb[4];                                          // the group loaded ahead
        auto load_group = [&]() {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                na[k] = ldg64(pa + 32 * k);
                if (DOT) nb[k] = GATHER ? __ldg(prm.b + __ldg(pg + 32 * k)) : ldg64(pb + 32 * k);
            }
            pa += 128;
            if (DOT && !GATHER) pb += 128;
            if (GATHER) pg += 128;
        };
        if (full > 0) load_group();
        int since_norm = 0;
        for (long long g = 0; g < full; ++g) {
            const double a0 = na[0], a1 = na[1], a2 = na[2], a3 = na[3];
            const double b0 = DOT ? nb[0] : 0.0, b1 = DOT ? nb[1] : 0.0, b2 = DOT ? nb[2] : 0.0, b3 = DOT ? nb[3] : 0.0;
            if (g + 1 < full) load_group();
            if (DOT) {
                const double p0 = __dmul_rn(a0, b0), p1 = __dmul_rn(a1, b1), p2 = __dmul_rn(a2, b2), p3 = __dmul_rn(a3, b3);
                const unsigned k0 = ((unsigned)__double2hiint(p0) & 0x7fffffffu) - w.key0;
                const unsigned k1 = ((unsigned)__double2hiint(p1) & 0x7fffffffu) - w.key0;
                const unsigned k2 = ((unsigned)__double2hiint(p2) & 0x7fffffffu) - w.key0;
                const unsigned k3 = ((unsigned)__double2hiint(p3) & 0x7fffffffu) - w.key0;
                const bool mine = max(max(k0, k1), max(k2, k3)) < w.span;
                if (__all_sync(0xffffffffu, mine)) {
                    win_add_product(w, p0, __fma_rn(a0, b0, -p0));
                    win_add_product(w, p1, __fma_rn(a1, b1, -p1));
                    win_add_product(w, p2, __fma_rn(a2, b2, -p2));
                    win_add_product(w, p3, __fma_rn(a3, b3, -p3));
                    w.cnt += 4u;
                } else {
                    w = prod_slow_group(w, col, stride, a0, a1, a2, a3, b0, b1, b2, b3, mine, true);
                    since_norm += 12;
                }
                if (w.cnt > (unsigned)(kWinFlushEvery - 4)) {
                    w = win_flush_products(w, col, stride);
                    since_norm += 4;
                }
            } else {
                const unsigned key = w.key0 - (1u << 20);            // single summands: exponent in [b, b + 50]
                const unsigned span = w.span ? kWinSpanSum : 0u;
                const unsigned k0 = ((unsigned)__double2hiint(a0) & 0x7fffffffu) - key;
                const unsigned k1 = ((unsigned)__double2hiint(a1) & 0x7fffffffu) - key;
                const unsigned k2 = ((unsigned)__double2hiint(a2) & 0x7fffffffu) - key;
                const unsigned k3 = ((unsigned)__double2hiint(a3) & 0x7fffffffu) - key;
                const bool mine = max(max(k0, k1), max(k2, k3)) < span;
                if (__all_sync(0xffffffffu, mine)) {
                    win_add_single(w, a0);
                    win_add_single(w, a1);
                    win_add_single(w, a2);
                    win_add_single(w, a3);
                    w.cnt += 4u;
                } else {
                    w = sum_slow_group(w, col, stride, a0, a1, a2, a3, mine);
                    since_norm += 8;
                }
                if (w.cnt > (unsigned)(kWinFlushEvery - 4)) {
                    w = win_flush_singles(w, col, stride);
                    since_norm += 4;
                }
            }
            if (since_norm > kMaxDepositsPerNormalize - 32) {
                bound_column(col, stride);
                since_norm = 0;
            }
        }
        // the last (< 128) elements of the segment: at most 4 per lane, ordinary path
        {
            unsigned status = w.st;
            double none[1][expansions(0)];
            for (long long i = lo + full * 128 + lane; i < hi; i += 32) {
                if (DOT) mul_add1<0, false>(col, stride, none, status, prm.a[i], GATHER ? prm.b[prm.gather[i]] : prm.b[i]);
                else add1<0, false>(col, stride, none, status, prm.a[i]);
            }
            w.st = status;
        }
        // ---- end of the segment: drain, warp-sum the 32 private columns, round, store ----
        w = DOT ? win_flush_products(w, col, stride) : win_flush_singles(w, col, stride);
        // each limb of a column has taken at most since_norm + 12 digits (<= 8 from the tail, 4 from the drain)
        // since it was last bounded or zeroed; up to 30 digits per lane the sum over 32 lanes stays below 2^62
        if (since_norm + 12 > 30) bound_column(col, stride);
        const unsigned st_all = __reduce_or_sync(0xffffffffu, w.st);
        w.st = 0u;
        for (int jl = 0; jl < kLimbs; ++jl) {
            const unsigned addr = col + jl * stride;
            long long v = (long long)lds64(addr);
            if (__any_sync(0xffffffffu, v != 0)) {
                sts64(addr, 0ull);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            }
            if (lane == 0) wl[jl] = v;
        }
        __syncwarp();
        if (lane == 0)
            segment_store(wl, st_all, prm.results + s, prm.statuses ? prm.statuses + s : nullptr, prm.round_mode, &prm.ws->status);
        __syncwarp();
    }
}

}  // namespace exb
