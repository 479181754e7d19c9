// segments_kernel.cuh -- batched ("segmented") ExSUM / ExDOT: many independent exact reductions in ONE
// launch.
//
// The reference's callers outside its tests are applications that reduce MANY SHORT vectors: k-means
// distance sums, SpMV rows, MRI gridding bins.  They gather each vector on the host and call exsum()
// once per vector (src/cpu/examples/kmeans/kmeans_clustering.cpp:213, spmv/main.cpp:85,
// mri-gridding/CPU_kernels.cpp:293) -- one OpenCL context, JIT build and two launches per call in the
// GPU build (src/gpu/blas/blas1/ExSUM.cpp:86-209).  Here results[s] = exact sum (dot product) of
// segment s = [seg[s], seg[s+1]) for all segments at once.
//
// Work decomposition (round 2; round 1 gave every segment to one warp, which left 2^20-element segments at
// 0.26 TB/s and 16-element segments at 0.3 G segments/s): the ELEMENT range [seg[0], seg[nseg]) is cut into one
// contiguous, equal-sized range per warp of the grid -- whatever the segment lengths are, every warp streams the
// same number of bytes.  A warp finds the segment its range starts in by a 32-way search and walks the segments
// that overlap its range:
//   * SHORT segments (<= kSegShort elements) that lie inside the range are taken 32 at a time, ONE LANE PER
//     SEGMENT: every lane deposits its own segment into its private shared-memory superaccumulator column
//     (superacc.cuh) and records the exponent range it has touched; the 32 columns are then finished one after
//     the other by the whole warp (limb j in lane j: warp-parallel normalise + round, reduce_kernel.cuh);
//   * longer segments are streamed by the whole warp (lanes stride over the elements with coalesced 8-byte
//     loads; register window (window.cuh) + private column per lane), the 32 columns are summed row by row
//     (split low / carry parts, REDUX) and finished the same way;
//   * a segment that crosses range boundaries is reduced piecewise: every warp that holds a piece adds its
//     (bounded) limbs to the scratch accumulator of the range in which the segment STARTS -- at most one
//     segment per range can cross its end -- with native 64-bit REDs, and the last piece to arrive (ticket
//     counter with acquire / release) normalises, rounds, stores and clears the scratch.  Integer limb sums
//     commute, so the result is the same bits whatever the decomposition.
// ExDOT segments optionally read their second operand through an index array, b[gather[i]], which
// makes a CSR sparse matrix-vector product y = A x one call: a = values, gather = column indices,
// seg = row pointers, b = x.
#pragma once
#include "gemv_kernel.cuh"

namespace exb {

constexpr int kSegShort = 64;             // segments up to this length take the lane-per-segment path

struct SegScratch {                       // one per warp range: the segment that starts in it and crosses its end
    unsigned long long limbs[2][kLimbs];  // pieces of even / odd warps (two accumulators: see seg_contribute)
    unsigned status;
    unsigned count;                       // pieces that have arrived
};

struct SegParams {
    const double* a;
    const double* b;              // nullptr for sums
    const int* gather;            // optional: second operand is b[gather[i]]
    const long long* seg;         // nseg + 1 non-decreasing offsets
    long long nseg;
    double* results;
    unsigned* statuses;           // per segment; may be nullptr
    Workspace* ws;
    SegScratch* scratch;          // [gridDim.x * warps per CTA], all zero between launches
    int round_mode;
};

// normalise, round and store one finished segment (limb j of its exact sum in lane j)
EXB_D void seg_finish(const SegParams& prm, WarpLimbs x, unsigned st, long long s, unsigned lane) {
    const bool neg = warp_normalize(x, lane);
    const double v = warp_value(x, neg, st, prm.round_mode, lane);
    if (lane == 0) {
        prm.results[s] = v;
        if (prm.statuses) prm.statuses[s] = st;
        if (st) atomicOr(&prm.ws->status, st);
    }
}

// one piece of a segment that crosses range boundaries: add to the owner range's scratch; the last piece finishes.
// One local carry split bounds every limb of a piece below 2^52 + 2^11 in magnitude -- and that bound IS reached by
// ordinary signed data (a piece with a negative sum leaves 2^52 - k in the limb above its top digit) -- so at most 2047
// pieces may meet in one 64-bit scratch limb.  Pieces of even and odd warps therefore use separate accumulators (the
// host launches fewer than 4094 warps), which the last piece normalises separately before adding them.
EXB_D void seg_contribute(const SegParams& prm, WarpLimbs x, unsigned st, long long s, long long owner, unsigned expected,
                          unsigned which, unsigned lane) {
    WarpLimbs c, u;
    c.a = x.a >> kDigits;
    x.a &= kLimbMask;
    c.b = lane < 6u ? (x.b >> kDigits) : 0;
    if (lane < 6u) x.b &= kLimbMask;
    u = wl_shift_up(c, lane);
    x.a += u.a;
    if (lane < 7u) x.b += u.b;
    SegScratch* sc = prm.scratch + owner;
    unsigned long long* dst = sc->limbs[which & 1u];
    if (x.a != 0) atomicAdd(&dst[lane], (unsigned long long)x.a);
    if (lane < 7u && x.b != 0) atomicAdd(&dst[32 + lane], (unsigned long long)x.b);
    if (lane == 0 && st) atomicOr(&sc->status, st);
    __syncwarp();
    unsigned ticket = 0;
    if (lane == 0) ticket = ticket_acq_rel(&sc->count);
    ticket = __shfl_sync(kFullWarp, ticket, 0);
    if (ticket != expected - 1u) return;
    WarpLimbs y;
    x.a = (long long)atomicExch(&sc->limbs[0][lane], 0ull);
    x.b = lane < 7u ? (long long)atomicExch(&sc->limbs[0][32 + lane], 0ull) : 0ll;
    y.a = (long long)atomicExch(&sc->limbs[1][lane], 0ull);
    y.b = lane < 7u ? (long long)atomicExch(&sc->limbs[1][32 + lane], 0ull) : 0ll;
    unsigned stt = 0;
    if (lane == 0) {
        stt = atomicExch(&sc->status, 0u);
        sc->count = 0u;
    }
    stt = __shfl_sync(kFullWarp, stt, 0);
    warp_normalize(x, lane);
    warp_normalize(y, lane);
    x.a += y.a;
    x.b += y.b;
    seg_finish(prm, x, stt, s, lane);
}

// one summand / product of lane-per-segment mode into the lane's own column (+ the touched-row range)
template <bool DOT>
EXB_D void seg_lane_element(unsigned col, unsigned stride, RowRange& rr, unsigned& status, double av, double bv) {
    double none[1][expansions(0)];
    if (DOT) {
        rr_note_product(rr, av, bv);
        mul_add1<0, false>(col, stride, none, status, av, bv);
    } else {
        rr_note(rr, av);
        deposit(col, stride, av, status);
    }
}

template <bool DOT, bool GATHER, int MAXT>
__global__ void __launch_bounds__(MAXT, 2) exblas_segments_kernel(const __grid_constant__ SegParams prm) {
    extern __shared__ long long smem[];
    const unsigned T = blockDim.x;                                    // <= MAXT
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, nwarps = T >> 5;
    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem);
    constexpr unsigned stride = 8u * MAXT;
    const unsigned col = smem_base + 8u * tid;
    const unsigned col0 = smem_base + 8u * (tid - lane);               // column of lane 0 of this warp
#pragma unroll
    for (int j = 0; j < kLimbs; ++j) sts64(col + j * stride, 0ull);

    const long long nseg = prm.nseg;
    const long long N0 = prm.seg[0], N1 = prm.seg[nseg];
    const long long N = N1 - N0;
    const long long gw = (long long)blockIdx.x * nwarps + warp;        // this warp's index in the grid
    const long long nw = (long long)gridDim.x * nwarps;
    if (N <= 0) {                                                     // nothing but empty segments
        for (long long s = gw * 32 + lane; s < nseg; s += nw * 32) {
            prm.results[s] = 0.0;
            if (prm.statuses) prm.statuses[s] = 0u;
        }
        return;
    }
    // equal element ranges, multiples of 32 elements
    long long R = (N + nw - 1) / nw;
    R = (R + 31) & ~31ll;
    const long long start = N0 + gw * R;
    if (start >= N1) return;                                          // more warps than work
    const bool last_range = start + R >= N1;
    const long long end = last_range ? N1 : start + R;

    // first segment of this range: the smallest s with pred(s) = seg[s+1] > start || seg[s] >= start.  pred is monotone
    // (false ... false true ... true) and pred(nseg - 1) holds (seg[nseg] = N1 > start): a 32-way search, one probe per lane.
    long long s_lo = 0, s_hi = nseg - 1;                              // answer in [s_lo, s_hi]; everything below s_lo is false
    while (s_lo < s_hi) {
        const long long step = (s_hi - s_lo + 31) / 32;               // >= 1: probes s_lo, s_lo + step, ... reach s_hi
        const long long probe = s_lo + (long long)lane * step;
        bool pred = true;                                             // probes at or beyond s_hi are true
        if (probe < s_hi) pred = (prm.seg[probe + 1] > start) || (prm.seg[probe] >= start);
        const unsigned m = __ballot_sync(kFullWarp, pred);
        const int f = m ? __ffs((int)m) - 1 : 32;                     // first true probe
        if (f == 0) {
            s_hi = s_lo;
            break;
        }
        const long long first_true = f < 32 ? s_lo + (long long)f * step : s_hi;
        s_lo = s_lo + (long long)(f - 1) * step + 1;                  // one above the last false probe
        if (first_true < s_hi) s_hi = first_true;
    }
    long long s = s_lo;

    Window w;
    win_reset(w);
    int skip_window = 0;                                              // pieces left to run without the register window
    // The segment bounds are read 32 at a time: lane k keeps seg[cbase + k] and seg[cbase + k + 1], and the bounds of the
    // segment at hand come out of that cache by shuffle -- read one by one they were two or three DEPENDENT global loads
    // per segment, most of the time of segments of a few hundred elements.
    long long cl0 = 0, cl1 = 0, cbase = -64;
    auto load_bounds = [&](long long at) {
        const long long i0 = at + lane < nseg ? at + lane : nseg, i1 = at + lane + 1 < nseg ? at + lane + 1 : nseg;
        cl0 = prm.seg[i0];
        cl1 = prm.seg[i1];
        cbase = at;
    };
    while (s < nseg) {
        if (s - cbase >= 32) load_bounds(s);
        const long long b0 = __shfl_sync(kFullWarp, cl0, (int)(s - cbase));
        if (b0 >= end && !(last_range && b0 == N1)) break;            // starts in a later range
        // ---------------- lane-per-segment mode: up to 32 short segments that lie inside this range ----------------
        if (__shfl_sync(kFullWarp, cl1, (int)(s - cbase)) - b0 <= kSegShort) {
            if (s != cbase) load_bounds(s);                           // a full batch of 32 candidates
            const long long sl = s + lane;
            const long long l0 = cl0, l1 = cl1;
            bool ok = false;
            if (sl < nseg) ok = l0 >= start && l1 <= end && l1 - l0 <= kSegShort && (l0 < end || (last_range && l0 == N1));
            const unsigned okm = __ballot_sync(kFullWarp, ok);
            const int batch = (okm == 0xffffffffu) ? 32 : __ffs((int)~okm) - 1;   // leading run of short, complete segments
            if (batch >= 4) {
                unsigned status = 0;
                RowRange rr = rr_empty();
                if ((int)lane < batch) {
                    for (long long i = l0; i < l1; ++i) {
                        const double av = prm.a[i];
                        const double bv = DOT ? (GATHER ? prm.b[prm.gather[i]] : prm.b[i]) : 0.0;
                        seg_lane_element<DOT>(col, stride, rr, status, av, bv);
                    }
                }
                unsigned my_lo, my_hi;
                rr_rows(rr, 0u, my_lo, my_hi);
                for (int q = 0; q < batch; ++q) {                     // finish column q: limb j in lane j
                    const unsigned rlo = __shfl_sync(kFullWarp, my_lo, q), rhi = __shfl_sync(kFullWarp, my_hi, q);
                    const unsigned st = __shfl_sync(kFullWarp, status, q);
                    const unsigned cq = col0 + 8u * (unsigned)q;
                    WarpLimbs x;
                    x.a = 0;
                    x.b = 0;
                    if (lane >= rlo && lane <= rhi) {
                        x.a = (long long)lds64(cq + lane * stride);
                        sts64(cq + lane * stride, 0ull);
                    }
                    if (lane < 7u && 32u + lane >= rlo && 32u + lane <= rhi) {
                        x.b = (long long)lds64(cq + (32u + lane) * stride);
                        sts64(cq + (32u + lane) * stride, 0ull);
                    }
                    seg_finish(prm, x, st, s + q, lane);
                }
                s += batch;
                continue;
            }
        }
        // ---------------- warp mode: this warp's piece [lo, hi) of segment s ----------------
        const long long b1 = __shfl_sync(kFullWarp, cl1, (int)(s - cbase));
        const long long lo = b0 > start ? b0 : start, hi = b1 < end ? b1 : end;
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
        // Register-window mode until two groups in a row miss (wide-range data); then plain inlined deposits for the rest of
        // this piece and for the next 15 pieces of this warp (the out-of-line miss path costs ~4x a direct deposit).
        bool windowed = skip_window == 0;
        if (!windowed) --skip_window;
        int bad = 0;
        for (long long g = 0; g < full; ++g) {
            const double a0 = na[0], a1 = na[1], a2 = na[2], a3 = na[3];
            const double b0v = DOT ? nb[0] : 0.0, b1v = DOT ? nb[1] : 0.0, b2v = DOT ? nb[2] : 0.0, b3v = DOT ? nb[3] : 0.0;
            if (g + 1 < full) load_group();
            if (!windowed) {
                unsigned status = w.st;
                if (DOT) {
                    const double xa[4] = {a0, a1, a2, a3}, xb[4] = {b0v, b1v, b2v, b3v};
                    double none[1][expansions(0)];
                    mul_add4<0, false, true>(col, stride, none, status, xa, xb);
                    since_norm += 8;
                } else {
                    deposit4<false>(col, stride, a0, a1, a2, a3, status);
                    since_norm += 4;
                }
                w.st = status;
            } else if (DOT) {
                const double p0 = __dmul_rn(a0, b0v), p1 = __dmul_rn(a1, b1v), p2 = __dmul_rn(a2, b2v), p3 = __dmul_rn(a3, b3v);
                const unsigned k0 = ((unsigned)__double2hiint(p0) & 0x7fffffffu) - w.key0;
                const unsigned k1 = ((unsigned)__double2hiint(p1) & 0x7fffffffu) - w.key0;
                const unsigned k2 = ((unsigned)__double2hiint(p2) & 0x7fffffffu) - w.key0;
                const unsigned k3 = ((unsigned)__double2hiint(p3) & 0x7fffffffu) - w.key0;
                const bool mine = max(max(k0, k1), max(k2, k3)) < w.span;
                if (__all_sync(0xffffffffu, mine)) {
                    win_add_product(w, p0, __fma_rn(a0, b0v, -p0));
                    win_add_product(w, p1, __fma_rn(a1, b1v, -p1));
                    win_add_product(w, p2, __fma_rn(a2, b2v, -p2));
                    win_add_product(w, p3, __fma_rn(a3, b3v, -p3));
                    w.cnt += 4u;
                    bad = 0;
                } else {
                    w = prod_slow_group(w, col, stride, a0, a1, a2, a3, b0v, b1v, b2v, b3v, mine, true);
                    since_norm += 12;
                    if (++bad >= 2) {
                        windowed = false;
                        skip_window = 15;
                    }
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
                    bad = 0;
                } else {
                    w = sum_slow_group(w, col, stride, a0, a1, a2, a3, mine);
                    since_norm += 8;
                    if (++bad >= 2) {
                        windowed = false;
                        skip_window = 15;
                    }
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
        // the last (< 128) elements of the piece: at most 4 per lane, ordinary path
        {
            unsigned status = w.st;
            double none[1][expansions(0)];
            for (long long i = lo + full * 128 + lane; i < hi; i += 32) {
                if (DOT) mul_add1<0, false>(col, stride, none, status, prm.a[i], GATHER ? prm.b[prm.gather[i]] : prm.b[i]);
                else add1<0, false>(col, stride, none, status, prm.a[i]);
            }
            w.st = status;
        }
        // ---- end of the piece: drain, sum the 32 private columns row by row (split low / carry parts: the columns need no
        // bounding first), limb j of the sum in lane j ----
        w = DOT ? win_flush_products(w, col, stride) : win_flush_singles(w, col, stride);
        const unsigned st_all = __reduce_or_sync(0xffffffffu, w.st);
        w.st = 0u;
        WarpLimbs x;
        x.a = 0;
        x.b = 0;
        // which rows has any lane touched?  13 independent loads per step and ONE warp reduction of the 13 flags (a row at a
        // time this scan was a chain of 39 dependent load -> vote -> branch steps, ~2000 cycles per piece: for segments of a
        // few hundred elements most of the time)
        unsigned long long touched = 0ull;
#pragma unroll
        for (int b3 = 0; b3 < 3; ++b3) {
            unsigned m = 0u;
#pragma unroll
            for (int q = 0; q < 13; ++q) m |= lds64(col + (unsigned)(13 * b3 + q) * stride) != 0ull ? (1u << q) : 0u;
            touched |= (unsigned long long)__reduce_or_sync(kFullWarp, m) << (13 * b3);
        }
        auto place = [&](int j, long long v) {                        // limb j of the sum lives in lane j (.a) / lane j - 32 (.b)
            if (j < 32) {
                if ((int)lane == j) x.a = v;
            } else if ((int)lane == j - 32) {
                x.b = v;
            }
        };
        long long carry_prev = 0;                                     // carry-save part of row `prev`: belongs to row prev + 1
        int prev = -2;
        while (touched) {
            const int jl = __ffsll((long long)touched) - 1;
            touched &= touched - 1ull;
            if (carry_prev != 0 && prev + 1 != jl) place(prev + 1, carry_prev);   // an untouched row receives the carry alone
            const unsigned addr = col + (unsigned)jl * stride;
            const long long v = (long long)lds64(addr);
            sts64(addr, 0ull);
            const unsigned long long vlo = (unsigned long long)(v & kLimbMask);
            const unsigned r0 = __reduce_add_sync(kFullWarp, (unsigned)vlo & 0x3ffffu);            // 3 x 18 bits: 32 lane sums < 2^23
            const unsigned r1 = __reduce_add_sync(kFullWarp, (unsigned)(vlo >> 18) & 0x3ffffu);
            const unsigned r2 = __reduce_add_sync(kFullWarp, (unsigned)(vlo >> 36));
            const int rh = __reduce_add_sync(kFullWarp, (int)(v >> kDigits));
            long long rowsum = (long long)((unsigned long long)r0 + ((unsigned long long)r1 << 18) + ((unsigned long long)r2 << 36));
            if (prev + 1 == jl) rowsum += carry_prev;
            if (jl == kLimbs - 1) {
                rowsum += ((long long)rh) << kDigits;                 // the top limb keeps its own carry-save bits
                carry_prev = 0;
            } else {
                carry_prev = rh;
            }
            place(jl, rowsum);
            prev = jl;
        }
        if (carry_prev != 0) place(prev + 1, carry_prev);             // (prev < 38 here: the top row leaves no carry)
        const bool complete = b0 >= start && b1 <= end;
        if (complete) {
            seg_finish(prm, x, st_all, s, lane);
        } else {
            const long long owner = (b0 - N0) / R;
            const unsigned expected = (unsigned)((b1 - 1 - N0) / R - owner + 1);
            seg_contribute(prm, x, st_all, s, owner, expected, (unsigned)(gw & 1), lane);
        }
        if (b1 > end) break;                                          // the segment continues in the next range: done here
        ++s;
    }
}

}  // namespace exb
