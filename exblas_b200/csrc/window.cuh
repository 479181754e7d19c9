// window.cuh -- register-resident window of the superaccumulator (superaccumulator-only mode, F == 0).
//
// The reference's "superacc only" kernels (ExSUM.Superacc.cl:212-294, ExDOT.Superacc.cl:218-320,
// ExGEMV.Superacc.cl:192-290) send every summand to the 39-limb accumulator in local memory.  On
// B200 the shared-memory read-modify-write of a deposit costs ~25 issue slots, which makes ExDOT and
// ExGEMV (two deposits per element) issue-bound well below the HBM rate.  Most real vectors are
// narrow: all their products sit within a few dozen binades of each other.  For those a thread keeps
// a WINDOW of three consecutive 52-bit digits of the accumulator in registers, anchored at a
// per-thread exponent `b`:
//
//      digit k (k = 0, 1, 2) counts units of u_k = 2^(b - 52 k)
//
// A product p (+ its TwoProd error e, |e| <= ulp(p)/2) whose exponent lies in [b+1, b+50] is split
// with magic-constant adds, all in the FP64 pipe:
//      t1 = p  + M0,  x1 = t1 - M0,  r1 = p - x1        M_k = 1.5 * 2^52 * u_k
//      t2 = r1 + M1                                      (r1 is a multiple of ulp(p) >= 2 u_1: no remainder)
//      t3 = e  + M1,  x3 = t3 - M1,  r3 = e - x3
//      t4 = r3 + M2                                      (e is a multiple of 2^(exp(p)-105) >= u_2: no remainder)
// and because t_i and M_k share a binade, bits(t_i) - bits(M_k) IS the signed digit.  The thread
// adds the raw bit patterns into three 64-bit registers and subtracts count * bits(M_k) when the
// window is flushed (every <= 1024 elements, so no digit sum can leave 63 bits): 10 FP64 + 8 integer
// instructions per product instead of two shared-memory deposits.  Summands outside the window take
// the ordinary deposit; the window is flushed into the same thread-private column as four ordinary
// deposits, so the result stays the exact sum whatever mix of paths was taken.
//
// Everything here is plain IEEE double arithmetic and integer bit casts, host + device, so the host
// unit test (tests/test_window_host.py, which compiles tests/window_host_check.cpp with g++) exercises the very same code.
#pragma once
#include "superacc.cuh"

namespace exb {

#if defined(__CUDA_ARCH__)
#define EXB_ADD(a, b) __dadd_rn((a), (b))
#define EXB_SUB(a, b) __dsub_rn((a), (b))
EXB_HD unsigned long long win_bits(double x) { return (unsigned long long)__double_as_longlong(x); }
EXB_HD double win_from_bits(unsigned long long u) { return __longlong_as_double((long long)u); }
#else
#define EXB_ADD(a, b) ((a) + (b))
#define EXB_SUB(a, b) ((a) - (b))
EXB_HD unsigned long long win_bits(double x) { unsigned long long u; std::memcpy(&u, &x, 8); return u; }
EXB_HD double win_from_bits(unsigned long long u) { double x; std::memcpy(&x, &u, 8); return x; }
#endif

constexpr int kWinFlushEvery = 1024;        // products per window between flushes
constexpr int kWinBMin = -880;              // u_2 = 2^(b-104) and every TwoProd error stay normal
constexpr int kWinBMax = 900;               // 2^11 * 2^52 * u_0 stays below 2^988
constexpr unsigned kWinSpanProd = 50u << 20;   // exponent of p in [b+1, b+50]
constexpr unsigned kWinSpanSum = 51u << 20;    // exponent of x in [b, b+50]   (single summands: two digits)

struct Window {
    unsigned long long a0, a1, a2;   // sums of raw bit patterns, one per digit
    double M0, M1, M2;               // magic constants of the three digits
    unsigned key0;                   // high word of 2^(b+1): in-window  <=>  (|hi| - key0) < span
    unsigned span;                   // kWinSpanProd when anchored, 0 when the window is not in use
    unsigned cnt;                    // products accumulated since the last flush
    unsigned misses;                 // consecutive groups in which this thread had a summand outside
    unsigned st;                     // status flags raised by the ordinary deposits of this thread
};

EXB_HD void win_reset(Window& w) {
    w.a0 = w.a1 = w.a2 = 0ull;
    w.M0 = w.M1 = w.M2 = 0.0;
    w.key0 = 0u;
    w.span = 0u;
    w.cnt = 0u;
    w.misses = 0u;
    w.st = 0u;
}

// Anchor an EMPTY window so that a value with high word `hi` sits in the middle of it.
// Returns false (window left unused) for zeros, specials and values too close to the layout edges.
EXB_HD bool win_anchor(Window& w, unsigned hi) {
    const int E = (int)((hi >> 20) & 0x7ffu);
    if (E == 0 || E == 0x7ff) return false;
    int b = E - 1023 - 25;
    if (b < kWinBMin || b > kWinBMax) return false;
    const unsigned long long m0 = ((unsigned long long)(unsigned)(b + 52 + 1023) << 52) | 0x0008000000000000ull;   // 1.5 * 2^(b+52)
    w.M0 = win_from_bits(m0);
    w.M1 = win_from_bits(m0 - (52ull << 52));
    w.M2 = win_from_bits(m0 - (104ull << 52));
    w.a0 = w.a1 = w.a2 = 0ull;
    w.key0 = (unsigned)(b + 1 + 1023) << 20;
    w.span = kWinSpanProd;
    w.cnt = 0u;
    return true;
}

// exponent of the double with high word `hi` in [b+1, b+50]
EXB_HD bool win_holds(const Window& w, unsigned hi) { return ((hi & 0x7fffffffu) - w.key0) < w.span; }

// p + e (TwoProductFMA of two doubles, p in the window) into the three digits
EXB_HD void win_add_product(Window& w, double p, double e) {
    const double t1 = EXB_ADD(p, w.M0);
    const double x1 = EXB_SUB(t1, w.M0);
    const double r1 = EXB_SUB(p, x1);
    const double t2 = EXB_ADD(r1, w.M1);
    const double t3 = EXB_ADD(e, w.M1);
    const double x3 = EXB_SUB(t3, w.M1);
    const double r3 = EXB_SUB(e, x3);
    const double t4 = EXB_ADD(r3, w.M2);
    w.a0 += win_bits(t1);
    w.a1 += win_bits(t2) + win_bits(t3);
    w.a2 += win_bits(t4);
}

// The window's content as four doubles whose exact sum it is (each an integer below 2^52, or the
// signed top carry, times a power of two), and the window emptied.  out[] entries may be zero.
EXB_HD void win_drain(Window& w, double (&out)[4]) {
    out[0] = out[1] = out[2] = out[3] = 0.0;
    if (w.cnt == 0u) return;
    const unsigned long long c = w.cnt;
    long long d0 = (long long)(w.a0 - c * win_bits(w.M0));
    long long d1 = (long long)(w.a1 - 2ull * c * win_bits(w.M1));
    long long d2 = (long long)(w.a2 - c * win_bits(w.M2));
    d1 += d2 >> kDigits;
    d2 &= kLimbMask;
    d0 += d1 >> kDigits;
    d1 &= kLimbMask;
    const long long top = d0 >> kDigits;                       // |top| <= 2^10
    d0 &= kLimbMask;
    // units: u_k = M_k / (1.5 * 2^52), built by exponent-field edits of the magic constants
    const unsigned long long e0 = (win_bits(w.M0) >> 52) - 52ull;       // biased exponent of u_0
    const double u0 = win_from_bits(e0 << 52), u1 = win_from_bits((e0 - 52ull) << 52), u2 = win_from_bits((e0 - 104ull) << 52);
    const double utop = win_from_bits((e0 + 52ull) << 52);
    out[0] = (double)d2 * u2;                                   // exact: integer < 2^52 times a power of two, all normal
    out[1] = (double)d1 * u1;
    out[2] = (double)d0 * u0;
    out[3] = (double)top * utop;
    w.a0 = w.a1 = w.a2 = 0ull;
    w.cnt = 0u;
}

struct Window;
EXB_HD void win_drain_single(Window& w, double (&out)[4]);

// Bookkeeping of one thread after a group of summands that took the ordinary path (the warp vote
// failed).  `mine`: this thread's own summands were all inside its window.  A thread whose summands
// fall outside in two groups in a row drains its window through `emit` (ordinary deposits) and
// anchors it at the first usable summand of the group (high words in hi[]).
template <int G, class Emit>
EXB_HD void win_after_slow_group(Window& w, bool mine, const unsigned (&hi)[G], bool single, Emit&& emit) {
    if (mine) {
        w.misses = 0u;
        return;
    }
    if (++w.misses < 2u) return;
    double out[4];
    if (single) win_drain_single(w, out); else win_drain(w, out);
    for (int k = 0; k < 4; ++k)
        if (out[k] != 0.0) emit(out[k]);
    w.span = 0u;
    for (int k = 0; k < G; ++k)
        if (win_anchor(w, hi[k])) break;
    w.misses = 0u;
}

// Single summands (ExSUM): x with exponent in [b, b+50] needs two digits only.
//      t1 = x + M0, x1 = t1 - M0, r1 = x - x1, t2 = r1 + M1      (r1 multiple of ulp(x) >= u_1)
// Uses digits 0 and 1 of the same window; a1 receives ONE pattern per summand here, so a window is
// used either for products or for single summands, never both (cnt1 = cnt).
EXB_HD bool win_holds_sum(const Window& w, unsigned hi) { return ((hi & 0x7fffffffu) - (w.key0 - (1u << 20))) < (w.span ? kWinSpanSum : 0u); }
EXB_HD void win_add_single(Window& w, double x) {
    const double t1 = EXB_ADD(x, w.M0);
    const double x1 = EXB_SUB(t1, w.M0);
    const double r1 = EXB_SUB(x, x1);
    const double t2 = EXB_ADD(r1, w.M1);
    w.a0 += win_bits(t1);
    w.a1 += win_bits(t2);
}
EXB_HD void win_drain_single(Window& w, double (&out)[4]) {
    out[0] = out[1] = out[2] = out[3] = 0.0;
    if (w.cnt == 0u) return;
    const unsigned long long c = w.cnt;
    long long d0 = (long long)(w.a0 - c * win_bits(w.M0));
    long long d1 = (long long)(w.a1 - c * win_bits(w.M1));
    d0 += d1 >> kDigits;
    d1 &= kLimbMask;
    const long long top = d0 >> kDigits;
    d0 &= kLimbMask;
    const unsigned long long e0 = (win_bits(w.M0) >> 52) - 52ull;
    out[1] = (double)d1 * win_from_bits((e0 - 52ull) << 52);
    out[2] = (double)d0 * win_from_bits(e0 << 52);
    out[3] = (double)top * win_from_bits((e0 + 52ull) << 52);
    w.a0 = w.a1 = w.a2 = 0ull;
    w.cnt = 0u;
}


// ---------------------------------------------------------------------------------------------
// W-digit window for PRODUCTS (W >= 3): the same split carried over more digits, for data whose
// products span more than 50 binades (e.g. the ill-conditioned dot product of BASELINE config 3:
// factors over ~60 binades each, products over ~125).  Digits k = 0 .. W-1 count units of
// u_k = 2^(b - 52 k); a product p with TwoProd error e is admitted when
//        b - 52 (W - 3) + 1  <=  exponent(p)  <=  b + 50          (50 + 52 (W - 3) binades)
// p is split over digits 0 .. W-2 and e over digits 1 .. W-1, each with the chain
//        t_k = r + M_k;  x_k = t_k - M_k;  r = r - x_k            (last digit: t only, no remainder)
// which is exact for the same reasons as in the 3-digit case (|r| <= u_(k-1) / 2 = 2^51 u_k going in;
// p is a multiple of 2^(exp(p) - 52) >= u_(W-2), e of 2^(exp(p) - 105) >= u_(W-1)).
// Cost: 2 + 2 (3 (W - 2) + 1) FP64 instructions per product (W = 5: 22) against two deposits.
// ---------------------------------------------------------------------------------------------
template <int W>
struct WindowP {
    unsigned long long a[W];         // sums of raw bit patterns, one per digit
    double M0;                       // magic constant of digit 0; M_k = M0 / 2^(52 k) is one integer subtract away (winp_M)
    unsigned key0, span, cnt, misses, st;
    int emin, emax;                  // exponents of the products that missed so far (winp_cover)
};

// magic constant of digit k: the exponent field of M0 lowered by 52 k (the low word of 1.5 * 2^e is zero).  Kept out of
// the window state on purpose: five constants in registers made the 5-digit loop spill its prefetch slots.
template <int W>
EXB_HD double winp_M(const WindowP<W>& w, int k) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(w.M0) - k * (52 << 20), 0);
#else
    return win_from_bits(win_bits(w.M0) - ((unsigned long long)(52 * k) << 52));
#endif
}

template <int W>
EXB_HD void winp_reset(WindowP<W>& w) {
    for (int k = 0; k < W; ++k) w.a[k] = 0ull;
    w.M0 = 0.0;
    w.key0 = w.span = w.cnt = w.misses = w.st = 0u;
    w.emin = 4096;
    w.emax = -4096;
}

// Anchor an EMPTY window so that a value with high word `hi` sits in the middle of the admitted range.
template <int W>
EXB_HD bool winp_anchor(WindowP<W>& w, unsigned hi) {
    const int E = (int)((hi >> 20) & 0x7ffu);
    if (E == 0 || E == 0x7ff) return false;
    const int b = E - 1023 - 25 + 26 * (W - 3);                  // the admitted range is [b - 52 (W-3) + 1, b + 50]
    if (b < -984 + 52 * (W - 1) || b > kWinBMax) return false;   // u_(W-1) and every TwoProd error stay normal
    const unsigned long long m0 = ((unsigned long long)(unsigned)(b + 52 + 1023) << 52) | 0x0008000000000000ull;
    w.M0 = win_from_bits(m0);
    for (int k = 0; k < W; ++k) w.a[k] = 0ull;
    w.key0 = (unsigned)(b - 52 * (W - 3) + 1 + 1023) << 20;
    w.span = (unsigned)(50 + 52 * (W - 3)) << 20;
    w.cnt = 0u;
    return true;
}

template <int W>
EXB_HD bool winp_holds(const WindowP<W>& w, unsigned hi) { return ((hi & 0x7fffffffu) - w.key0) < w.span; }

template <int W>
EXB_HD void winp_add_product(WindowP<W>& w, double p, double e) {
    double r = p;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < W - 1; ++k) {                            // p: digits 0 .. W-2
        const double t = EXB_ADD(r, winp_M(w, k));
        w.a[k] += win_bits(t);
        if (k < W - 2) r = EXB_SUB(r, EXB_SUB(t, winp_M(w, k)));
    }
    r = e;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 1; k < W; ++k) {                                // e: digits 1 .. W-1
        const double t = EXB_ADD(r, winp_M(w, k));
        w.a[k] += win_bits(t);
        if (k < W - 1) r = EXB_SUB(r, EXB_SUB(t, winp_M(w, k)));
    }
}

// The window's content as W + 1 doubles whose exact sum it is, and the window emptied.
template <int W>
EXB_HD void winp_drain(WindowP<W>& w, double (&out)[W + 1]) {
    for (int k = 0; k <= W; ++k) out[k] = 0.0;
    if (w.cnt == 0u) return;
    const unsigned long long c = w.cnt;
    long long d[W];
    for (int k = 0; k < W; ++k) {
        const unsigned long long per = (k == 0 || k == W - 1) ? 1ull : 2ull;     // bit patterns added per product
        d[k] = (long long)(w.a[k] - per * c * win_bits(winp_M(w, k)));
    }
    for (int k = W - 1; k > 0; --k) {
        d[k - 1] += d[k] >> kDigits;
        d[k] &= kLimbMask;
    }
    const long long top = d[0] >> kDigits;
    d[0] &= kLimbMask;
    const unsigned long long e0 = (win_bits(w.M0) >> 52) - 52ull;             // biased exponent of u_0
    for (int k = 0; k < W; ++k) out[k] = (double)d[k] * win_from_bits((e0 - (unsigned long long)(52 * k)) << 52);
    out[W] = (double)top * win_from_bits((e0 + 52ull) << 52);
    for (int k = 0; k < W; ++k) w.a[k] = 0ull;
    w.cnt = 0u;
}

template <int W, int G, class Emit>
EXB_HD void winp_after_slow_group(WindowP<W>& w, bool mine, const unsigned (&hi)[G], Emit&& emit) {
    if (mine) {
        w.misses = 0u;
        return;
    }
    if (++w.misses < 2u) return;
    double out[W + 1];
    winp_drain(w, out);
    for (int k = 0; k <= W; ++k)
        if (out[k] != 0.0) emit(out[k]);
    w.span = 0u;
    for (int k = 0; k < G; ++k)
        if (winp_anchor(w, hi[k])) break;
    w.misses = 0u;
}

// Anchor an EMPTY window so that it admits every exponent in [emin, emax] (unbiased exponents of products seen so
// far), with the slack split evenly on both sides.  False when the range does not fit or lies too close to the
// layout edges (window left unused).
template <int W>
EXB_HD bool winp_anchor_range(WindowP<W>& w, int emin, int emax) {
    constexpr int width = 50 + 52 * (W - 3);                     // admitted exponents: lo .. lo + width - 1
    const int need = emax - emin + 1;
    if (need > width) return false;
    const int lo = emin - (width - need) / 2;
    const int b = lo + 52 * (W - 3) - 1;
    if (b < -984 + 52 * (W - 1) || b > kWinBMax) return false;
    const unsigned long long m0 = ((unsigned long long)(unsigned)(b + 52 + 1023) << 52) | 0x0008000000000000ull;
    w.M0 = win_from_bits(m0);
    for (int k = 0; k < W; ++k) w.a[k] = 0ull;
    w.key0 = (unsigned)(lo + 1023) << 20;
    w.span = (unsigned)width << 20;
    w.cnt = 0u;
    return true;
}

// admitted exponent range of an anchored window
template <int W>
EXB_HD int winp_lo(const WindowP<W>& w) { return (int)(w.key0 >> 20) - 1023; }
template <int W>
EXB_HD int winp_hi(const WindowP<W>& w) { return (int)((w.key0 + w.span) >> 20) - 1024; }

// Re-anchoring policy of the W-digit window: [gmin, gmax] is the exponent range of the products of a group that did
// not take the window path (on the device: over the whole warp, so that every lane makes the same decision and all
// lanes' windows stay identical).  The window is moved -- drained through `emit` first -- so that it admits every
// exponent seen in such groups so far; once that range no longer fits, the window is left unused and its loop ends.
template <int W, class Emit>
EXB_HD void winp_cover(WindowP<W>& w, int gmin, int gmax, Emit&& emit) {
    if (gmin > gmax) return;                                     // only zeros / specials in the group
    if (gmin < w.emin) w.emin = gmin;
    if (gmax > w.emax) w.emax = gmax;
    if (w.span != 0u && winp_lo(w) <= w.emin && w.emax <= winp_hi(w)) return;
    double out[W + 1];
    winp_drain(w, out);
    for (int k = 0; k <= W; ++k)
        if (out[k] != 0.0) emit(out[k]);
    w.span = 0u;
    winp_anchor_range(w, w.emin, w.emax);
}

// ---------------------------------------------------------------------------------------------
// W-digit window for SINGLE summands (ExSUM), W >= 2, on the same WindowP<W> state: x with
//        b - 52 (W - 2)  <=  exponent(x)  <=  b + 50          (51 + 52 (W - 2) binades; W = 3: 103)
// is split over digits 0 .. W-1 (x is a multiple of 2^(exp(x) - 52) >= u_(W-1): the last digit has no
// remainder): 3 (W - 1) + 1 FP64 instructions per summand, one bit pattern per digit.
// ---------------------------------------------------------------------------------------------
template <int W>
EXB_HD bool wins_anchor_range(WindowP<W>& w, int emin, int emax) {
    constexpr int width = 51 + 52 * (W - 2);
    const int need = emax - emin + 1;
    if (need > width) return false;
    const int lo = emin - (width - need) / 2;
    const int b = lo + 52 * (W - 2);
    if (b < -984 + 52 * (W - 1) || b > kWinBMax) return false;
    const unsigned long long m0 = ((unsigned long long)(unsigned)(b + 52 + 1023) << 52) | 0x0008000000000000ull;
    w.M0 = win_from_bits(m0);
    for (int k = 0; k < W; ++k) w.a[k] = 0ull;
    w.key0 = (unsigned)(lo + 1023) << 20;
    w.span = (unsigned)width << 20;
    w.cnt = 0u;
    return true;
}

template <int W>
EXB_HD void wins_add(WindowP<W>& w, double x) {
    double r = x;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < W; ++k) {
        const double t = EXB_ADD(r, winp_M(w, k));
        w.a[k] += win_bits(t);
        if (k < W - 1) r = EXB_SUB(r, EXB_SUB(t, winp_M(w, k)));
    }
}

template <int W>
EXB_HD void wins_drain(WindowP<W>& w, double (&out)[W + 1]) {
    for (int k = 0; k <= W; ++k) out[k] = 0.0;
    if (w.cnt == 0u) return;
    const unsigned long long c = w.cnt;
    long long d[W];
    for (int k = 0; k < W; ++k) d[k] = (long long)(w.a[k] - c * win_bits(winp_M(w, k)));
    for (int k = W - 1; k > 0; --k) {
        d[k - 1] += d[k] >> kDigits;
        d[k] &= kLimbMask;
    }
    const long long top = d[0] >> kDigits;
    d[0] &= kLimbMask;
    const unsigned long long e0 = (win_bits(w.M0) >> 52) - 52ull;
    for (int k = 0; k < W; ++k) out[k] = (double)d[k] * win_from_bits((e0 - (unsigned long long)(52 * k)) << 52);
    out[W] = (double)top * win_from_bits((e0 + 52ull) << 52);
    for (int k = 0; k < W; ++k) w.a[k] = 0ull;
    w.cnt = 0u;
}

template <int W, class Emit>
EXB_HD void wins_cover(WindowP<W>& w, int gmin, int gmax, Emit&& emit) {
    if (gmin > gmax) return;
    if (gmin < w.emin) w.emin = gmin;
    if (gmax > w.emax) w.emax = gmax;
    if (w.span != 0u && winp_lo(w) <= w.emin && w.emax <= winp_hi(w)) return;
    double out[W + 1];
    wins_drain(w, out);
    for (int k = 0; k <= W; ++k)
        if (out[k] != 0.0) emit(out[k]);
    w.span = 0u;
    wins_anchor_range(w, w.emin, w.emax);
}

}  // namespace exb
