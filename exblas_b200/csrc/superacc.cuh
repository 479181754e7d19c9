// superacc.cuh -- the 39-limb Kulisch superaccumulator used by the B200 ExSUM / ExDOT path.
//
// Layout (identical to the reference's GPU layout, src/gpu/blas/blas1/ExSUM.FPE.cl:14-18 and
// include/common.hpp:43):  39 signed 64-bit limbs, radix 2^52 (12 carry-save bits per limb),
// f_words = 20, so limb j weighs 2^(52*(j-20)): limb 0 starts at 2^-1040, limb 38 at 2^936.
// The reference's CPU layout (41 limbs, f_words = 21, superaccumulator.cpp:14-17) is the same
// array shifted by one limb, so Round() gives the same double on both.
//
// What is here:
//   * normalize()          -- carry propagation to the unique normal form
//                             (restates Superaccumulator::Normalize, superaccumulator.cpp:138-162)
//   * round_ref_compat()   -- bit-for-bit restatement of Superaccumulator::Round
//                             (superaccumulator.cpp:80-134 == ExSUM.FPE.cl:119-162) on this layout;
//                             NOT always correctly rounded -- it is the
//                             reference-parity finaliser
//   * round_exact()        -- true round-to-nearest-even of the accumulator value
//   * deposit_*()          -- device side: exact split of one double into two signed limb digits
//                             and a non-atomic read-modify-write of a THREAD-PRIVATE accumulator
//                             column in shared memory (replaces Accumulate/AccumulateWord/xadd,
//                             ExSUM.FPE.cl:168-228, which use 64-bit local atomics)
#pragma once
#include <cstdint>
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define EXB_HD __host__ __device__ __forceinline__
#define EXB_D __device__ __forceinline__
#else
#define EXB_HD inline
#endif

namespace exb {

constexpr int kLimbs = 39;    // BIN_COUNT
constexpr int kFWords = 20;   // f_words
constexpr int kDigits = 52;   // digits = 64 - K, K = 12
constexpr long long kLimbMask = (1ll << kDigits) - 1;

// A double x = +-1.m * 2^(E-1023) with exponent field E has its mantissa LSB at 2^(E-1075).
// It lies entirely inside the accumulator when E-1075 >= -1040 (E >= 35) and its top digit
// lands at or below limb 38 (E < 2011, i.e. |x| < 2^988).  This is the reference GPU kernels'
// domain (outside it they write out of bounds).
constexpr unsigned kEMin = 35;
constexpr unsigned kELim = 2011;

// Status flags (OR-ed together), reported through the C ABI.
enum : unsigned {
    kStNaN = 1u,         // a NaN was met (input, or product in ExDOT)
    kStPosInf = 2u,      // +Inf met
    kStNegInf = 4u,      // -Inf met
    kStTooLarge = 8u,    // finite |x| >= 2^988: above the 39-limb layout; element dropped
    kStTooSmall = 16u,   // bits below 2^-1040 were truncated: result no longer exact
    kStPeerTimeout = 32u,  // fused multi-GPU exchange: a peer's contribution never arrived (result is partial)
};

// Limb deposits between two boundings of one accumulator column (bound_column): each deposit adds a
// digit of magnitude <= 2^52 to a limb that starts below 2^52 + 2^11 in magnitude, so 2040 deposits
// keep |limb| < 2^63.
constexpr int kMaxDepositsPerNormalize = 2040;

// ---------------------------------------------------------------------------------------------
// host + device: normal form and the two finalisers
// ---------------------------------------------------------------------------------------------

// Carry-propagate so that limbs 0..37 are in [0, 2^52) and limb 38 keeps the signed remainder.
// Returns true when the value is negative.  (superaccumulator.cpp:138-162 with imin = 0)
EXB_HD bool normalize(long long* acc) {
    long long carry = acc[0] >> kDigits;
    acc[0] -= carry << kDigits;
    for (int i = 1; i < kLimbs; ++i) {
        long long v = acc[i] + carry;
        carry = v >> kDigits;
        acc[i] = v - (carry << kDigits);
    }
    acc[kLimbs - 1] += (long long)((unsigned long long)carry << kDigits);
    return carry < 0;
}

// x * 2^e for |e| <= 1040 and |x| an integer below 2^64 (what the finalisers need).  On the device the scaling is
// two multiplications by exact powers of two assembled from the exponent field -- the first one is exact (the
// intermediate stays normal), so the result is rounded once, exactly like ldexp(), at a fraction of its code size
// (the epilogue runs once per launch from a cold instruction cache: every instruction counts there).
EXB_HD double exb_ldexp(double x, int e) {
#if defined(__CUDA_ARCH__)
    const int e1 = e / 2, e2 = e - e1;
    const double p1 = __longlong_as_double((long long)(e1 + 1023) << 52);
    const double p2 = __longlong_as_double((long long)(e2 + 1023) << 52);
    return __dmul_rn(__dmul_rn(x, p1), p2);
#else
    return std::ldexp(x, e);
#endif
}

// Restatement of the reference Round() (superaccumulator.cpp:80-134), split in two so that the serial
// (host, one thread) and the warp-parallel (reduce_kernel.cuh: warp_round) scans share the arithmetic:
//   * the SCAN finds i = the limb Round() starts from (:91-101), and whether the limbs below i-1 hold a
//     non-zero sticky contribution (:116-119);
//   * round_ref_parts() is everything else (:102-134, mylibm.hpp:156-171), from the two limbs i, i-1.
EXB_HD double round_ref_parts(int i, long long acc_i, long long acc_im1, bool sticky_nonzero, bool negative) {
    if (i < 0) return 0.0;                                               // :102-104
    long long hiword = negative ? kLimbMask - acc_i : acc_i;             // :106 (one's complement)
    double rounded = (double)hiword;
    double hi = exb_ldexp(rounded, (i - kFWords) * kDigits);             // :108
    if (i == 0) return negative ? -hi : hi;                              // :109-111
    hiword -= (long long)rounded;                                        // :112 (rounded is integral)
    double mid = exb_ldexp((double)hiword, (i - kFWords) * kDigits);     // :113
    long long loword = negative ? (1ll << kDigits) - acc_im1 : acc_im1;  // :121
    loword |= (long long)sticky_nonzero;                                 // :122
    double lo = exb_ldexp((double)loword, (i - 1 - kFWords) * kDigits);  // :123
    if (mid != 0) {                                                      // :128-130, mylibm.hpp:156-171
        double d = mid + lo;
        long long l;
#if defined(__CUDA_ARCH__)
        l = __double_as_longlong(d);
        l |= (lo != 0.0);
        d = __longlong_as_double(l);
#else
        std::memcpy(&l, &d, 8);
        l |= (lo != 0.0);
        std::memcpy(&d, &l, 8);
#endif
        lo = d;
    }
    hi = hi + lo;                                                        // :132 the only rounding
    return negative ? -hi : hi;
}

// `acc` must already be in normal form; `negative` is normalize()'s return value.
EXB_HD double round_ref_compat(const long long* acc, bool negative) {
    int i = kLimbs - 1;
    while (i >= 0 && acc[i] == 0) --i;                                   // :91-94
    if (negative) {
        while (i >= 0 && (acc[i] & kLimbMask) == kLimbMask) --i;         // :95-101
    }
    if (i <= 0) return round_ref_parts(i, i == 0 ? acc[0] : 0, 0, false, negative);
    long long sticky = 0;
    for (int j = 0; j != i - 1; ++j)                                     // :116-119
        sticky |= negative ? (1ll << kDigits) - acc[j] : acc[j];
    return round_ref_parts(i, acc[i], acc[i - 1], sticky != 0, negative);
}

EXB_HD int exb_clzll(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}

// Correctly rounded (nearest, ties to even) value of normalised limbs; handles subnormal
// results and overflow to +-inf.  Not part of the reference; checked against math.fsum / MPFR.
// Split like round_ref_compat: the scan produces the MAGNITUDE's top three radix-2^52 digits
// (m_top != 0 at limb index `top`; m1, m2 the two digits below, zero where they do not exist) and
// whether anything non-zero lies below them; round_exact_parts() does the rest.
EXB_HD double round_exact_parts(int top, unsigned long long m_top, unsigned long long m1, unsigned long long m2,
                                bool sticky, bool negative) {
    if (top < 0) return 0.0;
    const int width = 64 - exb_clzll(m_top);             // significant bits of the top digit (<= 63)
    const int P = kDigits * top + width - 1;             // MSB position above the limb-0 LSB
    const int e = P - kDigits * kFWords;                 // exponent of the MSB
    // 64-bit window w = bits [P-63, P] of the magnitude (zero filled), sticky = anything below it
    unsigned long long w = m_top << (64 - width);
    int filled = width;
    const unsigned long long lower[2] = {m1, m2};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 2; ++k) {
        const unsigned long long mj = lower[k];
        const int room = 64 - filled;
        if (room >= kDigits) {
            w |= mj << (room - kDigits);
            filled += kDigits;
        } else if (room > 0) {
            w |= mj >> (kDigits - room);
            sticky |= (mj & ((1ull << (kDigits - room)) - 1ull)) != 0;
            filled = 64;
        } else {
            sticky |= (mj != 0);
        }
    }
    int keep = 53;                                       // result bits; fewer when subnormal
    if (e < -1022) keep = 53 - (-1022 - e);
    unsigned long long q = 0;
    bool round_bit = false;
    if (keep >= 1) {
        q = w >> (64 - keep);
        round_bit = (w >> (63 - keep)) & 1ull;
        sticky |= (w & ((1ull << (63 - keep)) - 1ull)) != 0;
    } else if (keep == 0) {
        round_bit = true;                                // the MSB itself is the round bit
        sticky |= (w << 1) != 0;
    } else {
        sticky = true;                                   // below half of the smallest subnormal
    }
    if (round_bit && (sticky || (q & 1ull))) q += 1;
    double r = exb_ldexp((double)q, (P - keep + 1) - kDigits * kFWords);
    return negative ? -r : r;
}

EXB_HD double round_exact(const long long* acc, bool negative) {
    unsigned long long m[kLimbs];      // magnitude, 52-bit digits, top digit unbounded
    if (!negative) {
        for (int i = 0; i < kLimbs; ++i) m[i] = (unsigned long long)acc[i];
    } else {
        long long borrow = 0;
        for (int i = 0; i < kLimbs - 1; ++i) {
            long long v = borrow - acc[i];               // in (-2^52 - 1, 0]
            borrow = v >> kDigits;
            m[i] = (unsigned long long)(v - borrow * (1ll << kDigits));
        }
        m[kLimbs - 1] = (unsigned long long)(borrow - acc[kLimbs - 1]);
    }
    int top = kLimbs - 1;
    while (top >= 0 && m[top] == 0) --top;
    if (top < 0) return 0.0;
    bool sticky = false;
    for (int j = top - 3; j >= 0; --j) sticky |= (m[j] != 0);
    return round_exact_parts(top, m[top], top >= 1 ? m[top - 1] : 0ull, top >= 2 ? m[top - 2] : 0ull, sticky, negative);
}

// Exact accumulation of one double into a plain limb array (not the hot path: used where a value
// is added once per output, e.g. beta*y in ExGEMV, and on the host).  Integer split of the 53-bit
// mantissa into <= 3 limb digits.  Returns status flags.
EXB_HD unsigned accumulate_double(long long* acc, double x) {
    unsigned long long u;
#if defined(__CUDA_ARCH__)
    u = (unsigned long long)__double_as_longlong(x);
#else
    std::memcpy(&u, &x, 8);
#endif
    const unsigned E = (unsigned)((u >> 52) & 0x7ffu);
    unsigned long long mant = u & ((1ull << 52) - 1ull);
    const bool neg = (u >> 63) != 0;
    if (E == 0 && mant == 0) return 0u;
    if (E == 0x7ffu) return mant ? kStNaN : (neg ? kStNegInf : kStPosInf);
    if (E >= kELim) return kStTooLarge;
    unsigned st = 0u;
    int pos;                                    // position of the mantissa LSB above the accumulator LSB
    if (E == 0) pos = 1 - 1075 + 1040; else { mant |= 1ull << 52; pos = (int)E - 1075 + 1040; }
    if (pos < 0) {
        const int sh = -pos;
        if (sh >= 64) { return kStTooSmall; }
        if (mant & ((1ull << sh) - 1ull)) st |= kStTooSmall;
        mant >>= sh;
        pos = 0;
    }
    const int j = pos / kDigits, s = pos % kDigits;
    // mant << s spans up to 105 bits: digits for limbs j, j+1, j+2
    const unsigned long long lo = (mant << s) & (unsigned long long)kLimbMask;
    const unsigned long long rest = s ? (mant >> (kDigits - s)) : (mant >> kDigits);   // bits above limb j
    const unsigned long long mid = rest & (unsigned long long)kLimbMask;
    const unsigned long long hi = rest >> kDigits;
    if (neg) {
        acc[j] -= (long long)lo;
        if (j + 1 < kLimbs) acc[j + 1] -= (long long)mid;
        if (j + 2 < kLimbs) acc[j + 2] -= (long long)hi;
    } else {
        acc[j] += (long long)lo;
        if (j + 1 < kLimbs) acc[j + 1] += (long long)mid;
        if (j + 2 < kLimbs) acc[j + 2] += (long long)hi;
    }
    return st;
}

// Final value from limbs + status flags (IEEE semantics for the specials the kernel met).
EXB_HD double finalize_value(long long* acc, unsigned status, int round_mode) {
    bool neg = normalize(acc);
    if ((status & kStNaN) || ((status & kStPosInf) && (status & kStNegInf))) {
#if defined(__CUDA_ARCH__)
        return __longlong_as_double(0x7ff8000000000000ll);
#else
        return std::nan("");
#endif
    }
    if (status & kStPosInf) return HUGE_VAL;
    if (status & kStNegInf) return -HUGE_VAL;
    return round_mode ? round_exact(acc, neg) : round_ref_compat(acc, neg);
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------------
// device: deposits into a thread-private accumulator column
//
// Shared-memory layout: limb j of thread t lives at byte address col + j * stride, where
// col = smem_base + 8 * t and stride = 8 * T (T = blockDim.x, a multiple of 32).  A 64-bit
// access by lane l of a warp then falls in bank pair (l mod 16) whatever j is, so every
// warp-wide LDS.64 / STS.64 is conflict free even though each lane indexes a different limb.
// No atomics are needed because no two threads share a column.  Addresses are 32-bit
// shared-window addresses and the accesses are explicit ld.shared / st.shared, so the compiler
// can never demote them to generic loads.
// ---------------------------------------------------------------------------------------------
EXB_D unsigned long long lds64(unsigned addr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
EXB_D void sts64(unsigned addr, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// Fast path: x is finite, non-zero and inside the layout (kEMin <= E < kELim).
//
// Let J1 = floor((E-35)/52) + 1 and xs = |x| / 2^(52*J1 - 1040), an exponent-field edit.  Then
// xs is in [2^s, 2^(s+1)) with s = (E-35) mod 52 <= 51, and its LSB is >= 2^-52, so
//     D1 = rint(xs)            in [0, 2^52]     (digit for limb J1)
//     D0 = (xs - D1) * 2^52    in [-2^51, 2^51] (digit for limb J1-1), an exact integer.
// Both come out of the FP64 pipe with the add-a-magic-constant trick: bits(xs + 2^52) - bits(2^52)
// = D1 and bits(rem + 1.5) - bits(1.5) = rem * 2^52, which keeps the integer pipe (the busier
// one here) to ~15 instructions per element.  The sign of x is applied to D0 by flipping rem's
// sign bit and to D1 by a two's-complement negate folded into the 3-input accumulate add.
#ifdef EXB_INT_SPLIT
// Integer-only variant of the split (experiment, profiles/README.md "deposit split"): x = m * 2^(E-1075) with the
// 53-bit mantissa m; in units of the accumulator LSB (2^-1040) that is m << p, p = E - 35.  With J0 = p / 52 and
// s = p % 52 the two digits are D0 = (m << s) mod 2^52 (limb J0) and D1 = m >> (52 - s) (limb J0 + 1 = J1), both
// non-negative: eleven shift / logic instructions and no FP64-pipe work at all.
EXB_D void split_int(unsigned lo, unsigned ahi, unsigned& J1, unsigned long long& d0, unsigned long long& d1) {
    const unsigned E = ahi >> 20;
    J1 = __umulhi(E + 17u, 82595525u);                            // floor((E+17)/52) = J0 + 1, in [1, 38]
    const unsigned s = (E + 17u) - J1 * 52u;                      // (E - 35) mod 52
    const unsigned long long m = ((unsigned long long)((ahi & 0xfffffu) | 0x100000u) << 32) | lo;
    d0 = (m << s) & (unsigned long long)kLimbMask;
    d1 = m >> (52u - s);
}
EXB_D void deposit_fast(unsigned col, unsigned stride, unsigned lo, unsigned hi) {
    unsigned J1;
    unsigned long long d0, d1;
    split_int(lo, hi & 0x7fffffffu, J1, d0, d1);
    const unsigned long long sm = (unsigned long long)((long long)(int)hi >> 31);   // all ones when x < 0
    const unsigned a1 = col + J1 * stride;
    const unsigned a0 = (col - stride) + J1 * stride;
    const unsigned long long v0 = lds64(a0), v1 = lds64(a1);
    sts64(a0, v0 + (d0 ^ sm) - sm);
    sts64(a1, v1 + (d1 ^ sm) - sm);
}
EXB_D void deposit_fast_pos(unsigned col, unsigned stride, unsigned lo, unsigned hi) {
    unsigned J1;
    unsigned long long d0, d1;
    split_int(lo, hi, J1, d0, d1);
    const unsigned a1 = col + J1 * stride;
    const unsigned a0 = (col - stride) + J1 * stride;
    const unsigned long long v0 = lds64(a0), v1 = lds64(a1);
    sts64(a0, v0 + d0);
    sts64(a1, v1 + d1);
}
#else
EXB_D void deposit_fast(unsigned col, unsigned stride, unsigned lo, unsigned hi) {
    const unsigned ahi = hi & 0x7fffffffu;
    const unsigned E = ahi >> 20;
    const unsigned J1 = __umulhi(E + 17u, 82595525u);            // floor((E+17)/52), in [1, 38]
    const unsigned xhi = ahi - J1 * (52u << 20) + (1040u << 20);  // exponent field -> s + 1023
    const double xs = __hiloint2double((int)xhi, (int)lo);
    const double t = __dadd_rn(xs, 4503599627370496.0);           // 2^52 + D1
    const double xr = __dsub_rn(t, 4503599627370496.0);           // D1 as a double
    double rem = __dsub_rn(xs, xr);                               // exact, in [-0.5, 0.5]
    rem = __hiloint2double(__double2hiint(rem) ^ (int)(hi & 0x80000000u), __double2loint(rem));
    const double t2 = __dadd_rn(rem, 1.5);
    const unsigned long long d0 = (unsigned long long)__double_as_longlong(t2) - 0x3FF8000000000000ull;
    const unsigned m = (unsigned)((int)hi >> 31);                 // all ones when x < 0
    const unsigned long long tm = (unsigned long long)__double_as_longlong(t) ^ (((unsigned long long)m << 32) | m);
    // +D1 = bits(t) - K, -D1 = ~bits(t) + K + 1 with K = bits(2^52) = 0x4330000000000000
    const unsigned long long ksel = ((unsigned long long)(0xBCD00000u ^ (m & 0xFFE00000u)) << 32) | (0u - m);
    const unsigned a1 = col + J1 * stride;
    const unsigned a0 = (col - stride) + J1 * stride;
    const unsigned long long v0 = lds64(a0), v1 = lds64(a1);       // distinct limbs: both loads first
    sts64(a0, v0 + d0);
    sts64(a1, v1 + tm + ksel);
}

// Same for x > 0 only: no sign fix-ups at all (7 fewer integer instructions per element).  Used when
// a whole warp's vector is positive, which is the reference generator's case (init_fpuniform,
// common.cpp:18-33, draws positive values only) and that of norms, energies, histograms ...
EXB_D void deposit_fast_pos(unsigned col, unsigned stride, unsigned lo, unsigned hi) {
    const unsigned E = hi >> 20;                                   // sign bit is clear
    const unsigned J1 = __umulhi(E + 17u, 82595525u);
    const unsigned xhi = hi - J1 * (52u << 20) + (1040u << 20);
    const double xs = __hiloint2double((int)xhi, (int)lo);
    const double t = __dadd_rn(xs, 4503599627370496.0);
    const double xr = __dsub_rn(t, 4503599627370496.0);
    const double t2 = __dadd_rn(__dsub_rn(xs, xr), 1.5);
    const unsigned long long d0 = (unsigned long long)__double_as_longlong(t2) - 0x3FF8000000000000ull;
    const unsigned long long d1 = (unsigned long long)__double_as_longlong(t) - 0x4330000000000000ull;
    const unsigned a1 = col + J1 * stride;
    const unsigned a0 = (col - stride) + J1 * stride;
    const unsigned long long v0 = lds64(a0), v1 = lds64(a1);
    sts64(a0, v0 + d0);
    sts64(a1, v1 + d1);
}
#endif  // EXB_INT_SPLIT

// distance of |x| above the lower edge of the fast range, as an unsigned 32-bit key on the high
// word (so that one unsigned compare tests both edges, and a max over several keys tests them all)
EXB_D unsigned range_key(unsigned hi) { return (hi & 0x7fffffffu) - (kEMin << 20); }
constexpr unsigned kRangeSpan = (kELim - kEMin) << 20;
EXB_D bool in_fast_range(unsigned hi) { return range_key(hi) < kRangeSpan; }

// Slow path for everything else: zeros, specials, values outside the layout, and tiny values
// (E < 35) whose set bits may still all lie at or above 2^-1040.  Returns status flags.
__device__ __noinline__ unsigned deposit_slow(unsigned col, unsigned stride, unsigned lo, unsigned hi) {
    const unsigned ahi = hi & 0x7fffffffu;
    if ((ahi | lo) == 0u) return 0u;                              // +-0
    const unsigned E = ahi >> 20;
    if (E == 0x7ffu) return ((ahi & 0xfffffu) | lo) ? kStNaN : ((hi >> 31) ? kStNegInf : kStPosInf);
    if (E >= kELim) return kStTooLarge;
    // E < 35: mantissa LSB is below 2^-1040.  Integer value in units of 2^-1040 = m >> sh.
    unsigned long long mant = ((unsigned long long)(ahi & 0xfffffu) << 32) | lo;
    unsigned Eeff = E;
    if (E == 0u) Eeff = 1u; else mant |= (1ull << 52);
    const unsigned sh = kEMin - Eeff;                             // 1..34
    unsigned st = 0u;
    if (mant & ((1ull << sh) - 1ull)) st = kStTooSmall;           // truncated toward zero
    const unsigned long long v = mant >> sh;                      // < 2^52
    const unsigned long long d = (hi >> 31) ? (0ull - v) : v;
    sts64(col, lds64(col) + d);
    return st;
}

// Checked deposit of one double (any value).  Out of line on purpose: it serves the rare paths
// (expansion residuals, tails, mixed vectors) and keeping it a call keeps the hot loops small
// enough for the instruction cache.
__device__ __noinline__ unsigned deposit_any(unsigned col, unsigned stride, unsigned lo, unsigned hi) {
    if (in_fast_range(hi)) {
        deposit_fast(col, stride, lo, hi);
        return 0u;
    }
    return deposit_slow(col, stride, lo, hi);
}
EXB_D void deposit(unsigned col, unsigned stride, double x, unsigned& status) {
    status |= deposit_any(col, stride, (unsigned)__double2loint(x), (unsigned)__double2hiint(x));
}

// Deposit of an INTERNAL partial sum (an expansion level at flush time, a drained window digit).  Every input is
// below 2^988, but the sum of several of them inside one expansion need not be; the superaccumulator-only path
// adds the same inputs into limb 38 digit by digit without complaint, so the result must not depend on fpe here.
// Finite values in [2^988, 2^995) are therefore added to limb 38 (weight 2^936) as integers (< 2^59 each, and a
// flush deposits at most eight of them); only beyond that -- where limb 38 itself is about to overflow -- is the
// value dropped and flagged.
__device__ __noinline__ unsigned deposit_sum_any(unsigned col, unsigned stride, unsigned lo, unsigned hi) {
    if (in_fast_range(hi)) {
        deposit_fast(col, stride, lo, hi);
        return 0u;
    }
    const unsigned ahi = hi & 0x7fffffffu;
    const unsigned E = ahi >> 20;
    if (E >= kELim && E < kELim + 7u) {
        const unsigned long long mant = ((unsigned long long)((ahi & 0xfffffu) | 0x100000u) << 32) | lo;
        const unsigned long long v = mant << (E - kELim);             // units of 2^936
        const unsigned a38 = col + (unsigned)(kLimbs - 1) * stride;
        sts64(a38, (hi >> 31) ? lds64(a38) - v : lds64(a38) + v);
        return 0u;
    }
    return deposit_slow(col, stride, lo, hi);
}
EXB_D void deposit_sum(unsigned col, unsigned stride, double x, unsigned& status) {
    status |= deposit_sum_any(col, stride, (unsigned)__double2loint(x), (unsigned)__double2hiint(x));
}

// Four independent doubles; one range test for all of them (the common case is all-fast).
// POS: the caller has established (warp vote over a whole tile) that no value is negative.
template <bool POS>
EXB_D void deposit4(unsigned col, unsigned stride, double x0, double x1, double x2, double x3, unsigned& status) {
    const unsigned h0 = (unsigned)__double2hiint(x0), h1 = (unsigned)__double2hiint(x1);
    const unsigned h2 = (unsigned)__double2hiint(x2), h3 = (unsigned)__double2hiint(x3);
    if (max(max(range_key(h0), range_key(h1)), max(range_key(h2), range_key(h3))) < kRangeSpan) {
        if (POS) {
            deposit_fast_pos(col, stride, (unsigned)__double2loint(x0), h0);
            deposit_fast_pos(col, stride, (unsigned)__double2loint(x1), h1);
            deposit_fast_pos(col, stride, (unsigned)__double2loint(x2), h2);
            deposit_fast_pos(col, stride, (unsigned)__double2loint(x3), h3);
        } else {
            deposit_fast(col, stride, (unsigned)__double2loint(x0), h0);
            deposit_fast(col, stride, (unsigned)__double2loint(x1), h1);
            deposit_fast(col, stride, (unsigned)__double2loint(x2), h2);
            deposit_fast(col, stride, (unsigned)__double2loint(x3), h3);
        }
    } else {
        deposit(col, stride, x0, status);
        deposit(col, stride, x1, status);
        deposit(col, stride, x2, status);
        deposit(col, stride, x3, status);
    }
}

// A product with its TwoProd error term, p + e = a * b exactly, into THREE adjacent limbs instead of two deposits of two
// limbs each (round 2): e is at most half an ulp of p, so its digits fall on the two limbs below p's top limb and the
// middle limb is shared -- 3 instead of 4 read-modify-writes per product (12 instead of 16 shared-memory wavefronts per
// warp-product), which is what wide-range ExDOT / ExGEMV data are bound by.
// Grid: J = floor((E + 16) / 52), so that xs = |p| / 2^(52 J - 1040) lies in [2^s, 2^(s+1)) with s in [1, 52]
// (deposit_fast uses s in [0, 51]; s >= 1 is what makes the last digit of e an integer, see below).
//   p:  D1 = rint(xs) via the 2^52 magic add for s <= 51; for s = 52 xs IS a 53-bit integer: D1 = its mantissa, no
//       remainder (selected without a branch);  D0 = (xs - D1) 2^52
//   e:  es = e / 2^(52 J - 1040), |es| <= 2^(s-53) <= 1/2:   E1 = rint(es 2^52) = bits(es + 1.5) - bits(1.5),
//       r3 = es - E1 2^-52 (exact, |r3| <= 2^-53),           E0 = r3 2^104 = bits(r3 + 1.5 2^-52) - bits(1.5 2^-52)
//   limbs:  J += +-D1 (< 2^53: two digits' worth, which is what the ExDOT deposit budget counts per product),
//           J - 1 += +-D0 + E1,   J - 2 += E0
// E0 is an integer because e is a multiple of 2^(exp(p) - 105), i.e. of 2^(s - 105) on this grid, and s >= 1.
// p must be large enough for e to be a normal number: product3_ok().  Checked against exact rational arithmetic on
// 2 x 10^5 random products (all exponents, exact products, short mantissas) before it went to the GPU.
EXB_D bool product3_ok(unsigned phi, unsigned plo) {
    const unsigned ahi = phi & 0x7fffffffu;
    return ((ahi >> 20) - 200u < kELim - 200u) | ((ahi | plo) == 0u);     // 2^-823 <= |p| < 2^988, or an exact zero (adds zeros)
}
EXB_D void deposit_product3(unsigned col, unsigned stride, double p, double e) {
    const unsigned hi = (unsigned)__double2hiint(p), lo = (unsigned)__double2loint(p);
    const unsigned ahi = hi & 0x7fffffffu;
    const unsigned E = ahi >> 20;
    const bool pz = ahi == 0u;                                            // an exact zero (its error term is zero too): digits 0, 0, 0 into limbs 0..2
    const unsigned J = pz ? 2u : __umulhi(E + 16u, 82595525u);            // >= 4 otherwise
    const unsigned shift = J * (52u << 20) - (1040u << 20);               // exponent-field distance to the limb-J grid
    const unsigned xhi = pz ? 0u : ahi - shift;                           // exponent field s + 1023, s in [1, 52]
    const double xs = __hiloint2double((int)xhi, (int)lo);
    const bool wide = xhi >= ((52u + 1023u) << 20);                       // s == 52: xs is an integer of 53 bits
    const double t = __dadd_rn(xs, 4503599627370496.0);                   // 2^52 + D1 (s <= 51)
    const double xr = __dsub_rn(t, 4503599627370496.0);
    double rem = wide ? 0.0 : __dsub_rn(xs, xr);                          // exact, in [-0.5, 0.5]
    rem = __hiloint2double(__double2hiint(rem) ^ (int)(hi & 0x80000000u), __double2loint(rem));
    const double t2 = __dadd_rn(rem, 1.5);
    // D1 as an unsigned integer: bits(t) - bits(2^52), or the 53-bit mantissa itself
    const unsigned long long mant = ((unsigned long long)((ahi & 0xfffffu) | 0x100000u) << 32) | lo;
    const unsigned long long d1 = wide ? mant : (unsigned long long)__double_as_longlong(t) - 0x4330000000000000ull;
    // e on the same grid (its own sign stays in place; a zero stays a zero)
    const unsigned ehi = (unsigned)__double2hiint(e), elo = (unsigned)__double2loint(e);
    const bool ez = ((ehi & 0x7fffffffu) | elo) == 0u;
    const double es = ez ? 0.0 : __hiloint2double((int)(ehi - shift), (int)elo);
    const double t3 = __dadd_rn(es, 1.5);
    const double r3 = __dsub_rn(es, __dsub_rn(t3, 1.5));
    const double t4 = __dadd_rn(r3, 3.3306690738754696e-16);              // 1.5 * 2^-52
    const unsigned long long dmid = (unsigned long long)__double_as_longlong(t2) + (unsigned long long)__double_as_longlong(t3) -
                                    2ull * 0x3FF8000000000000ull;
    const unsigned long long dlow = (unsigned long long)__double_as_longlong(t4) - 0x3CB8000000000000ull;
    const unsigned long long sm = (unsigned long long)((long long)(int)hi >> 31);     // all ones when p < 0
    const unsigned a2 = col + J * stride;
    const unsigned a1 = a2 - stride, a0 = a1 - stride;
    const unsigned long long v0 = lds64(a0), v1 = lds64(a1), v2 = lds64(a2);
    sts64(a0, v0 + dlow);
    sts64(a1, v1 + dmid);
    sts64(a2, v2 + (d1 ^ sm) - sm);
}

// Bound a thread-private column in place WITHOUT a carry chain: every limb keeps its low 52 bits
// and receives the carry-save bits of the limb below (one step, no propagation).  The value is
// unchanged and afterwards |limb| < 2^52 + 2^11, which is all the deposit budget (2046 more
// digits) and the block merge (sum of <= 1024 columns) need.  Unlike a full normalisation the 39
// steps are independent, so the shared-memory latency pipelines instead of adding up.
EXB_D void bound_column(unsigned col, unsigned stride) {
    long long carry = 0;
    unsigned a = col;
#pragma unroll
    for (int j = 0; j < kLimbs - 1; ++j, a += stride) {
        const long long v = (long long)lds64(a);
        sts64(a, (unsigned long long)((v & kLimbMask) + carry));
        carry = v >> kDigits;
    }
    sts64(a, lds64(a) + (unsigned long long)carry);
}
#endif  // __CUDACC__

}  // namespace exb
