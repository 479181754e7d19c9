/* oracle/exblas_oracle.c -- CPU restatement of the reference's ExSUM / ExDOT algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's library.  The product
 * path (exblas_b200/csrc) never links, calls or falls back to anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 * (a) the unmodified reference CPU code compiled into oracle/_ref/libexblas_ref.so
 *     (exsum(), Superaccumulator limbs, Round()), and
 * (b) the golden vectors under tests/golden/ that were generated from that library
 *     (tests/golden/make_golden.py), so the pin also holds where /root/reference is absent.
 * The reference ships no golden vectors of its own.
 *
 * Plain scalar C: each function names the reference file:line it restates
 * (paths relative to /root/reference).  Nothing is copied: the reference is C++ with x86
 * inline asm / AVX intrinsics (CPU) and OpenCL C (GPU); this is portable C99 written from
 * the algorithm.
 *
 * Two limb layouts exist in the reference:
 *   layout 0 "cpu": 41 limbs, f_words = 21   (src/cpu/blas/blas1/superaccumulator.cpp:14-17)
 *   layout 1 "gpu": 39 limbs, f_words = 20   (src/gpu/blas/blas1/ExSUM.FPE.cl:14-18)
 * limb i weighs 2^(52*(i - f_words)); gpu limb j == cpu limb j+1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OB_DIGITS 52                 /* superaccumulator.hpp:118-119: K = 12, digits = 64 - K */
#define OB_K 12
#define OB_MAXLIMBS 41
#define OB_MASK ((((int64_t)1) << OB_DIGITS) - 1)

typedef struct {
    int64_t limb[OB_MAXLIMBS];
    int nl;      /* number of limbs */
    int fw;      /* f_words */
    int imin, imax;
    int overflow;
} ob_acc;

void ob_init(ob_acc* s, int layout) {
    /* superaccumulator.cpp:14-22 (cpu) / ExSUM.FPE.cl:14-18 (gpu) */
    memset(s, 0, sizeof(*s));
    if (layout == 0) { s->nl = 41; s->fw = 21; } else { s->nl = 39; s->fw = 20; }
    s->imin = 0;
    s->imax = s->nl - 1;
}

/* mylibm.hpp:182-198 xadd + seto: returns the old word, reports signed overflow */
static int64_t ob_xadd(int64_t* mem, int64_t x, int* of) {
    int64_t old = *mem;
    int64_t sum;
    *of = __builtin_add_overflow(old, x, &sum);
    *mem = sum;
    return old;
}

/* superaccumulator.hpp:132-171 AccumulateWord (TSAFE == 0) */
static void ob_accumulate_word(ob_acc* s, int64_t x, int i) {
    int64_t carry = x;
    int of;
    int64_t oldword = ob_xadd(&s->limb[i], x, &of);
    while (of) {
        carry = (oldword + carry) >> OB_DIGITS;           /* arithmetic shift; wraps like the asm */
        int positive = oldword > 0;
        int64_t carrybit = positive ? ((int64_t)1 << OB_K) : -((int64_t)1 << OB_K);
        int of2;
        ob_xadd(&s->limb[i], (int64_t)(0 - ((uint64_t)carry << OB_DIGITS)), &of2);
        carry += carrybit;
        ++i;
        if (i >= s->nl) { s->overflow = 1; return; }
        oldword = ob_xadd(&s->limb[i], carry, &of);
    }
}

/* mylibm.hpp:107-118 exponent(): unbiased exponent field */
static int ob_exponent(double x) {
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int)((u >> 52) & 0x7ff) - 0x3ff;
}

/* mylibm.hpp:130-141 myldexp(): add e to the exponent field, no checks */
static double ob_ldexp_bits(double x, int e) {
    uint64_t u;
    memcpy(&u, &x, 8);
    u += (uint64_t)(int64_t)e << 52;
    memcpy(&x, &u, 8);
    return x;
}

/* superaccumulator.hpp:173-194 Accumulate(double): scale / rint / subtract digit loop */
void ob_accumulate(ob_acc* s, double x) {
    if (x == 0) return;
    int e = ob_exponent(x);
    int exp_word = e / OB_DIGITS;                 /* C division truncates toward zero, as in the reference */
    int iup = exp_word + s->fw;
    double xscaled = ob_ldexp_bits(x, -OB_DIGITS * exp_word);
    for (int i = iup; xscaled != 0; --i) {
        double xrounded = rint(xscaled);          /* roundsd imm 0: nearest-even (mylibm.hpp:95) */
        int64_t xint = llrint(xscaled);           /* cvtsd2si (mylibm.hpp:60-62) */
        if (i < 0 || i >= s->nl) { s->overflow = 1; return; }   /* reference: out-of-bounds write (undefined) */
        ob_accumulate_word(s, xint, i);
        xscaled -= xrounded;
        xscaled *= 4503599627370496.0;            /* deltaScale = 2^52 */
    }
}

/* superaccumulator.cpp:138-162 Normalize(): returns the sign (1 = negative) */
int ob_normalize(ob_acc* s) {
    if (s->imin > s->imax) return 0;
    int64_t carry_in = s->limb[s->imin] >> OB_DIGITS;
    s->limb[s->imin] -= carry_in << OB_DIGITS;
    int i;
    for (i = s->imin + 1; i < s->nl; ++i) {
        s->limb[i] += carry_in;
        int64_t carry_out = s->limb[i] >> OB_DIGITS;
        s->limb[i] -= carry_out << OB_DIGITS;
        carry_in = carry_out;
    }
    s->imax = i - 1;
    s->limb[s->imax] += (int64_t)((uint64_t)carry_in << OB_DIGITS);   /* top limb keeps the signed carry */
    return carry_in < 0;
}

/* superaccumulator.cpp:68-78 Accumulate(Superaccumulator&) */
void ob_merge(ob_acc* dst, ob_acc* src) {
    ob_normalize(dst);
    ob_normalize(src);
    if (src->imin < dst->imin) dst->imin = src->imin;
    if (src->imax > dst->imax) dst->imax = src->imax;
    for (int i = dst->imin; i <= dst->imax; ++i) dst->limb[i] += src->limb[i];
}

/* mylibm.hpp:156-171 */
static double ob_odd_round_sum_nonnegative(double th, double tl) {
    double d = th + tl;
    int64_t l;
    memcpy(&l, &d, 8);
    l |= (tl != 0.0);
    memcpy(&d, &l, 8);
    return d;
}

/* superaccumulator.cpp:80-134 Round(), line by line -- including the behaviour SURVEY.md
 * section 0.2 documents (sticky bit inside the 53-bit result when the top limb is tiny,
 * one's-complement hiword for negative sums).  This is the reference-parity finaliser. */
double ob_round_ref(ob_acc* s) {
    if (s->imin > s->imax) return 0;
    int negative = ob_normalize(s);
    int i;
    for (i = s->imax; i >= s->imin && s->limb[i] == 0; --i) {}
    if (negative) {
        for (; i >= s->imin && (s->limb[i] & OB_MASK) == OB_MASK; --i) {}
    }
    if (i < 0) return 0.0;
    int64_t hiword = negative ? OB_MASK - s->limb[i] : s->limb[i];
    double rounded = (double)hiword;
    double hi = ldexp(rounded, (i - s->fw) * OB_DIGITS);
    if (i == 0) return negative ? -hi : hi;
    hiword -= llrint(rounded);
    double mid = ldexp((double)hiword, (i - s->fw) * OB_DIGITS);
    int64_t sticky = 0;
    for (int j = s->imin; j != i - 1; ++j)
        sticky |= negative ? ((int64_t)1 << OB_DIGITS) - s->limb[j] : s->limb[j];
    int64_t loword = negative ? ((int64_t)1 << OB_DIGITS) - s->limb[i - 1] : s->limb[i - 1];
    loword |= !!sticky;
    double lo = ldexp((double)loword, (i - 1 - s->fw) * OB_DIGITS);
    if (mid != 0) lo = ob_odd_round_sum_nonnegative(mid, lo);
    hi = hi + lo;
    return negative ? -hi : hi;
}

/* Exact round-to-nearest-even of the accumulator value.  NOT in the reference (its Round() above
 * is not always correctly rounded); pinned against math.fsum / MPFR in tests/test_oracle.py.
 * Works on a sign-magnitude copy of the normalised limbs. */
double ob_round_exact(ob_acc* s) {
    int negative = ob_normalize(s);
    uint64_t m[OB_MAXLIMBS];
    int nl = s->nl;
    /* magnitude: two's-complement negate across limbs when negative */
    if (!negative) {
        for (int i = 0; i < nl; ++i) m[i] = (uint64_t)s->limb[i];
    } else {
        int64_t borrow = 0;
        for (int i = 0; i < nl - 1; ++i) {
            int64_t v = -s->limb[i] + borrow;        /* in (-2^52, 0] */
            borrow = v >> OB_DIGITS;                 /* 0 or -1 */
            m[i] = (uint64_t)(v - borrow * ((int64_t)1 << OB_DIGITS));
        }
        m[nl - 1] = (uint64_t)(-s->limb[nl - 1] + borrow);
    }
    int top = nl - 1;
    while (top >= 0 && m[top] == 0) --top;
    if (top < 0) return 0.0;
    int lsb_exp = -OB_DIGITS * s->fw;
    int P = OB_DIGITS * top + (63 - __builtin_clzll(m[top]));    /* MSB position above limb-0 LSB */
    int e = P + lsb_exp;                                        /* exponent of the MSB */
    int keep = 53;
    if (e < -1022) keep = 53 - (-1022 - e);                     /* subnormal result: fewer bits */
    int c = P - keep + 1;                                       /* lowest kept bit position */
    /* gather bits [c, P] into q (keep <= 53 bits), round bit c-1, sticky below */
    uint64_t q = 0;
    int rnd = 0, sticky = 0;
    for (int pos = P; pos >= 0; --pos) {
        int limb = pos / OB_DIGITS, bit = pos % OB_DIGITS;
        int b;
        if (limb >= nl - 1) { limb = nl - 1; bit = pos - OB_DIGITS * (nl - 1); }
        b = (int)((m[limb] >> bit) & 1);
        if (pos >= c) q = (q << 1) | (uint64_t)b;
        else if (pos == c - 1) rnd = b;
        else if (b) { sticky = 1; break; }
    }
    /* keep <= 0 (deep subnormal): c - 1 >= P, so the loop above already put the MSB in rnd
     * (keep == 0) or everything in sticky (keep < 0), with q == 0. */
    if (rnd && (sticky || (q & 1))) q += 1;
    double r = ldexp((double)q, c + lsb_exp);
    return negative ? -r : r;
}

/* ExSUM.FPE.cl:27-32 KnuthTwoSum (USE_KNUTH build, ExSUM.Launcher.cpp:41).
 * volatile keeps gcc from contracting or re-associating; -ffp-contract=off is also set. */
static double ob_two_sum(double a, double b, double* s) {
    volatile double r = a + b;
    volatile double z = r - a;
    volatile double t1 = r - z;
    volatile double t2 = a - t1;
    volatile double t3 = b - z;
    *s = t2 + t3;
    return r;
}

/* ExDOT.FPE.cl:25-29 TwoProductFMA */
static double ob_two_prod(double a, double b, double* r) {
    double x = a * b;
    *r = fma(a, b, -x);
    return x;
}

/* one element through the expansion: ExSUM.FPE.cl:257-295 (plain) / ExSUM.FPE.EX.4.cl:253-291 (early exit).
 * first_level lets ExDOT feed the product error at level fpe-3 (ExDOT.FPE.cl:254-258). */
static void ob_fpe_push(ob_acc* s, double* a, int fpe, int early_exit, int first_level, double x) {
    for (int i = first_level; i < fpe; ++i) {
        double r;
        a[i] = ob_two_sum(a[i], x, &r);
        x = r;
        if (early_exit && x == 0.0) break;
    }
    if (x != 0.0) {
        ob_accumulate(s, x);
        for (int i = 0; i < fpe; ++i) { ob_accumulate(s, a[i]); a[i] = 0.0; }
    }
}

static int ob_fpe_size(int fpe, int early_exit, int min_fpe) {
    /* cpu ExSUM.cpp:72-99, gpu ExSUM.cpp:70-83, ExDOT.cpp:78-89.  fpe > 8 is clamped to 8
     * (documented deviation: the reference returns 0.0 there). */
    if (fpe < min_fpe) return 0;
    if (early_exit) return fpe <= 4 ? 4 : (fpe <= 6 ? 6 : 8);
    return fpe > 8 ? 8 : fpe;
}

static void ob_export39(ob_acc* s, int64_t* limbs39) {
    if (!limbs39) return;
    ob_normalize(s);
    for (int i = 0; i < 39; ++i) limbs39[i] = s->limb[i];
}

/* exsum with the GPU argument meaning: n elements a[offset + i*inca].
 * round_mode 0 = reference Round(), 1 = exact RN-even.  Uses the 39-limb layout. */
double oracle_exsum(int64_t n, const double* a, int64_t inca, int64_t offset, int fpe, int early_exit,
                    int round_mode, int64_t* limbs39) {
    ob_acc s;
    ob_init(&s, 1);
    int f = ob_fpe_size(fpe, early_exit, 2);
    double e[8] = {0};
    for (int64_t i = 0; i < n; ++i) {
        double x = a[offset + i * inca];
        if (f == 0) ob_accumulate(&s, x);                       /* ExSUM.Superacc.cl:232-247 */
        else ob_fpe_push(&s, e, f, early_exit, 0, x);
    }
    for (int i = 0; i < f; ++i) ob_accumulate(&s, e[i]);        /* ExSUM.FPE.cl:355-357 */
    ob_export39(&s, limbs39);
    return round_mode ? ob_round_exact(&s) : ob_round_ref(&s);
}

/* exdot: ExDOT.Superacc.cl:244-253 (fpe < 3), ExDOT.FPE.cl:226-271, ExDOT.FPE.EX.4.cl */
double oracle_exdot(int64_t n, const double* a, int64_t inca, int64_t offa, const double* b, int64_t incb,
                    int64_t offb, int fpe, int early_exit, int round_mode, int64_t* limbs39) {
    ob_acc s;
    ob_init(&s, 1);
    if (n <= 0) { ob_export39(&s, limbs39); return 0.0; }       /* ExDOT.cpp:70-71 */
    int f = ob_fpe_size(fpe, early_exit, 3);
    double e[8] = {0};
    for (int64_t i = 0; i < n; ++i) {
        double r;
        double x = ob_two_prod(a[offa + i * inca], b[offb + i * incb], &r);
        if (f == 0) {
            ob_accumulate(&s, x);
            if (r != 0.0) ob_accumulate(&s, r);
        } else {
            ob_fpe_push(&s, e, f, early_exit, 0, x);
            if (r != 0.0) ob_fpe_push(&s, e, f, early_exit, f - 3, r);
        }
    }
    for (int i = 0; i < f; ++i) ob_accumulate(&s, e[i]);
    ob_export39(&s, limbs39);
    return round_mode ? ob_round_exact(&s) : ob_round_ref(&s);
}

/* exgemv: y := alpha*op(A)*x + beta*y, A column-major with leading dimension lda.
 * Per output element the reference kernel (src/gpu/blas/blas2/ExGEMV.FPE.cl:244-288; 'T':
 * :420-470) runs TwoProductFMA, feeds both parts through ALL levels of the row's expansion
 * (first level 0, :272-276), flushes like ExSUM, then adds beta*y exactly (:346-377: beta == 0
 * nothing, beta == 1 the value, else TwoProductFMA) and rounds.  The reference's non-transpose
 * FPE kernel ignores alpha (:246) and its tests only use alpha == 1; here alpha is applied exactly
 * (alpha*a = p1 + e1 by TwoProd, then (p1 + e1)*x by two TwoProds), which coincides with the
 * reference for alpha == 1.  fpe <= 1 -> superaccumulators only (fpe == 1 is a plain DGEMV in the
 * reference, ExGEMV.cpp:92-94: not an exact algorithm, not restated). */
static void ob_gemv_term(ob_acc* s, double* e, int f, int ee, double av, double xv) {
    double r;
    double p = ob_two_prod(av, xv, &r);
    if (f == 0) {
        ob_accumulate(s, p);
        if (r != 0.0) ob_accumulate(s, r);
    } else {
        ob_fpe_push(s, e, f, ee, 0, p);
        if (r != 0.0) ob_fpe_push(s, e, f, ee, 0, r);
    }
}

int oracle_exgemv(char trans, int64_t m, int64_t n, double alpha, const double* a, int64_t lda, const double* x,
                  int64_t incx, double beta, double* y, int64_t incy, int fpe, int early_exit, int round_mode) {
    const int t = (trans == 'T' || trans == 't');
    const int64_t nout = t ? n : m, nin = t ? m : n;
    int f = fpe <= 1 ? 0 : ob_fpe_size(fpe, early_exit, 2);
    for (int64_t i = 0; i < nout; ++i) {
        ob_acc s;
        ob_init(&s, 1);
        double e[8] = {0};
        for (int64_t j = 0; j < nin; ++j) {
            const double av = t ? a[i * lda + j] : a[j * lda + i];
            const double xv = x[j * incx];
            if (alpha == 1.0) {
                ob_gemv_term(&s, e, f, early_exit, av, xv);
            } else {
                double e1;
                double p1 = ob_two_prod(alpha, av, &e1);
                ob_gemv_term(&s, e, f, early_exit, p1, xv);
                if (e1 != 0.0) ob_gemv_term(&s, e, f, early_exit, e1, xv);
            }
        }
        for (int k = 0; k < f; ++k) ob_accumulate(&s, e[k]);
        if (beta != 0.0) {
            if (beta == 1.0) {
                ob_accumulate(&s, y[i * incy]);
            } else {
                double r;
                double p = ob_two_prod(beta, y[i * incy], &r);
                ob_accumulate(&s, p);
                if (r != 0.0) ob_accumulate(&s, r);
            }
        }
        y[i * incy] = round_mode ? ob_round_exact(&s) : ob_round_ref(&s);
    }
    return 0;
}

/* Exact accumulation on a chosen layout, exporting all limbs (for comparison with the reference class). */
double oracle_superacc_limbs(int64_t n, const double* a, int layout, int64_t* limbs_out, int round_mode) {
    ob_acc s;
    ob_init(&s, layout);
    for (int64_t i = 0; i < n; ++i) ob_accumulate(&s, a[i]);
    ob_normalize(&s);
    if (limbs_out) for (int i = 0; i < s.nl; ++i) limbs_out[i] = s.limb[i];
    return round_mode ? ob_round_exact(&s) : ob_round_ref(&s);
}

double oracle_round_limbs(const int64_t* limbs, int layout, int round_mode) {
    ob_acc s;
    ob_init(&s, layout);
    for (int i = 0; i < s.nl; ++i) s.limb[i] = limbs[i];
    return round_mode ? ob_round_exact(&s) : ob_round_ref(&s);
}

/* Multi-rank merge as the reference's MPI path does it (cpu ExSUM.cpp:266-273): limb-wise
 * integer sum of normalised per-rank limbs, then Round() on the sum. */
double oracle_merge_round(const int64_t* limbs, int nranks, int layout, int round_mode, int64_t* merged_out) {
    ob_acc s;
    ob_init(&s, layout);
    for (int r = 0; r < nranks; ++r)
        for (int i = 0; i < s.nl; ++i) s.limb[i] += limbs[(size_t)r * s.nl + i];
    ob_normalize(&s);
    if (merged_out) for (int i = 0; i < s.nl; ++i) merged_out[i] = s.limb[i];
    return round_mode ? ob_round_exact(&s) : ob_round_ref(&s);
}

/* ---- input generators (common.cpp) restated on a private LCG-free interface: the reference uses
 * libc rand(); for portable fixed-seed inputs the tests generate with numpy instead and only use
 * these for the exact formulas. ---- */

/* common.cpp:147-150 */
void oracle_init_naive(int64_t n, double* a) {
    for (int64_t i = 0; i < n; ++i) a[i] = 1.1;
}

/* OpenMP-parallel version of oracle_exsum used only as the "port" CPU baseline in bench.py when
 * oracle/_ref is unavailable: contiguous slices, one accumulator per thread, limb-wise merge
 * (cpu ExSUM.cpp:235-264). */
#ifdef _OPENMP
#include <omp.h>
int oracle_max_threads(void) { return omp_get_max_threads(); }
void oracle_set_threads(int n) { omp_set_num_threads(n); }
double oracle_exsum_parallel(int64_t n, const double* a, int fpe, int early_exit, int round_mode) {
    int T = omp_get_max_threads();
    ob_acc* accs = (ob_acc*)malloc(sizeof(ob_acc) * (size_t)T);
    int f = ob_fpe_size(fpe, early_exit, 2);
    for (int k = 0; k < T; ++k) ob_init(&accs[k], 1);
#pragma omp parallel num_threads(T)
    {
        int t = omp_get_thread_num();
        int nt = omp_get_num_threads();
        ob_acc* s = &accs[t];
        int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        double e[8] = {0};
        for (int64_t i = lo; i < hi; ++i) {
            if (f == 0) ob_accumulate(s, a[i]);
            else ob_fpe_push(s, e, f, early_exit, 0, a[i]);
        }
        for (int i = 0; i < f; ++i) ob_accumulate(s, e[i]);
        ob_normalize(s);
    }
    for (int t = 1; t < T; ++t) ob_merge(&accs[0], &accs[t]);
    double r = round_mode ? ob_round_exact(&accs[0]) : ob_round_ref(&accs[0]);
    free(accs);
    return r;
}
#else
int oracle_max_threads(void) { return 1; }
void oracle_set_threads(int n) { (void)n; }
double oracle_exsum_parallel(int64_t n, const double* a, int fpe, int early_exit, int round_mode) {
    return oracle_exsum(n, a, 1, 0, fpe, early_exit, round_mode, 0);
}
#endif
