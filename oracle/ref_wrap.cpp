// oracle/ref_wrap.cpp -- extern "C" handles onto the UNMODIFIED reference CPU code.
//
// TEST INFRASTRUCTURE ONLY (never linked into, or called by, the product path).
// Compiled by oracle/Makefile together with the reference's own sources, taken
// from where they lie under /root/reference, into oracle/_ref/libexblas_ref.so:
//     src/cpu/blas/blas1/ExSUM.cpp, src/cpu/blas/blas1/superaccumulator.cpp,
//     src/common/common.cpp
// Nothing of the reference is copied into this repository; this file only calls it.
//
// What is exposed:
//   ref_exsum            -> exsum()                       include/blas1.hpp:48
//   ref_superacc_limbs   -> Superaccumulator::Accumulate(double)/Normalize()
//                                                         superaccumulator.hpp:173-194, .cpp:138-162
//   ref_round_limbs      -> Superaccumulator::Round()      superaccumulator.cpp:80-134
//   ref_exdot_superacc   -> the reference's only exact statement of ExDOT
//                           (src/gpu/blas/blas1/ExDOT.Superacc.cl:244-253) executed on the
//                           reference's CPU Superaccumulator class (no CPU exdot exists in the fork)
//   ref_exsum_mpfr / ref_exdot_mpfr -> the reference tests' MPFR checkers
//                           (tests/test.exsum.cpu.cpp:24-38, tests/test.exdot.gpu.cpp:24-46),
//                           MPFR prototypes declared by hand because mpfr.h is absent.
//   ref_init_*           -> the reference input generators  src/common/common.cpp:30,113,147
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <omp.h>

#include "blas1.hpp"
#include "common.hpp"
#include "superaccumulator.hpp"

// ---- hand-declared MPFR 4.x ABI (libmpfr.so.6) --------------------------------------------
extern "C" {
typedef struct {
    long _mpfr_prec;
    int _mpfr_sign;
    long _mpfr_exp;
    unsigned long* _mpfr_d;
} ob_mpfr_struct;
typedef ob_mpfr_struct ob_mpfr_t[1];
enum { OB_MPFR_RNDN = 0 };
void mpfr_init2(ob_mpfr_struct*, long);
void mpfr_clear(ob_mpfr_struct*);
void mpfr_set_zero(ob_mpfr_struct*, int);
int mpfr_set_d(ob_mpfr_struct*, double, int);
int mpfr_add_d(ob_mpfr_struct*, const ob_mpfr_struct*, double, int);
int mpfr_mul_d(ob_mpfr_struct*, const ob_mpfr_struct*, double, int);
int mpfr_add(ob_mpfr_struct*, const ob_mpfr_struct*, const ob_mpfr_struct*, int);
double mpfr_get_d(const ob_mpfr_struct*, int);
void mpfr_free_cache(void);
}

extern "C" {

int ref_omp_max_threads(void) { return omp_get_max_threads(); }
void ref_omp_set_threads(int n) { omp_set_num_threads(n); }

int ref_limb_count(void) {
    Superaccumulator s(e_bits, f_bits);
    return s.get_f_words() + s.get_e_words();
}

int ref_f_words(void) {
    Superaccumulator s(e_bits, f_bits);
    return s.get_f_words();
}

double ref_exsum(int n, double* a, int inca, int offset, int fpe, int early_exit, int parallel) {
    return exsum(n, a, inca, offset, fpe, early_exit != 0, parallel != 0);
}

// Exact accumulation of a[0..n) with the reference class; writes the normalised limbs.
// Returns Round() of the same accumulator.
double ref_superacc_limbs(const double* a, long n, int64_t* limbs_out) {
    Superaccumulator acc(e_bits, f_bits);
    for (long i = 0; i < n; ++i) acc.Accumulate(a[i]);
    acc.Normalize();
    std::vector<int64_t> v = acc.get_accumulator();
    for (size_t i = 0; i < v.size(); ++i) limbs_out[i] = v[i];
    return acc.Round();
}

double ref_round_limbs(const int64_t* limbs, int count) {
    std::vector<int64_t> v(limbs, limbs + count);
    Superaccumulator acc(v, e_bits, f_bits);
    return acc.Round();
}

// ExDOT.Superacc.cl:244-253 on the CPU class: x = a*b, r = fma(a,b,-x), both accumulated exactly.
double ref_exdot_superacc(long n, const double* a, const double* b, int64_t* limbs_out) {
    Superaccumulator acc(e_bits, f_bits);
    for (long i = 0; i < n; ++i) {
        double x = a[i] * b[i];
        double r = std::fma(a[i], b[i], -x);
        acc.Accumulate(x);
        if (r != 0.0) acc.Accumulate(r);
    }
    acc.Normalize();
    if (limbs_out) {
        std::vector<int64_t> v = acc.get_accumulator();
        for (size_t i = 0; i < v.size(); ++i) limbs_out[i] = v[i];
    }
    return acc.Round();
}

double ref_exsum_mpfr(long n, const double* a) {
    ob_mpfr_t acc;
    mpfr_init2(acc, 2098);
    mpfr_set_zero(acc, 0);
    for (long i = 0; i < n; ++i) mpfr_add_d(acc, acc, a[i], OB_MPFR_RNDN);
    double d = mpfr_get_d(acc, OB_MPFR_RNDN);
    mpfr_clear(acc);
    return d;
}

double ref_exdot_mpfr(long n, const double* a, const double* b) {
    ob_mpfr_t sum, dot, op;
    mpfr_init2(op, 64);
    mpfr_init2(dot, 128);
    mpfr_init2(sum, 4196);
    mpfr_set_zero(dot, 0);
    mpfr_set_zero(sum, 0);
    for (long i = 0; i < n; ++i) {
        mpfr_set_d(op, a[i], OB_MPFR_RNDN);
        mpfr_mul_d(dot, op, b[i], OB_MPFR_RNDN);
        mpfr_add(sum, sum, dot, OB_MPFR_RNDN);
    }
    double d = mpfr_get_d(sum, OB_MPFR_RNDN);
    mpfr_clear(op);
    mpfr_clear(dot);
    mpfr_clear(sum);
    mpfr_free_cache();
    return d;
}

// tests/test.exgemv.gpu.cpp:35-78 exgemvVsMPFR (column-major), returning the per-element MPFR
// values instead of the norm: dot = a (128 bit) * alpha * x, sum (2098 bit) += dot, + beta*y.
// Exact (hence correctly rounded) whenever alpha == 1 (a*x needs 106 <= 128 bits).
void ref_exgemv_mpfr(char trans, int m, int n, double alpha, const double* a, int lda, const double* x, int incx,
                     double beta, const double* y, int incy, double* out) {
    ob_mpfr_t sum, dot;
    mpfr_init2(dot, 128);
    mpfr_init2(sum, 2098);
    const bool t = (trans == 'T');
    const int nout = t ? n : m, nin = t ? m : n;
    for (int i = 0; i < nout; ++i) {
        mpfr_set_d(sum, 0.0, OB_MPFR_RNDN);
        for (int j = 0; j < nin; ++j) {
            mpfr_set_d(dot, t ? a[i * lda + j] : a[j * lda + i], OB_MPFR_RNDN);
            mpfr_mul_d(dot, dot, alpha, OB_MPFR_RNDN);
            mpfr_mul_d(dot, dot, x[j * incx], OB_MPFR_RNDN);
            mpfr_add(sum, sum, dot, OB_MPFR_RNDN);
        }
        mpfr_set_d(dot, y[i * incy], OB_MPFR_RNDN);
        mpfr_mul_d(dot, dot, beta, OB_MPFR_RNDN);
        mpfr_add(sum, sum, dot, OB_MPFR_RNDN);
        out[i] = mpfr_get_d(sum, OB_MPFR_RNDN);
    }
    mpfr_clear(dot);
    mpfr_clear(sum);
    mpfr_free_cache();
}

void ref_srand(unsigned seed) { srand(seed); }
void ref_init_naive(int n, double* a) { init_naive(n, a); }
void ref_init_fpuniform(int n, double* a, int range, int emax) { init_fpuniform(n, a, range, emax); }
void ref_init_ill_cond(int n, double* a, double c) { init_ill_cond(n, a, c); }

}  // extern "C"
