/* Hand-declared subset of the MPFR 4.x API (mpfr.h is not installed; libmpfr.so.6 is), enough for
 * the reference tests' -DEXBLAS_VS_MPFR checkers.  TEST INFRASTRUCTURE ONLY. */
#ifndef EXBLAS_B200_ORACLE_MPFR_H
#define EXBLAS_B200_ORACLE_MPFR_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct {
    long _mpfr_prec;
    int _mpfr_sign;
    long _mpfr_exp;
    unsigned long* _mpfr_d;
} __mpfr_struct;
typedef __mpfr_struct mpfr_t[1];
typedef __mpfr_struct* mpfr_ptr;
typedef const __mpfr_struct* mpfr_srcptr;
typedef enum { MPFR_RNDN = 0, MPFR_RNDZ, MPFR_RNDU, MPFR_RNDD, MPFR_RNDA } mpfr_rnd_t;
void mpfr_init2(mpfr_ptr, long);
void mpfr_clear(mpfr_ptr);
void mpfr_set_zero(mpfr_ptr, int);
int mpfr_set_d(mpfr_ptr, double, mpfr_rnd_t);
int mpfr_add_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_mul_d(mpfr_ptr, mpfr_srcptr, double, mpfr_rnd_t);
int mpfr_add(mpfr_ptr, mpfr_srcptr, mpfr_srcptr, mpfr_rnd_t);
double mpfr_get_d(mpfr_srcptr, mpfr_rnd_t);
void mpfr_free_cache(void);
#ifdef __cplusplus
}
#endif
#endif
