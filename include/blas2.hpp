/*
 * blas2.hpp -- drop-in replacement for the reference's include/blas2.hpp, ExGEMV only
 * (reference include/blas2.hpp:95).  transa == 'N' (BASELINE config 5) and 'T'; extrsv() is not part
 * of this hot path.
 *
 * y := alpha*A*x + beta*y with every element the rounded exact value.  A is column-major, m x n,
 * leading dimension lda.  Pointers may be host (as in the reference) or device pointers.
 */
#ifndef BLAS2_HPP_
#define BLAS2_HPP_

int exgemv(const char transa, const int m, const int n, const double alpha, double *a, const int lda,
           const int offseta, double *x, const int incx, const int offsetx, const double beta, double *y,
           const int incy, const int offsety, const int fpe, const bool early_exit = false);

#endif  // BLAS2_HPP_
