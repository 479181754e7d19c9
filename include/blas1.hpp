/*
 * blas1.hpp -- drop-in replacement for the reference's include/blas1.hpp (ExSUM / ExDOT only).
 *
 * Same free functions, argument order and defaults as reference include/blas1.hpp:48 and :74, so
 * code written against the reference (tests/test.exsum.gpu.cpp, tests/test.exdot.gpu.cpp, the
 * examples) compiles unchanged and links against libexblas_b200.so, which implements them on top
 * of the C ABI in exblas_b200.h with hand-written sm_100a kernels.
 *
 * Semantics follow the reference's GPU implementation (src/gpu/blas/blas1/ExSUM.cpp:64-84,
 * ExDOT.cpp:69-92): Ng elements ag[offset + i*inca]; fpe < 2 (exsum) / fpe < 3 (exdot) uses
 * superaccumulators only; early_exit buckets fpe to 4 / 6 / 8; `parallel` is ignored.  ag / bg may
 * be host pointers (as in the reference) or device pointers.  fpe < 0 prints the reference's
 * message and exits with status 1 (cpu ExSUM.cpp:25-28).  The value is rounded exactly like the
 * reference's Superaccumulator::Round(); set EXBLAS_B200_ROUND=exact for correct rounding.
 */
#ifndef BLAS1_HPP_
#define BLAS1_HPP_

double exsum(const int Ng, double *ag, const int inca, const int offset, const int fpe,
             const bool early_exit = false, const bool parallel = true);

double exdot(const int Ng, double *ag, const int inca, const int offseta, double *bg, const int incb,
             const int offsetb, const int fpe, const bool early_exit = false);

#endif  // BLAS1_HPP_
