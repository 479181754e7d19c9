// microbench.cuh -- the two measured ceilings the ExSUM / ExDOT roofline is taken against (SURVEY section 8d:
// "the roofline is the slower of the two"), measured on THIS GPU in THIS run, under the clocks of the moment:
//   * FP64 pipe: DADD lane-instructions per second (8 independent chains per thread, one 1024-thread CTA per SM);
//   * HBM read-only stream: the same 256-bit L1-bypassing loads as the reduction kernels, one DADD per element.
// bench.py calls them through exblas_b200_microbench(); they are diagnostics, not part of the reduction path.
#pragma once
#include "reduce_kernel.cuh"

namespace exb {

__global__ void __launch_bounds__(1024, 1) mb_dadd_kernel(double* out, int iters, double seed) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + k + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __dadd_rn(a[k], seed);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.678) out[0] = s;          // never true: keeps the chains alive
}

__global__ void __launch_bounds__(512) mb_read_kernel(const double* __restrict__ a, long long nvec, double* out) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const Vec4 v0 = ldg256(a + 4 * i), v1 = ldg256(a + 4 * (i + stride)), v2 = ldg256(a + 4 * (i + 2 * stride)),
                   v3 = ldg256(a + 4 * (i + 3 * stride));
        s0 += v0.x + v0.y + v0.z + v0.w;
        s1 += v1.x + v1.y + v1.z + v1.w;
        s2 += v2.x + v2.y + v2.z + v2.w;
        s3 += v3.x + v3.y + v3.z + v3.w;
    }
    for (; i < nvec; i += stride) {
        const Vec4 v = ldg256(a + 4 * i);
        s0 += v.x + v.y + v.z + v.w;
    }
    const double s = s0 + s1 + s2 + s3;
    if (s == 12345.678) out[0] = s;
}

}  // namespace exb
