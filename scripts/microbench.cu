// Development helper: measured ceilings on the B200 that bound the ExSUM/ExDOT kernels.
//   1. FP64 pipe: DADD instructions/s (8 independent chains per thread, 148 x 1024 threads)
//   2. HBM read-only stream: 256-bit L1-bypassing loads, 8 GiB, one DADD per element
//   3. shared-memory read-modify-write: conflict-free LDS.64 + STS.64 pairs per second
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench scripts/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
struct alignas(32) V4 { double x, y, z, w; };
__device__ __forceinline__ V4 ldg256(const double* p) {
    V4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__global__ void dadd_kernel(double* out, int iters, double seed) {
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = seed + k + threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __dadd_rn(a[k], seed);
    }
    double s = 0;
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 12345.678) out[0] = s;
}
__global__ void read_kernel(const double* __restrict__ a, long long nvec, double* out) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        V4 v0 = ldg256(a + 4 * i), v1 = ldg256(a + 4 * (i + stride)), v2 = ldg256(a + 4 * (i + 2 * stride)), v3 = ldg256(a + 4 * (i + 3 * stride));
        s0 += v0.x + v0.y + v0.z + v0.w; s1 += v1.x + v1.y + v1.z + v1.w; s2 += v2.x + v2.y + v2.z + v2.w; s3 += v3.x + v3.y + v3.z + v3.w;
    }
    for (; i < nvec; i += stride) { V4 v = ldg256(a + 4 * i); s0 += v.x + v.y + v.z + v.w; }
    double s = s0 + s1 + s2 + s3;
    if (s == 12345.678) out[0] = s;
}
extern __shared__ long long sm[];
__global__ void rmw_kernel(int iters, long long* out) {
    const unsigned T = blockDim.x;
    for (int j = 0; j < 39; ++j) sm[j * T + threadIdx.x] = 0;
    unsigned j = threadIdx.x % 37;
    for (int i = 0; i < iters; ++i) {
        j = (j * 5 + 3) % 38;
        long long* p = sm + j * T + threadIdx.x;
        long long v0 = p[0], v1 = p[T];
        p[0] = v0 + i; p[T] = v1 - i;
        asm volatile("" ::: "memory");
    }
    if (sm[threadIdx.x] == 0x123456789) out[0] = 1;
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    double* d; cudaMalloc(&d, (size_t)1 << 33); cudaMemset(d, 0, (size_t)1 << 33);
    long long* lo; cudaMalloc(&lo, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // 1. DADD
    int iters = 20000;
    dadd_kernel<<<sms, 1024>>>(d, 100, 1.0); cudaDeviceSynchronize();
    cudaEventRecord(e0); dadd_kernel<<<sms, 1024>>>(d, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double dadd = (double)sms * 1024 * 8 * iters / (ms * 1e-3);
    printf("{\"fp64_dadd_per_s\": %.4e, \"fp64_dadd_per_clk_per_sm_at_1965MHz\": %.2f, \"ms\": %.3f}\n", dadd, dadd / sms / 1.965e9, ms);
    // 2. read stream
    long long nvec = ((size_t)1 << 33) / 32;
    for (int T : {256, 512, 1024}) {
        int blocks = sms * (2048 / T);
        read_kernel<<<blocks, T>>>(d, nvec, d); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) read_kernel<<<blocks, T>>>(d, nvec, d);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"read_stream_GBs\": %.1f, \"threads\": %d, \"blocks\": %d}\n", 5.0 * (double)((size_t)1 << 33) / (ms * 1e-3) / 1e9, T, blocks);
    }
    // 2b. the same stream sustained for ~0.5 s (power cap / clock effects), T = 512
    {
        int blocks = sms * 4;
        cudaEventRecord(e0);
        for (int r = 0; r < 400; ++r) read_kernel<<<blocks, 512>>>(d, nvec, d);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"read_stream_sustained_GBs\": %.1f, \"seconds\": %.3f}\n", 400.0 * (double)((size_t)1 << 33) / (ms * 1e-3) / 1e9, ms * 1e-3);
    }
    // 3. shared RMW
    cudaFuncSetAttribute(rmw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 39 * 8);
    rmw_kernel<<<sms, 512, 512 * 39 * 8>>>(100, lo); cudaDeviceSynchronize();
    iters = 20000;
    cudaEventRecord(e0); rmw_kernel<<<sms, 512, 512 * 39 * 8>>>(iters, lo); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double rmw = (double)sms * 512 * iters / (ms * 1e-3);
    printf("{\"smem_rmw_pairs_per_s\": %.4e, \"pairs_per_clk_per_sm_at_1965MHz\": %.3f, \"equiv_exsum_GBs\": %.1f}\n", rmw, rmw / sms / 1.965e9, rmw * 8 / 1e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
