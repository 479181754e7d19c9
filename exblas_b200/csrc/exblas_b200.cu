// exblas_b200.cu -- host side of the C ABI (include/exblas_b200.h) and the blas1.hpp wrappers.
//
// Replaces the reference's OpenCL host plumbing for this path:
//   src/gpu/blas/blas1/ExSUM.cpp:64-209, ExSUM.Launcher.cpp:47-238,
//   src/gpu/blas/blas1/ExDOT.cpp:69-223, ExDOT.Launcher.cpp:43-199
// (per-call context / queue / JIT build / whole-vector copy / two kernel launches / blocking
// read) by a persistent handle: pre-compiled sm_100a kernels, one launch per reduction, a
// self-cleaning device workspace, chunked H2D streaming for host inputs, and an exact NCCL limb
// all-reduce for the multi-GPU path (the reference's MPI_Reduce of limbs, cpu ExSUM.cpp:266-273).
//
// There is NO CPU fallback: without a CUDA device every compute entry point fails with
// EXBLAS_B200_ENOGPU / EXBLAS_B200_ECUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cctype>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include <sched.h>

#include "../../include/exblas_b200.h"
#include "reduce_kernel.cuh"
#include "gemv_kernel.cuh"
#include "segments_kernel.cuh"
#include "microbench.cuh"

using namespace exb;

// launch-shape constants (compile-time so that registers per thread can be bounded)
#ifndef EXB_MAXT
#define EXB_MAXT 512
#endif
constexpr int kMaxT = EXB_MAXT;   // largest CTA (bounds registers per thread: 65536 / kMaxT)

// 256-bit vectors in flight per thread and input stream.  Measured on B200 (profiles/): the direct
// and small-expansion kernels want a deep window (DRAM latency under load is long), the large
// expansions need the registers for a[F][2] instead.
constexpr int vectors_in_flight(int f, bool ee, bool dot) {
#ifdef EXB_USUM
    return dot ? EXB_UDOT : EXB_USUM;
#else
    // round 2, sustained (1 s, power-cap clocks) A/B at n = 2^30 (profiles/ab_sustained_r02.jsonl): six vectors per
    // thread beat eight by 2-3 % on every ExSUM variant (fewer live registers, same bytes in flight per SM as the
    // latency needs), four beat three by 6 % for the large ExDOT expansions; four vectors lose 2-4 %.
    (void)f;
    return dot ? 4 : 6;
#endif
}

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (no link-time dependency; inside a torch process this binds to the NCCL that
// torch already loaded, in a plain C program to the system libnccl.so.2)
// ------------------------------------------------------------------------------------------------
namespace {

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ struct Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
struct Id128 { char bytes[128]; };

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) return;
        api.GetUniqueId = (int (*)(void*))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(api.lib, "ncclCommInitRank");
        api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
        api.CommDestroy = (int (*)(void*))dlsym(api.lib, "ncclCommDestroy");
        api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
    });
    return api;
}
constexpr int kNcclInt64 = 4;   // ncclInt64
constexpr int kNcclSum = 0;     // ncclSum

}  // namespace

// ------------------------------------------------------------------------------------------------
// Host copy pool: PAGEABLE host input (what the reference's callers pass: new[] / _mm_malloc, tests/test.exsum.cpu.cpp:79,
// src/gpu/blas/blas1/ExSUM.cpp:126) cannot be read by the GPU's copy engines directly; the driver's own staging of a
// pageable cudaMemcpy runs at ~11 GB/s.  Here a few worker threads copy each chunk into pinned bounce buffers in
// parallel while the previous chunk is in flight over PCIe, which brings the pageable path to several times that.
// Workers are pinned to the CPUs local to the GPU (sysfs local_cpulist) when the process is allowed to run there.
// ------------------------------------------------------------------------------------------------
namespace {

class CopyPool {
public:
    CopyPool(int nthreads, const std::vector<int>& cpus) : n_(nthreads < 1 ? 1 : nthreads) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this, i, cpus] { run(i, cpus); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int threads() const { return n_; }
    // blocking parallel memcpy of up to two segments (the calling thread takes part as worker 0)
    void copy(void* d0, const void* s0, size_t b0, void* d1 = nullptr, const void* s1 = nullptr, size_t b1 = 0) {
        seg_[0] = {(char*)d0, (const char*)s0, b0};
        seg_[1] = {(char*)d1, (const char*)s1, b1};
        if (n_ == 1 || b0 + b1 < (size_t)1 << 20) {
            work(0, 1);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            pending_ = n_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        work(0, n_);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    struct Seg { char* d; const char* s; size_t b; };
    void work(int i, int n) {
        for (const Seg& g : seg_) {
            if (!g.b) continue;
            const size_t per = ((g.b + n - 1) / n + 4095) & ~(size_t)4095;
            const size_t lo = (size_t)i * per, hi = lo + per < g.b ? lo + per : g.b;
            if (lo < hi) memcpy(g.d + lo, g.s + lo, hi - lo);
        }
    }
    void run(int i, std::vector<int> cpus) {
        if (!cpus.empty()) {
            cpu_set_t set;
            CPU_ZERO(&set);
            for (int c : cpus) if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &set);
            sched_setaffinity(0, sizeof(set), &set);      // best effort: fails harmlessly outside the allowed cpuset
        }
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work(i, n_);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
    Seg seg_[2] = {};
};

// CPUs local to a PCI device: /sys/bus/pci/devices/<domain:bus:dev.fn>/local_cpulist ("0-31,64-95")
std::vector<int> local_cpus_of_device(int device) {
    std::vector<int> cpus;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
        cudaGetLastError();
        return cpus;
    }
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return cpus;
    char line[1024] = {0};
    if (fgets(line, sizeof line, f)) {
        for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
            int a = 0, b = 0;
            const int k = sscanf(tok, "%d-%d", &a, &b);
            if (k == 1) b = a;
            if (k >= 1)
                for (int c = a; c <= b && cpus.size() < 4096; ++c) cpus.push_back(c);
        }
    }
    fclose(f);
    return cpus;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct exblas_b200_handle_s {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;          // stream in use: the legacy default stream (0) or the user's
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr};
    cudaEvent_t consumed[2] = {nullptr, nullptr};
    Workspace* d_ws = nullptr;
    Result* d_res = nullptr;
    Result* h_res = nullptr;                // pinned
    double* d_stage[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [buffer][a|b]
    int64_t stage_elems = 0;
    void* d_seg_scratch = nullptr;          // batched reductions: one SegScratch per warp range (self-cleaning)
    size_t seg_scratch_bytes = 0;
    long long* d_gemv_scratch = nullptr;    // [parts][39][m] limbs + [parts][m] status words
    size_t gemv_scratch_bytes = 0;
    int64_t opt_block_threads = kMaxT;
    int64_t opt_blocks = 0;
    int64_t opt_host_chunk = (int64_t)1 << 23;
    int64_t opt_host_threads = 0;           // pageable host input: copy threads (0 = automatic, 1 = none: plain cudaMemcpy staging)
    int64_t opt_pageable_chunk = (int64_t)1 << 21;   // elements per chunk through the pinned bounce buffers
    CopyPool* pool = nullptr;
    double* h_bounce[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};   // pinned [ring][a|b]
    int64_t bounce_elems = 0;
    cudaEvent_t bounce_free[3] = {nullptr, nullptr, nullptr};
    int64_t opt_adaptive = 1;
    int64_t opt_gemv_parts = 0;
    int64_t opt_gemv_t_shape = 2;
    int64_t opt_gemv_n_shape = 1;
    int64_t opt_reduce_prefetch = 0;        // expansion kernel: L2 bulk-prefetch distance in tiles (0 = off)
    int64_t opt_track_max = (int64_t)1 << 20;   // ExSUM vectors up to this length: superaccumulator-only kernel, no unrolled body, touched rows only
    int64_t opt_dot_handoff = 16;           // ExDOT expansion kernels: tiles per CTA from which a thrashing warp hands its rows to the 5-digit window loop (0 = never)
    int64_t opt_gemv_tma = 1;               // ExGEMV 'T': stage x with TMA bulk copies (0 = plain copies by a warp; for A/B and tests)
    int64_t opt_gemv_prefetch = 2;          // ExGEMV window kernels: L2 bulk-prefetch distance in rounds (0 = off)
    int64_t opt_window = 2;                 // register window in the superaccumulator-only kernels (performance only)
    bool opt_shape_fixed = false;           // "block_threads" / "blocks" were set by hand: no size-dependent launch shape
    int64_t opt_solo_max = (int64_t)1 << 13;    // vectors up to this length: one CTA, published from shared memory
    int64_t opt_small_max = (int64_t)1 << 22;   // vectors up to this length: 256-thread CTAs
    int64_t opt_world_size = 0;             // declared number of ranks of a multi-GPU job (0 = not declared)
    int64_t opt_peer_timeout_ms = 600000;   // fused exchange: how long the closing warp waits for a peer (0 = for ever)
    unsigned long long* d_phase = nullptr;  // diagnostics: per-CTA phase stamps of the last reduction kernel
    int phase_blocks = 0;
    void* comm = nullptr;
    int nranks = 1;
    // fused peer-memory exchange (exblas_b200_peer_export / peer_attach)
    Mailbox* d_mailbox = nullptr;           // this rank's mailbox (cudaMalloc: exportable through CUDA IPC)
    Mailbox* peer_box[kMaxPeers] = {};      // every rank's mailbox mapped here (own entry = d_mailbox)
    int peer_ranks = 0, peer_rank = 0;
    unsigned long long epoch = 0;
    int64_t opt_fused = 1;
    int64_t launches = 0;
    bool acc_pending = false;               // workspace accumulator holds an unfinished (chunked) reduction
    uint32_t last_status = 0;
    std::string last_kernel;                // name of the reduction kernel launched last (bench.py reports it)
    std::string err;
};

namespace {

// cudaFuncSetAttribute is a slow driver call; do it once per kernel (per device) with the largest
// shared-memory size this library ever asks for.
cudaError_t allow_big_smem(const void* fn, int device) {
    static std::mutex mu;
    static std::unordered_set<uint64_t> done;
    const uint64_t key = (uint64_t)(uintptr_t)fn * 131u + (uint64_t)device;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count(key)) return cudaSuccess;
    // the opt-in limit (227 KB on sm_100) covers static + dynamic shared memory together
    cudaFuncAttributes attr;
    cudaError_t e = cudaFuncGetAttributes(&attr, fn);
    if (e != cudaSuccess) return e;
    int optin = 227 * 1024;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)attr.sharedSizeBytes);
    if (e == cudaSuccess) done.insert(key);
    return e;
}

int fail_cuda(exblas_b200_handle_t h, cudaError_t e, const char* what) {
    if (h) h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return EXBLAS_B200_ECUDA;
}
#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #call);     \
    } while (0)

// fpe / early_exit -> expansion size actually run (cpu ExSUM.cpp:72-99, gpu ExSUM.cpp:70-83,
// ExDOT.cpp:78-89).  Deviation: fpe > 8 is clamped to 8 instead of returning 0.0.
int effective_fpe(int fpe, int early_exit, int min_fpe) {
    if (fpe < min_fpe) return 0;
    if (early_exit) return fpe <= 4 ? 4 : (fpe <= 6 ? 6 : 8);
    return fpe > 8 ? 8 : fpe;
}

typedef void (*kernel_fn)(const ReduceParams);

template <int F, bool EE, bool DOT>
kernel_fn kernel_ptr() {
    return exblas_reduce_kernel<F, EE, DOT, vectors_in_flight(F, EE, DOT), kMaxT>;
}

template <bool DOT>
kernel_fn select_kernel(int f, bool ee) {
    if (ee) {
        switch (f) {
            case 4: return kernel_ptr<4, true, DOT>();
            case 6: return kernel_ptr<6, true, DOT>();
            default: return kernel_ptr<8, true, DOT>();
        }
    }
    switch (f) {
        case 2: return kernel_ptr<2, false, DOT>();
        case 3: return kernel_ptr<3, false, DOT>();
        case 4: return kernel_ptr<4, false, DOT>();
        case 5: return kernel_ptr<5, false, DOT>();
        case 6: return kernel_ptr<6, false, DOT>();
        case 7: return kernel_ptr<7, false, DOT>();
        default: return kernel_ptr<8, false, DOT>();
    }
}

// One kernel launch over DEVICE data.  finalize = 0 leaves the partial sum in the workspace.
int launch_reduce(exblas_b200_handle_t h, bool dot, int f, bool ee, const double* a, const double* b, int64_t n,
                  int64_t inca, int64_t incb, int finalize, int round_mode) {
    ReduceParams p;
    memset(&p, 0, sizeof(p));
    p.a = a;
    p.b = b;
    p.n = n;
    p.inca = inca;
    p.incb = incb;
    p.ws = h->d_ws;
    p.out = h->d_res;
    p.finalize = finalize;
    p.round_mode = round_mode;
    p.keep = 0;
    p.adaptive = h->opt_adaptive ? 1 : 0;
    p.window = (int)h->opt_window;                   // 0 off, 1 three-digit window, 2 (default) + five-digit window for ExDOT
    p.fresh = h->acc_pending ? 0 : 1;
    p.nranks = 0;
    if (finalize && h->peer_ranks > 1 && h->opt_fused) {          // closing launch: exchange limbs inside the kernel
        p.nranks = h->peer_ranks;
        p.rank = h->peer_rank;
        p.epoch = ++h->epoch;
        for (int r = 0; r < h->peer_ranks; ++r) p.peers[r] = h->peer_box[r];
    }
    h->acc_pending = finalize ? false : true;      // an open (chunked) reduction leaves its partial sum in gacc

    // vector region: unit strides and, for ExDOT, the same 32-byte phase on both streams
    p.head = 0;
    p.nvec = 0;
    if (inca == 1 && (!dot || incb == 1) && ((uintptr_t)a % 8 == 0) && (!dot || (uintptr_t)b % 8 == 0)) {
        const int64_t mis_a = (int64_t)(((uintptr_t)a % 32) / 8);
        const int64_t head = (4 - mis_a) % 4;
        bool ok = true;
        if (dot) ok = ((uintptr_t)b % 32) / 8 == (uintptr_t)mis_a;
        if (ok && n >= head + 4) {
            p.head = head;
            p.nvec = (n - head) / 4;
        }
    }
    // launch shape.  Large vectors: one 512-thread CTA per SM.  Below that the fixed costs of a launch (clearing and
    // merging T columns per CTA, the global merge round trip) dominate, so the CTAs shrink with n, and the
    // smallest vectors take ONE CTA that publishes straight from shared memory (no global round trip at all).
    int T = (int)h->opt_block_threads;
    int64_t max_blocks = h->opt_blocks > 0 ? h->opt_blocks : h->num_sms;
    const int64_t work = p.nvec > 0 ? p.nvec : (n + 7) / 8;      // 256-bit vectors (scalar path: groups of 8 loads)
    p.peer_timeout_ns = (unsigned long long)h->opt_peer_timeout_ms * 1000000ull;
    p.phase = h->d_phase;
    p.l2_prefetch = (int)h->opt_reduce_prefetch;
    p.handoff_tiles = (int)h->opt_dot_handoff;
    if (!h->opt_shape_fixed && n <= h->opt_solo_max && p.fresh && finalize) {
        // latency regime: the small single-CTA kernel (every fpe value: fpe never changes the result)
        const int64_t per_thread4 = p.nvec > 0 ? (p.nvec + 3) / 4 : (n + 3) / 4;
        T = per_thread4 <= 128 ? 128 : (per_thread4 <= 256 ? 256 : 512);
        if (p.nvec > 4 * (int64_t)T) p.nvec = 4 * (int64_t)T;    // (only if solo_max_elems was raised beyond 2^13: rest is scalar)
        p.iters = 0;
        if (h->d_phase) h->phase_blocks = 1;
        kernel_fn sfn = dot ? exblas_small_kernel<true> : exblas_small_kernel<false>;
        CK(allow_big_smem((const void*)sfn, h->device));
        void* sargs[] = {(void*)&p};
        CK(cudaLaunchKernel((const void*)sfn, dim3(1), dim3((unsigned)T), sargs, (size_t)T * kLimbs * sizeof(long long), h->stream));
        h->launches += 1;
        h->last_kernel = std::string(dot ? "exblas_small_kernel<DOT=1>" : "exblas_small_kernel<DOT=0>") + " grid=1 block=" + std::to_string(T);
        return EXBLAS_B200_OK;
    }
    // Mid sizes (ExSUM up to "track_max_elems", default 2^20): what such a launch costs is its fixed part, and the cheapest
    // fixed part is the superaccumulator-only kernel WITHOUT its unrolled body -- every summand then passes through the
    // loops of reduce_finish that record the exponent range, so the merge sums only the limb rows that were touched (the
    // analogue of the reference's imin / imax).  Measured in graph replay for 2^14 ... 2^20 elements, narrow and wide data:
    // 7.1-10.2 us against 8.3-14.1 us for the kernels fpe would select (profiles/track_sweep_r02.jsonl).  Every fpe value
    // takes it: fpe never changes a result.
    const bool mid = !h->opt_shape_fixed && !dot && n <= h->opt_track_max;
    if (mid) {
        f = 0;
        ee = false;
    }
    if (!h->opt_shape_fixed && n <= h->opt_small_max) T = 256;
    int64_t blocks = (work + 4 * (int64_t)T - 1) / (4 * (int64_t)T);   // four vectors per thread keep the loads overlapped
    if (blocks < 1) blocks = 1;
    if (blocks > max_blocks) blocks = max_blocks;
    // unrolled vector body: the same number of tiles in every CTA; the rest is spread over all threads (reduce_finish)
    const int U = f == 0 ? 1 : vectors_in_flight(f, ee, dot);     // exblas_reduce0_kernel tiles by rows of T * 4 elements
    p.iters = p.nvec / ((int64_t)T * U * blocks);
    if (p.iters > 0x7fffffff) p.iters = 0x7fffffff;               // (2^31 tiles per CTA: beyond any memory)
    if (mid) p.iters = 0;
    if (h->d_phase) h->phase_blocks = (int)blocks;

    // superaccumulator-only mode: window loop with 6 (ExSUM) / 2 (ExDOT) rows in flight, direct loop with 8 / 4
    // (measured at n = 2^30 against 4/8, 8/8 and 3/4, 4/4: profiles/f0_window_r01.jsonl)
    kernel_fn fn = f == 0 ? (dot ? exblas_reduce0_kernel<true, 2, 4, kMaxT> : exblas_reduce0_kernel<false, 6, 8, kMaxT>)
                          : (dot ? select_kernel<true>(f, ee) : select_kernel<false>(f, ee));
    const size_t smem = (size_t)T * kLimbs * sizeof(long long);
    CK(allow_big_smem((const void*)fn, h->device));
    void* args[] = {(void*)&p};
    CK(cudaLaunchKernel((const void*)fn, dim3((unsigned)blocks), dim3((unsigned)T), args, smem, h->stream));
    h->launches += 1;
    {
        char nm[160];
        if (f == 0)
            snprintf(nm, sizeof nm, "exblas_reduce0_kernel<DOT=%d,DW=%d,DD=%d,MAXT=%d> grid=%lld block=%d", dot ? 1 : 0,
                     dot ? 2 : 6, dot ? 4 : 8, kMaxT, (long long)blocks, T);
        else
            snprintf(nm, sizeof nm, "exblas_reduce_kernel<F=%d,EE=%d,DOT=%d,U=%d,MAXT=%d> grid=%lld block=%d", f, ee ? 1 : 0,
                     dot ? 1 : 0, U, kMaxT, (long long)blocks, T);
        h->last_kernel = nm;
    }
    return EXBLAS_B200_OK;
}

bool is_device_pointer(const void* p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int ensure_stage(exblas_b200_handle_t h, int64_t elems, bool dot) {
    if (h->stage_elems >= elems && (!dot || h->d_stage[0][1])) return EXBLAS_B200_OK;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            if (h->d_stage[i][j]) {
                cudaFree(h->d_stage[i][j]);
                h->d_stage[i][j] = nullptr;
            }
    h->stage_elems = 0;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < (dot ? 2 : 1); ++j) CK(cudaMalloc(&h->d_stage[i][j], (size_t)elems * sizeof(double)));
    h->stage_elems = elems;
    return EXBLAS_B200_OK;
}

// Host inputs: chunked H2D copies on the copy stream, overlapped with the reduction of the previous
// chunk; every chunk adds into the same device accumulator, the last launch publishes.
int reduce_from_host_impl(exblas_b200_handle_t h, bool dot, int f, bool ee, const double* a, int64_t inca, const double* b,
                          int64_t incb, int64_t n, int round_mode);

int reduce_from_host(exblas_b200_handle_t h, bool dot, int f, bool ee, const double* a, int64_t inca, const double* b,
                     int64_t incb, int64_t n, int round_mode) {
    int rc = reduce_from_host_impl(h, dot, f, ee, a, inca, b, incb, n, round_mode);
    if (rc != EXBLAS_B200_OK && h->acc_pending) {
        // a failed chunked reduction must not leave its partial sum behind
        cudaStreamSynchronize(h->stream);
        cudaMemset(h->d_ws, 0, sizeof(Workspace));
        cudaGetLastError();
        h->acc_pending = false;
    }
    return rc;
}

bool is_pageable_host_pointer(const void* p) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// pinned bounce ring + copy threads for pageable input (allocated on first use)
int ensure_bounce(exblas_b200_handle_t h, int64_t elems, bool dot) {
    if (!h->pool) {
        int nt = (int)h->opt_host_threads;
        if (nt <= 0) {
            const unsigned hw = std::thread::hardware_concurrency();
            const int64_t ranks = h->opt_world_size > 1 ? h->opt_world_size : 1;
            nt = (int)((hw ? hw : 8u) / (ranks > 1 ? ranks : 2));   // a multi-rank job shares the cores; alone, leave half to the caller
            if (nt > 8) nt = 8;
            if (nt < 2) nt = 2;
        }
        h->pool = new CopyPool(nt, local_cpus_of_device(h->device));
    }
    if (h->bounce_elems >= elems && (!dot || h->h_bounce[0][1])) return EXBLAS_B200_OK;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 2; ++j)
            if (h->h_bounce[i][j]) {
                cudaFreeHost(h->h_bounce[i][j]);
                h->h_bounce[i][j] = nullptr;
            }
    h->bounce_elems = 0;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < (dot ? 2 : 1); ++j) CK(cudaHostAlloc(&h->h_bounce[i][j], (size_t)elems * sizeof(double), cudaHostAllocDefault));
        if (!h->bounce_free[i]) CK(cudaEventCreateWithFlags(&h->bounce_free[i], cudaEventDisableTiming));
    }
    h->bounce_elems = elems;
    return EXBLAS_B200_OK;
}

int reduce_from_host_impl(exblas_b200_handle_t h, bool dot, int f, bool ee, const double* a, int64_t inca, const double* b,
                          int64_t incb, int64_t n, int round_mode) {
    const int64_t max_inc = dot ? (inca > incb ? inca : incb) : inca;
    // PAGEABLE input of some size: worker threads copy chunk k + 1 into a pinned bounce buffer while chunk k is on the bus
    const bool bounce = h->opt_host_threads != 1 && n * max_inc >= ((int64_t)1 << 19) &&
                        (is_pageable_host_pointer(a) || (dot && is_pageable_host_pointer(b)));
    int64_t chunk = bounce ? h->opt_pageable_chunk : h->opt_host_chunk;
    if (chunk > n) chunk = n;
    const int64_t span = (chunk - 1) * max_inc + 1;            // doubles copied per chunk and stream
    int rc = ensure_stage(h, span, dot);
    if (rc) return rc;
    if (bounce) {
        rc = ensure_bounce(h, span, dot);
        if (rc) return rc;
    }
    int k = 0;
    for (int64_t i0 = 0; i0 < n; i0 += chunk, ++k) {
        const int64_t cnt = (n - i0 < chunk) ? n - i0 : chunk;
        const int s = k & 1;
        const size_t bytes_a = (size_t)((cnt - 1) * inca + 1) * sizeof(double);
        const size_t bytes_b = dot ? (size_t)((cnt - 1) * incb + 1) * sizeof(double) : 0;
        const double* src_a = a + i0 * inca;
        const double* src_b = dot ? b + i0 * incb : nullptr;
        if (bounce) {
            const int r = k % 3;
            if (k >= 3) CK(cudaEventSynchronize(h->bounce_free[r]));        // its previous H2D has left the buffer
            h->pool->copy(h->h_bounce[r][0], src_a, bytes_a, dot ? h->h_bounce[r][1] : nullptr, src_b, bytes_b);
            src_a = h->h_bounce[r][0];
            src_b = h->h_bounce[r][1];
        }
        if (k >= 2) CK(cudaStreamWaitEvent(h->copy_stream, h->consumed[s], 0));
        CK(cudaMemcpyAsync(h->d_stage[s][0], src_a, bytes_a, cudaMemcpyHostToDevice, h->copy_stream));
        if (dot) CK(cudaMemcpyAsync(h->d_stage[s][1], src_b, bytes_b, cudaMemcpyHostToDevice, h->copy_stream));
        if (bounce) CK(cudaEventRecord(h->bounce_free[k % 3], h->copy_stream));
        CK(cudaEventRecord(h->copied[s], h->copy_stream));
        CK(cudaStreamWaitEvent(h->stream, h->copied[s], 0));
        const int last = (i0 + cnt >= n);
        rc = launch_reduce(h, dot, f, ee, h->d_stage[s][0], h->d_stage[s][1], cnt, inca, incb, last, round_mode);
        if (rc) return rc;
        CK(cudaEventRecord(h->consumed[s], h->stream));
    }
    return EXBLAS_B200_OK;
}

int reduce_any(exblas_b200_handle_t h, bool dot, const double* a, int64_t inca, int64_t offa, const double* b,
               int64_t incb, int64_t offb, int64_t n, int fpe, int early_exit, int round_mode, bool device_only) {
    if (!h) return EXBLAS_B200_EINVAL;
    if (n < 0 || fpe < 0 || inca < 1 || (dot && incb < 1) || offa < 0 || offb < 0) {
        h->err = "invalid argument (n < 0, fpe < 0, inc < 1 or offset < 0)";
        return EXBLAS_B200_EINVAL;
    }
    if (n > 0 && (!a || (dot && !b))) {
        h->err = "null data pointer";
        return EXBLAS_B200_EINVAL;
    }
    CK(cudaSetDevice(h->device));
    const int f = effective_fpe(fpe, early_exit, dot ? 3 : 2);
    const bool ee = early_exit && f > 0;
    if (n == 0) {
        // publish an exact zero without touching the inputs (ExDOT.cpp:70-71 returns 0.0 for Ng <= 0)
        static const double zero = 0.0;
        (void)zero;
        return launch_reduce(h, dot, 0, false, (const double*)h->d_res, (const double*)h->d_res, 0, 1, 1, 1, round_mode);
    }
    const double* pa = a + offa;
    const double* pb = dot ? b + offb : nullptr;
    const bool dev_a = is_device_pointer(pa);
    const bool dev_b = dot ? is_device_pointer(pb) : dev_a;
    if (dev_a && dev_b) return launch_reduce(h, dot, f, ee, pa, pb, n, inca, incb, 1, round_mode);
    if (device_only) {
        h->err = "the *_async entry points need device (or managed) pointers";
        return EXBLAS_B200_EINVAL;
    }
    if (dev_a != dev_b) {
        h->err = "exdot: a and b must both be host or both be device pointers";
        return EXBLAS_B200_EINVAL;
    }
    return reduce_from_host(h, dot, f, ee, pa, inca, pb, incb, n, round_mode);
}

int fetch_result(exblas_b200_handle_t h, double* result, int64_t* limbs, uint32_t* status) {
    CK(cudaMemcpyAsync(h->h_res, h->d_res, sizeof(Result), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->last_status = h->h_res->status;
    if (result) *result = h->h_res->value;
    if (status) *status = h->h_res->status;
    if (limbs)
        for (int j = 0; j < kLimbs; ++j) limbs[j] = h->h_res->limbs[j];
    if (h->h_res->status & kStPeerTimeout) {
        // the ranks no longer agree on what has been summed: this is an error, not a data property
        h->err = "fused multi-GPU exchange: a peer's limbs did not arrive within peer_timeout_ms; the value (NaN) and limbs are "
                 "partial, and the peers must be re-attached (exblas_b200_peer_export / peer_attach) before the next reduction";
        return EXBLAS_B200_EPEER;
    }
    return EXBLAS_B200_OK;
}

typedef void (*gemv_fn)(const GemvParams);

// ExGEMV CTAs are 384 threads (one row each): the kernel needs ~150 registers per thread (pointers,
// a deep window of 8-byte loads, the expansion), which a 512-thread CTA cannot have without spills.
constexpr int kGemvT = 384;
constexpr int gemv_groups_in_flight(int f) { return f <= 4 ? 4 : 3; }

template <int F, bool EE, bool A1>
gemv_fn gemv_ptr() {
    return exgemv_n_kernel<F, EE, A1, gemv_groups_in_flight(F), kGemvT>;
}

template <bool A1>
gemv_fn select_gemv(int f, bool ee) {
    if (ee) {
        switch (f) {
            case 4: return gemv_ptr<4, true, A1>();
            case 6: return gemv_ptr<6, true, A1>();
            default: return gemv_ptr<8, true, A1>();
        }
    }
    switch (f) {
        case 0: return gemv_ptr<0, false, A1>();
        case 2: return gemv_ptr<2, false, A1>();
        case 3: return gemv_ptr<3, false, A1>();
        case 4: return gemv_ptr<4, false, A1>();
        case 5: return gemv_ptr<5, false, A1>();
        case 6: return gemv_ptr<6, false, A1>();
        case 7: return gemv_ptr<7, false, A1>();
        default: return gemv_ptr<8, false, A1>();
    }
}

// Column split: enough CTAs to fill every SM for whole waves, as little limb scratch as possible.
int choose_gemv_parts(int64_t row_blocks, int64_t n, int num_sms, int pmin = 1) {
    int best = pmin;
    double best_cost = 1e300;
    int64_t pmax = n / 64 > 0 ? (n / 64 < 64 ? n / 64 : 64) : 1;
    if (pmax < pmin) pmax = pmin;
    for (int p = pmin; p <= pmax; ++p) {
        const double ctas = (double)row_blocks * p;
        const double waves = ceil(ctas / num_sms);
        const double eff = ctas / (waves * num_sms);
        const double cost = (1.0 / eff) * (1.0 + 78.0 * p / (double)n) + 1e-9 * p;
        if (cost < best_cost) {
            best_cost = cost;
            best = p;
        }
    }
    return best;
}

int gemv_device(exblas_b200_handle_t h, int64_t m, int64_t n, double alpha, const double* a, int64_t rs, int64_t cs,
                const double* x, int64_t incx, double beta, double* y, int64_t incy, int f, bool ee, int round_mode,
                bool transposed) {
    // 'T' with alpha == 1 and columns long enough to amortise the per-output warp merge: one warp per
    // output, register windows, no scratch and no second kernel (exgemv_t_win_kernel).  Every fpe value
    // takes it -- fpe only selects HOW the exact sum is accumulated, never the result.
    if (transposed && alpha == 1.0 && h->opt_window && n >= 256 && n < (1ll << 31) - 65536 && m < (1ll << 31)) {   // (32-bit row / set counters)
        GemvParams p;
        memset(&p, 0, sizeof(p));
        p.a = a;
        p.x = x;
        p.y = y;
        p.m = m;
        p.n = n;
        p.rs = rs;
        p.cs = cs;
        p.incx = incx;
        p.incy = incy;
        p.alpha = alpha;
        p.beta = beta;
        p.ws = h->d_ws;
        p.round_mode = round_mode;
        p.l2_prefetch = (((uintptr_t)a % 16) == 0 && (rs % 2) == 0) ? (int)h->opt_gemv_prefetch : 0;
        p.x_vec_ok = (incx == 1 && ((uintptr_t)x % 16) == 0 && h->opt_gemv_tma) ? 1 : 0;   // x chunks by cp.async.bulk (else copied by a warp)
        // launch shapes (option "gemv_t_shape"): threads x groups of 4 rows in flight per lane x rows of x per buffer
        struct TShape { int T, chunk, nbuf; gemv_fn fn; };
        static const TShape shapes[] = {
            {256, 8192, 2, exgemv_t_win_kernel<8, 256, 8192, 2>},
            {384, 6144, 2, exgemv_t_win_kernel<6, 384, 6144, 2>},
            {512, 4096, 2, exgemv_t_win_kernel<4, 512, 4096, 2>},
            {512, 3072, 2, exgemv_t_win_kernel<3, 512, 3072, 2>},
            {512, 2048, 4, exgemv_t_win_kernel<4, 512, 2048, 4>},
            {512, 1024, 8, exgemv_t_win_kernel<4, 512, 1024, 8>},
        };
        const TShape& sh = shapes[h->opt_gemv_t_shape];
        const int T = sh.T, nwarps = T / 32, chunk = sh.chunk;
        const int64_t nsets = (m + nwarps - 1) / nwarps;
        const unsigned grid = (unsigned)(nsets < h->num_sms ? nsets : h->num_sms);
        const size_t smem = ((size_t)T * kLimbs + (size_t)sh.nbuf * (size_t)chunk + 40 * (size_t)nwarps) * sizeof(long long) + 12 * (size_t)sh.nbuf + 8;   // + the x pipeline's mbarriers / counters
        gemv_fn fn = sh.fn;
        CK(allow_big_smem((const void*)fn, h->device));
        void* args[] = {(void*)&p};
        CK(cudaLaunchKernel((const void*)fn, dim3(grid), dim3((unsigned)T), args, smem, h->stream));
        h->launches += 1;
        return EXBLAS_B200_OK;
    }
    // superaccumulator-only mode, alpha == 1, unit row stride: the register-window kernel (window.cuh);
    // it stages its slice of x in shared memory, so a part holds at most kGemvXsMax columns
    // (a part holds at most kGemvXsMax columns; beyond 1024 parts the per-part limbs could overflow when summed)
    // Like 'T' above, every fpe value takes it: fpe only selects HOW the exact sum is accumulated, never the result, and
    // the window kernel streams at 2.8x the rate of the expansion kernels (round 1: 5.6 against 2.0 TB/s at 32768^2).
    // Option "window" = 0 keeps the expansion kernels (exgemv_n_kernel<F, EE, ...>) reachable.
    const bool windowed = alpha == 1.0 && rs == 1 && h->opt_window && (n + kGemvXsMax - 1) / kGemvXsMax <= 1024;
    // launch shapes of the window kernel (option "gemv_n_shape"): rows per CTA x column groups in flight
    struct NShape { int T; gemv_fn fn; };
    static const NShape nshapes[] = {
        {384, exgemv_n_win_kernel<8, 384>},
        {512, exgemv_n_win_kernel<4, 512>},
        {512, exgemv_n_win_kernel<5, 512>},
    };
    const NShape& nsh = nshapes[h->opt_gemv_n_shape];
    const int Tmax = windowed ? nsh.T : kGemvT;
    int T = Tmax;
    if (m < T) T = (int)((m + 31) / 32 * 32);
    if (T < 32) T = 32;
    const int64_t row_blocks = (m + T - 1) / T;
    const int pmin = windowed ? (int)((n + kGemvXsMax - 1) / kGemvXsMax) : 1;
    int parts = h->opt_gemv_parts > 0 ? (int)h->opt_gemv_parts : choose_gemv_parts(row_blocks, n, h->num_sms, pmin > 0 ? pmin : 1);
    if (parts < pmin) parts = pmin;
    int64_t cpp = ((n + parts - 1) / parts + 3) / 4 * 4;
    if (cpp < 4) cpp = 4;
    parts = (int)((n + cpp - 1) / cpp);
    if (parts < 1) parts = 1;
    const size_t need = (size_t)parts * m * (kLimbs * sizeof(long long) + sizeof(unsigned));
    if (need > h->gemv_scratch_bytes) {
        if (h->d_gemv_scratch) cudaFree(h->d_gemv_scratch);
        h->d_gemv_scratch = nullptr;
        h->gemv_scratch_bytes = 0;
        CK(cudaMalloc(&h->d_gemv_scratch, need));
        h->gemv_scratch_bytes = need;
    }
    GemvParams p;
    memset(&p, 0, sizeof(p));
    p.a = a;
    p.x = x;
    p.y = y;
    p.m = m;
    p.n = n;
    p.rs = rs;
    p.cs = cs;
    p.incx = incx;
    p.incy = incy;
    p.alpha = alpha;
    p.beta = beta;
    p.cols_per_part = cpp;
    p.parts = parts;
    p.scratch = h->d_gemv_scratch;
    p.row_status = (unsigned*)(h->d_gemv_scratch + (size_t)parts * kLimbs * m);
    p.ws = h->d_ws;
    p.round_mode = round_mode;
    p.adaptive = h->opt_adaptive ? 1 : 0;
    p.x_vec_ok = (incx == 1 && ((uintptr_t)x % 32) == 0) ? 1 : 0;       // part starts are multiples of 4 columns
    p.l2_prefetch = (windowed && ((uintptr_t)a % 16) == 0 && (cs % 2) == 0 && (T % 2) == 0) ? (int)h->opt_gemv_prefetch : 0;
    gemv_fn fn = windowed ? nsh.fn : (alpha == 1.0 ? select_gemv<true>(f, ee) : select_gemv<false>(f, ee));
    const size_t smem = windowed ? (size_t)Tmax * kLimbs * sizeof(long long) + (size_t)cpp * sizeof(double)   // fixed limb stride
                                 : (size_t)T * kLimbs * sizeof(long long);
    CK(allow_big_smem((const void*)fn, h->device));
    void* args[] = {(void*)&p};
    CK(cudaLaunchKernel((const void*)fn, dim3((unsigned)row_blocks, (unsigned)parts), dim3((unsigned)T), args, smem,
                        h->stream));
    exgemv_finish_kernel<<<(unsigned)((m + 31) / 32), kFinishT, 0, h->stream>>>(p);
    CK(cudaGetLastError());
    h->launches += 2;
    return EXBLAS_B200_OK;
}

// Batched reductions: one warp per segment (segments_kernel.cuh).  All pointers are device pointers.
typedef void (*seg_fn)(const SegParams);
constexpr int kSegT = 256;

int segments_device(exblas_b200_handle_t h, const double* a, const double* b, const int* gather, const int64_t* seg,
                    int64_t nseg, int64_t total_hint, int round_mode, double* results, uint32_t* statuses) {
    const int nwarps = kSegT / 32;
    // every warp of the grid streams an equal share of the ELEMENTS; enough CTAs for two per SM unless the input is small
    // (total_hint = number of elements when the caller knows it, else -1)
    int64_t grid = 2 * (int64_t)h->num_sms;
    // A segment longer than a range is reduced piecewise, one piece per warp; every piece adds limbs of magnitude
    // below 2^52 + 2^11 to one of two 64-bit scratch accumulators (even / odd warps), so at most 2 x 2047 warps may run.
    if (grid * nwarps > 4080) grid = 4080 / nwarps;
    if (total_hint >= 0) {
        int64_t want = (total_hint + nwarps * 2048 - 1) / (nwarps * 2048);      // ~2048 elements per warp at least
        const int64_t by_seg = (nseg + nwarps * 256 - 1) / (nwarps * 256);       // (many empty segments still need writers)
        if (want < by_seg) want = by_seg;
        if (want < 1) want = 1;
        if (grid > want) grid = want;
    }
    const size_t need = (size_t)grid * nwarps * sizeof(SegScratch);
    if (need > h->seg_scratch_bytes) {
        if (h->d_seg_scratch) cudaFree(h->d_seg_scratch);
        h->d_seg_scratch = nullptr;
        h->seg_scratch_bytes = 0;
        const size_t cap = (size_t)2 * h->num_sms * nwarps * sizeof(SegScratch);
        CK(cudaMalloc(&h->d_seg_scratch, cap));
        CK(cudaMemsetAsync(h->d_seg_scratch, 0, cap, h->stream));       // self-cleaning afterwards (the last piece of a segment clears its slot)
        h->seg_scratch_bytes = cap;
    }
    SegParams p;
    memset(&p, 0, sizeof(p));
    p.a = a;
    p.b = b;
    p.gather = gather;
    p.seg = (const long long*)seg;
    p.nseg = nseg;
    p.results = results;
    p.statuses = statuses;
    p.ws = h->d_ws;
    p.scratch = (SegScratch*)h->d_seg_scratch;
    p.round_mode = round_mode;
    seg_fn fn = b ? (gather ? exblas_segments_kernel<true, true, kSegT> : exblas_segments_kernel<true, false, kSegT>)
                  : exblas_segments_kernel<false, false, kSegT>;
    const size_t smem = (size_t)kSegT * kLimbs * sizeof(long long);
    CK(allow_big_smem((const void*)fn, h->device));
    void* args[] = {(void*)&p};
    CK(cudaLaunchKernel((const void*)fn, dim3((unsigned)grid), dim3((unsigned)kSegT), args, smem, h->stream));
    h->launches += 1;
    return EXBLAS_B200_OK;
}

int segments_any(exblas_b200_handle_t h, const double* a, const double* b, const int32_t* gather, int64_t nb,
                 const int64_t* seg, int64_t nseg, int round_mode, double* results, uint32_t* statuses) {
    if (!h) return EXBLAS_B200_EINVAL;
    if (nseg < 0 || (nseg > 0 && (!seg || !results)) || (gather && !b)) {
        h->err = "segments: invalid argument";
        return EXBLAS_B200_EINVAL;
    }
    if (nseg == 0) return EXBLAS_B200_OK;
    CK(cudaSetDevice(h->device));
    const bool dev = is_device_pointer(results);
    if (dev) {
        int rc = segments_device(h, a, b, (const int*)gather, seg, nseg, -1, round_mode, results, statuses);
        if (rc) return rc;
    } else {
        // host operands (what the reference's callers hold): validate the offsets, stage everything once
        int64_t total = 0;
        for (int64_t s = 0; s < nseg; ++s) {
            if (seg[s] < 0 || seg[s + 1] < seg[s]) {
                h->err = "segments: offsets must be non-negative and non-decreasing";
                return EXBLAS_B200_EINVAL;
            }
        }
        total = seg[nseg];
        if (total > 0 && (!a || ((b != nullptr) && !gather && !b))) return EXBLAS_B200_EINVAL;
        if (gather) {
            if (nb <= 0) {
                h->err = "segments: nb (length of b) is required with a gather index on host operands";
                return EXBLAS_B200_EINVAL;
            }
            for (int64_t i = seg[0]; i < total; ++i)
                if (gather[i] < 0 || gather[i] >= nb) {
                    h->err = "segments: gather index out of range";
                    return EXBLAS_B200_EINVAL;
                }
        }
        double *da = nullptr, *db = nullptr, *dr = nullptr;
        int* dg = nullptr;
        int64_t* ds = nullptr;
        uint32_t* dst = nullptr;
        const size_t nbv = b ? (gather ? (size_t)nb : (size_t)total) : 0;
        int rc = [&]() -> int {
            if (total) CK(cudaMalloc(&da, (size_t)total * sizeof(double)));
            if (nbv) CK(cudaMalloc(&db, nbv * sizeof(double)));
            if (gather && total) CK(cudaMalloc(&dg, (size_t)total * sizeof(int)));
            CK(cudaMalloc(&ds, (size_t)(nseg + 1) * sizeof(int64_t)));
            CK(cudaMalloc(&dr, (size_t)nseg * sizeof(double)));
            CK(cudaMalloc(&dst, (size_t)nseg * sizeof(uint32_t)));
            if (total) CK(cudaMemcpyAsync(da, a, (size_t)total * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            if (nbv) CK(cudaMemcpyAsync(db, b, nbv * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            if (dg) CK(cudaMemcpyAsync(dg, gather, (size_t)total * sizeof(int), cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(ds, seg, (size_t)(nseg + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, h->stream));
            int r2 = segments_device(h, da, b ? db : nullptr, dg, ds, nseg, total - seg[0], round_mode, dr, dst);
            if (r2) return r2;
            CK(cudaMemcpyAsync(results, dr, (size_t)nseg * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            if (statuses) CK(cudaMemcpyAsync(statuses, dst, (size_t)nseg * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            return EXBLAS_B200_OK;
        }();
        if (da) cudaFree(da);
        if (db) cudaFree(db);
        if (dg) cudaFree(dg);
        if (ds) cudaFree(ds);
        if (dr) cudaFree(dr);
        if (dst) cudaFree(dst);
        if (rc) return rc;
    }
    // status word (OR over all segments): read and reset, stream ordered
    CK(cudaMemcpyAsync(&h->h_res->status, &h->d_ws->status, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemsetAsync(&h->d_ws->status, 0, sizeof(unsigned), h->stream));
    if (!dev) {
        CK(cudaStreamSynchronize(h->stream));
        h->last_status = h->h_res->status;
    }
    return EXBLAS_B200_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int exblas_b200_version(void) { return 100; }

const char* exblas_b200_strerror(int code) {
    switch (code) {
        case EXBLAS_B200_OK: return "ok";
        case EXBLAS_B200_EINVAL: return "invalid argument";
        case EXBLAS_B200_ECUDA: return "CUDA error";
        case EXBLAS_B200_ENOGPU: return "no usable CUDA device";
        case EXBLAS_B200_ENCCL: return "NCCL error";
        case EXBLAS_B200_ENOMEM: return "out of memory";
        case EXBLAS_B200_EPEER: return "multi-GPU peer exchange timed out";
        default: return "unknown error";
    }
}

int exblas_b200_create(exblas_b200_handle_t* out, int device) {
    if (!out) return EXBLAS_B200_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return EXBLAS_B200_ENOGPU;
    }
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return EXBLAS_B200_ENOGPU;
    if (device >= count) return EXBLAS_B200_EINVAL;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return EXBLAS_B200_ECUDA;
    if (prop.major != 10) {
        fprintf(stderr, "exblas_b200: device %d is sm_%d%d; this library is built for sm_100a only\n", device,
                prop.major, prop.minor);
        return EXBLAS_B200_ENOGPU;
    }
    exblas_b200_handle_t h = new exblas_b200_handle_s();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    int rc = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        h->stream = nullptr;                    // legacy default stream, ordered with other default-stream work
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&h->copied[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->consumed[i], cudaEventDisableTiming));
        }
        CK(cudaMalloc(&h->d_ws, sizeof(Workspace)));
        CK(cudaMemset(h->d_ws, 0, sizeof(Workspace)));
        CK(cudaMalloc(&h->d_res, sizeof(Result)));
        CK(cudaMemset(h->d_res, 0, sizeof(Result)));
        CK(cudaMallocHost(&h->h_res, sizeof(Result)));
        return EXBLAS_B200_OK;
    }();
    if (rc) {
        fprintf(stderr, "exblas_b200_create: %s\n", h->err.c_str());
        exblas_b200_destroy(h);
        return rc;
    }
    *out = h;
    return EXBLAS_B200_OK;
}

int exblas_b200_destroy(exblas_b200_handle_t h) {
    if (!h) return EXBLAS_B200_OK;
    cudaSetDevice(h->device);
    if (h->comm && nccl().ok) nccl().CommDestroy(h->comm);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            if (h->d_stage[i][j]) cudaFree(h->d_stage[i][j]);
    for (int r = 0; r < h->peer_ranks; ++r)
        if (h->peer_box[r] && h->peer_box[r] != h->d_mailbox) cudaIpcCloseMemHandle(h->peer_box[r]);
    delete h->pool;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 2; ++j)
            if (h->h_bounce[i][j]) cudaFreeHost(h->h_bounce[i][j]);
        if (h->bounce_free[i]) cudaEventDestroy(h->bounce_free[i]);
    }
    if (h->d_mailbox) cudaFree(h->d_mailbox);
    if (h->d_gemv_scratch) cudaFree(h->d_gemv_scratch);
    if (h->d_seg_scratch) cudaFree(h->d_seg_scratch);
    if (h->d_phase) cudaFree(h->d_phase);
    if (h->d_ws) cudaFree(h->d_ws);
    if (h->d_res) cudaFree(h->d_res);
    if (h->h_res) cudaFreeHost(h->h_res);
    for (int i = 0; i < 2; ++i) {
        if (h->copied[i]) cudaEventDestroy(h->copied[i]);
        if (h->consumed[i]) cudaEventDestroy(h->consumed[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    return EXBLAS_B200_OK;
}

int exblas_b200_set_stream(exblas_b200_handle_t h, void* stream) {
    if (!h) return EXBLAS_B200_EINVAL;
    h->stream = (cudaStream_t)stream;
    return EXBLAS_B200_OK;
}

int exblas_b200_set_option(exblas_b200_handle_t h, const char* name, int64_t value) {
    if (!h || !name) return EXBLAS_B200_EINVAL;
    if (!strcmp(name, "block_threads")) {
        if (value < 32 || value > kMaxT || value % 32) return EXBLAS_B200_EINVAL;
        h->opt_block_threads = value;
        h->opt_shape_fixed = true;
    } else if (!strcmp(name, "blocks")) {
        if (value < 0 || value > 2040) return EXBLAS_B200_EINVAL;    // 2040 partials of magnitude < 2^52 + 2^10 (+ a pending normalised sum) fit a 64-bit limb
        h->opt_blocks = value;
        h->opt_shape_fixed = true;
    } else if (!strcmp(name, "auto_shape")) {                        // back to the size-dependent launch shape
        h->opt_shape_fixed = value == 0;
        if (value) {
            h->opt_blocks = 0;
            h->opt_block_threads = kMaxT;
        }
    } else if (!strcmp(name, "solo_max_elems")) {
        if (value < 0) return EXBLAS_B200_EINVAL;
        h->opt_solo_max = value;
    } else if (!strcmp(name, "small_max_elems")) {
        if (value < 0) return EXBLAS_B200_EINVAL;
        h->opt_small_max = value;
    } else if (!strcmp(name, "world_size")) {
        if (value < 0) return EXBLAS_B200_EINVAL;
        h->opt_world_size = value;
    } else if (!strcmp(name, "peer_timeout_ms")) {
        if (value < 0) return EXBLAS_B200_EINVAL;
        h->opt_peer_timeout_ms = value;
    } else if (!strcmp(name, "phase_timing")) {                      // diagnostics: globaltimer stamps per CTA and phase
        if (value && !h->d_phase) {
            if (cudaSetDevice(h->device) != cudaSuccess) return EXBLAS_B200_ECUDA;
            if (cudaMalloc(&h->d_phase, (size_t)2048 * kPhaseSlots * sizeof(unsigned long long)) != cudaSuccess) return EXBLAS_B200_ECUDA;
            cudaMemset(h->d_phase, 0, (size_t)2048 * kPhaseSlots * sizeof(unsigned long long));
        } else if (!value && h->d_phase) {
            cudaFree(h->d_phase);
            h->d_phase = nullptr;
        }
    } else if (!strcmp(name, "fused_allreduce")) {
        h->opt_fused = value != 0;
    } else if (!strcmp(name, "gemv_parts")) {
        if (value < 0 || value > 1024) return EXBLAS_B200_EINVAL;   // per-part limbs are bounded by 2^52 + 2^11, not normalised: 1024 of them fit
        h->opt_gemv_parts = value;
    } else if (!strcmp(name, "gemv_t_shape")) {
        if (value < 0 || value > 5) return EXBLAS_B200_EINVAL;
        h->opt_gemv_t_shape = value;
    } else if (!strcmp(name, "dot_handoff_tiles")) {
        if (value < 0 || value > 0x7fffffff) return EXBLAS_B200_EINVAL;
        h->opt_dot_handoff = value;
    } else if (!strcmp(name, "track_max_elems")) {
        h->opt_track_max = value;
    } else if (!strcmp(name, "gemv_tma")) {
        h->opt_gemv_tma = value ? 1 : 0;
    } else if (!strcmp(name, "reduce_prefetch")) {
        if (value < 0 || value > 16) return EXBLAS_B200_EINVAL;
        h->opt_reduce_prefetch = value;
    } else if (!strcmp(name, "gemv_prefetch")) {
        if (value < 0 || value > 16) return EXBLAS_B200_EINVAL;
        h->opt_gemv_prefetch = value;
    } else if (!strcmp(name, "gemv_n_shape")) {
        if (value < 0 || value > 2) return EXBLAS_B200_EINVAL;
        h->opt_gemv_n_shape = value;
    } else if (!strcmp(name, "window")) {
        if (value < 0 || value > 3) return EXBLAS_B200_EINVAL;
        h->opt_window = value;
    } else if (!strcmp(name, "adaptive")) {
        h->opt_adaptive = value != 0;
    } else if (!strcmp(name, "host_chunk_elems")) {
        if (value < 1024) return EXBLAS_B200_EINVAL;
        h->opt_host_chunk = value;
    } else if (!strcmp(name, "pageable_chunk_elems")) {
        if (value < 1024) return EXBLAS_B200_EINVAL;
        h->opt_pageable_chunk = value;
    } else if (!strcmp(name, "host_threads")) {
        if (value < 0 || value > 64) return EXBLAS_B200_EINVAL;
        if (h->pool && value != h->opt_host_threads) {
            delete h->pool;
            h->pool = nullptr;
        }
        h->opt_host_threads = value;
    } else {
        return EXBLAS_B200_EINVAL;
    }
    return EXBLAS_B200_OK;
}

int exblas_b200_exsum_async(exblas_b200_handle_t h, const double* d_a, int64_t n, int64_t inca, int64_t offset, int fpe,
                            int early_exit, int round_mode) {
    return reduce_any(h, false, d_a, inca, offset, nullptr, 1, 0, n, fpe, early_exit, round_mode, true);
}

int exblas_b200_exdot_async(exblas_b200_handle_t h, const double* d_a, int64_t inca, int64_t offseta, const double* d_b,
                            int64_t incb, int64_t offsetb, int64_t n, int fpe, int early_exit, int round_mode) {
    return reduce_any(h, true, d_a, inca, offseta, d_b, incb, offsetb, n, fpe, early_exit, round_mode, true);
}

int exblas_b200_fetch(exblas_b200_handle_t h, double* result, int64_t* limbs, uint32_t* status) {
    if (!h) return EXBLAS_B200_EINVAL;
    CK(cudaSetDevice(h->device));
    return fetch_result(h, result, limbs, status);
}

int exblas_b200_result_ptr(exblas_b200_handle_t h, void** d_result) {
    if (!h || !d_result) return EXBLAS_B200_EINVAL;
    *d_result = h->d_res;
    return EXBLAS_B200_OK;
}

int exblas_b200_exsum_limbs(exblas_b200_handle_t h, const double* a, int64_t n, int64_t inca, int64_t offset, int fpe,
                            int early_exit, int round_mode, int64_t* limbs, double* result) {
    int rc = reduce_any(h, false, a, inca, offset, nullptr, 1, 0, n, fpe, early_exit, round_mode, false);
    if (rc) return rc;
    return fetch_result(h, result, limbs, nullptr);
}

int exblas_b200_exdot_limbs(exblas_b200_handle_t h, const double* a, int64_t inca, int64_t offseta, const double* b,
                            int64_t incb, int64_t offsetb, int64_t n, int fpe, int early_exit, int round_mode,
                            int64_t* limbs, double* result) {
    int rc = reduce_any(h, true, a, inca, offseta, b, incb, offsetb, n, fpe, early_exit, round_mode, false);
    if (rc) return rc;
    return fetch_result(h, result, limbs, nullptr);
}

int exblas_b200_exsum(exblas_b200_handle_t h, const double* a, int64_t n, int64_t inca, int64_t offset, int fpe,
                      int early_exit, int round_mode, double* result) {
    if (!result) return EXBLAS_B200_EINVAL;
    return exblas_b200_exsum_limbs(h, a, n, inca, offset, fpe, early_exit, round_mode, nullptr, result);
}

int exblas_b200_exdot(exblas_b200_handle_t h, const double* a, int64_t inca, int64_t offseta, const double* b,
                      int64_t incb, int64_t offsetb, int64_t n, int fpe, int early_exit, int round_mode,
                      double* result) {
    if (!result) return EXBLAS_B200_EINVAL;
    return exblas_b200_exdot_limbs(h, a, inca, offseta, b, incb, offsetb, n, fpe, early_exit, round_mode, nullptr,
                                   result);
}

int exblas_b200_exsum_segments(exblas_b200_handle_t h, const double* a, const int64_t* seg, int64_t nseg, int fpe,
                               int early_exit, int round_mode, double* results, uint32_t* statuses) {
    if (fpe < 0) return EXBLAS_B200_EINVAL;
    (void)early_exit;                      // every segment runs the superaccumulator + register window; fpe never changes a result
    return segments_any(h, a, nullptr, nullptr, 0, seg, nseg, round_mode, results, statuses);
}

int exblas_b200_exdot_segments(exblas_b200_handle_t h, const double* a, const double* b, const int32_t* gather, int64_t nb,
                               const int64_t* seg, int64_t nseg, int fpe, int early_exit, int round_mode, double* results,
                               uint32_t* statuses) {
    if (fpe < 0 || !b) return EXBLAS_B200_EINVAL;
    (void)early_exit;
    return segments_any(h, a, b, gather, nb, seg, nseg, round_mode, results, statuses);
}

int exblas_b200_exgemv(exblas_b200_handle_t h, char trans, int64_t m, int64_t n, double alpha, const double* a,
                       int64_t lda, int64_t offseta, const double* x, int64_t incx, int64_t offsetx, double beta,
                       double* y, int64_t incy, int64_t offsety, int fpe, int early_exit, int round_mode) {
    if (!h) return EXBLAS_B200_EINVAL;
    const bool tr = (trans == 'T' || trans == 't');
    if (!tr && trans != 'N' && trans != 'n') {
        h->err = "exgemv: transa must be 'N' or 'T'";
        return EXBLAS_B200_EINVAL;
    }
    if (m < 0 || n < 0 || lda < (m > 1 ? m : 1) || incx < 1 || incy < 1 || fpe < 0 || offseta < 0 || offsetx < 0 ||
        offsety < 0 || (m > 0 && (!y || (n > 0 && (!a || !x))))) {
        h->err = "exgemv: invalid argument";
        return EXBLAS_B200_EINVAL;
    }
    const int64_t mo = tr ? n : m, ni = tr ? m : n;          // outputs, summands per output
    const int64_t rs = tr ? lda : 1, cs = tr ? 1 : lda;
    if (mo == 0) return EXBLAS_B200_OK;
    CK(cudaSetDevice(h->device));
    // fpe: 0 superaccumulators only; 1 is the reference's plain (non-reproducible) DGEMV comparator
    // (ExGEMV.cpp:92-94) -- here it also runs the exact superaccumulator kernel; early exit buckets 4/6/8.
    int f = fpe <= 1 ? 0 : (early_exit ? (fpe <= 4 ? 4 : (fpe <= 6 ? 6 : 8)) : (fpe > 8 ? 8 : fpe));
    const bool ee = early_exit && f > 0;
    const double* pa = a ? a + offseta : a;
    const double* px = x ? x + offsetx : x;
    double* py = y + offsety;
    const bool dev = is_device_pointer(py) && (ni == 0 || m == 0 || n == 0 || (is_device_pointer(pa) && is_device_pointer(px)));
    int rc;
    if (dev) {
        rc = gemv_device(h, mo, ni, alpha, pa, rs, cs, px, incx, beta, py, incy, f, ee, round_mode, tr);
        if (rc) return rc;
    } else {
        // host operands (what the reference's exgemv takes, ExGEMV.cpp:109-234): stage on the device
        double *da = nullptr, *dx = nullptr, *dy = nullptr;
        const size_t na = (n > 0 && m > 0) ? (size_t)lda * (n - 1) + m : 0, nx = ni > 0 ? (size_t)(ni - 1) * incx + 1 : 0,
                     ny = (size_t)(mo - 1) * incy + 1;
        rc = [&]() -> int {
            if (na) CK(cudaMalloc(&da, na * sizeof(double)));
            if (nx) CK(cudaMalloc(&dx, nx * sizeof(double)));
            CK(cudaMalloc(&dy, ny * sizeof(double)));
            if (na) CK(cudaMemcpyAsync(da, pa, na * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            if (nx) CK(cudaMemcpyAsync(dx, px, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            CK(cudaMemcpyAsync(dy, py, ny * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            int r2 = gemv_device(h, mo, ni, alpha, da, rs, cs, dx, incx, beta, dy, incy, f, ee, round_mode, tr);
            if (r2) return r2;
            CK(cudaMemcpyAsync(py, dy, ny * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            return EXBLAS_B200_OK;
        }();
        if (da) cudaFree(da);
        if (dx) cudaFree(dx);
        if (dy) cudaFree(dy);
        if (rc) return rc;
    }
    // status word: read and reset (stream ordered)
    CK(cudaMemcpyAsync(&h->h_res->status, &h->d_ws->status, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemsetAsync(&h->d_ws->status, 0, sizeof(unsigned), h->stream));
    if (!dev) {
        CK(cudaStreamSynchronize(h->stream));
        h->last_status = h->h_res->status;
    }
    return EXBLAS_B200_OK;
}

int exblas_b200_sync(exblas_b200_handle_t h) {
    if (!h) return EXBLAS_B200_EINVAL;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->last_status = h->h_res->status;
    return EXBLAS_B200_OK;
}

int exblas_b200_round(const int64_t* limbs, int round_mode, double* result) {
    if (!limbs || !result) return EXBLAS_B200_EINVAL;
    long long acc[kLimbs];
    for (int j = 0; j < kLimbs; ++j) acc[j] = limbs[j];
    *result = finalize_value(acc, 0, round_mode);
    return EXBLAS_B200_OK;
}

int exblas_b200_normalize(int64_t* limbs, int* negative) {
    if (!limbs) return EXBLAS_B200_EINVAL;
    long long acc[kLimbs];
    for (int j = 0; j < kLimbs; ++j) acc[j] = limbs[j];
    bool neg = normalize(acc);
    for (int j = 0; j < kLimbs; ++j) limbs[j] = acc[j];
    if (negative) *negative = neg ? 1 : 0;
    return EXBLAS_B200_OK;
}

int exblas_b200_merge_limbs(int64_t* dst, const int64_t* src) {
    if (!dst || !src) return EXBLAS_B200_EINVAL;
    long long a[kLimbs], b[kLimbs];
    for (int j = 0; j < kLimbs; ++j) {
        a[j] = dst[j];
        b[j] = src[j];
    }
    normalize(a);
    normalize(b);
    for (int j = 0; j < kLimbs; ++j) a[j] += b[j];
    normalize(a);
    for (int j = 0; j < kLimbs; ++j) dst[j] = a[j];
    return EXBLAS_B200_OK;
}

int exblas_b200_nccl_unique_id(void* id128) {
    if (!id128) return EXBLAS_B200_EINVAL;
    if (!nccl().ok) return EXBLAS_B200_ENCCL;
    return nccl().GetUniqueId(id128) == 0 ? EXBLAS_B200_OK : EXBLAS_B200_ENCCL;
}

int exblas_b200_comm_init(exblas_b200_handle_t h, int nranks, int rank, const void* id128) {
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return EXBLAS_B200_EINVAL;
    if (!nccl().ok) {
        h->err = "libnccl.so.2 could not be loaded";
        return EXBLAS_B200_ENCCL;
    }
    CK(cudaSetDevice(h->device));
    Id128 id;
    memcpy(id.bytes, id128, sizeof(id.bytes));
    int rc = nccl().CommInitRank(&h->comm, nranks, id, rank);
    if (rc != 0) {
        h->err = std::string("ncclCommInitRank: ") + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
        h->comm = nullptr;
        return EXBLAS_B200_ENCCL;
    }
    h->nranks = nranks;
    return EXBLAS_B200_OK;
}

int exblas_b200_peer_export(exblas_b200_handle_t h, void* handle64) {
    if (!h || !handle64) return EXBLAS_B200_EINVAL;
    CK(cudaSetDevice(h->device));
    if (!h->d_mailbox) CK(cudaMalloc(&h->d_mailbox, sizeof(Mailbox)));
    // (re-)initialise: epochs restart at 1 after every attach, so no stale sequence number may survive
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemset(h->d_mailbox, 0, sizeof(Mailbox)));
    CK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t ipc;
    CK(cudaIpcGetMemHandle(&ipc, h->d_mailbox));
    static_assert(sizeof(ipc) == 64, "CUDA IPC handles are 64 bytes");
    memcpy(handle64, &ipc, 64);
    return EXBLAS_B200_OK;
}

int exblas_b200_peer_attach(exblas_b200_handle_t h, int nranks, int rank, const void* handles) {
    if (!h || !handles || nranks < 1 || nranks > kMaxPeers || rank < 0 || rank >= nranks) return EXBLAS_B200_EINVAL;
    if (!h->d_mailbox) {
        h->err = "exblas_b200_peer_export must be called first";
        return EXBLAS_B200_EINVAL;
    }
    CK(cudaSetDevice(h->device));
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            h->peer_box[r] = h->d_mailbox;
            continue;
        }
        cudaIpcMemHandle_t ipc;
        memcpy(&ipc, (const char*)handles + 64 * (size_t)r, 64);
        void* ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess));
        h->peer_box[r] = (Mailbox*)ptr;
    }
    h->peer_ranks = nranks;
    h->peer_rank = rank;
    h->epoch = 0;
    return EXBLAS_B200_OK;
}

int exblas_b200_allreduce_async(exblas_b200_handle_t h, int round_mode) {
    if (!h) return EXBLAS_B200_EINVAL;
    CK(cudaSetDevice(h->device));
    // with the fused exchange active the closing kernel has already merged and rounded
    if (h->peer_ranks > 1 && h->opt_fused) return EXBLAS_B200_OK;
    if (h->opt_world_size > 1 && h->nranks != h->opt_world_size) {
        // the caller declared a multi-rank job (option "world_size") but no transport joins that many ranks:
        // finishing locally would silently return this rank's shard only
        h->err = "allreduce: world_size > 1 but neither exblas_b200_comm_init nor an attached fused peer exchange covers it";
        return EXBLAS_B200_ENCCL;
    }
    if (h->nranks > 1) {
        if (!h->comm) {
            h->err = "exblas_b200_comm_init was not called";
            return EXBLAS_B200_ENCCL;
        }
        // limbs[39] + flag counters[5] are contiguous int64 in the result slot
        int rc = nccl().AllReduce(h->d_res->limbs, h->d_res->limbs, kLimbs + kFlagSlots, kNcclInt64, kNcclSum, h->comm,
                                  h->stream);
        if (rc != 0) {
            h->err = std::string("ncclAllReduce: ") + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "?");
            return EXBLAS_B200_ENCCL;
        }
        h->launches += 1;
    }
    exblas_finalize_kernel<<<1, 32, 0, h->stream>>>(h->d_res, round_mode);
    CK(cudaGetLastError());
    h->launches += 1;
    return EXBLAS_B200_OK;
}

int exblas_b200_last_status(exblas_b200_handle_t h, uint32_t* status_flags) {
    if (!h || !status_flags) return EXBLAS_B200_EINVAL;
    *status_flags = h->last_status;
    return EXBLAS_B200_OK;
}

const char* exblas_b200_last_error(exblas_b200_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int64_t exblas_b200_phase_times(exblas_b200_handle_t h, uint64_t* out, int64_t capacity) {
    if (!h || !h->d_phase || !out) return 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return 0;
    cudaStreamSynchronize(h->stream);
    int64_t words = (int64_t)h->phase_blocks * kPhaseSlots;
    if (words > capacity) words = capacity / kPhaseSlots * kPhaseSlots;
    if (words <= 0) return 0;
    if (cudaMemcpy(out, h->d_phase, (size_t)words * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    cudaMemset(h->d_phase, 0, (size_t)2048 * kPhaseSlots * sizeof(unsigned long long));
    return words / kPhaseSlots;
}

int64_t exblas_b200_launch_count(exblas_b200_handle_t h) { return h ? h->launches : 0; }

const char* exblas_b200_last_kernel(exblas_b200_handle_t h) { return h ? h->last_kernel.c_str() : ""; }

int exblas_b200_microbench(exblas_b200_handle_t h, int what, const double* d_buf, int64_t n, double* result) {
    if (!h || !result) return EXBLAS_B200_EINVAL;
    CK(cudaSetDevice(h->device));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms = 0.f;
    int rc = [&]() -> int {
        if (what == 0) {                    // FP64 pipe: DADD lane-instructions per second
            const int iters = 20000;
            mb_dadd_kernel<<<h->num_sms, 1024, 0, h->stream>>>((double*)h->d_res, 100, 1.0);
            CK(cudaEventRecord(e0, h->stream));
            mb_dadd_kernel<<<h->num_sms, 1024, 0, h->stream>>>((double*)h->d_res, iters, 1.0);
            CK(cudaEventRecord(e1, h->stream));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            *result = (double)h->num_sms * 1024.0 * 8.0 * iters / (ms * 1e-3);
            h->launches += 2;
            return EXBLAS_B200_OK;
        }
        if (what == 1) {                    // read-only stream over d_buf[0, n): GB/s
            if (!d_buf || n < 4 || ((uintptr_t)d_buf % 32) != 0 || !is_device_pointer(d_buf)) return EXBLAS_B200_EINVAL;
            const long long nvec = n / 4;
            const int blocks = h->num_sms * 4, reps = 3;
            mb_read_kernel<<<blocks, 512, 0, h->stream>>>(d_buf, nvec, (double*)h->d_res);
            CK(cudaEventRecord(e0, h->stream));
            for (int r = 0; r < reps; ++r) mb_read_kernel<<<blocks, 512, 0, h->stream>>>(d_buf, nvec, (double*)h->d_res);
            CK(cudaEventRecord(e1, h->stream));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            *result = (double)reps * (double)nvec * 32.0 / (ms * 1e-3) / 1e9;
            h->launches += 1 + reps;
            return EXBLAS_B200_OK;
        }
        return EXBLAS_B200_EINVAL;
    }();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// blas1.hpp drop-in wrappers (same C++ signatures as the reference's include/blas1.hpp:48,74)
// ------------------------------------------------------------------------------------------------
namespace {
exblas_b200_handle_t default_handle() {
    static exblas_b200_handle_t h = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        int rc = exblas_b200_create(&h, -1);
        if (rc != EXBLAS_B200_OK) {
            // the reference's GPU path prints and exits when no platform / device is found
            // (src/gpu/blas/blas1/ExSUM.cpp:98-108)
            fprintf(stderr, "exblas_b200: cannot create a GPU handle: %s\n", exblas_b200_strerror(rc));
            exit(EXIT_FAILURE);
        }
    });
    return h;
}
// The reference's examples call exsum() from several pthreads (RNGExample.cpp:549-556, with
// parallel = false); the default handle shares one workspace, so the wrappers serialise on it.
std::mutex& default_mutex() {
    static std::mutex m;
    return m;
}
int default_round_mode() {
    const char* e = getenv("EXBLAS_B200_ROUND");
    return (e && (!strcmp(e, "exact") || !strcmp(e, "1"))) ? EXBLAS_B200_ROUND_EXACT : EXBLAS_B200_ROUND_REFERENCE;
}
}  // namespace

double exsum(const int Ng, double* ag, const int inca, const int offset, const int fpe, const bool early_exit,
             const bool /*parallel: ignored like the reference GPU path, ExSUM.cpp:61*/) {
    if (fpe < 0) {   // cpu ExSUM.cpp:25-28
        fprintf(stderr, "Size of floating-point expansion should be a positive number. Preferably, it should be in the interval [2, 8]\n");
        exit(1);
    }
    exblas_b200_handle_t h = default_handle();
    std::lock_guard<std::mutex> lock(default_mutex());
    double r = 0.0;
    int rc = exblas_b200_exsum(h, ag, Ng < 0 ? 0 : Ng, inca, offset, fpe, early_exit ? 1 : 0, default_round_mode(), &r);
    if (rc != EXBLAS_B200_OK) {
        fprintf(stderr, "exsum: %s (%s)\n", exblas_b200_strerror(rc), exblas_b200_last_error(h));
        exit(EXIT_FAILURE);
    }
    return r;
}

double exdot(const int Ng, double* ag, const int inca, const int offseta, double* bg, const int incb, const int offsetb,
             const int fpe, const bool early_exit) {
    if (Ng <= 0) return 0.0;   // ExDOT.cpp:70-71
    if (fpe < 0) {
        fprintf(stderr, "Size of floating-point expansion should be a positive number. Preferably, it should be in the interval [3, 8]\n");
        exit(1);
    }
    exblas_b200_handle_t h = default_handle();
    std::lock_guard<std::mutex> lock(default_mutex());
    double r = 0.0;
    int rc = exblas_b200_exdot(h, ag, inca, offseta, bg, incb, offsetb, Ng, fpe, early_exit ? 1 : 0,
                               default_round_mode(), &r);
    if (rc != EXBLAS_B200_OK) {
        fprintf(stderr, "exdot: %s (%s)\n", exblas_b200_strerror(rc), exblas_b200_last_error(h));
        exit(EXIT_FAILURE);
    }
    return r;
}

int exgemv(const char transa, const int m, const int n, const double alpha, double* a, const int lda, const int offseta,
           double* x, const int incx, const int offsetx, const double beta, double* y, const int incy, const int offsety,
           const int fpe, const bool early_exit) {
    exblas_b200_handle_t h = default_handle();
    std::lock_guard<std::mutex> lock(default_mutex());
    int rc = exblas_b200_exgemv(h, transa, m, n, alpha, a, lda, offseta, x, incx, offsetx, beta, y, incy, offsety, fpe,
                                early_exit ? 1 : 0, default_round_mode());
    if (rc == EXBLAS_B200_OK) rc = exblas_b200_sync(h);
    if (rc != EXBLAS_B200_OK) {
        fprintf(stderr, "exgemv: %s (%s)\n", exblas_b200_strerror(rc), exblas_b200_last_error(h));
        exit(EXIT_FAILURE);      // the reference prints and exits on OpenCL errors (ExGEMV.cpp:120-160)
    }
    return 0;
}
