// pvs_common.cuh -- shared plumbing for the C-ABI library (status, errors, model block).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/pvs_b200.h"

namespace pvs {

// ---- thread-local error string ------------------------------------------------------
std::string& last_error();
int fail(int status, const char* fmt, ...);
extern std::atomic<long long> g_launches;
extern std::atomic<int> g_path;

#define PVS_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return ::pvs::fail(PVS_ERR_CUDA, "%s failed: %s (%s:%d)", #call,            \
                               cudaGetErrorString(e__), __FILE__, __LINE__);            \
    } while (0)

#define PVS_CHECK(cond, status, ...)                                                    \
    do {                                                                                \
        if (!(cond)) return ::pvs::fail(status, __VA_ARGS__);                           \
    } while (0)

// every kernel launch goes through this so gpu_launches is a count, not a guess
#define PVS_LAUNCH(kernel, grid, block, smem, stream, ...)                              \
    do {                                                                                \
        kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                             \
        ::pvs::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess)                                                         \
            return ::pvs::fail(PVS_ERR_CUDA, "launch %s failed: %s (%s:%d)", #kernel,   \
                               cudaGetErrorString(e__), __FILE__, __LINE__);            \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Function attributes (dynamic shared-memory limit) and device capabilities are per DEVICE: one bit per
// device ordinal, so a process that drives several GPUs configures every kernel on each of them.
// need() is true the first time it is asked about the current device (a duplicate set under a race is harmless).
struct PerDeviceOnce {
    std::atomic<unsigned long long> done{0};
    bool need(int* dev_out = nullptr)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev_out) *dev_out = dev;
        const unsigned long long bit = 1ull << (dev & 63);
        return (done.load(std::memory_order_acquire) & bit) == 0;
    }
    void mark()
    {
        int dev = 0;
        cudaGetDevice(&dev);
        done.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
};
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

}  // namespace pvs

// ---- model block --------------------------------------------------------------------
// One device allocation per model; pointers below point into it.
struct pvs_model {
    int kind = 0;
    int k = 0;      // clusters / components (0 for PCA)
    int d = 0;      // feature dim the model consumes (PCA: output dim)
    int d_in = 0;   // PCA: input dim
    void* block = nullptr;  // device allocation
    size_t block_bytes = 0;

    // K-Means: centers [k,d], c2 [k] = ||c||^2 (fp32, as sklearn's row_norms)
    const float* centers = nullptr;
    const float* c2 = nullptr;

    // GMM (diag): logit[t,j] = cst[j] + sum_d x^2 * wq[j,d] + x * wl[j,d]
    //   wcat [k, 2d] = [ -0.5 P | mu P ],  P = precisions_cholesky^2
    //   cst  [k]     = -0.5 (d ln 2pi + sum mu^2 P) + sum ln pc + ln pi
    const float* wcat = nullptr;
    const float* cst = nullptr;
    const float* mu = nullptr;      // [k,d]
    const float* var = nullptr;     // [k,d] covariances_
    const float* pi = nullptr;      // [k]
    const float* g_pi = nullptr;    // [k]   1/sqrt(pi)
    const float* g_mu = nullptr;    // [k,d] 1/(sqrt(pi) sqrt(var))
    const float* g_sig = nullptr;   // [k,d] 1/(sqrt2 sqrt(pi) var)

    // PCA: comp [d, d_in], bias [d] = -(mean @ comp^T)
    const float* comp = nullptr;
    const float* bias = nullptr;

    // tensor-core operand copies (tf32 hi/lo splits), filled by the tcgen05 path
    const float* tc0 = nullptr;
    const float* tc1 = nullptr;
    int tc_ld = 0;
    // generic tensor path (pvs_tc_gemmnt.cu): zero-padded hi / lo copies of comp (PCA) or wcat (GMM)
    const float* tcg0 = nullptr;
    const float* tcg1 = nullptr;
    int tcg_ld = 0;
    // fp16x2 tensor path (GMM, K = 256 / D = 64): y is scaled by 2^-h_exp so that |mu| + 6 sigma <= 128,
    // th0 / th1 = fp16 hi / lo parts of the interleaved weights (-P/2 * 4^h_exp, mu P * 2^h_exp) [k, 2d]
    bool h_ok = false;
    int h_exp = 0;
    const void* th0 = nullptr;
    const void* th1 = nullptr;
    // K-Means: th0 / th1 = fp16 hi / lo parts of the centres * 2^-h_exp, zero-padded to th_ld (multiple of 64)
    // columns; h_flags = ring of device-side range flags (one per call in flight)
    std::vector<float> cst_host;    // GMM: host copy of cst (passed to the fused kernel in its parameter block)
    int th_ld = 0;
    int* h_flags = nullptr;
    mutable std::atomic<unsigned> h_flag_next{0};
};
