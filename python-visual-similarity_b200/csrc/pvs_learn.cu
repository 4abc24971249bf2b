// pvs_learn.cu -- the accumulators of learn() (pyvisim/encoders/_base_encoder.py:311-342) on the device.
//
// The reference fits the visual vocabulary with scikit-learn: KMeans (Lloyd) for VLAD, a diagonal
// GaussianMixture (EM) for Fisher vectors.  Both alternate the two steps the encode path already has
// as kernels:
//   K-Means  E-step = hard assignment (pvs_kmeans_assign: tcgen05 scores + fused arg-min),
//            M-step = per-cluster sums of the members
//   GMM      E-step = posteriors (pvs_gmm_posterior: tcgen05 logits + softmax) + the log-likelihood,
//            M-step = zeroth / first / second order statistics  sum_t q, q^T x, q^T x^2
// One call = one pass over the training descriptors; the sums leave in fp64 and the (tiny) parameter update
// -- division, variance floor, convergence test -- is host arithmetic in the Python mirror, exactly where
// scikit-learn does it.
#include "pvs_kernels.cuh"

namespace pvs {
namespace {

// sums[label, :] += x[row, :] (fp64), counts[label] += 1, inertia += ||x - c_label||^2.  One warp per row;
// within a CTA rows are visited in order, the accumulation order across CTAs is whatever the atomics give
// (fp64: the spread is ~1e-16 relative, far below the fp32 parameters that are derived from the sums).
__global__ void __launch_bounds__(256)
cluster_sums_kernel(const float* __restrict__ x, const int32_t* __restrict__ labels, int64_t rows, int d, int k,
                    const float* __restrict__ centers, double* __restrict__ sums, unsigned long long* __restrict__ counts,
                    double* __restrict__ inertia)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warps = (int64_t)gridDim.x * 8;
    double in_acc = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += warps) {
        const int l = labels ? labels[r] : 0;
        if (l < 0 || l >= k) continue;
        const float* xr = x + r * d;
        const float* c = centers ? centers + (int64_t)l * d : nullptr;
        for (int j = lane; j < d; j += 32) {
            const float v = xr[j];
            atomicAdd(&sums[(int64_t)l * d + j], (double)v);
            if (c) { const double df = (double)v - (double)c[j]; in_acc += df * df; }
        }
        if (lane == 0) atomicAdd(&counts[l], 1ull);
    }
    if (inertia) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) in_acc += __shfl_xor_sync(0xffffffffu, in_acc, o);
        if (lane == 0 && in_acc != 0.0) atomicAdd(inertia, in_acc);
    }
}

__global__ void rows_sub_kernel(float* __restrict__ x, int64_t rows, int d, const float* __restrict__ v)
{
    const int64_t n = rows * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] -= v[i % d];
}

__global__ void slab_offsets_kernel(int64_t* __restrict__ offs, int64_t n_slabs, int64_t slab, int64_t rows)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_slabs) offs[i] = i * slab < rows ? i * slab : rows;
}

// logits [rows, k] -> posteriors in place + sum over rows of logsumexp (fp64).  One warp per row.
__global__ void __launch_bounds__(256)
softmax_lse_kernel(float* __restrict__ L, int64_t rows, int k, double* __restrict__ loglik)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    double lse_d = 0.0;
    if (row < rows) {
        float* p = L + row * k;
        float mx = -INFINITY;
        for (int j = lane; j < k; j += 32) mx = fmaxf(mx, p[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float base = isfinite(mx) ? mx : 0.f;
        float s = 0.f;
        for (int j = lane; j < k; j += 32) s += expf(p[j] - base);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float lse = logf(s) + base;
        for (int j = lane; j < k; j += 32) p[j] = expf(p[j] - lse);
        lse_d = (double)lse;
    }
    __shared__ double red[8];
    if (lane == 0) red[threadIdx.x >> 5] = lse_d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        if (loglik) atomicAdd(loglik, t);
    }
}

// per-slab statistics S [n_slabs, k, 2d+1] = ( q^T [x | x^2], sum q ) / T_slab  ->  fp64 totals
__global__ void __launch_bounds__(256)
em_fold_kernel(const float* __restrict__ S, const int64_t* __restrict__ offs, int64_t n_slabs, int k, int d,
               double* __restrict__ s0, double* __restrict__ s1, double* __restrict__ s2)
{
    const int ld = 2 * d + 1;
    const int64_t n = (int64_t)k * ld;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(e / ld), c = (int)(e - (int64_t)j * ld);
        double acc = 0.0;
        for (int64_t s = 0; s < n_slabs; ++s) {
            const double T = (double)(offs[s + 1] - offs[s]);
            if (T > 0.0) acc += (double)S[s * n + e] * T;
        }
        if (c < d) s1[(int64_t)j * d + c] += acc;
        else if (c < 2 * d) s2[(int64_t)j * d + (c - d)] += acc;
        else s0[j] += acc;
    }
}

constexpr int64_t EM_CHUNK_ROWS = 1 << 17;       // rows per E/M pass (bounds the posterior scratch)
constexpr int64_t EM_SLAB_ROWS = 512;            // pseudo-"images" of the statistics kernel (fp32 partial sums)

struct EmWs { size_t q, S, offs, total; int64_t chunk, slabs; };
EmWs em_ws(const pvs_model* g, int64_t rows)
{
    EmWs w{};
    w.chunk = rows < EM_CHUNK_ROWS ? rows : EM_CHUNK_ROWS;
    w.slabs = ceil_div(w.chunk, EM_SLAB_ROWS);
    size_t off = 0;
    w.q = off;    off += align_up((size_t)w.chunk * g->k * 4, 256);
    w.S = off;    off += align_up((size_t)w.slabs * g->k * (2 * g->d + 1) * 4, 256);
    w.offs = off; off += align_up((size_t)(w.slabs + 1) * 8, 256);
    w.total = off + 256;
    return w;
}
}  // namespace
}  // namespace pvs

using namespace pvs;

extern "C" int pvs_rows_sub(float* x_dev, int64_t rows, int d, const float* v_dev, void* stream)
{
    PVS_CHECK(rows >= 0 && d > 0 && (rows == 0 || (x_dev && v_dev)), PVS_ERR_BAD_ARG, "pvs_rows_sub: bad arguments");
    if (rows == 0) return PVS_OK;
    PVS_LAUNCH(rows_sub_kernel, 148 * 8, 256, 0, (cudaStream_t)stream, x_dev, rows, d, v_dev);
    return PVS_OK;
}

extern "C" int pvs_cluster_sums(const float* x_dev, const int32_t* labels_dev, int64_t rows, int d, int k,
                                const float* centers_dev, double* sums_dev, int64_t* counts_dev, double* inertia_dev,
                                void* stream)
{
    PVS_CHECK(rows >= 0 && d > 0 && k > 0, PVS_ERR_BAD_ARG, "pvs_cluster_sums: bad sizes");
    PVS_CHECK(sums_dev && counts_dev && (rows == 0 || x_dev), PVS_ERR_BAD_ARG, "pvs_cluster_sums: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    PVS_CUDA(cudaMemsetAsync(sums_dev, 0, (size_t)k * d * sizeof(double), st));
    PVS_CUDA(cudaMemsetAsync(counts_dev, 0, (size_t)k * sizeof(int64_t), st));
    if (inertia_dev) PVS_CUDA(cudaMemsetAsync(inertia_dev, 0, sizeof(double), st));
    if (rows == 0) return PVS_OK;
    const int64_t grid = ceil_div(rows, 8) < 148 * 8 ? ceil_div(rows, 8) : 148 * 8;
    PVS_LAUNCH(cluster_sums_kernel, (unsigned)grid, 256, 0, st, x_dev, labels_dev, rows, d, k, centers_dev, sums_dev,
               (unsigned long long*)counts_dev, inertia_dev);
    return PVS_OK;
}

extern "C" size_t pvs_kmeans_lloyd_workspace_bytes(const pvs_model* km, int64_t rows)
{
    return pvs_kmeans_assign_workspace_bytes(km, rows);
}

extern "C" int pvs_kmeans_lloyd_step(const pvs_model* km, const float* x_dev, int64_t rows, int32_t* labels_dev,
                                     double* sums_dev, int64_t* counts_dev, double* inertia_dev, void* workspace,
                                     size_t workspace_bytes, void* stream)
{
    PVS_CHECK(km && km->kind == PVS_MODEL_KMEANS, PVS_ERR_BAD_ARG, "pvs_kmeans_lloyd_step: not a K-Means model");
    PVS_CHECK(labels_dev || rows == 0, PVS_ERR_BAD_ARG, "pvs_kmeans_lloyd_step: NULL labels buffer");
    if (int rc = pvs_kmeans_assign(km, x_dev, rows, labels_dev, workspace, workspace_bytes, stream)) return rc;
    return pvs_cluster_sums(x_dev, labels_dev, rows, km->d, km->k, km->centers, sums_dev, counts_dev, inertia_dev, stream);
}

extern "C" size_t pvs_gmm_em_workspace_bytes(const pvs_model* g, int64_t rows)
{
    if (!g || g->kind != PVS_MODEL_GMM_DIAG || rows <= 0) return 0;
    return em_ws(g, rows).total;
}

extern "C" int pvs_gmm_em_step(const pvs_model* g, const float* x_dev, int64_t rows, double* s0_dev, double* s1_dev,
                               double* s2_dev, double* loglik_dev, void* workspace, size_t workspace_bytes, void* stream)
{
    PVS_CHECK(g && g->kind == PVS_MODEL_GMM_DIAG, PVS_ERR_BAD_ARG, "pvs_gmm_em_step: not a GMM model");
    PVS_CHECK(rows >= 0 && s0_dev && s1_dev && s2_dev && loglik_dev && (rows == 0 || x_dev), PVS_ERR_BAD_ARG,
              "pvs_gmm_em_step: NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t kd = (size_t)g->k * g->d;
    PVS_CUDA(cudaMemsetAsync(s0_dev, 0, (size_t)g->k * sizeof(double), st));
    PVS_CUDA(cudaMemsetAsync(s1_dev, 0, kd * sizeof(double), st));
    PVS_CUDA(cudaMemsetAsync(s2_dev, 0, kd * sizeof(double), st));
    PVS_CUDA(cudaMemsetAsync(loglik_dev, 0, sizeof(double), st));
    if (rows == 0) return PVS_OK;
    const EmWs w = em_ws(g, rows);
    PVS_CHECK(workspace && workspace_bytes >= w.total, PVS_ERR_WORKSPACE, "pvs_gmm_em_step: workspace %zu < required %zu",
              workspace_bytes, w.total);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float* q = (float*)(ws + w.q);
    float* S = (float*)(ws + w.S);
    int64_t* offs = (int64_t*)(ws + w.offs);
    for (int64_t r = 0; r < rows; r += w.chunk) {
        const int64_t n = rows - r < w.chunk ? rows - r : w.chunk;
        const float* y = x_dev + r * g->d;
        // E-step: logits on the tensor cores when the device has them (3xTF32), softmax + log-likelihood
        if (g_path.load() != PVS_PATH_SIMT && g->tcg0 && tc_gemm_nt_supported(n, g->k)) {
            if (int rc = tc_gemm_nt(y, g->d, g->d, true, g->tcg0, g->tcg1, g->tcg_ld, g->k, q, g->k, n, 1.f, g->cst, st)) return rc;
        } else if (int rc = launch_gemm_nt(y, g->d, g->wcat, 2 * g->d, q, g->k, n, g->k, g->d, 1, 1.f, g->cst, st)) return rc;
        PVS_LAUNCH(softmax_lse_kernel, (unsigned)ceil_div(n, 8), 256, 0, st, q, n, g->k, loglik_dev);
        // M-step accumulators: fp32 partial statistics per slab of rows, folded into the fp64 totals
        const int64_t slabs = ceil_div(n, EM_SLAB_ROWS);
        PVS_LAUNCH(slab_offsets_kernel, (unsigned)ceil_div(slabs + 1, 256), 256, 0, st, offs, slabs, EM_SLAB_ROWS, n);
        if (int rc = launch_fv_stats(q, y, g->d, g->k, offs, slabs, S, st)) return rc;
        PVS_LAUNCH(em_fold_kernel, (unsigned)ceil_div((int64_t)g->k * (2 * g->d + 1), 256), 256, 0, st, S, offs, slabs, g->k, g->d,
                   s0_dev, s1_dev, s2_dev);
    }
    return PVS_OK;
}
