// pvs_simt.cu -- fp32 CUDA-core kernels for the encode-and-compare path (any K, any D).
//
// These are the shape-generic kernels: every stage of the path exists here in plain fp32
// so that odd shapes (k=32 vocabularies, D=514 descriptors, ragged images) always have a
// GPU path.  The tcgen05 tensor-core kernels in pvs_tc_*.cu take over the contractions
// when the shape allows; the aggregation / normalisation / top-k kernels below are
// HBM- or latency-bound by nature and stay on CUDA cores.
//
// Reference routines restated (see include/pvs_b200.h for the per-export mapping):
//   gemm_nt           sklearn PCA.transform / KMeans scores / GMM log-prob contractions
//   row_softmax       sklearn mixture/_base.py:552-582 (logsumexp + exp)
//   row_argmin        sklearn _k_means_lloyd.pyx:_update_chunk_dense arg-min scan
//   vlad_aggregate    pyvisim/encoders/vlad.py:98-111
//   fv_stats/finalize pyvisim/encoders/fisher_vector.py:102-133
//   l2_normalize      sklearn preprocessing.normalize as used by cosine_similarity
//   topk_rows         pyvisim/eval.py:40-43 (np.argsort(-scores)[:k])
#include <cuda_fp16.h>
#include "pvs_kernels.cuh"

namespace pvs {

constexpr unsigned FULL = 0xffffffffu;

// =====================================================================================
// generic NT contraction
// =====================================================================================
namespace {
constexpr int G_BM = 128, G_BK = 16;

// C = alpha * A' B^T + bias, fp32 FMA on CUDA cores (the round-to-nearest path: long contractions, shapes the tensor kernels
// do not take, PVS_PATH_SIMT).  CTA tile 128 x (16 TN), 256 threads, thread tile 8 x TN as 4-wide groups 64 apart (the float4
// reads of a k-row are then conflict-free and the row reads broadcast), k-slabs of 16 double-buffered in shared memory with
// the global loads of the next slab in flight during the FMAs.  Every output element is ONE accumulator that takes its
// products in increasing k -- the same sequence whatever the tile width, so the wide and the narrow variant (and the earlier
// single-buffered kernel) give bit-identical results.
template <int TN, bool SQ>
__global__ void __launch_bounds__(256, TN == 8 ? 2 : 3)
gemm_nt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
               float* __restrict__ C, int64_t ldc, int64_t M, int N, int n_base, int a_cols, float alpha,
               const float* __restrict__ bias)
{
    constexpr int BN = 16 * TN, NB = BN / 16, CG = TN >= 4 ? TN / 4 : 1, CW = TN >= 4 ? 4 : TN;
    __shared__ __align__(16) float As[2][G_BK][G_BM + 4];
    __shared__ __align__(16) float Bs[2][G_BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * G_BM;
    const int n0 = n_base + blockIdx.y * BN;
    const int kdim = SQ ? 2 * a_cols : a_cols;
    const int ty = tid >> 4, tx = tid & 15;
    const int lc = tid & 15, lr = tid >> 4;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // this thread's rows of the two operands: one base pointer each, a constant step between them, validity as bit masks
    // (the loads below are predicated, not branched over)
    const float* arow = A + (m0 + lr) * lda;
    const float* brow = B + (int64_t)(n0 + lr) * ldb;
    const int64_t astep = 16 * lda, bstep = 16 * ldb;
    unsigned amask = 0, bmask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) amask |= (m0 + lr + 16 * i < M ? 1u : 0u) << i;
#pragma unroll
    for (int i = 0; i < NB; ++i) bmask |= (n0 + lr + 16 * i < N ? 1u : 0u) << i;
    float pa[8], pb[NB];
    auto gload = [&](int k0) {
        const int kk = k0 + lc;
        const bool kin = kk < kdim;
        const bool sqr = SQ && kk < a_cols;                  // A' = [a^2 | a]: the first a_cols columns are squares
        const int kc = (SQ && kk >= a_cols) ? kk - a_cols : kk;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = 0.f;
            if (kin && ((amask >> i) & 1u)) v = __ldg(arow + i * astep + kc);
            pa[i] = sqr ? v * v : v;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            float v = 0.f;
            if (kin && ((bmask >> i) & 1u)) v = __ldg(brow + i * bstep + kk);
            pb[i] = v;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][lc][lr + 16 * i] = pa[i];
#pragma unroll
        for (int i = 0; i < NB; ++i) Bs[buf][lc][lr + 16 * i] = pb[i];
    };
    const int n_slab = (kdim + G_BK - 1) / G_BK;
    gload(0);
    sstore(0);
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < n_slab; ++s) {
        const int buf = s & 1;
        if (s + 1 < n_slab) gload((s + 1) * G_BK);
#pragma unroll
        for (int k = 0; k < G_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bb[TN];
            if constexpr (TN >= 4) {
#pragma unroll
                for (int g = 0; g < CG; ++g) {
                    const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
                    bb[4 * g] = b.x; bb[4 * g + 1] = b.y; bb[4 * g + 2] = b.z; bb[4 * g + 3] = b.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) bb[j] = Bs[buf][k][tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        if (s + 1 < n_slab) sstore(buf ^ 1);                   // nobody reads that buffer between the two barriers
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + (TN >= 4 ? (j / CW) * 64 + tx * 4 + (j % CW) : tx * TN + j);
            if (n < N) C[m * ldc + n] = alpha * acc[i][j] + (bias ? bias[n] : 0.f);
        }
    }
}

// The last R <= 4 columns of an N that is a few columns past a multiple of 128 (the VGG16 PCA has N = 257): a 32-wide tile
// for one column costs as much as for 32.  Here a thread owns a row and its R accumulators, the k-slabs of A go through
// shared memory so that the global loads stay coalesced, and the products are added in increasing k like everywhere else.
template <int R, bool SQ>
__global__ void __launch_bounds__(256)
gemm_nt_skinny_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                      float* __restrict__ C, int64_t ldc, int64_t M, int N, int n_base, int a_cols, float alpha,
                      const float* __restrict__ bias)
{
    __shared__ float As[G_BK][256 + 1];
    __shared__ float Bs[G_BK][R];
    const int tid = threadIdx.x, lc = tid & 15, lr = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * 256;
    const int kdim = SQ ? 2 * a_cols : a_cols;
    const float* arow = A + (m0 + lr) * lda;
    const int64_t astep = 16 * lda;
    unsigned amask = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) amask |= (m0 + lr + 16 * i < M ? 1u : 0u) << i;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    for (int k0 = 0; k0 < kdim; k0 += G_BK) {
        const int kk = k0 + lc;
        const bool kin = kk < kdim;
        const bool sqr = SQ && kk < a_cols;
        const int kc = (SQ && kk >= a_cols) ? kk - a_cols : kk;
        float pa[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float v = 0.f;
            if (kin && ((amask >> i) & 1u)) v = __ldg(arow + i * astep + kc);
            pa[i] = sqr ? v * v : v;
        }
        __syncthreads();                                       // the previous slab has been consumed
#pragma unroll
        for (int i = 0; i < 16; ++i) As[lc][lr + 16 * i] = pa[i];
        if (tid < G_BK * R) {
            const int k = tid / R, r = tid - k * R, n = n_base + r;
            Bs[k][r] = (n < N && k0 + k < kdim) ? __ldg(B + (int64_t)n * ldb + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < G_BK; ++k) {
            const float a = As[k][tid];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fmaf(a, Bs[k][r], acc[r]);
        }
    }
    const int64_t m = m0 + tid;
    if (m < M) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = n_base + r;
            if (n < N) C[m * ldc + n] = alpha * acc[r] + (bias ? bias[n] : 0.f);
        }
    }
}

template <bool SQ>
int launch_gemm_nt_t(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                     int64_t M, int N, int a_cols, float alpha, const float* bias, cudaStream_t st)
{
    // full 128-column tiles on the wide variant, the remaining columns on a 64- or 32-wide tile or, if there are at most
    // four of them, on the row-per-thread kernel (N = 257: 2 wide tiles + 1 column)
    const int n_wide = N / 128, rem = N - 128 * n_wide, nb = 128 * n_wide;
    const unsigned gm = (unsigned)ceil_div(M, G_BM);
    if (n_wide > 0)
        PVS_LAUNCH((gemm_nt_kernel<8, SQ>), dim3(gm, (unsigned)n_wide), 256, 0, st, A, lda, B, ldb, C, ldc, M, N, 0, a_cols, alpha, bias);
    if (rem > 32)
        PVS_LAUNCH((gemm_nt_kernel<4, SQ>), dim3(gm, (unsigned)ceil_div(rem, 64)), 256, 0, st, A, lda, B, ldb, C, ldc, M, N, nb, a_cols,
                   alpha, bias);
    else if (rem > 4)
        PVS_LAUNCH((gemm_nt_kernel<2, SQ>), dim3(gm, 1), 256, 0, st, A, lda, B, ldb, C, ldc, M, N, nb, a_cols, alpha, bias);
    else if (rem > 1)
        PVS_LAUNCH((gemm_nt_skinny_kernel<4, SQ>), dim3((unsigned)ceil_div(M, 256)), 256, 0, st, A, lda, B, ldb, C, ldc, M, N, nb, a_cols,
                   alpha, bias);
    else if (rem == 1)
        PVS_LAUNCH((gemm_nt_skinny_kernel<1, SQ>), dim3((unsigned)ceil_div(M, 256)), 256, 0, st, A, lda, B, ldb, C, ldc, M, N, nb, a_cols,
                   alpha, bias);
    return PVS_OK;
}
}  // namespace

int launch_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                   int64_t M, int N, int a_cols, int square_cat, float alpha, const float* bias,
                   cudaStream_t st)
{
    if (M <= 0 || N <= 0) return PVS_OK;
    return square_cat ? launch_gemm_nt_t<true>(A, lda, B, ldb, C, ldc, M, N, a_cols, alpha, bias, st)
                      : launch_gemm_nt_t<false>(A, lda, B, ldb, C, ldc, M, N, a_cols, alpha, bias, st);
}

// =====================================================================================
// row softmax / arg-min (one warp per row)
// =====================================================================================
namespace {
__global__ void __launch_bounds__(256)
row_softmax_kernel(float* __restrict__ L, int64_t rows, int k, int32_t* __restrict__ argmax_out)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    float* p = L + row * k;
    float mx = -INFINITY;
    int mi = 0x7fffffff;
    for (int j = lane; j < k; j += 32) {
        const float v = p[j];
        if (v > mx) { mx = v; mi = j; }      // increasing j per lane: first max kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, mx, o);
        const int oi = __shfl_xor_sync(FULL, mi, o);
        if (ov > mx || (ov == mx && oi < mi)) { mx = ov; mi = oi; }
    }
    const float base = isfinite(mx) ? mx : 0.f;
    float s = 0.f;
    for (int j = lane; j < k; j += 32) s += expf(p[j] - base);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    const float lse = logf(s) + base;
    for (int j = lane; j < k; j += 32) p[j] = expf(p[j] - lse);
    if (argmax_out && lane == 0) argmax_out[row] = mi;
}

__global__ void __launch_bounds__(256)
row_argmin_kernel(const float* __restrict__ S, int64_t rows, int k, int32_t* __restrict__ labels)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* p = S + row * k;
    float mn = INFINITY;
    int mi = 0x7fffffff;
    for (int j = lane; j < k; j += 32) {
        const float v = p[j];
        if (v < mn) { mn = v; mi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL, mn, o);
        const int oi = __shfl_xor_sync(FULL, mi, o);
        if (ov < mn || (ov == mn && oi < mi)) { mn = ov; mi = oi; }
    }
    if (lane == 0) labels[row] = (mi == 0x7fffffff) ? 0 : mi;
}
}  // namespace

int launch_row_softmax(float* L, int64_t rows, int k, int32_t* argmax_out, cudaStream_t st)
{
    if (rows <= 0) return PVS_OK;
    PVS_LAUNCH(row_softmax_kernel, (unsigned)ceil_div(rows, 8), 256, 0, st, L, rows, k, argmax_out);
    return PVS_OK;
}

int launch_row_argmin(const float* S, int64_t rows, int k, int32_t* labels, cudaStream_t st)
{
    if (rows <= 0) return PVS_OK;
    PVS_LAUNCH(row_argmin_kernel, (unsigned)ceil_div(rows, 8), 256, 0, st, S, rows, k, labels);
    return PVS_OK;
}

// =====================================================================================
// shared normalisation helpers
// =====================================================================================
__device__ __forceinline__ float signed_pow(float v, float p)
{
    if (p == 1.f) return v;
    const float a = fabsf(v);
    const float r = (p == 0.5f) ? sqrtf(a) : powf(a, p);
    return v > 0.f ? r : (v < 0.f ? -r : (isnan(v) ? v : 0.f));   // np.sign(0) * x = 0
}
// contribution of one element to an ord-norm accumulator
__device__ __forceinline__ float norm_term(float v, float ord)
{
    const float a = fabsf(v);
    if (ord == 2.f) return a * a;
    if (ord == 1.f) return a;
    if (isinf(ord)) return a;          // combined with max
    return powf(a, ord);
}
__device__ __forceinline__ float norm_combine(float x, float y, float ord) { return isinf(ord) ? fmaxf(x, y) : x + y; }
__device__ __forceinline__ float norm_finish(float acc, float ord)
{
    if (ord == 2.f) return sqrtf(acc);
    if (ord == 1.f || isinf(ord)) return acc;
    return powf(acc, 1.f / ord);
}

// =====================================================================================
// VLAD aggregation (vlad.py:98-111).  One CTA per image:
//   1. histogram of the image's labels in shared memory, exclusive scan -> start[k]
//   2. stable placement (warp 0, __match_any_sync): members[] = descriptor indices grouped
//      by cluster, in descriptor order inside each cluster
//   3. one warp per cluster walks its member list and accumulates (x_t - c) in registers in
//      exactly the reference's sequential order, then applies the signed power and the
//      per-cluster ord-norm and writes the K x D block once (empty clusters: zeros).
// Every descriptor row is read once (coalesced, VEC floats per lane) and every output row
// written once; the label scan that the previous kernel repeated per cluster is gone.
// Images with more descriptors than the shared-memory member list holds fall back to a
// ballot scan of the labels per cluster (same arithmetic, same order).
// =====================================================================================
namespace {
template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float (&v)[VEC])
{
    if constexpr (VEC == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VEC == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y; }
    else v[0] = __ldg(p);
}
template <int VEC>
__device__ __forceinline__ void st_vec(float* p, const float (&v)[VEC])
{
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else *p = v[0];
}

// acc += row - cv for one member row; lane owns elements (lane + 32 j) * VEC + v
template <int VEC, int N>
__device__ __forceinline__ void add_row(const float* __restrict__ row, int d, int lane, const float (&cv)[N][VEC], float (&acc)[N][VEC])
{
    float x[N][VEC];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const int e = (lane + 32 * j) * VEC;
        if (e < d) ld_vec<VEC>(row + e, x[j]);
        else
#pragma unroll
            for (int v = 0; v < VEC; ++v) x[j][v] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[j][v] += x[j][v] - cv[j][v];
}

constexpr int VA_THREADS = 128;      // four warps: registers (up to ~100 with N * VEC = 18) still allow 5 CTAs per SM
template <int VEC, int N, bool FAST>
__global__ void __launch_bounds__(VA_THREADS)
vlad_aggregate_kernel(const float* __restrict__ y, int d, const int32_t* __restrict__ labels,
                      const int64_t* __restrict__ offsets, const float* __restrict__ centers, int k,
                      int k_per_cta, int t_cap, float power, float ord, float eps, float* __restrict__ out)
{
    extern __shared__ int sm_i[];
    constexpr int NW = VA_THREADS / 32;
    int* count = sm_i;                 // [k]     members per cluster
    int* start = sm_i + k;             // [k+1]   exclusive scan
    int* wcur = start + k + 1;         // [NW][k] per-warp histogram, then per-warp fill cursor
    int* members = wcur + NW * k;      // [t_cap]
    const int64_t img = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = NW;
    const int64_t r0 = offsets[img];
    const int T = (int)(offsets[img + 1] - r0);
    const int kbeg = blockIdx.y * k_per_cta;
    const int kend = min(k, kbeg + k_per_cta);
    const int32_t* lab = labels + r0;
    const bool sorted = T <= t_cap;

    if (sorted) {
        // every warp owns a contiguous range of the image's descriptors; because the ranges
        // are ordered and each warp places its own range in order, the member lists come
        // out in descriptor order without any warp waiting for another
        const int chunk = ((T + NW - 1) / NW + 31) & ~31;
        const int tb = warp * chunk, te = min(T, tb + chunk);
        for (int j = tid; j < NW * k; j += VA_THREADS) wcur[j] = 0;
        __syncthreads();
        for (int t = tb + lane; t < te; t += 32) atomicAdd(&wcur[warp * k + lab[t]], 1);
        __syncthreads();
        for (int j = tid; j < k; j += VA_THREADS) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) tot += wcur[w * k + j];
            count[j] = tot;
        }
        __syncthreads();
        if (warp == 0) {
            int carry = 0;
            for (int base = 0; base < k; base += 32) {
                const int v = base + lane < k ? count[base + lane] : 0;
                int incl = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += u; }
                if (base + lane < k) start[base + lane] = carry + incl - v;
                carry += __shfl_sync(FULL, incl, 31);
            }
            if (lane == 0) start[k] = carry;
        }
        __syncthreads();
        for (int j = tid; j < k; j += VA_THREADS) {            // histogram -> first free slot per (warp, cluster)
            int run = start[j];
#pragma unroll
            for (int w = 0; w < NW; ++w) { const int c = wcur[w * k + j]; wcur[w * k + j] = run; run += c; }
        }
        __syncthreads();
        int* cur = wcur + warp * k;
        for (int t0 = tb; t0 < te; t0 += 32) {                  // stable placement, 32 descriptors at a time
            const int t = t0 + lane;
            const bool valid = t < te;
            const int l = valid ? lab[t] : -1 - lane;             // invalid lanes match nobody
            const unsigned m = __match_any_sync(FULL, l);
            const int rank = __popc(m & ((1u << lane) - 1u));
            int base = 0;
            if (valid) base = cur[l];
            __syncwarp();
            if (valid) {
                members[base + rank] = t;
                if (rank == 0) cur[l] = base + __popc(m);
            }
            __syncwarp();
        }
        __syncthreads();
    }

    for (int c = kbeg + warp; c < kend; c += nw) {
        float acc[N][VEC], cv[N][VEC];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int e = (lane + 32 * j) * VEC;
#pragma unroll
            for (int v = 0; v < VEC; ++v) { acc[j][v] = 0.f; cv[j][v] = 0.f; }
            if (e < d) ld_vec<VEC>(centers + (int64_t)c * d + e, cv[j]);
        }
        int n_members;
        if (sorted) {
            const int s0 = start[c];
            n_members = start[c + 1] - s0;
            // up to RIF member rows in flight per warp (more for short rows): one memory round
            // trip per RIF members, and the rows are still added strictly in order
            constexpr int RIF = N * VEC <= 4 ? 8 : N * VEC <= 8 ? 4 : 2;
            for (int i = 0; i < n_members; i += RIF) {
                float x[RIF][N][VEC];
#pragma unroll
                for (int u = 0; u < RIF; ++u) {
                    if (i + u < n_members) {                                 // warp-uniform
                        const float* row = y + (r0 + members[s0 + i + u]) * (int64_t)d;
#pragma unroll
                        for (int j = 0; j < N; ++j) {
                            const int e = (lane + 32 * j) * VEC;
#pragma unroll
                            for (int v = 0; v < VEC; ++v) x[u][j][v] = 0.f;
                            if (e < d) ld_vec<VEC>(row + e, x[u][j]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < RIF; ++u) {
                    if (i + u < n_members) {
#pragma unroll
                        for (int j = 0; j < N; ++j)
#pragma unroll
                            for (int v = 0; v < VEC; ++v) acc[j][v] += x[u][j][v] - cv[j][v];
                    }
                }
            }
        } else {
            n_members = 0;
            for (int t0 = 0; t0 < T; t0 += 32) {
                const int l = (t0 + lane < T) ? lab[t0 + lane] : -1;
                unsigned m = __ballot_sync(FULL, l == c);
                n_members += __popc(m);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    add_row<VEC, N>(y + (r0 + t0 + b) * (int64_t)d, d, lane, cv, acc);
                }
            }
        }
        float* orow = out + (img * k + c) * (int64_t)d;
        if (n_members == 0) {                                              // 0 / (0 + eps) = 0
            const float z[VEC] = {};
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const int e = (lane + 32 * j) * VEC;
                if (e < d) st_vec<VEC>(orow + e, z);
            }
            continue;
        }
        float nrm = 0.f;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int e = (lane + 32 * j) * VEC;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if constexpr (FAST) {
                    nrm = fmaf(acc[j][v], acc[j][v], nrm);                 // padding lanes hold 0
                } else {
                    acc[j][v] = signed_pow(acc[j][v], power);
                    if (e < d) nrm = norm_combine(nrm, norm_term(acc[j][v], ord), ord);
                }
            }
        }
        float den;
        if constexpr (FAST) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(FULL, nrm, o);
            den = 1.f / (sqrtf(nrm) + eps);                                // one reciprocal, then multiplies
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nrm = norm_combine(nrm, __shfl_xor_sync(FULL, nrm, o), ord);
            den = norm_finish(nrm, ord) + eps;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const int e = (lane + 32 * j) * VEC;
            float o[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) o[v] = FAST ? acc[j][v] * den : acc[j][v] / den;
            if (e < d) st_vec<VEC>(orow + e, o);
        }
    }
}

template <int VEC, int N>
int launch_vlad_agg(dim3 grid, size_t smem, cudaStream_t st, bool fast, const float* y, int d, const int32_t* labels,
                    const int64_t* offsets, const float* centers, int k, int k_per_cta, int t_cap, float power, float ord,
                    float eps, float* out)
{
    if (fast) {
        if (smem > 48 * 1024) PVS_CUDA(cudaFuncSetAttribute(vlad_aggregate_kernel<VEC, N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PVS_LAUNCH((vlad_aggregate_kernel<VEC, N, true>), grid, VA_THREADS, smem, st, y, d, labels, offsets, centers, k, k_per_cta, t_cap, power, ord, eps, out);
    } else {
        if (smem > 48 * 1024) PVS_CUDA(cudaFuncSetAttribute(vlad_aggregate_kernel<VEC, N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PVS_LAUNCH((vlad_aggregate_kernel<VEC, N, false>), grid, VA_THREADS, smem, st, y, d, labels, offsets, centers, k, k_per_cta, t_cap, power, ord, eps, out);
    }
    return PVS_OK;
}
}  // namespace

int launch_vlad_aggregate(const float* y, int d, const int32_t* labels, const int64_t* offsets,
                          int64_t n_images, int64_t total_rows, const float* centers, int k, float power,
                          float norm_order, float eps, float* out, cudaStream_t st)
{
    if (n_images <= 0) return PVS_OK;
    PVS_CHECK(d <= 2048, PVS_ERR_UNSUPPORTED, "VLAD aggregation supports d <= 2048 (got %d)", d);
    PVS_CHECK(k <= 8192, PVS_ERR_UNSUPPORTED, "VLAD aggregation supports k <= 8192 (got %d)", k);
    // enough CTAs to fill the machine when there are few images (README quick start: 2)
    // cluster ranges per image: enough CTAs (4 warps each) for ~32 warps per SM even when
    // there are few images (README quick start: 2)
    int groups = 1;
    while (groups < 32 && n_images * groups * (VA_THREADS / 32) < 148 * 32 && (k / (groups * 2)) >= 8) groups *= 2;
    const int k_per_cta = (int)ceil_div(k, groups);
    dim3 grid((unsigned)n_images, (unsigned)ceil_div(k, k_per_cta));
    // member-list capacity: a few times the mean image size, so ordinary images take the
    // sorted path and shared memory still leaves several CTAs per SM
    int64_t avg = total_rows / n_images + 1;
    int t_cap = 1024;
    while (t_cap < 4 * avg && t_cap < 32768) t_cap <<= 1;
    const size_t smem = ((size_t)(2 + VA_THREADS / 32) * k + 1 + t_cap) * sizeof(int);
    const bool fast = power == 1.f && norm_order == 2.f;
    const bool a16 = d % 4 == 0 && (((uintptr_t)y | (uintptr_t)centers | (uintptr_t)out) & 15) == 0;
    const bool a8 = d % 2 == 0 && (((uintptr_t)y | (uintptr_t)centers | (uintptr_t)out) & 7) == 0;
#define VA(V, NN) return launch_vlad_agg<V, NN>(grid, smem, st, fast, y, d, labels, offsets, centers, k, k_per_cta, t_cap, power, norm_order, eps, out)
    if (a16) {
        const int n = (int)ceil_div(d / 4, 32);
        if (n <= 1) VA(4, 1);
        if (n <= 2) VA(4, 2);
        if (n <= 4) VA(4, 4);
        if (n <= 8) VA(4, 8);
        VA(4, 16);
    }
    if (a8) {
        const int n = (int)ceil_div(d / 2, 32);
        if (n <= 1) VA(2, 1);
        if (n <= 2) VA(2, 2);
        if (n <= 4) VA(2, 4);
        if (n <= 9) VA(2, 9);
        if (n <= 16) VA(2, 16);
        VA(2, 32);
    }
    const int n = (int)ceil_div(d, 32);
    if (n <= 2) VA(1, 2);
    if (n <= 4) VA(1, 4);
    if (n <= 8) VA(1, 8);
    if (n <= 17) VA(1, 17);
    if (n <= 32) VA(1, 32);
    VA(1, 64);
#undef VA
}

// =====================================================================================
// Fisher-vector statistics: per image S = q^T [y | y*y] / T  and  s0 = sum_t q / T
// =====================================================================================
namespace {
constexpr int S_BM = 64, S_BN = 64, S_BK = 16;

__global__ void __launch_bounds__(256)
fv_stats_kernel(const float* __restrict__ q, const float* __restrict__ y, int d, int k,
                const int64_t* __restrict__ offsets, float* __restrict__ S)
{
    __shared__ __align__(16) float As[S_BK][S_BM + 4];
    __shared__ __align__(16) float Bs[S_BK][S_BN + 4];
    const int64_t img = blockIdx.z;
    const int64_t r0 = offsets[img];
    const int T = (int)(offsets[img + 1] - r0);
    const int j0 = blockIdx.y * S_BM, n0 = blockIdx.x * S_BN;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lc = tid & 63, lr = tid >> 6;            // loader: column, row (0..3)
    const int nd = 2 * d;

    float acc[4][4];
    float s0[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t0 = 0; t0 < T; t0 += S_BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int kk = lr + 4 * i;
            const int t = t0 + kk;
            const int j = j0 + lc, n = n0 + lc;
            float a = 0.f, b = 0.f;
            if (t < T) {
                if (j < k) a = q[(r0 + t) * (int64_t)k + j];
                if (n < d) b = y[(r0 + t) * (int64_t)d + n];
                else if (n < nd) { const float v = y[(r0 + t) * (int64_t)d + (n - d)]; b = v * v; }
            }
            As[kk][lc] = a;
            Bs[kk][lc] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < S_BK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                s0[i] += a[i];
            }
        }
        __syncthreads();
    }
    const float inv_t = 1.f / (float)T;          // T == 0 -> inf -> NaN, as the reference
    const int ld = nd + 1;
    float* Simg = S + img * (int64_t)k * ld;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = j0 + ty * 4 + i;
        if (j >= k) continue;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int n = n0 + tx * 4 + jj;
            if (n < nd) Simg[(int64_t)j * ld + n] = acc[i][jj] * inv_t;
        }
        if (blockIdx.x == 0 && tx == 0) Simg[(int64_t)j * ld + nd] = s0[i] * inv_t;
    }
}

// gradients wrt (pi, mu, sigma), analytic normalisation, signed power, global ord-norm
// (fisher_vector.py:107-132).  One CTA of 1024 threads per image.  The kernel is bound by
// the latency of its loads (statistics + four model tables per element), so it is written
// for memory-level parallelism: s0 goes to shared memory first, then every thread handles
// four independent elements per step with all their loads issued before the arithmetic.
// Two passes over the statistics (the second one hits L1/L2): the first only accumulates
// the norm, the second recomputes each element and writes it scaled, so the output is
// written exactly once and never read back.  FAST = (power 0.5, ord 2), the defaults:
// |sqrt|x||^2 = |x|, so the norm pass needs no square roots at all.
template <bool FAST>
__global__ void __launch_bounds__(1024)
fv_finalize_kernel(const float* __restrict__ S, int ld, const float* __restrict__ s0part, int parts,
                   const int64_t* __restrict__ offsets, int k, int d, const float* __restrict__ mu,
                   const float* __restrict__ var, const float* __restrict__ pi,
                   const float* __restrict__ g_pi, const float* __restrict__ g_mu,
                   const float* __restrict__ g_sig, float power, float ord, float eps,
                   float* __restrict__ out, float raw1, float raw2, const int* __restrict__ raw_gate, float rawg)
{
    extern __shared__ float s0s[];                          // [k]
    __shared__ float red[32];
    __shared__ float s_den;
    const int64_t img = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int kd = k * d;
    const float* Simg = S + img * (int64_t)k * ld;
    float* o = out + img * (int64_t)(2 * kd + k);
    // raw1 != 0: S holds raw sums in operand units (segment-folded by the fused kernel): first-order columns still need
    // raw1 / T, second-order columns raw2 / T -- unless the range flag went up and the 3xTF32 kernels wrote S / T instead
    // With the range flag up the 3xTF32 kernels wrote S instead: raw as well, in units of rawg (0 = already S / T).
    const bool gated = raw_gate && *raw_gate != 0;
    const float r1 = gated ? rawg : raw1, r2 = gated ? rawg : raw2;
    const bool raw = r1 != 0.f;
    const float f1 = raw ? r1 / (float)(offsets[img + 1] - offsets[img]) : 1.f;
    const float f2 = raw ? r2 / (float)(offsets[img + 1] - offsets[img]) : 1.f;
    if (s0part) {
        const float inv_t = 1.f / (float)(offsets[img + 1] - offsets[img]);   // T == 0 -> inf -> NaN, as the reference
        for (int j = tid; j < k; j += nt) {
            float t = 0.f;
            for (int q = 0; q < parts; ++q) t += s0part[(img * parts + q) * (int64_t)k + j];
            s0s[j] = t * inv_t;
        }
    } else {
        for (int j = tid; j < k; j += nt) s0s[j] = Simg[(int64_t)j * ld + 2 * d];
    }
    __syncthreads();
    float den = 0.f;
    // element e = j * d + dd.  When d divides the block size a thread keeps dd fixed and only
    // steps j, so the index needs no division (d = 64 / 128 in every bundled model);
    // otherwise e is strided and split with one integer division.
    const bool fixed_dd = (nt % d) == 0;
    const int dd_f = tid % d, j_f = tid / d, j_step = nt / d;
    const int n_it = fixed_dd ? (k - j_f + j_step - 1) / j_step : (kd - tid + nt - 1) / nt;
    auto fast_root = [](float x) {                       // sign(x) sqrt|x| without the IEEE sqrt sequence
        const float a = fabsf(x);
        return a > 0.f ? copysignf(a * rsqrtf(a), x) : x;  // keeps +-0 and propagates NaN
    };
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        float part = 0.f;
        const float inv_den = pass ? 1.f / den : 0.f;
        for (int it0 = 0; it0 < n_it; it0 += 4) {
            float s1[4], s2[4], m[4], v[4], gm[4], gs[4], s0[4];
            int ee[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u;
                int j, dd;
                if (fixed_dd) { j = j_f + it * j_step; dd = dd_f; }
                else { const int e = tid + it * nt; j = e / d; dd = e - j * d; }
                ee[u] = it < n_it ? j * d + dd : -1;
                if (ee[u] >= 0) {
                    const float* Sj = Simg + (unsigned)(j * ld);
                    s1[u] = Sj[dd] * f1;
                    s2[u] = Sj[d + dd] * f2;
                    m[u] = mu[ee[u]]; v[u] = var[ee[u]]; gm[u] = g_mu[ee[u]]; gs[u] = g_sig[ee[u]];
                    s0[u] = s0s[j];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = ee[u];
                if (e >= 0) {
                    const float dm = (s1[u] - s0[u] * m[u]) * gm[u];
                    const float ds = (-s2[u] - s0[u] * m[u] * m[u] + s0[u] * v[u] + 2.f * s1[u] * m[u]) * gs[u];
                    if (pass == 0) {
                        if (FAST) part += fabsf(dm) + fabsf(ds);
                        else {
                            part = norm_combine(part, norm_term(signed_pow(dm, power), ord), ord);
                            part = norm_combine(part, norm_term(signed_pow(ds, power), ord), ord);
                        }
                    } else if (FAST) {
                        o[k + e] = fast_root(dm) * inv_den;
                        o[k + kd + e] = fast_root(ds) * inv_den;
                    } else {
                        o[k + e] = signed_pow(dm, power) / den;
                        o[k + kd + e] = signed_pow(ds, power) / den;
                    }
                }
            }
        }
        for (int j = tid; j < k; j += nt) {
            const float dp = (s0s[j] - pi[j]) * g_pi[j];
            if (pass == 0) part = FAST ? part + fabsf(dp) : norm_combine(part, norm_term(signed_pow(dp, power), ord), ord);
            else o[j] = FAST ? fast_root(dp) * inv_den : signed_pow(dp, power) / den;
        }
        if (pass == 0) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part = norm_combine(part, __shfl_xor_sync(FULL, part, off), ord);
            if ((tid & 31) == 0) red[tid >> 5] = part;
            __syncthreads();
            if (tid == 0) {
                float t = red[0];
                for (int w = 1; w < (nt >> 5); ++w) t = norm_combine(t, red[w], ord);
                s_den = (FAST ? sqrtf(t) : norm_finish(t, ord)) + eps;
            }
            __syncthreads();
            den = s_den;
        }
    }
}

// The same for the headline shape (K = 256, D = 64, power 0.5, ord 2, zeroth-order sums given as
// partials): one 1024-thread CTA per image, thread = (dimension dd, sixteen components j0 + 16 it).
// Compile-time strides and a fully unrolled component loop: all loads of the image are issued up
// front, the gradients stay in registers between the norm pass and the write pass (no second read
// of the statistics or the model tables), ~4x fewer instructions than the generic kernel.
__global__ void __launch_bounds__(1024)
fv_finalize_k256_d64_kernel(const float* __restrict__ S, const float* __restrict__ s0part, int parts,
                            const int64_t* __restrict__ offsets, const float* __restrict__ mu,
                            const float* __restrict__ var, const float* __restrict__ pi,
                            const float* __restrict__ g_pi, const float* __restrict__ g_mu,
                            const float* __restrict__ g_sig, float eps, float* __restrict__ out, float raw1, float raw2,
                            const int* __restrict__ raw_gate, float rawg)
{
    constexpr int K = 256, D = 64, KD = K * D, NJ = 16;
    __shared__ float s0s[K];
    __shared__ float red[32];
    __shared__ float s_den;
    const int64_t img = blockIdx.x;
    const int tid = threadIdx.x, dd = tid & (D - 1), j0 = tid >> 6;
    const float* Simg = S + img * (int64_t)(K * 2 * D);
    float* o = out + img * (int64_t)(2 * KD + K);
    auto fast_root = [](float x) {                       // sign(x) sqrt|x| without the IEEE sqrt sequence
        const float a = fabsf(x);
        return a > 0.f ? copysignf(a * rsqrtf(a), x) : x;  // keeps +-0 and propagates NaN
    };
    // issue the loads of this thread's 16 (component, dimension) elements before anything waits
    float s1[NJ], s2[NJ];
#pragma unroll
    for (int it = 0; it < NJ; ++it) {
        const int j = j0 + 16 * it;
        s1[it] = Simg[j * (2 * D) + dd];
        s2[it] = Simg[j * (2 * D) + D + dd];
    }
    const bool gated = raw_gate && *raw_gate != 0;
    const float r1 = gated ? rawg : raw1, r2 = gated ? rawg : raw2;
    if (r1 != 0.f) {                                     // raw sums in operand units (see fv_finalize_kernel)
        const float inv_t = 1.f / (float)(offsets[img + 1] - offsets[img]);
        const float f1 = r1 * inv_t, f2 = r2 * inv_t;
#pragma unroll
        for (int it = 0; it < NJ; ++it) { s1[it] *= f1; s2[it] *= f2; }
    }
    float dp = 0.f;
    if (tid < K) {
        const float inv_t = 1.f / (float)(offsets[img + 1] - offsets[img]);   // T == 0 -> inf -> NaN, as the reference
        float t = 0.f;
        for (int q = 0; q < parts; ++q) t += s0part[(img * parts + q) * (int64_t)K + tid];
        t *= inv_t;
        s0s[tid] = t;
        dp = (t - pi[tid]) * g_pi[tid];
    }
    __syncthreads();
    float part = tid < K ? fabsf(dp) : 0.f;
#pragma unroll
    for (int it = 0; it < NJ; ++it) {
        const int j = j0 + 16 * it, e = j * D + dd;
        const float m = mu[e], v = var[e], s0 = s0s[j];
        const float dm = (s1[it] - s0 * m) * g_mu[e];
        const float ds = (-s2[it] - s0 * m * m + s0 * v + 2.f * s1[it] * m) * g_sig[e];
        s1[it] = dm;
        s2[it] = ds;
        part += fabsf(dm) + fabsf(ds);                   // |sign(x) sqrt|x||^2 = |x|
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(FULL, part, off);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid < 32) {
        float t = red[tid];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(FULL, t, off);
        if (tid == 0) s_den = sqrtf(t) + eps;
    }
    __syncthreads();
    const float inv_den = 1.f / s_den;
    if (tid < K) o[tid] = fast_root(dp) * inv_den;
#pragma unroll
    for (int it = 0; it < NJ; ++it) {
        const int e = (j0 + 16 * it) * D + dd;
        o[K + e] = fast_root(s1[it]) * inv_den;
        o[K + KD + e] = fast_root(s2[it]) * inv_den;
    }
}
}  // namespace

int launch_fv_stats(const float* q, const float* y, int d, int k, const int64_t* offsets,
                    int64_t n_images, float* S, cudaStream_t st)
{
    if (n_images <= 0) return PVS_OK;
    for (int64_t i0 = 0; i0 < n_images; i0 += 65535) {       // gridDim.z limit
        const int64_t n = n_images - i0 < 65535 ? n_images - i0 : 65535;
        dim3 grid((unsigned)ceil_div(2 * d, S_BN), (unsigned)ceil_div(k, S_BM), (unsigned)n);
        PVS_LAUNCH(fv_stats_kernel, grid, 256, 0, st, q, y, d, k, offsets + i0, S + i0 * (int64_t)k * (2 * d + 1));
    }
    return PVS_OK;
}

int launch_fv_finalize(const float* S, int ld, const float* s0part, int parts, const int64_t* offsets,
                       const pvs_model* g, int64_t n_images, float power, float norm_order, float eps, float* out,
                       cudaStream_t st, float raw1, float raw2, const int* raw_gate, float rawg)
{
    if (n_images <= 0) return PVS_OK;
    PVS_CHECK(g->k <= 12000, PVS_ERR_UNSUPPORTED, "fv_finalize supports k <= 12000 (got %d)", g->k);
    if (power == 0.5f && norm_order == 2.f && g->k == 256 && g->d == 64 && ld == 128 && s0part)
        PVS_LAUNCH(fv_finalize_k256_d64_kernel, (unsigned)n_images, 1024, 0, st, S, s0part, parts, offsets, g->mu, g->var, g->pi,
                   g->g_pi, g->g_mu, g->g_sig, eps, out, raw1, raw2, raw_gate, rawg);
    else if (power == 0.5f && norm_order == 2.f)
        PVS_LAUNCH(fv_finalize_kernel<true>, (unsigned)n_images, 1024, (size_t)g->k * sizeof(float), st, S, ld, s0part, parts, offsets, g->k, g->d,
                   g->mu, g->var, g->pi, g->g_pi, g->g_mu, g->g_sig, power, norm_order, eps, out, raw1, raw2, raw_gate, rawg);
    else
        PVS_LAUNCH(fv_finalize_kernel<false>, (unsigned)n_images, 1024, (size_t)g->k * sizeof(float), st, S, ld, s0part, parts, offsets, g->k, g->d,
                   g->mu, g->var, g->pi, g->g_pi, g->g_mu, g->g_sig, power, norm_order, eps, out, raw1, raw2, raw_gate, rawg);
    return PVS_OK;
}

// =====================================================================================
// row L2 normalisation (fp32 -> fp32 / bf16), zero rows stay zero
// =====================================================================================
namespace {
template <typename OutT>
__global__ void __launch_bounds__(256)
l2_normalize_kernel(const float* __restrict__ x, int64_t d, OutT* __restrict__ out)
{
    __shared__ float red[8];
    __shared__ float s_inv;
    const float* p = x + (int64_t)blockIdx.x * d;
    OutT* o = out + (int64_t)blockIdx.x * d;
    float s = 0.f;
    for (int64_t j = threadIdx.x; j < d; j += blockDim.x) { const float v = p[j]; s = fmaf(v, v, s); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        const float n = sqrtf(t);
        s_inv = n == 0.f ? 1.f : n;
    }
    __syncthreads();
    const float n = s_inv;
    for (int64_t j = threadIdx.x; j < d; j += blockDim.x) {
        const float v = p[j] / n;
        if constexpr (sizeof(OutT) == 2) o[j] = __float2bfloat16_rn(v);
        else o[j] = v;
    }
}

// Row L2 normalisation into the operand format of the fp32-accurate tensor-core similarity (pvs_tc_sim3.cu):
// v = x / |x| in fp32 like the other variants, then v * 2^15 = hi + lo with both parts fp16 (|v| <= 1, so the
// scaled value fits fp16; lo stays a normal fp16 number down to |v| ~ 4e-6).  hi plane [n, d], lo plane right
// behind it (plane stride n * d).  hi + lo carries 22 mantissa bits of v.
__global__ void __launch_bounds__(256)
l2_normalize_split_kernel(const float* __restrict__ x, int64_t d, __half* __restrict__ hi, __half* __restrict__ lo)
{
    __shared__ float red[8];
    __shared__ float s_inv;
    const float* p = x + (int64_t)blockIdx.x * d;
    float s = 0.f;
    for (int64_t j = threadIdx.x; j < d; j += blockDim.x) { const float v = p[j]; s = fmaf(v, v, s); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        const float n = sqrtf(t);
        s_inv = n == 0.f ? 1.f : n;
    }
    __syncthreads();
    const float n = s_inv;
    __half* oh = hi + (int64_t)blockIdx.x * d;
    __half* ol = lo + (int64_t)blockIdx.x * d;
    for (int64_t j = threadIdx.x; j < d; j += blockDim.x) {
        const float v = (p[j] / n) * 32768.f;
        const __half h = __float2half_rn(v);
        oh[j] = h;
        ol[j] = __float2half_rn(v - __half2float(h));
    }
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, int64_t n, float* __restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __bfloat162float(x[i]);
}
}  // namespace

namespace {
// Small outputs (similarity_score of a few images): one CTA per (i, j) pair streams both rows
// once and reduces dot, |x|^2, |y|^2 together -- the tiled contraction above would put the
// whole K = d reduction of a 1 x 1 output on a single CTA.  Zero rows give 0, as sklearn's
// normalize leaves them zero.
__global__ void __launch_bounds__(256)
cosine_small_kernel(const float* __restrict__ x, const float* __restrict__ y, int m, int64_t d, float* __restrict__ s)
{
    __shared__ float red[3][8];
    const int i = blockIdx.x / m, j = blockIdx.x - i * m;
    const float* px = x + (int64_t)i * d;
    const float* py = y + (int64_t)j * d;
    float dot = 0.f, xx = 0.f, yy = 0.f;
    for (int64_t t = threadIdx.x; t < d; t += blockDim.x) {
        const float a = px[t], b = py[t];
        dot = fmaf(a, b, dot);
        xx = fmaf(a, a, xx);
        yy = fmaf(b, b, yy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dot += __shfl_xor_sync(FULL, dot, o);
        xx += __shfl_xor_sync(FULL, xx, o);
        yy += __shfl_xor_sync(FULL, yy, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dot; red[1][threadIdx.x >> 5] = xx; red[2][threadIdx.x >> 5] = yy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; }
        const float nx = sqrtf(b), ny = sqrtf(c);
        s[blockIdx.x] = (nx == 0.f || ny == 0.f) ? 0.f : (a / nx) / ny;
    }
}
}  // namespace

int launch_cosine_small(const float* x, int64_t n, const float* y, int64_t m, int64_t d, float* s, cudaStream_t st)
{
    if (n <= 0 || m <= 0) return PVS_OK;
    PVS_LAUNCH(cosine_small_kernel, (unsigned)(n * m), 256, 0, st, x, y, (int)m, d, s);
    return PVS_OK;
}

int launch_l2_normalize(const float* x, int64_t n, int64_t d, void* out, int out_dtype, cudaStream_t st)
{
    if (n <= 0) return PVS_OK;
    PVS_CHECK(n < 2147483647LL, PVS_ERR_BAD_SHAPE, "too many rows");
    if (out_dtype == PVS_F32) PVS_LAUNCH(l2_normalize_kernel<float>, (unsigned)n, 256, 0, st, x, d, (float*)out);
    else if (out_dtype == PVS_BF16) PVS_LAUNCH(l2_normalize_kernel<__nv_bfloat16>, (unsigned)n, 256, 0, st, x, d, (__nv_bfloat16*)out);
    else if (out_dtype == PVS_F16X2) PVS_LAUNCH(l2_normalize_split_kernel, (unsigned)n, 256, 0, st, x, d, (__half*)out, (__half*)out + n * d);
    else return fail(PVS_ERR_BAD_ARG, "unknown dtype %d", out_dtype);
    return PVS_OK;
}

// ALREADY normalised fp32 values -> the PVS_F16X2 operand planes (hi plane [n], lo plane [n] right behind it)
namespace {
__global__ void f32_to_split_kernel(const float* __restrict__ x, int64_t n, __half* __restrict__ hi, __half* __restrict__ lo)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float v = x[i] * 32768.f;
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn(v - __half2float(h));
    }
}
}  // namespace

int launch_f32_to_split(const float* x, int64_t n, void* out, cudaStream_t st)
{
    if (n <= 0) return PVS_OK;
    PVS_LAUNCH(f32_to_split_kernel, 148 * 8, 256, 0, st, x, n, (__half*)out, (__half*)out + n);
    return PVS_OK;
}

int launch_bf16_to_f32(const void* x, int64_t n, float* out, cudaStream_t st)
{
    if (n <= 0) return PVS_OK;
    PVS_LAUNCH(bf16_to_f32_kernel, 148 * 8, 256, 0, st, (const __nv_bfloat16*)x, n, out);
    return PVS_OK;
}

// =====================================================================================
// top-k
//   key = (orderable(score) << 32) | (0xffffffff - index): a plain descending sort of the
//   64-bit keys gives "score descending, lowest index first on ties".
// =====================================================================================
__device__ __forceinline__ unsigned long long topk_key(float s, unsigned idx)
{
    unsigned u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ float key_score(unsigned long long key)
{
    unsigned u = (unsigned)(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned key_index(unsigned long long key) { return 0xffffffffu - (unsigned)(key & 0xffffffffu); }

// in-place descending bitonic sort of n (power of two) keys in shared memory
__device__ void bitonic_desc(unsigned long long* buf, int n)
{
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int l = i ^ stride;
                if (l > i) {
                    const unsigned long long a = buf[i], b = buf[l];
                    const bool desc = (i & size) == 0;
                    if (desc ? (a < b) : (a > b)) { buf[i] = b; buf[l] = a; }
                }
            }
        }
    }
    __syncthreads();
}

namespace {
constexpr int TK_CAP = 2048;

__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ S, int64_t lds, int64_t n_db, int k, int64_t idx_offset,
                 float* __restrict__ scores_out, int64_t* __restrict__ idx_out, int64_t out_ld,
                 const unsigned long long* __restrict__ upper_in, unsigned long long* __restrict__ last_out)
{
    // upper_in (optional, per row): only keys strictly below it take part -- pass p of a k > TOPK_MAX ranking
    // continues below the last key of pass p - 1 (last_out).  Keys are unique, so the passes tile the order.
    const unsigned long long upper = upper_in ? upper_in[blockIdx.x] : ~0ull;
    __shared__ unsigned long long buf[TK_CAP];
    __shared__ int count;
    __shared__ unsigned long long tau;
    const float* row = S + (int64_t)blockIdx.x * lds;
    for (int i = threadIdx.x; i < TK_CAP; i += blockDim.x) buf[i] = 0ull;
    if (threadIdx.x == 0) { count = 0; tau = 0ull; }
    __syncthreads();
    // `cnt` mirrors the shared `count` in a register of every thread: it only changes by the result of a
    // block-wide __syncthreads_count, so the decision to prune is uniform by construction (reading the
    // shared counter here would race with the appends of warps that are already past the test).
    int cnt = 0;
    for (int64_t base = 0; base < n_db; base += blockDim.x) {
        if (cnt + (int)blockDim.x > TK_CAP) {
            bitonic_desc(buf, TK_CAP);
            if (threadIdx.x == 0) { tau = buf[k - 1]; count = k; }
            cnt = k;
            __syncthreads();
            for (int i = k + threadIdx.x; i < TK_CAP; i += blockDim.x) buf[i] = 0ull;
            __syncthreads();
        }
        const int64_t i = base + threadIdx.x;
        bool appended = false;
        if (i < n_db) {
            const unsigned long long key = topk_key(row[i], (unsigned)i);
            if (key > tau && key < upper) { buf[atomicAdd(&count, 1)] = key; appended = true; }
        }
        cnt += __syncthreads_count(appended);
    }
    bitonic_desc(buf, TK_CAP);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const unsigned long long key = buf[j];
        const bool valid = key != 0ull;
        scores_out[(int64_t)blockIdx.x * out_ld + j] = valid ? key_score(key) : -INFINITY;
        idx_out[(int64_t)blockIdx.x * out_ld + j] = valid ? (int64_t)key_index(key) + idx_offset : -1;
    }
    if (last_out && threadIdx.x == 0) last_out[blockIdx.x] = buf[k - 1];     // 0 once the row is exhausted
}

__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int parts, int64_t n_q,
                  int k, int n_pow2, float* __restrict__ scores_out, int64_t* __restrict__ idx_out)
{
    extern __shared__ unsigned long long mbuf[];
    const int64_t q = blockIdx.x;
    const int total = parts * k;
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < total) {
            const int p = i / k, j = i - p * k;
            const int64_t id = idx[((int64_t)p * n_q + q) * k + j];
            if (id >= 0) key = topk_key(scores[((int64_t)p * n_q + q) * k + j], (unsigned)id);
        }
        mbuf[i] = key;
    }
    bitonic_desc(mbuf, n_pow2);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const unsigned long long key = mbuf[j];
        const bool valid = key != 0ull;
        scores_out[q * k + j] = valid ? key_score(key) : -INFINITY;
        idx_out[q * k + j] = valid ? (int64_t)key_index(key) : -1;
    }
}

__global__ void label_metrics_kernel(const int64_t* __restrict__ idx, const int32_t* __restrict__ db_labels,
                                     const int32_t* __restrict__ q_labels, int64_t n_q, int k,
                                     int32_t* __restrict__ hits, float* __restrict__ ap)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_q) return;
    const int32_t lbl = q_labels[q];
    int rel = 0;
    double psum = 0.0;                                       // the reference sums Python floats (eval.py:88-92)
    for (int j = 0; j < k; ++j) {
        const int64_t id = idx[q * (int64_t)k + j];
        if (id >= 0 && db_labels[id] == lbl) { ++rel; psum += (double)rel / (double)(j + 1); }
    }
    if (hits) hits[q] = rel > 0;
    if (ap) ap[q] = rel > 0 ? (float)(psum / (double)rel) : 0.f;      // quirk Q6: R counted inside the list
}
}  // namespace

int launch_topk_rows(const float* S, int64_t lds, int64_t rows, int64_t n_db, int k, int64_t idx_offset,
                     float* scores_out, int64_t* idx_out, cudaStream_t st, int64_t out_ld,
                     const unsigned long long* upper_in, unsigned long long* last_out)
{
    if (rows <= 0) return PVS_OK;
    PVS_CHECK(k >= 1 && k <= PVS_TOPK_MAX, PVS_ERR_BAD_ARG, "k must be in [1, %d] per pass (got %d)", PVS_TOPK_MAX, k);
    PVS_CHECK(n_db < 4294967295LL, PVS_ERR_BAD_SHAPE, "database shard must have < 2^32 rows");
    PVS_LAUNCH(topk_rows_kernel, (unsigned)rows, 256, 0, st, S, lds, n_db, k, idx_offset, scores_out, idx_out,
               out_ld > 0 ? out_ld : (int64_t)k, upper_in, last_out);
    return PVS_OK;
}

int launch_topk_merge(const float* scores, const int64_t* idx, int parts, int64_t n_q, int k,
                      float* scores_out, int64_t* idx_out, cudaStream_t st)
{
    if (n_q <= 0) return PVS_OK;
    PVS_CHECK(k >= 1 && k <= PVS_TOPK_MAX && parts >= 1, PVS_ERR_BAD_ARG, "bad k/parts");
    int n_pow2 = 1;
    while (n_pow2 < parts * k) n_pow2 <<= 1;
    const size_t smem = (size_t)n_pow2 * sizeof(unsigned long long);
    PVS_CHECK(smem <= 200 * 1024, PVS_ERR_UNSUPPORTED, "parts*k = %d too large to merge in one pass", parts * k);
    if (smem > 48 * 1024)
        PVS_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PVS_LAUNCH(topk_merge_kernel, (unsigned)n_q, 256, smem, st, scores, idx, parts, n_q, k, n_pow2, scores_out, idx_out);
    return PVS_OK;
}

int launch_label_metrics(const int64_t* idx, const int32_t* db_labels, const int32_t* q_labels,
                         int64_t n_q, int k, int32_t* hits, float* ap, cudaStream_t st)
{
    if (n_q <= 0) return PVS_OK;
    PVS_LAUNCH(label_metrics_kernel, (unsigned)ceil_div(n_q, 256), 256, 0, st, idx, db_labels, q_labels, n_q, k, hits, ap);
    return PVS_OK;
}

// =====================================================================================
// uint8 -> float32 widening (uint8 transport of the host entry points): 16 bytes in, 64 bytes out per thread step
// =====================================================================================
namespace {
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint4* __restrict__ in, float4* __restrict__ out, size_t n16,
                                                          const uint8_t* __restrict__ tail_in, float* __restrict__ tail_out, int n_tail)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = __ldg(in + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            out[4 * i + q] = make_float4((float)(w[q] & 0xffu), (float)((w[q] >> 8) & 0xffu), (float)((w[q] >> 16) & 0xffu), (float)(w[q] >> 24));
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail_out[threadIdx.x] = (float)tail_in[threadIdx.x];
}
}  // namespace

int launch_u8_to_f32(const uint8_t* in, float* out, size_t n, cudaStream_t st)
{
    if (n == 0) return PVS_OK;
    const size_t n16 = n / 16;
    const int n_tail = (int)(n - n16 * 16);
    const unsigned grid = (unsigned)std::min<size_t>(std::max<size_t>((n16 + 255) / 256, 1), 148 * 16);
    PVS_LAUNCH(u8_to_f32_kernel, grid, 256, 0, st, (const uint4*)in, (float4*)out, n16, in + n16 * 16, out + n16 * 16, n_tail);
    return PVS_OK;
}

}  // namespace pvs
