// pvs_tc_fv.cu -- Fisher-vector contractions on tcgen05 tensor cores (3xTF32).
//
// Shape handled here: K = 256 components, D = 64 dims after the optional PCA (the
// SIFT/RootSIFT-PCA-64 configurations, i.e. BASELINE.json configs[1]); anything else takes
// the CUDA-core kernels in pvs_simt.cu.  Three contractions per batch, all instantiations
// of the skeleton in pvs_tc.cuh with hi/lo tf32 operand pairs:
//
//   project   Y[rows,64]    = X[rows,d_in] . C^T + b        K-major x K-major, N = 64
//             epilogue      -> Yaug = [ y*y | y ] as hi/lo pairs            (fisher_vector.py:92)
//   posterior L[rows,256]   = Yaug . [ -P/2 | mu P ]^T + c  K-major x K-major, N = 256
//             epilogue      -> softmax over the 256 columns of each TMEM lane, q as hi/lo
//                              pairs into an image-padded layout             (fisher_vector.py:99)
//   stats     S[128,256]    = Yaug_i^T . Q_i  per image     MN-major x MN-major, K = T_i
//             epilogue      -> S/T in the [k][2d+1] layout fv_finalize reads  (fisher_vector.py:102-104)
//
// Q rows of one image are padded with zero rows to a multiple of 32 so that the stats
// kernel can fetch fixed 32-row TMA boxes: the descriptor rows a box picks up past the end
// of the image then multiply zeros.
#include "pvs_tc.cuh"
#include "pvs_kernels.cuh"

namespace pvs {
namespace tc {

constexpr int FV_K = 256;     // mixture components
constexpr int FV_D = 64;      // dims after PCA
constexpr int FV_2D = 128;
constexpr int Q_PAD = 32;     // contraction rows per stats stage

// ---------------------------------------------------------------------------------------
// small preparation kernels
// ---------------------------------------------------------------------------------------
__global__ void split_kernel(const float4* __restrict__ x, int64_t n4, float4* __restrict__ hi, float4* __restrict__ lo)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = x[i];
        float4 h, l;
        tf32_split(v.x, h.x, l.x);
        tf32_split(v.y, h.y, l.y);
        tf32_split(v.z, h.z, l.z);
        tf32_split(v.w, h.w, l.w);
        hi[i] = h;
        lo[i] = l;
    }
}

// Yaug = [ y*y | y ] split into hi/lo, for models without PCA (y = raw descriptors, d = 64)
__global__ void yaug_kernel(const float* __restrict__ y, int64_t rows, float* __restrict__ hi, float* __restrict__ lo)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n = rows * FV_D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = i / FV_D;
        const int c = (int)(i - r * FV_D);
        const float v = y[i];
        float h, l;
        tf32_split(v * v, h, l);
        hi[r * FV_2D + c] = h;
        lo[r * FV_2D + c] = l;
        tf32_split(v, h, l);
        hi[r * FV_2D + FV_D + c] = h;
        lo[r * FV_2D + FV_D + c] = l;
    }
}

// qoff[i] = sum_{j<i} round_up(T_j, 32); one block, chunked scan with a running carry
__global__ void __launch_bounds__(1024) qoff_kernel(const int64_t* __restrict__ offsets, int64_t n_images,
                                                    int64_t* __restrict__ qoff)
{
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n_images; base += 1024) {
        const int64_t i = base + threadIdx.x;
        int64_t v = 0;
        if (i < n_images) v = (offsets[i + 1] - offsets[i] + Q_PAD - 1) / Q_PAD * Q_PAD;
        int64_t s = v;                                   // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sums[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;                          // inclusive over warps
        }
        __syncthreads();
        const int64_t before = carry + (warp ? warp_sums[warp - 1] : 0) + s - v;
        if (i < n_images) qoff[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) qoff[n_images] = carry;
}

// image that owns the first row of every 128-row posterior tile
__global__ void tile_image_kernel(const int64_t* __restrict__ offsets, int64_t n_images, int n_tiles,
                                  int* __restrict__ tile_img)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int64_t row = (int64_t)t * 128;
    int64_t lo = 0, hi = n_images;                        // last i with offsets[i] <= row
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= row) lo = mid; else hi = mid;
    }
    tile_img[t] = (int)lo;
}

// zero the padding rows [qoff[i] + T_i, qoff[i+1]) of both Q parts; one warp per image
__global__ void q_pad_zero_kernel(const int64_t* __restrict__ offsets, const int64_t* __restrict__ qoff,
                                  int64_t n_images, float* __restrict__ q_hi, float* __restrict__ q_lo)
{
    const int64_t img = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (img >= n_images) return;
    const int lane = threadIdx.x & 31;
    const int64_t t = offsets[img + 1] - offsets[img];
    const int64_t r0 = qoff[img] + t, r1 = qoff[img + 1];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0; r < r1; ++r) {
        float4* a = reinterpret_cast<float4*>(q_hi + r * FV_K);
        float4* b = reinterpret_cast<float4*>(q_lo + r * FV_K);
        a[lane] = z; a[lane + 32] = z;
        b[lane] = z; b[lane + 32] = z;
    }
}

// ---------------------------------------------------------------------------------------
// project: Y = X C^T + b  ->  Yaug hi/lo
// ---------------------------------------------------------------------------------------
struct PcaParams {
    CUtensorMap x_hi, x_lo, c_hi, c_lo;
    const float* bias;
    float* y_hi;
    float* y_lo;
    int64_t rows;
    int m_blocks, nkb;
};
struct NoEpiState {};

struct PcaPolicy {
    using Params = PcaParams;
    using EpiState = NoEpiState;
    struct Tile { int nkb, mb; };
    static constexpr bool BF16 = false, A_MN = false, B_MN = false, EPI_READS_STAGES = false;
    static constexpr int PASSES = 3, BLOCK_N = FV_D, KSTEPS = 4, STAGES = 4;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = BLOCK_N * 128, A_LBO = 0, B_LBO = 0, SCRATCH_BYTES = 256;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.x_hi); tma_prefetch_desc(&p.c_hi); }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks; }
    __device__ static Tile tile(const Params& p, int i) { return {p.nkb, i}; }
    __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* a_hi, uint8_t* a_lo, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        tma_load_2d(a_hi, &p.x_hi, bar, kb * 32, t.mb * 128);
        tma_load_2d(a_lo, &p.x_lo, bar, kb * 32, t.mb * 128);
        tma_load_2d(b_hi, &p.c_hi, bar, kb * 32, 0);
        tma_load_2d(b_lo, &p.c_lo, bar, kb * 32, 0);
    }
    __device__ static void epi_init(const Params& p, uint8_t* scratch, int tid)
    {
        float* b = reinterpret_cast<float*>(scratch);
        if (tid < FV_D) b[tid] = p.bias[tid];
        epi_barrier();
    }
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane,
                                    uint8_t* scratch, EpiState&)
    {
        const float* bias = reinterpret_cast<const float*>(scratch);
        const int64_t row = (int64_t)t.mb * 128 + quarter * 32 + lane;
        const bool valid = row < p.rows;
        float4* oh = reinterpret_cast<float4*>(p.y_hi + row * FV_2D);
        float4* ol = reinterpret_cast<float4*>(p.y_lo + row * FV_2D);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld32(tmem + half * 32, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float y[4], h[4], l[4], sh[4], sl[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        y[e] = v[j + e] + bias[half * 32 + j + e];
                        tf32_split(y[e], h[e], l[e]);
                        tf32_split(y[e] * y[e], sh[e], sl[e]);
                    }
                    const int c4 = (half * 32 + j) >> 2;                 // float4 index inside a 64-wide half
                    oh[c4] = make_float4(sh[0], sh[1], sh[2], sh[3]);    // y*y -> columns [0, 64)
                    ol[c4] = make_float4(sl[0], sl[1], sl[2], sl[3]);
                    oh[16 + c4] = make_float4(h[0], h[1], h[2], h[3]);   // y   -> columns [64, 128)
                    ol[16 + c4] = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// posterior: logits + softmax -> Q hi/lo (image-padded rows)
// ---------------------------------------------------------------------------------------
struct PostParams {
    CUtensorMap y_hi, y_lo, w_hi, w_lo;
    const float* cst;
    const int64_t* offsets;
    const int64_t* qoff;
    const int* tile_img;
    float* q_hi;
    float* q_lo;
    int32_t* argmax;
    int64_t rows, n_images;
    int m_blocks;
};

struct PostPolicy {
    using Params = PostParams;
    using EpiState = NoEpiState;
    struct Tile { int nkb, mb; };
    static constexpr bool BF16 = false, A_MN = false, B_MN = false, EPI_READS_STAGES = false;
    static constexpr int PASSES = 3, BLOCK_N = FV_K, KSTEPS = 4, STAGES = 2;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = BLOCK_N * 128, A_LBO = 0, B_LBO = 0, SCRATCH_BYTES = 1024;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.y_hi); tma_prefetch_desc(&p.w_hi); }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks; }
    __device__ static Tile tile(const Params&, int i) { return {FV_2D / 32, i}; }
    __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* a_hi, uint8_t* a_lo, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        tma_load_2d(a_hi, &p.y_hi, bar, kb * 32, t.mb * 128);
        tma_load_2d(a_lo, &p.y_lo, bar, kb * 32, t.mb * 128);
        tma_load_2d(b_hi, &p.w_hi, bar, kb * 32, 0);
        tma_load_2d(b_lo, &p.w_lo, bar, kb * 32, 0);
    }
    __device__ static void epi_init(const Params& p, uint8_t* scratch, int tid)
    {
        float* c = reinterpret_cast<float*>(scratch);
        c[tid] = p.cst[tid];
        c[tid + 128] = p.cst[tid + 128];
        epi_barrier();
    }
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane,
                                    uint8_t* scratch, EpiState&)
    {
        const float* cst = reinterpret_cast<const float*>(scratch);
        const int64_t row = (int64_t)t.mb * 128 + quarter * 32 + lane;
        const bool valid = row < p.rows;
        // pass 1: row maximum (and arg-max, lowest index on ties)
        float mx = -INFINITY;
        int mi = 0;
#pragma unroll 1
        for (int c = 0; c < FV_K; c += 32) {
            float v[32];
            tmem_ld32(tmem + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float x = v[j] + cst[c + j];
                if (x > mx) { mx = x; mi = c + j; }
            }
        }
        const float base = (mx > -INFINITY && mx < INFINITY) ? mx : 0.f;
        // pass 2: e = exp(l - max), stashed back into the accumulator columns
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < FV_K; c += 32) {
            float v[32];
            tmem_ld32(tmem + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                v[j] = __expf(v[j] + cst[c + j] - base);
                sum += v[j];
            }
            tmem_st32(tmem + c, v);
        }
        tmem_st_wait();
        const float inv = 1.f / sum;
        // pass 3: q = e / sum, split and stored at the image-padded row
        int64_t qrow = 0;
        if (valid) {
            int64_t img = p.tile_img[t.mb];
            while (img + 1 < p.n_images && p.offsets[img + 1] <= row) ++img;
            qrow = p.qoff[img] + (row - p.offsets[img]);
            if (p.argmax) p.argmax[row] = mi;
        }
        float4* oh = reinterpret_cast<float4*>(p.q_hi + qrow * FV_K);
        float4* ol = reinterpret_cast<float4*>(p.q_lo + qrow * FV_K);
#pragma unroll 1
        for (int c = 0; c < FV_K; c += 32) {
            float v[32];
            tmem_ld32(tmem + c, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float h[4], l[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) tf32_split(v[j + e] * inv, h[e], l[e]);
                    oh[(c + j) >> 2] = make_float4(h[0], h[1], h[2], h[3]);
                    ol[(c + j) >> 2] = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// stats: S_i = Yaug_i^T Q_i / T_i  and  s0 = column means of Q_i
// ---------------------------------------------------------------------------------------
struct StatsParams {
    CUtensorMap y_hi, y_lo, q_hi, q_lo;       // MN-major (SWIZZLE_128B_ATOM_32B) maps, 32 x 32 boxes
    const int64_t* offsets;
    const int64_t* qoff;
    float* S;                                 // [n_images, 256, 129]: [k][ s1 (64) | s2 (64) | s0 ]
    int64_t n_images;
};
struct StatsState { float s0a, s0b; };

struct StatsPolicy {
    using Params = StatsParams;
    using EpiState = StatsState;
    struct Tile { int nkb; int t; int64_t img, r0, q0; };
    static constexpr bool BF16 = false, A_MN = true, B_MN = true, EPI_READS_STAGES = true;
    static constexpr int KT = Q_PAD;
    static constexpr int PASSES = 3, BLOCK_N = FV_K, STAGES = 2, KSTEPS = KT / 8;
    static constexpr int A_LBO = KT * 128, B_LBO = KT * 128;
    static constexpr int A_BYTES = (FV_2D / 32) * A_LBO, B_BYTES = (FV_K / 32) * B_LBO, SCRATCH_BYTES = 0;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.y_hi); tma_prefetch_desc(&p.q_hi); }
    __device__ static int num_tiles(const Params& p) { return (int)p.n_images; }
    __device__ static Tile tile(const Params& p, int i)
    {
        const int64_t r0 = p.offsets[i];
        const int t = (int)(p.offsets[i + 1] - r0);
        return {(t + KT - 1) / KT, t, (int64_t)i, r0, p.qoff[i]};
    }
    __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* a_hi, uint8_t* a_lo, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        const int ry = (int)(t.r0 + (int64_t)kb * KT), rq = (int)(t.q0 + (int64_t)kb * KT);
#pragma unroll
        for (int b = 0; b < FV_2D / 32; ++b) {
            tma_load_2d(a_hi + b * A_LBO, &p.y_hi, bar, 32 * b, ry);
            tma_load_2d(a_lo + b * A_LBO, &p.y_lo, bar, 32 * b, ry);
        }
#pragma unroll
        for (int b = 0; b < FV_K / 32; ++b) {
            tma_load_2d(b_hi + b * B_LBO, &p.q_hi, bar, 32 * b, rq);
            tma_load_2d(b_lo + b * B_LBO, &p.q_lo, bar, 32 * b, rq);
        }
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile&, EpiState& st, int, int) { st.s0a = st.s0b = 0.f; }
    // column sums of the Q tile in shared memory: epilogue thread j owns components 2j, 2j+1.
    // Tile layout (SWIZZLE_128B_ATOM_32B): 32-column block b at b * 4096; row r at r * 128;
    // the 32-B chunk index is XOR-ed with (r & 3).
    __device__ static void consume(const Params&, const Tile&, int, uint8_t*, uint8_t*, uint8_t* b_hi, uint8_t* b_lo,
                                   EpiState& st, int quarter, int lane)
    {
        const int j = (quarter * 32 + lane) * 2;          // component index (any bijection onto 0..254 works)
        const int blk = j >> 5, c = j & 31;
        const uint32_t col_off = (uint32_t)(blk * B_LBO + (c & 7) * 4);
        const int chunk = c >> 3;
        float a = 0.f, b = 0.f;
#pragma unroll 8
        for (int r = 0; r < KT; ++r) {
            const uint32_t off = col_off + (uint32_t)(r * 128 + ((chunk ^ (r & 3)) << 5));
            const float2 h = *reinterpret_cast<const float2*>(b_hi + off);
            const float2 l = *reinterpret_cast<const float2*>(b_lo + off);
            a += h.x + l.x;
            b += h.y + l.y;
        }
        st.s0a += a;
        st.s0b += b;
    }
    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane, uint8_t*,
                                    EpiState& st)
    {
        constexpr int LD = FV_2D + 1;
        const int e = quarter * 32 + lane;                 // row of Yaug^T: [0,64) = y*y, [64,128) = y
        const int n = e < FV_D ? FV_D + e : e - FV_D;      // column in the [ s1 | s2 | s0 ] layout
        float* Simg = p.S + t.img * (int64_t)FV_K * LD;
        const float inv_t = 1.f / (float)t.t;              // T == 0 -> NaN row, like the reference
        const bool empty = t.nkb == 0;
#pragma unroll 1
        for (int c = 0; c < FV_K; c += 32) {
            float v[32];
            tmem_ld32(tmem + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
                Simg[(int64_t)(c + jj) * LD + n] = empty ? __int_as_float(0x7fc00000) : v[jj] * inv_t;
        }
        const int j = e * 2;
        Simg[(int64_t)j * LD + FV_2D] = empty ? __int_as_float(0x7fc00000) : st.s0a * inv_t;
        Simg[(int64_t)(j + 1) * LD + FV_2D] = empty ? __int_as_float(0x7fc00000) : st.s0b * inv_t;
    }
};

}  // namespace tc

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
using namespace tc;

static int run_split(const float* x, int64_t n, float* hi, float* lo, cudaStream_t st)
{
    if (n <= 0) return PVS_OK;
    PVS_CHECK(n % 4 == 0, PVS_ERR_BAD_SHAPE, "split needs a multiple of 4 elements");
    PVS_LAUNCH(split_kernel, 148 * 8, 256, 0, st, (const float4*)x, n / 4, (float4*)hi, (float4*)lo);
    return PVS_OK;
}

// tf32 hi/lo copies of the weights the tensor-core kernels read (called from *_create)
int tc_prepare_model(pvs_model* m)
{
    if (!tc_available()) return PVS_OK;
    const float* src = nullptr;
    size_t n = 0;
    if (m->kind == PVS_MODEL_GMM_DIAG && m->k == FV_K && m->d == FV_D) { src = m->wcat; n = (size_t)FV_K * FV_2D; }
    else if (m->kind == PVS_MODEL_PCA && m->d == FV_D && m->d_in % 32 == 0) { src = m->comp; n = (size_t)m->d * m->d_in; }
    if (!src) return PVS_OK;
    float* buf = nullptr;
    PVS_CUDA(cudaMalloc((void**)&buf, 2 * n * sizeof(float)));
    int rc = run_split(src, (int64_t)n, buf, buf + n, nullptr);
    if (rc == PVS_OK && cudaStreamSynchronize(nullptr) != cudaSuccess) rc = fail(PVS_ERR_CUDA, "weight split failed");
    if (rc != PVS_OK) { cudaFree(buf); return rc; }
    m->tc0 = buf;
    m->tc1 = buf + n;
    return PVS_OK;
}

bool tc_fv_supported(const pvs_model* g, const pvs_model* pca, int64_t rows, int64_t n_images)
{
    if (!tc_available() || !g->tc0 || g->k != FV_K || g->d != FV_D) return false;
    if (pca && !pca->tc0) return false;
    // TMA coordinates are 32-bit
    return rows + (int64_t)(Q_PAD - 1) * n_images + 256 < 2147483647LL && n_images < 2147483647LL;
}

int tc_fv_plan(const pvs_model* g, const pvs_model* pca, int64_t rows, int64_t n_images, void* ws, TcFvPlan* pl)
{
    size_t off = 0;
    char* base = (char*)ws;
    auto take = [&](size_t bytes) { void* p = base ? base + off : nullptr; off += align_up(bytes ? bytes : 16, 1024); return p; };
    const int d_in = pca ? pca->d_in : g->d;
    pl->q_rows_cap = rows + (int64_t)(Q_PAD - 1) * n_images;
    pl->n_tiles = (int)ceil_div(rows, 128);
    pl->x_hi = pca ? (float*)take((size_t)rows * d_in * 4) : nullptr;
    pl->x_lo = pca ? (float*)take((size_t)rows * d_in * 4) : nullptr;
    pl->y_hi = (float*)take((size_t)rows * FV_2D * 4);
    pl->y_lo = (float*)take((size_t)rows * FV_2D * 4);
    pl->q_hi = (float*)take((size_t)pl->q_rows_cap * FV_K * 4);
    pl->q_lo = (float*)take((size_t)pl->q_rows_cap * FV_K * 4);
    pl->S = (float*)take((size_t)n_images * FV_K * (FV_2D + 1) * 4);
    pl->qoff = (int64_t*)take((size_t)(n_images + 1) * 8);
    pl->tile_img = (int*)take((size_t)(pl->n_tiles + 1) * 4);
    pl->total = off + 1024;
    return PVS_OK;
}

int tc_fv_prep(const TcFvPlan& pl, const int64_t* offsets, int64_t n_images, int64_t rows, cudaStream_t st)
{
    PVS_LAUNCH(qoff_kernel, 1, 1024, 0, st, offsets, n_images, pl.qoff);
    if (pl.n_tiles > 0)
        PVS_LAUNCH(tile_image_kernel, (unsigned)ceil_div(pl.n_tiles, 256), 256, 0, st, offsets, n_images, pl.n_tiles, pl.tile_img);
    PVS_LAUNCH(q_pad_zero_kernel, (unsigned)ceil_div(n_images, 8), 256, 0, st, offsets, pl.qoff, n_images, pl.q_hi, pl.q_lo);
    return PVS_OK;
}

int tc_fv_project(const TcFvPlan& pl, const pvs_model* g, const pvs_model* pca, const float* desc, int64_t rows,
                  cudaStream_t st)
{
    if (rows <= 0) return PVS_OK;
    if (!pca) {
        PVS_LAUNCH(yaug_kernel, 148 * 8, 256, 0, st, desc, rows, pl.y_hi, pl.y_lo);
        return PVS_OK;
    }
    if (int rc = run_split(desc, rows * pca->d_in, pl.x_hi, pl.x_lo, st)) return rc;
    PcaParams p{};
    int rc;
    if ((rc = make_tmap_2d(&p.x_hi, pl.x_hi, false, rows, pca->d_in, pca->d_in, 32, 128))) return rc;
    if ((rc = make_tmap_2d(&p.x_lo, pl.x_lo, false, rows, pca->d_in, pca->d_in, 32, 128))) return rc;
    if ((rc = make_tmap_2d(&p.c_hi, pca->tc0, false, pca->d, pca->d_in, pca->d_in, 32, FV_D))) return rc;
    if ((rc = make_tmap_2d(&p.c_lo, pca->tc1, false, pca->d, pca->d_in, pca->d_in, 32, FV_D))) return rc;
    p.bias = pca->bias; p.y_hi = pl.y_hi; p.y_lo = pl.y_lo; p.rows = rows;
    p.m_blocks = pl.n_tiles; p.nkb = pca->d_in / 32;
    return launch_tc<PcaPolicy>(p, p.m_blocks, st);
}

int tc_fv_posterior(const TcFvPlan& pl, const pvs_model* g, const int64_t* offsets, int64_t n_images, int64_t rows,
                    int32_t* argmax, cudaStream_t st)
{
    if (rows <= 0) return PVS_OK;
    PostParams p{};
    int rc;
    if ((rc = make_tmap_2d(&p.y_hi, pl.y_hi, false, rows, FV_2D, FV_2D, 32, 128))) return rc;
    if ((rc = make_tmap_2d(&p.y_lo, pl.y_lo, false, rows, FV_2D, FV_2D, 32, 128))) return rc;
    if ((rc = make_tmap_2d(&p.w_hi, g->tc0, false, FV_K, FV_2D, FV_2D, 32, FV_K))) return rc;
    if ((rc = make_tmap_2d(&p.w_lo, g->tc1, false, FV_K, FV_2D, FV_2D, 32, FV_K))) return rc;
    p.cst = g->cst; p.offsets = offsets; p.qoff = pl.qoff; p.tile_img = pl.tile_img;
    p.q_hi = pl.q_hi; p.q_lo = pl.q_lo; p.argmax = argmax; p.rows = rows; p.n_images = n_images;
    p.m_blocks = pl.n_tiles;
    return launch_tc<PostPolicy>(p, p.m_blocks, st);
}

int tc_fv_stats(const TcFvPlan& pl, const int64_t* offsets, int64_t n_images, int64_t rows, cudaStream_t st)
{
    if (n_images <= 0) return PVS_OK;
    StatsParams p{};
    int rc;
    if ((rc = make_tmap_2d(&p.y_hi, pl.y_hi, false, rows, FV_2D, FV_2D, 32, Q_PAD, true))) return rc;
    if ((rc = make_tmap_2d(&p.y_lo, pl.y_lo, false, rows, FV_2D, FV_2D, 32, Q_PAD, true))) return rc;
    if ((rc = make_tmap_2d(&p.q_hi, pl.q_hi, false, pl.q_rows_cap, FV_K, FV_K, 32, Q_PAD, true))) return rc;
    if ((rc = make_tmap_2d(&p.q_lo, pl.q_lo, false, pl.q_rows_cap, FV_K, FV_K, 32, Q_PAD, true))) return rc;
    p.offsets = offsets; p.qoff = pl.qoff; p.S = pl.S; p.n_images = n_images;
    return launch_tc<StatsPolicy>(p, (int)n_images, st);
}

}  // namespace pvs
