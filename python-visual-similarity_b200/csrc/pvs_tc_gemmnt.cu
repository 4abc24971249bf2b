// pvs_tc_gemmnt.cu -- generic fp32-accurate contraction on tcgen05 (3xTF32, CTA pairs):
//     C[m, n] = alpha * A'[m, :] . B[n, :] + bias[n],   A' = A  or  (a0*a0, a0, a1*a1, a1, ...)
// for any inner dimension and any m, n.  It carries the stages whose shape the specialised
// kernels (pvs_tc_fv.cu: K = 256, D = 64) do not cover: PCA projection of VGG16 descriptors
// (514 -> 257), GMM logits for D = 128 / 257 (fisher_vector.py:92,99), and the same routines
// when called on their own (pvs_pca_project, pvs_gmm_posterior).
//   A : fp32 rows in global memory, loaded by the producer warps (scalar loads: the concat
//       boundary and odd row pitches rule out vectors), split into tf32 hi/lo in registers
//   B : weights, hi/lo split and zero-padded to 32 columns once at model creation, streamed by
//       TMA (rows beyond n are zero-filled by the tensor map)
//   D : [256 x 256] pair tiles in TMEM; the epilogue applies alpha / bias and stores rows.
#include <string.h>
#include "pvs_tc2.cuh"
#include "pvs_kernels.cuh"

namespace pvs {
namespace tc2 {

struct GemmNtParams {
    CUtensorMap b_hi, b_lo;            // B [n, k_pad] fp32 hi / lo, box 32 cols x 128 rows
    const float* a;                    // A [m, a_cols], row pitch lda
    const float* bias;                 // [n] or NULL
    float* c;                          // C [m, n], row pitch ldc
    int64_t lda, ldc, m;
    float alpha;
    int n, a_cols, nkb, m_blocks, n_blocks;
    int segk, nseg;                    // k-blocks per accumulation segment, segments per output tile
};
struct GemmNtState {};
struct GRegs { float4 v[8]; };

template <bool SQUARE>
struct GemmNtPolicy {
    using Params = GemmNtParams;
    using EpiState = GemmNtState;
    using Regs = GRegs;
    // A tile of the skeleton is one K-SEGMENT of an output tile: the tensor core accumulates `segk` k-blocks from zero, the
    // epilogue adds the segment to C (the first one writes alpha * acc + bias, the others C += alpha * acc; the tile was
    // written by the same threads a moment ago and is still in L2).  tcgen05.mma truncates when it adds a K-slice to the
    // accumulator; over the 514-long contraction of the VGG16-PCA logits (195 accumulation steps) that put 9 of 4 096 images
    // 1e-4 .. 1.8e-4 off the fp64 result.  The segments of an output tile run back to back on one CTA pair.
    struct Tile { int nkb, mb, nb, kb0; bool first; };
    static constexpr bool BF16 = false, MANUAL = true, B_RESIDENT = false, ACC_INIT = false, TILE_SYNC = false;
    static constexpr int PASSES = 3, BLOCK_N = 256, KSTEPS = 4, NKB_RES = 0, STAGES = 3, PGROUPS = 3;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = 128 * 128, SCRATCH_BYTES = 0, TMA_BYTES = 2 * B_BYTES;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.b_hi); tma_prefetch_desc(&p.b_lo); }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks * p.n_blocks * p.nseg; }
    __device__ static int tile_at(const Params& p, int it, int pair, int n_pairs, int)
    {
        const long long u = (long long)pair + (long long)(it / p.nseg) * n_pairs;
        return u < (long long)p.m_blocks * p.n_blocks ? (int)(u * p.nseg + it % p.nseg) : -1;
    }
    __device__ static Tile tile(const Params& p, int i)
    {
        const int u = i / p.nseg, seg = i - u * p.nseg, kb0 = seg * p.segk;
        const int left = p.nkb - kb0;
        return {left < p.segk ? left : p.segk, u / p.n_blocks, u % p.n_blocks, kb0, seg == 0};
    }
    __device__ static void load(const Params& p, const Tile& t, int kb, int rank, uint8_t*, uint8_t*, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        tma_load_2d_pair(b_hi, &p.b_hi, bar, (t.kb0 + kb) * 32, t.nb * BLOCK_N + rank * 128);
        tma_load_2d_pair(b_lo, &p.b_lo, bar, (t.kb0 + kb) * 32, t.nb * BLOCK_N + rank * 128);
    }
    // element `col` of the (virtual) operand row A'
    __device__ static float a_elem(const Params& p, const float* row, int col)
    {
        if constexpr (SQUARE) {
            // interleaved (y_d^2, y_d) pairs: the quadratic and the linear term of one dimension are
            // accumulated next to each other, so the running sum stays at the scale of the final
            // logit instead of first growing to -sum(P y^2)/2 -- the tensor core truncates when it
            // accumulates, and that error is proportional to the running sum
            const int dd = col >> 1;
            if (dd >= p.a_cols) return 0.f;
            const float v = __ldg(row + dd);
            return (col & 1) ? v : v * v;
        } else {
            return col < p.a_cols ? __ldg(row + col) : 0.f;
        }
    }
    __device__ static void fetch(const Params& p, const Tile& t, int kb, int rank, int pw, int lane, Regs& g)
    {
        const int col = (t.kb0 + kb) * 32 + (lane & 7) * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t gr = (int64_t)t.mb * 256 + rank * 128 + pw * 32 + i * 4 + (lane >> 3);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < p.m) {
                const float* row = p.a + gr * p.lda;
                v.x = a_elem(p, row, col);
                v.y = a_elem(p, row, col + 1);
                v.z = a_elem(p, row, col + 2);
                v.w = a_elem(p, row, col + 3);
            }
            g.v[i] = v;
        }
    }
    __device__ static void store(const Params&, const Tile&, int, const Regs& g, uint8_t* hi, uint8_t* lo, int pw, int lane)
    {
        const int c = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = pw * 32 + i * 4 + (lane >> 3);
            const float4 v = g.v[i];
            float4 h, l;
            tf32_split(v.x, h.x, l.x);
            tf32_split(v.y, h.y, l.y);
            tf32_split(v.z, h.z, l.z);
            tf32_split(v.w, h.w, l.w);
            const uint32_t off = (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
            *reinterpret_cast<float4*>(hi + off) = h;
            *reinterpret_cast<float4*>(lo + off) = l;
        }
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, int rank, uint32_t tmem, int quarter, int lane, uint8_t*,
                                    EpiState&)
    {
        const int64_t row = (int64_t)t.mb * 256 + rank * 128 + quarter * 32 + lane;
        float* crow = p.c + row * p.ldc;
        const bool vec = (p.ldc & 3) == 0 && ((uintptr_t)p.c & 15) == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + c0, v);
            tmem_ld_wait();
            const int col0 = t.nb * BLOCK_N + c0;
            if (row >= p.m || col0 >= p.n) continue;
            if (t.first) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = col0 + j;
                    v[j] = p.alpha * v[j] + ((p.bias && col < p.n) ? p.bias[col] : 0.f);
                }
            } else if (vec && col0 + 32 <= p.n) {
                // later K-segments: fire-and-forget RED.ADD at L2 (round to nearest; this thread owns the row and its
                // operations on one address keep program order, so the sum is the same as with load + add + store)
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    atomicAdd(reinterpret_cast<float4*>(crow + col0 + j),
                              make_float4(p.alpha * v[j], p.alpha * v[j + 1], p.alpha * v[j + 2], p.alpha * v[j + 3]));
                continue;
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.n) atomicAdd(crow + col0 + j, p.alpha * v[j]);
                continue;
            }
            if (vec && col0 + 32 <= p.n) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + col0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.n) crow[col0 + j] = v[j];
            }
        }
    }
};

}  // namespace tc2

using namespace tc2;

// zero-padded [rows, ld = pad32(cols)] tf32 hi / lo copies of a device matrix (one allocation)
// interleave_halves: source row [u (cols/2) | v (cols/2)] is stored as (u0, v0, u1, v1, ...)
int tc_make_padded_hi_lo(const float* src_dev, int rows, int cols, bool interleave_halves, float** buf_out, int* ld_out)
{
    const int ld = (cols + 31) / 32 * 32;
    const size_t n = (size_t)rows * ld;
    std::vector<float> h((size_t)rows * cols), hi(n, 0.f), lo(n, 0.f);
    PVS_CUDA(cudaMemcpy(h.data(), src_dev, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int j = 0; j < rows; ++j)
        for (int i = 0; i < cols; ++i) {
            // round-to-nearest-away at 13 dropped mantissa bits == cvt.rna.tf32.f32 for finite values
            const int half = cols / 2;
            const int si = interleave_halves ? ((i & 1) ? half + (i >> 1) : (i >> 1)) : i;
            const float x = h[(size_t)j * cols + si];
            uint32_t u;
            memcpy(&u, &x, 4);
            const uint32_t uh = (u + 0x1000u) & 0xFFFFE000u;
            float a;
            memcpy(&a, &uh, 4);
            const float r = x - a;
            memcpy(&u, &r, 4);
            const uint32_t ul = (u + 0x1000u) & 0xFFFFE000u;
            float b;
            memcpy(&b, &ul, 4);
            hi[(size_t)j * ld + i] = a;
            lo[(size_t)j * ld + i] = b;
        }
    float* buf = nullptr;
    PVS_CUDA(cudaMalloc((void**)&buf, 2 * n * sizeof(float)));
    cudaError_t e = cudaMemcpy(buf, hi.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(buf + n, lo.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(buf); return fail(PVS_ERR_CUDA, "weight upload failed: %s", cudaGetErrorString(e)); }
    *buf_out = buf;
    *ld_out = ld;
    return PVS_OK;
}

// generic operand copies for PCA (components [d, d_in]) and GMM (wcat [k, 2d]) models
int tc_prepare_generic(pvs_model* m)
{
    if (!tc_available()) return PVS_OK;
    float* buf = nullptr;
    int ld = 0, rows = 0;
    if (m->kind == PVS_MODEL_PCA) { rows = m->d; if (int rc = tc_make_padded_hi_lo(m->comp, m->d, m->d_in, false, &buf, &ld)) return rc; }
    else if (m->kind == PVS_MODEL_GMM_DIAG) { rows = m->k; if (int rc = tc_make_padded_hi_lo(m->wcat, m->k, 2 * m->d, true, &buf, &ld)) return rc; }
    else return PVS_OK;
    m->tcg0 = buf;
    m->tcg1 = buf + (size_t)rows * ld;
    m->tcg_ld = ld;
    return PVS_OK;
}

bool tc_gemm_nt_supported(int64_t m, int n) { return tc_available() && m > 0 && n > 0 && m < 2147483000LL * 128; }

// C = alpha * A' B^T + bias with B given as padded hi / lo copies (b_rows x b_ld)
int tc_gemm_nt(const float* a, int64_t lda, int a_cols, bool square_cat, const float* b_hi, const float* b_lo, int b_ld,
               int n, float* c, int64_t ldc, int64_t m, float alpha, const float* bias, cudaStream_t st)
{
    if (m <= 0 || n <= 0) return PVS_OK;
    GemmNtParams p{};
    int rc;
    if ((rc = make_tmap_2d(&p.b_hi, b_hi, false, n, b_ld, b_ld, 32, 128))) return rc;
    if ((rc = make_tmap_2d(&p.b_lo, b_lo, false, n, b_ld, b_ld, 32, 128))) return rc;
    p.a = a; p.lda = lda; p.a_cols = a_cols; p.bias = bias; p.c = c; p.ldc = ldc; p.m = m; p.n = n; p.alpha = alpha;
    p.nkb = b_ld / 32;
    p.segk = 4;                                                // 128 K-elements = 48 accumulation steps per segment
    if (const char* e = getenv("PVS_GEMM_SEGK")) { const int v = atoi(e); if (v >= 1) p.segk = v; }
    p.nseg = (p.nkb + p.segk - 1) / p.segk;
    p.m_blocks = (int)ceil_div(m, 256);
    p.n_blocks = (int)ceil_div(n, 256);
    const int tiles = p.m_blocks * p.n_blocks;
    return square_cat ? launch_tc2<GemmNtPolicy<true>>(p, tiles, st) : launch_tc2<GemmNtPolicy<false>>(p, tiles, st);
}

}  // namespace pvs
