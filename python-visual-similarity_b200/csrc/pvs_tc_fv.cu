// pvs_tc_fv.cu -- Fisher-vector contractions on tcgen05 tensor cores (3xTF32).
//
// Shape handled here: K = 256 components, D = 64 dims after the optional PCA (the
// SIFT/RootSIFT-PCA-64 configurations, i.e. BASELINE.json configs[1]); anything else takes
// the CUDA-core kernels in pvs_simt.cu.  Three contractions per batch, all instantiations
// of the skeleton in pvs_tc.cuh.  The intermediates that travel through HBM are plain fp32
// (Y: 64 floats/row, Q: 256 floats/row); the tf32 hi/lo operand pairs only ever exist in
// shared memory: four producer warps load fp32 rows with coalesced 16-byte loads, split
// them (x = hi + lo, both tf32-exact), and store the swizzled UMMA tiles themselves.
// The weight operands were split once at model creation and arrive by TMA.
//
//   project   Y[rows,64]    = X[rows,d_in] . C^T + b        K-major x K-major, N = 64
//                                                              (fisher_vector.py:92)
//   posterior L[rows,256]   = [y*y | y] . [ -P/2 | mu P ]^T + c     K-major x K-major, N = 256
//             epilogue      -> softmax over the 256 columns held by each TMEM lane
//                                                              (fisher_vector.py:99)
//   stats     S[128,256]    = [y*y | y]_i^T . Q_i  per image   MN-major x MN-major, K = T_i
//             epilogue      -> S/T in the [k][2d+1] layout fv_finalize reads, s0 from the
//                              Q tiles while they sit in shared memory
//                                                              (fisher_vector.py:102-104)
#include "pvs_tc.cuh"
#include "pvs_tc2.cuh"
#include "pvs_kernels.cuh"
#include <cuda_fp16.h>

namespace pvs {
namespace tc {

constexpr int FV_K = 256;     // mixture components
constexpr int FV_D = 64;      // dims after PCA
constexpr int FV_2D = 128;
constexpr int ST_KT = 16;     // contraction rows (descriptors) per stats stage

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void split4(const float4& v, float4& h, float4& l)
{
    tf32_split(v.x, h.x, l.x);
    tf32_split(v.y, h.y, l.y);
    tf32_split(v.z, h.z, l.z);
    tf32_split(v.w, h.w, l.w);
}
// K-major SWIZZLE_128B tile: row r (128 B), 16-B chunk c
__device__ __forceinline__ uint32_t sw128_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
// MN-major SWIZZLE_128B_BASE32B block [rows x 128 B]: row r, 16-B chunk c (32-B chunk c>>1 is swizzled)
__device__ __forceinline__ uint32_t sw32_off(int r, int c) { return (uint32_t)(r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4))); }

// One warp owns rows [32 pw, 32 pw + 32) of a [128 rows x 32 floats] K-major operand tile.
// fetch: 8 coalesced 16-byte loads per lane from src[row * ld + col0 ...] (rows >= rows_total
// read as zero); store: optional squaring, tf32 hi/lo split, swizzled 16-byte stores.
struct KRegs { float4 v[8]; };
__device__ __forceinline__ void fetch_kmajor_32rows(const float* __restrict__ src, int64_t ld, int64_t row0, int64_t rows_total,
                                                    int col0, int pw, int lane, KRegs& r)
{
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gr = row0 + pw * 32 + i * 4 + (lane >> 3);
        r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < rows_total) r.v[i] = ldg4(src + gr * ld + col0 + c * 4);
    }
}
template <bool SQUARE>
__device__ __forceinline__ void store_kmajor_32rows(const KRegs& r, uint8_t* hi, uint8_t* lo, int pw, int lane)
{
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = pw * 32 + i * 4 + (lane >> 3);
        float4 v = r.v[i];
        if (SQUARE) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
        float4 h, l;
        split4(v, h, l);
        const uint32_t off = sw128_off(row, c);
        *reinterpret_cast<float4*>(hi + off) = h;
        *reinterpret_cast<float4*>(lo + off) = l;
    }
}

// ---------------------------------------------------------------------------------------
// project: Y = X C^T + b
// ---------------------------------------------------------------------------------------
struct PcaParams {
    CUtensorMap c_hi, c_lo, y_map;            // y_map: Y [rows, 64] fp32, box 32 cols x 32 rows (TMA store, pair kernel)
    const float* x;
    const float* bias;
    float* y;
    int64_t rows;
    int m_blocks, nkb, d_in;
    int* flag;                                // raised when some |y| > lim (fp16x2 operand range), may be NULL
    float lim;
};
struct NoEpiState {};

struct PcaPolicy {
    using Params = PcaParams;
    using EpiState = NoEpiState;
    struct Tile { int nkb, mb; };
    static constexpr bool BF16 = false, A_MN = false, B_MN = false, EPI_READS_STAGES = false, MANUAL = true;
    static constexpr int PASSES = 3, BLOCK_N = FV_D, KSTEPS = 4, STAGES = 4, PGROUPS = 2;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = BLOCK_N * 128, A_LBO = 0, B_LBO = 0, SCRATCH_BYTES = 256;
    static constexpr int TMA_BYTES = 2 * B_BYTES;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.c_hi); tma_prefetch_desc(&p.c_lo); }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks; }
    __device__ static int tile_at(const Params&, int it, int n) { return strided_tile(it, n); }
    __device__ static Tile tile(const Params& p, int i) { return {p.nkb, i}; }
    __device__ static void load(const Params& p, const Tile&, int kb, uint8_t*, uint8_t*, uint8_t* b_hi, uint8_t* b_lo,
                                uint64_t* bar)
    {
        tma_load_2d(b_hi, &p.c_hi, bar, kb * 32, 0);
        tma_load_2d(b_lo, &p.c_lo, bar, kb * 32, 0);
    }
    using Regs = KRegs;
    __device__ static void fetch(const Params& p, const Tile& t, int kb, int pw, int lane, Regs& r)
    {
        fetch_kmajor_32rows(p.x, p.d_in, (int64_t)t.mb * 128, p.rows, kb * 32, pw, lane, r);
    }
    struct PState {};
    __device__ static void store(const Params&, const Tile&, int, const Regs& r, uint8_t* a_hi, uint8_t* a_lo, uint8_t*,
                                 uint8_t*, int pw, int, int lane, PState&)
    {
        store_kmajor_32rows<false>(r, a_hi, a_lo, pw, lane);
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
        float4* o = reinterpret_cast<float4*>(p.y + row * FV_D);
        float amax = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            __syncwarp();
            tmem_ld32(tmem + half * 32, v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; ++j) { v[j] += bias[half * 32 + j]; amax = fmaxf(amax, fabsf(v[j])); }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    o[(half * 32 + j) >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        if (p.flag && amax > p.lim) *p.flag = 1;               // fp16x2 range guard (inf raises it, NaN just propagates)
    }
};

// ---------------------------------------------------------------------------------------
// posterior: logits + softmax -> Q
// ---------------------------------------------------------------------------------------
struct PostParams {
    CUtensorMap w_hi, w_lo, q_map;            // q_map: Q [rows, 256] fp32, box 32 cols x 32 rows (TMA store)
    CUtensorMap qh_map, ql_map;               // fp16x2 variant, planes == 1: Q * 2^14 as fp16 hi / lo planes [rows, 256],
                                              // box 64 cols x 32 rows -- the operand format of the statistics kernel
    const float* y;
    const float* cst;
    float* q;
    int32_t* argmax;
    int64_t rows;
    int m_blocks;
    const int* flag;                          // fp16x2 range flag (NULL: no gating): the fp16x2 variant runs when it is 0,
                                              // the 3xTF32 variant when it is not
    float sc_y;                               // 2^-e (fp16x2 variant)
    int planes;                               // fp16x2 variant: write the fp16 planes instead of fp32 Q
    float* rinv;                              // planes: [rows, 4] per-row normaliser 1 / sum_k exp(l - max) (slot 0 of 16 bytes)
};

// ---------------------------------------------------------------------------------------
// stats: S_i = [y*y | y]_i^T Q_i / T_i  and  s0 = column means of Q_i
// ---------------------------------------------------------------------------------------
struct StatsParams {
    const float* y;                           // [rows, 64]
    const float* q;                           // [rows, 256]
    const int64_t* offsets;
    float* S;                                 // [n_images, 256, 128]: [k][ s1 (64) | s2 (64) ], already / T
    float* s0part;                            // [n_images, 16, 256]: raw column sums of Q per producer warp
    int64_t n_images;
    const int* smax;                          // segments per image slot: max over the images of ceil(k-blocks / segk), >= 1 (device)
    int segk;                                 // k-blocks (ST_KT descriptors each) per statistics segment
};

// A tile of the statistics kernels is one SEGMENT of an image: `segk` k-blocks accumulated from zero in one of the two TMEM
// accumulators and folded into the image's S rows (global memory, L2-resident, fp32 round-to-nearest adds) while the next
// segment is being multiplied into the other accumulator.  tcgen05.mma adds every K-slice to the accumulator with truncation;
// accumulated over a whole 2 000-descriptor image that bias, amplified by the cancellation in d_sigma, put 46 (fp16x2) to
// 98 (3xTF32) of the 8 189 images of the C2 batch 1e-4 .. 3.2e-4 off the fp64 result.  Every image gets *smax consecutive
// tiles on ONE CTA (the folds of an image are ordered: same threads, same addresses); tiles past the image's last segment
// are empty and skipped.
struct SegTile { int nkb; int t; int64_t img, r0; int kb0; bool first, last, skip; };
__device__ __forceinline__ int seg_tile_at(const StatsParams& p, int it)
{
    const int smax = *p.smax;
    const long long img = (long long)blockIdx.x + (long long)(it / smax) * gridDim.x;
    return img < p.n_images ? (int)(img * smax + it % smax) : -1;
}
__device__ __forceinline__ SegTile seg_tile(const StatsParams& p, int i, int kt)
{
    const int smax = *p.smax;
    const int64_t img = i / smax;
    const int seg = i - (int)img * smax;
    const int64_t r0 = p.offsets[img];
    const int t = (int)(p.offsets[img + 1] - r0);
    const int total = (t + kt - 1) / kt, kb0 = seg * p.segk;
    const int left = total - kb0;
    SegTile tl;
    tl.nkb = left <= 0 ? 0 : (left < p.segk ? left : p.segk);
    tl.t = t; tl.img = img; tl.r0 = r0; tl.kb0 = kb0;
    tl.first = seg == 0;
    tl.last = left <= p.segk;                              // also true for an empty image (its only tile writes NaN)
    tl.skip = seg > 0 && left <= 0;
    return tl;
}
// fold of a finished segment by one warp (its TMEM lane quarter).  S keeps RAW sums (fv_finalize applies 1 / T, exactly the
// multiply the last segment used to apply here); the first segment is stored (NaN for T == 0, like the reference), the later
// ones are added with fire-and-forget RED.ADDs at L2 (round to nearest; one owner thread per address, program order per
// address: a fixed summation order)
__device__ __forceinline__ void seg_fold(const SegTile& t, float* Srow, uint32_t tmem, int64_t ld = FV_2D, bool live = true)
{
    const bool empty = t.t == 0;
    const float nanv = __int_as_float(0x7fc00000);
#pragma unroll 1
    for (int c = 0; c < FV_K; c += 32) {
        float v[32];
        tmem_ld32(tmem + c, v);
        tmem_ld_wait();
        if (!live) continue;
        if (t.first) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) __stcg(Srow + (int64_t)(c + jj) * ld, empty ? nanv : v[jj]);
        } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) atomicAdd(Srow + (int64_t)(c + jj) * ld, v[jj]);
        }
    }
}

struct StatsPolicy {
    using Params = StatsParams;
    using EpiState = NoEpiState;
    using Tile = SegTile;
    static constexpr bool BF16 = false, A_MN = true, B_MN = true, EPI_READS_STAGES = false, MANUAL = true;
    static constexpr int KT = ST_KT;
    static constexpr int PASSES = 3, BLOCK_N = FV_K, STAGES = 4, KSTEPS = KT / 8, PGROUPS = 4;
    static constexpr int A_LBO = KT * 128, B_LBO = KT * 128;          // one [KT rows x 128 B] block per 32 columns
    static constexpr int A_BYTES = (FV_2D / 32) * A_LBO, B_BYTES = (FV_K / 32) * B_LBO, SCRATCH_BYTES = 0;
    static constexpr int TMA_BYTES = 0;
    __device__ static void prefetch(const Params&) {}
    __device__ static int num_tiles(const Params& p) { return (int)p.n_images * *p.smax; }
    __device__ static int tile_at(const Params& p, int it, int) { return seg_tile_at(p, it); }
    __device__ static Tile tile(const Params& p, int i) { return seg_tile(p, i, KT); }
    __device__ static void load(const Params&, const Tile&, int, uint8_t*, uint8_t*, uint8_t*, uint8_t*, uint64_t*) {}
    // producer warp pw owns descriptor rows [RPW pw, RPW pw + RPW) of the stage; rows past the
    // end of the image are written as zeros (that is what makes ragged T exact)
    static constexpr int RPW = KT / 4;
    struct Regs { float4 q[RPW * 2]; float4 y[RPW / 2]; };
    __device__ static void fetch(const Params& p, const Tile& t, int kb, int pw, int lane, Regs& g)
    {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        // Q: 64 float4 per row -> two per lane
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int tt = (t.kb0 + kb) * KT + pw * RPW + rr;
            const float* qrow = p.q + (t.r0 + tt) * FV_K;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) g.q[rr * 2 + h2] = tt < t.t ? ldg4(qrow + (lane + 32 * h2) * 4) : z;
        }
        // Y: 16 float4 per row -> half a warp per row, two rows per pass
#pragma unroll
        for (int pr = 0; pr < RPW / 2; ++pr) {
            const int tt = (t.kb0 + kb) * KT + pw * RPW + pr * 2 + (lane >> 4);
            g.y[pr] = tt < t.t ? ldg4(p.y + (t.r0 + tt) * FV_D + (lane & 15) * 4) : z;
        }
    }
    // zeroth-order statistics: every producer lane keeps the running column sums of the Q
    // values it handles (8 components per lane) and writes them out once per image; the
    // eight per-warp partials are added up by fv_finalize.  (Reading the Q tiles back from
    // shared memory for this cost more shared-memory bandwidth than the MMAs had left.)
    struct PState { float4 s0[2]; };
    static_assert(PGROUPS * 4 == TC_FV_S0_PARTS, "one partial per producer warp");
    __device__ static void store(const Params& p, const Tile& t, int kb, const Regs& g, uint8_t* a_hi, uint8_t* a_lo,
                                 uint8_t* b_hi, uint8_t* b_lo, int pw, int grp, int lane, PState& ps)
    {
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = pw * RPW + rr;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int c4 = lane + 32 * h2;
                const float4 qv = g.q[rr * 2 + h2];
                ps.s0[h2].x += qv.x; ps.s0[h2].y += qv.y; ps.s0[h2].z += qv.z; ps.s0[h2].w += qv.w;
                float4 h, l;
                split4(qv, h, l);
                const uint32_t off = (uint32_t)((c4 >> 3) * B_LBO) + sw32_off(r, c4 & 7);
                *reinterpret_cast<float4*>(b_hi + off) = h;
                *reinterpret_cast<float4*>(b_lo + off) = l;
            }
        }
        // y*y goes into column blocks 0,1 and y into blocks 2,3
#pragma unroll
        for (int pr = 0; pr < RPW / 2; ++pr) {
            const int r = pw * RPW + pr * 2 + (lane >> 4);
            const int c4 = lane & 15;
            const float4 v = g.y[pr];
            const float4 s = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
            float4 h, l;
            const uint32_t off = (uint32_t)((c4 >> 3) * A_LBO) + sw32_off(r, c4 & 7);
            split4(s, h, l);
            *reinterpret_cast<float4*>(a_hi + off) = h;
            *reinterpret_cast<float4*>(a_lo + off) = l;
            split4(v, h, l);
            *reinterpret_cast<float4*>(a_hi + 2 * A_LBO + off) = h;
            *reinterpret_cast<float4*>(a_lo + 2 * A_LBO + off) = l;
        }
        const int kba = t.kb0 + kb;                        // k-block index inside the image (tiles are segments of an image)
        if (kba + PGROUPS >= (t.t + KT - 1) / KT) {        // this group's last k-block of the image
            // slot by the k-block's position in the image, not by the group: which rows meet in a
            // partial sum then does not depend on the tiles this CTA handled before (same bits for
            // any chunking of the batch)
            float4* dst = reinterpret_cast<float4*>(p.s0part + ((t.img * (PGROUPS * 4) + (kba % PGROUPS) * 4 + pw) * FV_K));
            dst[lane] = ps.s0[0];
            dst[lane + 32] = ps.s0[1];
            ps.s0[0] = ps.s0[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane, uint8_t*,
                                    EpiState&)
    {
        if (t.skip) return;                                // past the image's last segment
        const int e = quarter * 32 + lane;                 // operand row: [0,64) = y*y, [64,128) = y
        const int n = e < FV_D ? FV_D + e : e - FV_D;      // column in the [ s1 | s2 ] layout
        // the epilogue warps have nothing else to do here (the producers take the zeroth-order sums), so they fold;
        // S stays raw: fv_finalize applies 1 / T
        seg_fold(t, p.S + t.img * (int64_t)FV_K * FV_2D + n, tmem);
    }
};


// ---------------------------------------------------------------------------------------
// stats, fp16x2 operands: the same contraction with both operands scaled into fp16's range by
// exact powers of two (Q by 2^14; y by 2^-e, e from the mixture model so that |mu| + 6 sigma
// <= 128 * 2^e, hence y'^2 <= 2^14) and split x = hi + lo into two fp16 parts.  Three
// kind::f16 MMAs (hi*lo + lo*hi + hi*hi) keep 22 mantissa bits -- the class of 3xTF32 -- at
// twice the tensor rate and half the shared-memory bytes per stage, which moves this kernel
// from the tensor pipe to the HBM roofline (it has to read Q: 1 KB per descriptor).  Gated
// by a device flag the projection kernel raises when some |y| would leave the fp16 range; the
// 3xTF32 kernel is launched right behind with the opposite gate.
// ---------------------------------------------------------------------------------------
struct Stats16Params {
    CUtensorMap qh_map, ql_map;               // Q * 2^14 as fp16 hi / lo planes [rows, 256], box 64 cols x 16 rows
    CUtensorMap y_map;                        // Y [rows, 64] fp32, box 32 cols x 16 rows
    StatsParams b;
    const float* rinv;                        // [rows + 16, 4]: per-descriptor softmax normaliser (slot 0 of 16 bytes)
    const int* flag;                          // != 0: operands out of fp16 range -> this kernel does nothing
    float sc_y, un1, un2;                     // 2^-e, 2^(e-14), 2^(2e-14)
    int red;                                  // later segments are added to S with RED.ADD instead of load + add + store
};

// The posterior kernel leaves e' = 2^14 exp(l - max) split into two fp16 planes plus the per-descriptor
// normaliser r = 1 / sum_k exp(l - max) (q = r e' / 2^14): r is folded into the [y'^2 | y'] operand rows and
// into the zeroth-order sums here, which saved the posterior kernel a pass over its accumulator.
// The planes are already scaled and split, so the Q operand of a
// stage is eight TMA boxes straight into the swizzled MN-major tiles -- no registers, no conversion,
// as many stages in flight as shared memory holds.  Y (256 B per descriptor) arrives by TMA as fp32
// in a staging tile of the stage; four converter warps scale / square / split it into the A tiles,
// zero the Q rows past the end of the image in the image's last stage, and publish the stage.  The
// epilogue warps read the Q tiles of every stage for the zeroth-order sums while the MMAs run.
struct Stats16Policy {
    using Params = Stats16Params;
    struct EpiState { float2 s0; };
    using Tile = SegTile;                      // one segment of an image (see StatsParams)
    static constexpr bool BF16 = false, F16 = true, A_MN = true, B_MN = true, EPI_READS_STAGES = true, MANUAL = true;
    // The fold of a finished segment belongs to four warps of its own: the epilogue warps have to release every operand stage
    // after reading it for the zeroth-order sums, and with the fold on them the ring stalled (4.6 -> 5.9 ms for the kernel).
    static constexpr bool FOLD_WARPS = true;
    struct FoldState {};
    static constexpr bool TMA_OWN_BARRIER = true;
    static constexpr int KT = 16;             // descriptors per stage = one K=16 step
    static constexpr int PASSES = 3, BLOCK_N = FV_K, STAGES = 7, KSTEPS = 1, PGROUPS = 1;
    static constexpr int A_LBO = KT * 128, B_LBO = KT * 128;          // one [KT rows x 128 B] block per 64 columns
    static constexpr int A_BYTES = (FV_2D / 64) * A_LBO, B_BYTES = (FV_K / 64) * B_LBO, SCRATCH_BYTES = 0;
    static constexpr int Y_STAGE = KT * FV_D * 4;                     // two [16 x 32] fp32 boxes
    // extra region of a stage: Y staging | r staging (16 x 16 B) | cleaned r (16 floats) | the policy's mbarrier
    static constexpr int R_OFF = Y_STAGE, RC_OFF = Y_STAGE + KT * 16, BAR_OFF = Y_STAGE + 512;
    static constexpr int STAGE_EXTRA = Y_STAGE + 1024;
    static constexpr int TMA_BYTES = 2 * B_BYTES + Y_STAGE + KT * 16;
    __device__ static bool enabled(const Params& p) { return *p.flag == 0; }
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.qh_map); tma_prefetch_desc(&p.ql_map); tma_prefetch_desc(&p.y_map); }
    __device__ static void init_stage(uint8_t* extra) { mbar_init(reinterpret_cast<uint64_t*>(extra + BAR_OFF), 1); }
    __device__ static int num_tiles(const Params& p) { return (int)p.b.n_images * *p.b.smax; }
    __device__ static int tile_at(const Params& p, int it, int) { return seg_tile_at(p.b, it); }
    __device__ static Tile tile(const Params& p, int i) { return seg_tile(p.b, i, KT); }
    __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t*, uint8_t*, uint8_t* b_hi, uint8_t* b_lo, uint64_t*)
    {
        uint8_t* extra = b_lo + B_BYTES;
        uint64_t* bar = reinterpret_cast<uint64_t*>(extra + BAR_OFF);
        const int row = (int)(t.r0 + (int64_t)(t.kb0 + kb) * KT);
        mbar_expect_tx(bar, TMA_BYTES);
        bulk_load_1d(extra + R_OFF, p.rinv + (int64_t)row * 4, KT * 16, bar);     // (the array is padded by 16 rows)
#pragma unroll
        for (int cb = 0; cb < FV_K / 64; ++cb) {
            tma_load_2d(b_hi + cb * B_LBO, &p.qh_map, bar, cb * 64, row);
            tma_load_2d(b_lo + cb * B_LBO, &p.ql_map, bar, cb * 64, row);
        }
        tma_load_2d(extra, &p.y_map, bar, 0, row);
        tma_load_2d(extra + Y_STAGE / 2, &p.y_map, bar, 32, row);
    }
    struct Regs {};
    __device__ static void fetch(const Params&, const Tile&, int, int, int, Regs&) {}
    struct PState { uint32_t uses; };          // stages published so far by this warp (PGROUPS == 1: all of them)
    __device__ static void store(const Params& p, const Tile& t, int kb, const Regs&, uint8_t* a_hi, uint8_t* a_lo,
                                 uint8_t* b_hi, uint8_t* b_lo, int pw, int, int lane, PState& ps)
    {
        uint8_t* extra = b_lo + B_BYTES;
        mbar_wait(reinterpret_cast<uint64_t*>(extra + BAR_OFF), (ps.uses / STAGES) & 1u);
        ++ps.uses;
        const int valid = t.t - (t.kb0 + kb) * KT;         // rows of this stage that belong to the image
        // a lane owns 8 dims of one row: row r = 4 pw + lane / 8, dims [8 (lane % 8), +8)
        const int r = pw * 4 + (lane >> 3);
        const int blk = (lane & 7) >> 2, c0 = (lane & 3) * 2;
        const uint8_t* src = extra + blk * (Y_STAGE / 2);
        float4 y0 = *reinterpret_cast<const float4*>(src + sw128_off(r, c0));
        float4 y1 = *reinterpret_cast<const float4*>(src + sw128_off(r, c0 + 1));
        float rn = *reinterpret_cast<const float*>(extra + R_OFF + r * 16);
        if (r >= valid) { y0 = y1 = make_float4(0.f, 0.f, 0.f, 0.f); rn = 0.f; }
        if ((lane & 7) == 0) *reinterpret_cast<float*>(extra + RC_OFF + r * 4) = rn;   // for the zeroth-order sums
        const float w[8] = {y0.x * p.sc_y, y0.y * p.sc_y, y0.z * p.sc_y, y0.w * p.sc_y,
                            y1.x * p.sc_y, y1.y * p.sc_y, y1.z * p.sc_y, y1.w * p.sc_y};
        float sq[8], v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { sq[i] = w[i] * w[i] * rn; v[i] = w[i] * rn; }
        uint4 h, l;
        const uint32_t off = sw128_off(r, lane & 7);       // y'^2 -> column block 0, y' -> block 1
        split8_h(sq, h, l);
        *reinterpret_cast<uint4*>(a_hi + off) = h;
        *reinterpret_cast<uint4*>(a_lo + off) = l;
        split8_h(v, h, l);
        *reinterpret_cast<uint4*>(a_hi + A_LBO + off) = h;
        *reinterpret_cast<uint4*>(a_lo + A_LBO + off) = l;
        if (valid < KT) {
            // last stage of the image: the Q boxes also brought rows of the next image (or zeros past the
            // end of the batch); clear them so that neither the MMA (0 x NaN) nor the zeroth-order sums see them
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            for (int i = pw * 32 + lane; i < (KT - valid) * 64; i += 128) {
                const int rr = valid + (i >> 6), part = (i >> 5) & 1, cb = (i >> 3) & 3, ch = i & 7;
                *reinterpret_cast<uint4*>((part ? b_lo : b_hi) + cb * B_LBO + rr * 128 + (ch << 4)) = z;
            }
        }
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile& t, EpiState& st, int, int) { if (t.first) st.s0 = make_float2(0.f, 0.f); }
    // zeroth-order sums: thread (quarter, lane) owns components 2e, 2e + 1 (e = 32 quarter + lane) = one
    // 4-byte word of every row in column block `quarter`; a warp reads one 128-B row per instruction
    __device__ static void consume(const Params&, const Tile&, int, uint8_t*, uint8_t*, uint8_t* b_hi, uint8_t* b_lo,
                                   EpiState& st, int quarter, int lane)
    {
        const uint32_t base = (uint32_t)(quarter * B_LBO + (lane & 3) * 4);
        const int ch = lane >> 2;
        const float4* rc = reinterpret_cast<const float4*>(b_lo + B_BYTES + RC_OFF);   // cleaned normalisers of the 16 rows
        float2 acc = st.s0;
#pragma unroll
        for (int r4 = 0; r4 < KT / 4; ++r4) {
            const float4 rr = rc[r4];
            const float rv[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r4 * 4 + u;
                const uint32_t off = base + (uint32_t)(r * 128 + ((ch ^ (r & 7)) << 4));
                const float2 h = __half22float2(*reinterpret_cast<const __half2*>(b_hi + off));
                const float2 l = __half22float2(*reinterpret_cast<const __half2*>(b_lo + off));
                acc.x = fmaf(h.x + l.x, rv[u], acc.x);
                acc.y = fmaf(h.y + l.y, rv[u], acc.y);
            }
        }
        st.s0 = acc;
    }
    // epilogue warps: only the zeroth-order sums (collected stage by stage in consume())
    __device__ static void epilogue_host(const Params& p, const Tile& t, EpiState& st, int quarter, int lane)
    {
        if (t.skip || !t.last) return;
        const int e = quarter * 32 + lane;
        // raw zeroth-order sums go into partial slot 0 (the other slots stay zero, fv_finalize adds them up and divides by T)
        reinterpret_cast<float2*>(p.b.s0part + t.img * (int64_t)(TC_FV_S0_PARTS * FV_K))[e] =
            make_float2(st.s0.x * (1.f / 16384.f), st.s0.y * (1.f / 16384.f));
    }
    // fold warps: add the finished segment to the image's S rows.  S stays in raw operand units (fv_finalize applies the
    // operand scales and 1 / T, like for the cluster kernel).  The first segment of an image stores; the later ones either
    // go as fire-and-forget fp32 reductions (RED.ADD at L2: round to nearest, one owner thread per address and program
    // order per address, so the sum is the same ((s0 + s1) + s2) ... as with the loads) or, PVS_FV_RED=0, as 64 loads per
    // round trip to L2 that come back while the accumulator is read.
    __device__ static void fold(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane, FoldState&)
    {
        if (t.skip) return;                                // past the image's last segment
        const int e = quarter * 32 + lane;                 // operand row: [0,64) = y'^2, [64,128) = y'
        const int n = e < FV_D ? FV_D + e : e - FV_D;      // column in the [ s1 | s2 ] layout
        float* Simg = p.b.S + t.img * (int64_t)FV_K * FV_2D + n;
        const bool empty = t.t == 0;                       // T == 0: NaN, like the reference's division by zero
        const float nanv = __int_as_float(0x7fc00000);
        if (t.first || p.red) {
#pragma unroll 1
            for (int c = 0; c < FV_K; c += 32) {
                float v[32];
                tmem_ld32(tmem + c, v);
                tmem_ld_wait();
                if (t.first) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) __stcg(Simg + (int64_t)(c + jj) * FV_2D, empty ? nanv : v[jj]);
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) atomicAdd(Simg + (int64_t)(c + jj) * FV_2D, v[jj]);
                }
            }
            return;
        }
#pragma unroll 1
        for (int c = 0; c < FV_K; c += 64) {               // 64 running sums in flight per trip to L2, the accumulator in two halves
            float v[32], r[64];
#pragma unroll
            for (int jj = 0; jj < 64; ++jj) r[jj] = __ldcg(Simg + (int64_t)(c + jj) * FV_2D);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                __syncwarp();
                tmem_ld32(tmem + c + 32 * h, v);
                tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) __stcg(Simg + (int64_t)(c + 32 * h + jj) * FV_2D, v[jj] + r[32 * h + jj]);
            }
        }
    }
    __device__ static void epilogue(const Params&, const Tile&, uint32_t, int, int, uint8_t*, EpiState&) {}   // unused (FOLD_WARPS)
};
// the 3xTF32 kernel behind the opposite gate
struct StatsGatedPolicy : StatsPolicy {
    using Params = Stats16Params;
    __device__ static bool enabled(const Params& p) { return *p.flag != 0; }
    __device__ static int num_tiles(const Params& p) { return StatsPolicy::num_tiles(p.b); }
    __device__ static int tile_at(const Params& p, int it, int n) { return StatsPolicy::tile_at(p.b, it, n); }
    __device__ static Tile tile(const Params& p, int i) { return StatsPolicy::tile(p.b, i); }
    __device__ static void prefetch(const Params&) {}
    __device__ static void load(const Params&, const Tile&, int, uint8_t*, uint8_t*, uint8_t*, uint8_t*, uint64_t*) {}
    __device__ static void fetch(const Params& p, const Tile& t, int kb, int pw, int lane, Regs& g) { StatsPolicy::fetch(p.b, t, kb, pw, lane, g); }
    __device__ static void store(const Params& p, const Tile& t, int kb, const Regs& g, uint8_t* a_hi, uint8_t* a_lo,
                                 uint8_t* b_hi, uint8_t* b_lo, int pw, int grp, int lane, PState& ps)
    {
        StatsPolicy::store(p.b, t, kb, g, a_hi, a_lo, b_hi, b_lo, pw, grp, lane, ps);
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane, uint8_t* sc, EpiState& st)
    {
        StatsPolicy::epilogue(p.b, t, tmem, quarter, lane, sc, st);
    }
};


// ---------------------------------------------------------------------------------------
// stats, any feature dimension D (K = 256): the [2D x K] statistics are cut into 128-row M
// tiles of the augmented operand [y*y | y]; tile = (image, M tile).  The M tiles of an image
// run on neighbouring CTAs at the same time, so its posteriors are read from HBM once and
// served to the other tiles from L2.  Zeroth-order partials come from the M tile 0 producers.
// (VGG16-PCA: D = 257 -> 5 M tiles; RootSIFT without PCA: D = 128 -> 2.)
// ---------------------------------------------------------------------------------------
struct StatsGenParams {
    const float* y;                           // [rows, d]
    const float* q;                           // [rows, 256]
    const int64_t* offsets;
    float* S;                                 // [n_images, 256, 2d]: [k][ s1 (d) | s2 (d) ], already / T
    float* s0part;                            // [n_images, 16, 256]
    int64_t n_images;
    int d, n_mt;
    const int* smax;                          // segments per (image, M tile) slot, see StatsParams
    int segk;
};

struct StatsGenPolicy {
    using Params = StatsGenParams;
    using EpiState = NoEpiState;
    struct Tile : SegTile { int mt; };         // one segment of (image, M tile of 128 augmented columns)
    static constexpr bool BF16 = false, A_MN = true, B_MN = true, EPI_READS_STAGES = false, MANUAL = true;
    static constexpr int KT = ST_KT;
    static constexpr int PASSES = 3, BLOCK_N = FV_K, STAGES = 4, KSTEPS = KT / 8, PGROUPS = 4;
    static constexpr int A_LBO = KT * 128, B_LBO = KT * 128;
    static constexpr int A_BYTES = 4 * A_LBO, B_BYTES = (FV_K / 32) * B_LBO, SCRATCH_BYTES = 0, TMA_BYTES = 0;
    static constexpr int RPW = KT / 4;
    static_assert(PGROUPS * 4 == TC_FV_S0_PARTS, "one partial per producer warp");
    __device__ static void prefetch(const Params&) {}
    __device__ static int num_tiles(const Params& p) { return (int)(p.n_images * p.n_mt) * *p.smax; }
    __device__ static int tile_at(const Params& p, int it, int)
    {
        const int smax = *p.smax;
        const long long u = (long long)blockIdx.x + (long long)(it / smax) * gridDim.x;    // (image, M tile) unit
        return u < p.n_images * p.n_mt ? (int)(u * smax + it % smax) : -1;
    }
    __device__ static Tile tile(const Params& p, int i)
    {
        const int smax = *p.smax;
        const int u = i / smax, seg = i - u * smax;
        const int64_t img = u / p.n_mt;
        const int64_t r0 = p.offsets[img];
        const int t = (int)(p.offsets[img + 1] - r0);
        const int total = (t + KT - 1) / KT, kb0 = seg * p.segk;
        const int left = total - kb0;
        Tile tl;
        tl.nkb = left <= 0 ? 0 : (left < p.segk ? left : p.segk);
        tl.t = t; tl.img = img; tl.r0 = r0; tl.kb0 = kb0;
        tl.first = seg == 0; tl.last = left <= p.segk; tl.skip = seg > 0 && left <= 0;
        tl.mt = u - (int)img * p.n_mt;
        return tl;
    }
    __device__ static void load(const Params&, const Tile&, int, uint8_t*, uint8_t*, uint8_t*, uint8_t*, uint64_t*) {}
    struct Regs { float4 q[RPW * 2]; float4 y[RPW]; };
    struct PState { float4 s0[2]; };
    // augmented column a of a descriptor row: [0, d) -> y[a]^2, [d, 2d) -> y[a - d], beyond -> 0
    __device__ static float aug(const Params& p, const float* yrow, int a)
    {
        if (a < p.d) { const float v = __ldg(yrow + a); return v * v; }
        return a < 2 * p.d ? __ldg(yrow + a - p.d) : 0.f;
    }
    __device__ static void fetch(const Params& p, const Tile& t, int kb, int pw, int lane, Regs& g)
    {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const int a0 = t.mt * 128 + lane * 4;                  // this lane's four augmented columns
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int tt = (t.kb0 + kb) * KT + pw * RPW + rr;
            const bool valid = tt < t.t;
            const float* qrow = p.q + (t.r0 + tt) * FV_K;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) g.q[rr * 2 + h2] = valid ? ldg4(qrow + (lane + 32 * h2) * 4) : z;
            g.y[rr] = z;
            if (valid) {
                const float* yrow = p.y + (t.r0 + tt) * (int64_t)p.d;
                g.y[rr] = make_float4(aug(p, yrow, a0), aug(p, yrow, a0 + 1), aug(p, yrow, a0 + 2), aug(p, yrow, a0 + 3));
            }
        }
    }
    __device__ static void store(const Params& p, const Tile& t, int kb, const Regs& g, uint8_t* a_hi, uint8_t* a_lo,
                                 uint8_t* b_hi, uint8_t* b_lo, int pw, int grp, int lane, PState& ps)
    {
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = pw * RPW + rr;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int c4 = lane + 32 * h2;
                const float4 qv = g.q[rr * 2 + h2];
                ps.s0[h2].x += qv.x; ps.s0[h2].y += qv.y; ps.s0[h2].z += qv.z; ps.s0[h2].w += qv.w;
                float4 h, l;
                split4(qv, h, l);
                const uint32_t off = (uint32_t)((c4 >> 3) * B_LBO) + sw32_off(r, c4 & 7);
                *reinterpret_cast<float4*>(b_hi + off) = h;
                *reinterpret_cast<float4*>(b_lo + off) = l;
            }
            float4 h, l;
            split4(g.y[rr], h, l);
            const uint32_t off = (uint32_t)((lane >> 3) * A_LBO) + sw32_off(r, lane & 7);
            *reinterpret_cast<float4*>(a_hi + off) = h;
            *reinterpret_cast<float4*>(a_lo + off) = l;
        }
        const int kba = t.kb0 + kb;                        // k-block index inside the image (tiles are segments)
        if (kba + PGROUPS >= (t.t + KT - 1) / KT) {        // this group's last k-block of the (image, M tile)
            if (t.mt == 0) {
                float4* dst = reinterpret_cast<float4*>(p.s0part + ((t.img * (PGROUPS * 4) + (kba % PGROUPS) * 4 + pw) * FV_K));
                dst[lane] = ps.s0[0];
                dst[lane + 32] = ps.s0[1];
            }
            ps.s0[0] = ps.s0[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        (void)grp;
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane, uint8_t*,
                                    EpiState&)
    {
        if (t.skip) return;                                // past the last segment
        const int a = t.mt * 128 + quarter * 32 + lane;    // augmented column owned by this thread
        const int ld = 2 * p.d;
        const int col = a < p.d ? p.d + a : a - p.d;       // [ s1 | s2 ] layout
        // fold of the segment (the epilogue warps are otherwise idle); S stays raw, fv_finalize applies 1 / T
        seg_fold(t, p.S + t.img * (int64_t)FV_K * ld + col, tmem, ld, a < ld);
    }
};

// ---------------------------------------------------------------------------------------
// CTA-pair versions of project and posterior (pvs_tc2.cuh): the weight operand is resident
// in shared memory (half per CTA) instead of being re-streamed from L2 for every tile, and
// three producer groups keep three operand stages in flight.
// ---------------------------------------------------------------------------------------
}  // namespace tc
namespace tc2 {
using tc::FV_D;
using tc::FV_K;
using tc::FV_2D;
using tc::KRegs;
using tc::fetch_kmajor_32rows;
using tc::store_kmajor_32rows;
using tc::PcaParams;
using tc::PostParams;
using tc::NoEpiState;

// Y = X C^T + b, d_in == 128: pair tile [256 rows x 64], C resident (32 rows of C per CTA)
struct PcaPairPolicy {
    using Params = PcaParams;
    using EpiState = NoEpiState;
    struct Tile { int nkb, mb; };
    static constexpr bool BF16 = false, MANUAL = true, B_RESIDENT = true, ACC_INIT = false, TILE_SYNC = false;
    static constexpr int PASSES = 3, BLOCK_N = FV_D, KSTEPS = 4, NKB_RES = 4, STAGES = 4, PGROUPS = 4;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = (FV_D / 2) * 128, TMA_BYTES = 0;
    // scratch starts 256 B past a 1024-B boundary: pad, two 1024-aligned [32 x 32] fp32 TMA-store tiles per
    // epilogue warp (the two 32-column halves of its rows), then the bias
    static constexpr int STG_OFF = 768, BIAS_OFF = STG_OFF + 4 * 2 * 4096, SCRATCH_BYTES = BIAS_OFF + 256;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.c_hi); tma_prefetch_desc(&p.c_lo); tma_prefetch_desc(&p.y_map); }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks; }
    __device__ static int tile_at(const Params&, int it, int pair, int n_pairs, int n)
    {
        const long long t = (long long)pair + (long long)it * n_pairs;
        return t < n ? (int)t : -1;
    }
    __device__ static Tile tile(const Params& p, int i) { return {p.nkb, i}; }
    __device__ static void load_resident(const Params& p, int rank, uint8_t* res, uint64_t* bar)
    {
        for (int kb = 0; kb < NKB_RES; ++kb) {
            tma_load_2d_pair(res + (2 * kb) * B_BYTES, &p.c_hi, bar, kb * 32, rank * (FV_D / 2));
            tma_load_2d_pair(res + (2 * kb + 1) * B_BYTES, &p.c_lo, bar, kb * 32, rank * (FV_D / 2));
        }
    }
    __device__ static void load(const Params&, const Tile&, int, int, uint8_t*, uint8_t*, uint8_t*, uint8_t*, uint64_t*) {}
    using Regs = KRegs;
    __device__ static void fetch(const Params& p, const Tile& t, int kb, int rank, int pw, int lane, Regs& r)
    {
        fetch_kmajor_32rows(p.x, p.d_in, (int64_t)t.mb * 256 + rank * 128, p.rows, kb * 32, pw, lane, r);
    }
    __device__ static void store(const Params&, const Tile&, int, const Regs& r, uint8_t* a_hi, uint8_t* a_lo, int pw, int lane)
    {
        store_kmajor_32rows<false>(r, a_hi, a_lo, pw, lane);
    }
    __device__ static void epi_init(const Params& p, uint8_t* scratch, int tid)
    {
        float* b = reinterpret_cast<float*>(scratch + BIAS_OFF);
        if (tid < FV_D) b[tid] = p.bias[tid];
        epi_barrier();
    }
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, int rank, uint32_t tmem, int quarter, int lane,
                                    uint8_t* scratch, EpiState&)
    {
        const float* bias = reinterpret_cast<const float*>(scratch + BIAS_OFF);
        const int wrow0 = (int)((int64_t)t.mb * 256 + rank * 128 + quarter * 32);
        uint8_t* stg = scratch + STG_OFF + quarter * (2 * 4096);
        // a thread owns a row; the [32 rows x 32 cols] halves go through 128-byte-swizzled shared
        // tiles and leave as TMA stores (full 128-byte row segments, rows past the end clipped)
        if (lane == 0) tma_store_wait_read<0>();                 // the previous tile's stores have read their tiles
        __syncwarp();
        float amax = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float v[32];
            tmem_ld32(tmem + half * 32, v);
            tmem_ld_wait();
            float* tile = reinterpret_cast<float*>(stg + half * 4096);
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] += bias[half * 32 + j]; amax = fmaxf(amax, fabsf(v[j])); }
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(tile + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                    make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        }
        // fp16x2 range guard (rows past the end hold the bias only); inf raises it, NaN is dropped by fmaxf and
        // simply propagates through either precision variant
        if (p.flag && amax > p.lim) *p.flag = 1;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(&p.y_map, stg, 0, wrow0);
            tma_store_2d(&p.y_map, stg + 4096, 32, wrow0);
            tma_store_commit();
        }
    }
};

// logits + softmax -> Q: pair tile [256 rows x 256 components], W = [-P/2 | mu P] resident
// H = false: 3xTF32 operands (tf32 hi / lo).  H = true: fp16x2 -- y scaled by 2^-e and the weights by
// 4^e / 2^e (exact), both split into fp16 hi + lo: the same 22-bit products at twice the tensor rate;
// a k-block is then 32 dimensions (64 interleaved fp16 columns) and the resident weights take half the space.
template <bool H>
struct PostPairT {
    using Params = PostParams;
    using EpiState = NoEpiState;
    struct Tile { int nkb, mb; };
    static constexpr bool BF16 = false, F16 = H, MANUAL = true, B_RESIDENT = true, ACC_INIT = false, TILE_SYNC = false;
    static constexpr int KCOLS = H ? 64 : 32;                         // operand columns per 128-B row = per k-block
    static constexpr int PASSES = 3, BLOCK_N = FV_K, KSTEPS = 4, NKB_RES = FV_2D / KCOLS, STAGES = 2;
    // fp16x2: one producer group (a tile is only two k-blocks) keeps the CTA at 448 threads, i.e. enough registers
    // for the eight epilogue warps not to spill
    static constexpr int PGROUPS = H ? 1 : 2;
    static constexpr bool PREFETCH2 = H;                              // ... with both k-blocks of the next tile in flight
    static constexpr int A_BYTES = 128 * 128, B_BYTES = (FV_K / 2) * 128, TMA_BYTES = 0;
    // scratch starts 256 B past a 1024-B boundary (barrier block): 768 B pad, then two 1024-aligned 4-KB
    // TMA-store staging tiles per epilogue warp (two [32 x 32] fp32 tiles, or a (hi, lo) pair of
    // [32 x 64] fp16 tiles), then cst and the exchange area of the warp pairs
    static constexpr int EPI_WARPS = H ? 8 : 4;                       // (the 3xTF32 variant has no shared memory left for eight)
    static constexpr int NH = EPI_WARPS / 4, CW = FV_K / NH;          // column halves per lane quarter, columns per warp
    static constexpr int STG_OFF = 768, CST_OFF = STG_OFF + EPI_WARPS * 8192, XCH_OFF = CST_OFF + 1024,
                         SCRATCH_BYTES = XCH_OFF + (NH == 2 ? 3072 : 0);
    __device__ static bool enabled(const Params& p) { return !p.flag || (*p.flag == 0) == H; }
    __device__ static void prefetch(const Params& p)
    {
        tma_prefetch_desc(&p.w_hi); tma_prefetch_desc(&p.w_lo); tma_prefetch_desc(&p.q_map);
        if (H) { tma_prefetch_desc(&p.qh_map); tma_prefetch_desc(&p.ql_map); }
    }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks; }
    __device__ static int tile_at(const Params&, int it, int pair, int n_pairs, int n)
    {
        const long long t = (long long)pair + (long long)it * n_pairs;
        return t < n ? (int)t : -1;
    }
    __device__ static Tile tile(const Params&, int i) { return {NKB_RES, i}; }
    __device__ static void load_resident(const Params& p, int rank, uint8_t* res, uint64_t* bar)
    {
        for (int kb = 0; kb < NKB_RES; ++kb) {
            tma_load_2d_pair(res + (2 * kb) * B_BYTES, &p.w_hi, bar, kb * KCOLS, rank * (FV_K / 2));
            tma_load_2d_pair(res + (2 * kb + 1) * B_BYTES, &p.w_lo, bar, kb * KCOLS, rank * (FV_K / 2));
        }
    }
    __device__ static void load(const Params&, const Tile&, int, int, uint8_t*, uint8_t*, uint8_t*, uint8_t*, uint64_t*) {}
    // operand columns [0,64) are y*y, [64,128) are y
    // Operand row = interleaved (y_d^2, y_d) pairs (the weight copy is interleaved to match): the
    // quadratic and the linear term of a dimension are accumulated next to each other, so the
    // running sum in TMEM stays at the scale of the final logit -- the tensor core truncates when
    // it accumulates and that error is proportional to the running sum (3x lower FV error than
    // the [y*y | y] order, tools/probe_fv_err.py).  K-block kb holds dimensions [16 kb, 16 kb + 16).
    // a lane fills one 16-byte chunk of the operand row: two dimensions (tf32) or four (fp16)
    struct Regs { float4 v[8]; };
    __device__ static void fetch(const Params& p, const Tile& t, int kb, int rank, int pw, int lane, Regs& r)
    {
        const int64_t row0 = (int64_t)t.mb * 256 + rank * 128 + pw * 32;
        const int c = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t gr = row0 + i * 4 + (lane >> 3);
            r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gr < p.rows) {
                if constexpr (H) r.v[i] = tc::ldg4(p.y + gr * FV_D + kb * 32 + c * 4);
                else {
                    const float2 y2 = __ldg(reinterpret_cast<const float2*>(p.y + gr * FV_D + kb * 16 + c * 2));
                    r.v[i].x = y2.x; r.v[i].y = y2.y;
                }
            }
        }
    }
    __device__ static void store(const Params& p, const Tile&, int, const Regs& r, uint8_t* a_hi, uint8_t* a_lo, int pw, int lane)
    {
        const int c = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = pw * 32 + i * 4 + (lane >> 3);
            const uint32_t off = tc::sw128_off(row, c);
            if constexpr (H) {
                const float a = r.v[i].x * p.sc_y, b = r.v[i].y * p.sc_y, cc = r.v[i].z * p.sc_y, d = r.v[i].w * p.sc_y;
                const float x[8] = {a * a, a, b * b, b, cc * cc, cc, d * d, d};
                uint4 h, l;
                tc::split8_h(x, h, l);
                *reinterpret_cast<uint4*>(a_hi + off) = h;
                *reinterpret_cast<uint4*>(a_lo + off) = l;
            } else {
                const float4 y = r.v[i];
                float4 h, l;
                tc::split4(make_float4(y.x * y.x, y.x, y.y * y.y, y.y), h, l);
                *reinterpret_cast<float4*>(a_hi + off) = h;
                *reinterpret_cast<float4*>(a_lo + off) = l;
            }
        }
    }
    // Eight epilogue warps: the two warps of a TMEM lane quarter split the 256 components of their 32 rows
    // (half h takes columns [128 h, 128 h + 128)) and reconcile the row maximum and the row sum through shared
    // memory.  One warp per scheduler and row was latency-bound (dependent TMEM load -> reduce chains); two
    // independent warps per scheduler halve the time a tile spends in the epilogue.
    __device__ static void pair_barrier(int quarter) { asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory"); }
    __device__ static void epi_init(const Params& p, uint8_t* scratch, int tid)
    {
        float* c = reinterpret_cast<float*>(scratch + CST_OFF);
        for (int i = tid; i < FV_K; i += 32 * EPI_WARPS) c[i] = p.cst[i];
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    }
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, int rank, uint32_t tmem, int quarter, int lane,
                                    uint8_t* scratch, EpiState&)
    {
        const int half = NH == 2 ? (((int)threadIdx.x >> 5) - 2) >> 2 : 0;
        const int c0 = half * CW;
        const int64_t row = (int64_t)t.mb * 256 + rank * 128 + quarter * 32 + lane;
        const bool valid = row < p.rows;
        // The per-component constant is added here in fp32 rather than pre-loaded into the
        // accumulator: the tensor core truncates when it accumulates, relative to the running
        // sum, and starting that sum at |cst| ~ 10^2 cost a factor 3.6 in FV accuracy.
        // Every pass keeps the TMEM load of the next 32 columns in flight while it works on the
        // current ones and splits its reductions over four independent chains.
        float va[32], vb[32];
        const float* cstv = reinterpret_cast<const float*>(scratch + CST_OFF);
        // [2 halves][max, arg-max, sum][32 rows]: three slots, so that every write is separated from the
        // partner's last read of the same slot by a pair barrier
        float* xch = reinterpret_cast<float*>(scratch + XCH_OFF) + quarter * 192;
        auto addc = [&](float (&v)[32], int c) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += cstv[c + j];
        };
        const bool tme = quarter == 0 && lane == 0 && rank == 0 && half == 0;
        (void)tme;
        PVS_T0(tp1);
        // pass 1: maximum of this warp's 128 columns (arg-max only when the caller asked for it)
        float mx = -INFINITY;
        int mi = 0;
        if (p.argmax) {
#pragma unroll 1
            for (int c = c0; c < c0 + CW; c += 32) {
                tmem_ld32(tmem + c, va);
                tmem_ld_wait();
                addc(va, c);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (va[j] > mx) { mx = va[j]; mi = c + j; }   // strict >: lowest index on ties
            }
        } else {
            float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            tmem_ld32(tmem + c0, va);
#pragma unroll 1
            for (int c = c0; c < c0 + CW; c += 64) {
                tmem_ld_wait();
                tmem_ld32(tmem + c + 32, vb);
                addc(va, c);
#pragma unroll
                for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], va[j]);
                tmem_ld_wait();
                if (c + 64 < c0 + CW) tmem_ld32(tmem + c + 64, va);
                addc(vb, c + 32);
#pragma unroll
                for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], vb[j]);
            }
            mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
        if constexpr (NH == 2) {
            xch[half * 96 + lane] = mx;
            xch[half * 96 + 32 + lane] = __int_as_float(mi);
            pair_barrier(quarter);
            const float omx = xch[(half ^ 1) * 96 + lane];
            const int omi = __float_as_int(xch[(half ^ 1) * 96 + 32 + lane]);
            // the lower half wins ties (lowest index), like one scan over all 256 columns
            if (half == 0 ? omx > mx : omx >= mx) { mx = omx; mi = omi; }
        }
        PVS_TPHASE(8, tp1, tme);
        PVS_T0(tp2);
        const float base = (mx > -INFINITY && mx < INFINITY) ? mx : 0.f;
        // pass 2: e = exp(l - max) = 2^(l log2e - max log2e): one FFMA + one MUFU per logit;
        // e is stashed back into the accumulator columns
        constexpr float LOG2E = 1.4426950408889634f;
        const float nb = -base * LOG2E;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        if (H && p.planes) {
            // Plane output with the normalisation deferred to the consumer: e' = 2^14 exp(l - max) (<= 2^14, so it
            // sits in fp16's range by construction) leaves as fp16 hi + lo planes in the same pass that computes
            // it, and the per-row factor 1 / sum_k exp(l - max) goes to rinv; the statistics kernel folds it into
            // its [y'^2 | y'] operand rows and its zeroth-order sums.  One TMEM read pass less, no TMEM write.
            const int wrow0 = (int)((int64_t)t.mb * 256 + rank * 128 + quarter * 32);
            uint8_t* th = scratch + STG_OFF + (half * 4 + quarter) * 8192;
            uint8_t* tl = th + 4096;
            const float nb14 = nb + 14.f;
            auto exp8 = [&](float (&v)[32], int j8, uint4& h, uint4& l) {
                float x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x[u]) : "f"(fmaf(v[8 * j8 + u], LOG2E, nb14)));
                    s4[u & 3] += x[u];
                }
                tc::split8_h(x, h, l);
            };
            auto put8 = [&](int j, const uint4& h, const uint4& l) {
                const uint32_t off = (uint32_t)(lane * 128 + ((j ^ (lane & 7)) << 4));
                *reinterpret_cast<uint4*>(th + off) = h;
                *reinterpret_cast<uint4*>(tl + off) = l;
            };
            tmem_ld32(tmem + c0, va);
#pragma unroll 1
            for (int c = c0; c < c0 + CW; c += 64) {
                tmem_ld_wait();
                tmem_ld32(tmem + c + 32, vb);
                addc(va, c);
                // the first 32 columns are converted BEFORE waiting for the previous chunk's TMA stores to
                // have read the tiles, so that wait overlaps the exponentials
                uint4 ha[4], la[4];
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) exp8(va, j8, ha[j8], la[j8]);
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) put8(j8, ha[j8], la[j8]);
                tmem_ld_wait();
                if (c + 64 < c0 + CW) tmem_ld32(tmem + c + 64, va);
                addc(vb, c + 32);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    uint4 h, l;
                    exp8(vb, j8, h, l);
                    put8(4 + j8, h, l);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&p.qh_map, th, c, wrow0);
                    tma_store_2d(&p.ql_map, tl, c, wrow0);
                    tma_store_commit();
                }
            }
            const float spart = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            float tot = spart;
            if constexpr (NH == 2) {
                xch[half * 96 + 64 + lane] = spart;
                pair_barrier(quarter);
                tot = (half == 0 ? spart : xch[64 + lane]) + (half == 0 ? xch[96 + 64 + lane] : spart);
            }
            if (half == 0 && valid) {
                p.rinv[row * 4] = 16384.f / tot;             // = 1 / sum_k exp(l - max)
                if (p.argmax) p.argmax[row] = mi;
            }
            PVS_TPHASE(9, tp2, tme);
            return;                                            // (the staging tiles are waited for before their next use)
        }
        auto exp_chunk = [&](float (&v)[32]) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v[j], LOG2E, nb)));
                v[j] = e;
                s4[j & 3] += e;
            }
        };
        tmem_ld32(tmem + c0, va);
#pragma unroll 1
        for (int c = c0; c < c0 + CW; c += 64) {
            tmem_ld_wait();
            tmem_ld32(tmem + c + 32, vb);                      // (tcgen05.st reads its registers at issue)
            addc(va, c);
            exp_chunk(va);
            tmem_st32(tmem + c, va);
            tmem_ld_wait();
            if (c + 64 < c0 + CW) tmem_ld32(tmem + c + 64, va);
            addc(vb, c + 32);
            exp_chunk(vb);
            tmem_st32(tmem + c + 32, vb);
        }
        const float spart = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        float inv;
        if constexpr (NH == 2) {
            xch[half * 96 + 64 + lane] = spart;
            tmem_st_wait();
            pair_barrier(quarter);
            const float s_lo = half == 0 ? spart : xch[64 + lane], s_hi = half == 0 ? xch[96 + 64 + lane] : spart;
            inv = 1.f / (s_lo + s_hi);                         // same association in both warps
        } else {
            tmem_st_wait();
            inv = 1.f / spart;
        }
        PVS_TPHASE(9, tp2, tme);
        PVS_T0(tp3);
        if (valid && p.argmax && half == 0) p.argmax[row] = mi;
        // pass 3: q = e / sum.  A thread owns a row, so storing straight from registers would
        // scatter 16-byte pieces over 32 rows per instruction.  Chunks are written to 128-byte-swizzled
        // shared-memory tiles instead and leave through TMA stores (rows past the end of the batch are
        // clipped by the tensor map).
        uint8_t* stg = scratch + STG_OFF + (half * 4 + quarter) * 8192;
        static_assert(H || NH == 1, "the fp32 store path below assumes one warp per lane quarter");
        const int wrow0 = (int)((int64_t)t.mb * 256 + rank * 128 + quarter * 32);
        auto store_chunk = [&](const float (&v)[32], int c, int buf) {
            if (lane == 0) tma_store_wait_read<1>();             // the group that last used this tile has read it
            __syncwarp();
            float* tile = reinterpret_cast<float*>(stg + buf * 4096);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
                *reinterpret_cast<float4*>(tile + lane * 32 + ((j4 ^ (lane & 7)) << 2)) =
                    make_float4(v[4 * j4] * inv, v[4 * j4 + 1] * inv, v[4 * j4 + 2] * inv, v[4 * j4 + 3] * inv);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&p.q_map, tile, c, wrow0);
                tma_store_commit();
            }
        };
        tmem_ld32(tmem + c0, va);
#pragma unroll 1
        for (int c = c0; c < c0 + CW; c += 64) {
            tmem_ld_wait();
            tmem_ld32(tmem + c + 32, vb);
            store_chunk(va, c, 0);
            tmem_ld_wait();
            if (c + 64 < c0 + CW) tmem_ld32(tmem + c + 64, va);
            store_chunk(vb, c + 32, 1);
        }
        PVS_TPHASE(10, tp3, tme);
    }
};
using PostPairPolicy = PostPairT<false>;
using Post16PairPolicy = PostPairT<true>;
}  // namespace tc2
namespace tc {

// one-off tf32 split of a weight matrix (model creation)
__global__ void split_kernel(const float4* __restrict__ x, int64_t n4, float4* __restrict__ hi, float4* __restrict__ lo)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 h, l;
        split4(x[i], h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
using namespace tc;

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
    split_kernel<<<64, 256>>>((const float4*)src, (int64_t)(n / 4), (float4*)buf, (float4*)(buf + n));
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(buf); return fail(PVS_ERR_CUDA, "weight split failed: %s", cudaGetErrorString(e)); }
    m->tc0 = buf;
    m->tc1 = buf + n;
    if (m->kind == PVS_MODEL_GMM_DIAG && m->h_ok) {
        // fp16x2 weights: interleaved (-P/2 * 4^e, mu P * 2^e) columns, split into fp16 hi + lo on the host
        std::vector<float> w(n);
        if (cudaMemcpy(w.data(), m->wcat, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
            return fail(PVS_ERR_CUDA, "weight download failed");
        std::vector<__half> hl(2 * n);
        const float sq = ldexpf(1.f, 2 * m->h_exp), sl = ldexpf(1.f, m->h_exp);
        bool ok = true;
        for (int j = 0; j < FV_K; ++j)
            for (int i = 0; i < FV_2D; ++i) {
                const int dd = i >> 1;
                const float x = (i & 1) ? w[(size_t)j * FV_2D + FV_D + dd] * sl : w[(size_t)j * FV_2D + dd] * sq;
                const __half h = __float2half_rn(x);
                hl[(size_t)j * FV_2D + i] = h;
                hl[n + (size_t)j * FV_2D + i] = __float2half_rn(x - __half2float(h));
                ok = ok && fabsf(x) < 30000.f;
            }
        void* hb = nullptr;
        if (ok) {
            PVS_CUDA(cudaMalloc(&hb, 2 * n * sizeof(__half)));
            if (cudaMemcpy(hb, hl.data(), 2 * n * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFree(hb);
                return fail(PVS_ERR_CUDA, "fp16 weight upload failed");
            }
            m->th0 = hb;
            m->th1 = (const __half*)hb + n;
        } else {
            m->h_ok = false;
        }
    } else if (m->kind == PVS_MODEL_GMM_DIAG) {
        m->h_ok = false;
    }
    return PVS_OK;
}

bool tc_fv_supported(const pvs_model* g, const pvs_model* pca, int64_t rows, int64_t n_images)
{
    if (!tc_available() || !g->tc0 || g->k != FV_K || g->d != FV_D) return false;
    if (pca && !pca->tc0) return false;
    return rows < 2147483000LL && n_images < 2147483647LL;      // TMA coordinates are 32-bit
}

int tc_fv_plan(const pvs_model* g, const pvs_model* pca, int64_t rows, int64_t n_images, void* ws, TcFvPlan* pl)
{
    size_t off = 0;
    char* base = (char*)ws;
    auto take = [&](size_t bytes) { void* p = base ? base + off : nullptr; off += align_up(bytes ? bytes : 16, 1024); return p; };
    pl->n_tiles = (int)ceil_div(rows, 128);
    pl->rows = rows;
    pl->y = pca ? (float*)take((size_t)rows * FV_D * 4) : nullptr;
    pl->q = (float*)take((size_t)rows * FV_K * 4);
    pl->S = (float*)take((size_t)n_images * FV_K * FV_2D * 4);
    pl->rinv = (float*)take(((size_t)rows + 16) * 16);
    pl->s0part = (float*)take((size_t)n_images * TC_FV_S0_PARTS * FV_K * 4 + 16);
    pl->flag = pl->s0part ? (int*)(pl->s0part + (size_t)n_images * TC_FV_S0_PARTS * FV_K) : nullptr;
    pl->fp16x2 = pca && g->h_ok && g->th0 && !getenv("PVS_FV_NO_FP16X2");
    pl->total = off + 1024;
    return PVS_OK;
}

int tc_fv_begin(const TcFvPlan& pl, int64_t n_images, cudaStream_t st)
{
    // a producer group only writes its zeroth-order partial when it owned a k-block of the image
    PVS_CUDA(cudaMemsetAsync(pl.s0part, 0, (size_t)n_images * TC_FV_S0_PARTS * FV_K * 4 + 16, st));
    return PVS_OK;
}

int tc_fv_project(const TcFvPlan& pl, const pvs_model* g, const pvs_model* pca, const float* desc, int64_t rows, cudaStream_t st)
{
    if (rows <= 0 || !pca) return PVS_OK;
    PVS_CHECK(((uintptr_t)desc & 15) == 0, PVS_ERR_BAD_ARG, "descriptor buffer must be 16-byte aligned");
    PcaParams p{};
    int rc;
    if ((rc = make_tmap_2d(&p.c_hi, pca->tc0, false, pca->d, pca->d_in, pca->d_in, 32, FV_D))) return rc;
    if ((rc = make_tmap_2d(&p.c_lo, pca->tc1, false, pca->d, pca->d_in, pca->d_in, 32, FV_D))) return rc;
    p.x = desc; p.bias = pca->bias; p.y = pl.y; p.rows = rows;
    p.m_blocks = pl.n_tiles; p.nkb = pca->d_in / 32; p.d_in = pca->d_in;
    p.flag = pl.fp16x2 ? pl.flag : nullptr;
    p.lim = ldexpf(255.f, g->h_exp);                          // (y 2^-e)^2 must stay below 65504
    if (pca->d_in == 128) {                                   // CTA pairs, C resident (32 rows per CTA)
        if ((rc = make_tmap_2d(&p.c_hi, pca->tc0, false, pca->d, pca->d_in, pca->d_in, 32, FV_D / 2))) return rc;
        if ((rc = make_tmap_2d(&p.c_lo, pca->tc1, false, pca->d, pca->d_in, pca->d_in, 32, FV_D / 2))) return rc;
        if ((rc = make_tmap_2d(&p.y_map, pl.y, false, rows, FV_D, FV_D, 32, 32))) return rc;
        p.m_blocks = (int)ceil_div(rows, 256);
        return tc2::launch_tc2<tc2::PcaPairPolicy>(p, p.m_blocks, st);
    }
    return launch_tc<PcaPolicy>(p, p.m_blocks, st);
}

int tc_fv_posterior(const TcFvPlan& pl, const pvs_model* g, const float* y, int64_t rows, int32_t* argmax, cudaStream_t st,
                    bool fallback_only)
{
    if (rows <= 0) return PVS_OK;
    PVS_CHECK(((uintptr_t)y & 15) == 0, PVS_ERR_BAD_ARG, "descriptor buffer must be 16-byte aligned");
    PostParams p{};
    int rc;
    // the interleaved (-P/2, mu P) hi / lo copies of pvs_tc_gemmnt.cu (row pitch 2D = 128 here)
    PVS_CHECK(g->tcg0 && g->tcg_ld == FV_2D, PVS_ERR_BAD_ARG, "GMM model lacks the interleaved weight copy");
    if ((rc = make_tmap_2d(&p.w_hi, g->tcg0, false, FV_K, FV_2D, FV_2D, 32, FV_K / 2))) return rc;
    if ((rc = make_tmap_2d(&p.w_lo, g->tcg1, false, FV_K, FV_2D, FV_2D, 32, FV_K / 2))) return rc;
    if ((rc = make_tmap_2d(&p.q_map, pl.q, false, rows, FV_K, FV_K, 32, 32))) return rc;
    p.y = y; p.cst = g->cst; p.q = pl.q; p.argmax = argmax; p.rows = rows;
    p.m_blocks = (int)ceil_div(rows, 256);                    // CTA pairs: 256-row tiles, W resident
    if (!pl.fp16x2) return tc2::launch_tc2<tc2::PostPairPolicy>(p, p.m_blocks, st);
    // fp16x2 kernel (writes Q as fp16 hi / lo planes for the statistics kernel), and behind it the 3xTF32
    // kernel (fp32 Q in the same buffer) that only runs when the projection raised the range flag
    PostParams h = p;
    h.flag = pl.flag; h.sc_y = ldexpf(1.f, -g->h_exp); h.planes = 1; h.rinv = pl.rinv;
    if ((rc = make_tmap_2d(&h.w_hi, g->th0, true, FV_K, FV_2D, FV_2D, 64, FV_K / 2))) return rc;
    if ((rc = make_tmap_2d(&h.w_lo, g->th1, true, FV_K, FV_2D, FV_2D, 64, FV_K / 2))) return rc;
    if ((rc = make_tmap_2d(&h.qh_map, pl.q, true, rows, FV_K, FV_K, 64, 32))) return rc;
    if ((rc = make_tmap_2d(&h.ql_map, (const __half*)pl.q + (size_t)rows * FV_K, true, rows, FV_K, FV_K, 64, 32))) return rc;
    if (!fallback_only && (rc = tc2::launch_tc2<tc2::Post16PairPolicy>(h, h.m_blocks, st))) return rc;
    p.flag = pl.flag;
    return tc2::launch_tc2<tc2::PostPairPolicy>(p, p.m_blocks, st);
}

// segments per image slot of the statistics kernels: max over the images of ceil(k-blocks / segk), at least 1, for two
// segment lengths at once (smax[0] for segk_a, smax[1] for segk_b)
__global__ void fv_smax_kernel(const int64_t* __restrict__ offsets, int64_t n_images, int segk_a, int segk_b, int* __restrict__ smax)
{
    int ma = 1, mb = 1;
    for (int64_t i = threadIdx.x; i < n_images; i += blockDim.x) {
        const int t = (int)(offsets[i + 1] - offsets[i]);
        const int nkb = (t + ST_KT - 1) / ST_KT;
        ma = max(ma, (nkb + segk_a - 1) / segk_a);
        mb = max(mb, (nkb + segk_b - 1) / segk_b);
    }
    ma = __reduce_max_sync(0xffffffffu, ma);
    mb = __reduce_max_sync(0xffffffffu, mb);
    __shared__ int part[2][32];
    if ((threadIdx.x & 31) == 0) { part[0][threadIdx.x >> 5] = ma; part[1][threadIdx.x >> 5] = mb; }
    __syncthreads();
    if (threadIdx.x < 32) {
        ma = threadIdx.x < (blockDim.x >> 5) ? part[0][threadIdx.x] : 1;
        mb = threadIdx.x < (blockDim.x >> 5) ? part[1][threadIdx.x] : 1;
        ma = __reduce_max_sync(0xffffffffu, ma);
        mb = __reduce_max_sync(0xffffffffu, mb);
        if (threadIdx.x == 0) { smax[0] = ma; smax[1] = mb; }
    }
}

int tc_fv_stats(const TcFvPlan& pl, const pvs_model* g, const float* y, const int64_t* offsets, int64_t n_images, cudaStream_t st,
                bool fallback_only)
{
    if (n_images <= 0) return PVS_OK;
    StatsParams p{};
    p.y = y; p.q = pl.q; p.offsets = offsets; p.S = pl.S; p.s0part = pl.s0part; p.n_images = n_images;
    int seg = 2;                                               // 128-descriptor tiles per segment (table in DESIGN.md)
    if (const char* e = getenv("PVS_FV_SEG")) { const int v = atoi(e); if (v >= 1) seg = v; }
    // fp16x2: K = 16 per MMA, `seg` tiles of 128 descriptors per segment; 3xTF32: K = 8 per MMA, i.e. twice the accumulation
    // steps per descriptor, so half the descriptors per segment.  Two ints behind the range flag hold the segment counts.
    p.segk = seg * (128 / ST_KT);
    p.smax = pl.flag + 1;
    StatsParams pt = p;
    pt.segk = p.segk > 1 ? p.segk / 2 : 1;
    pt.smax = pl.flag + 2;
    fv_smax_kernel<<<1, 1024, 0, st>>>(offsets, n_images, p.segk, pt.segk, pl.flag + 1);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!pl.fp16x2) return launch_tc<StatsPolicy>(pt, (int)n_images, st);
    // fp16x2 kernel, and behind it the 3xTF32 kernel that only runs when the range flag was raised
    Stats16Params h{};
    h.b = p; h.flag = pl.flag; h.rinv = pl.rinv;
    h.sc_y = ldexpf(1.f, -g->h_exp); h.un1 = ldexpf(1.f, g->h_exp - 14); h.un2 = ldexpf(1.f, 2 * g->h_exp - 14);
    h.red = 1;
    if (const char* e = getenv("PVS_FV_RED")) h.red = atoi(e) != 0;
    const int64_t rows = pl.rows;
    int rc;
    if ((rc = make_tmap_2d(&h.qh_map, pl.q, true, rows, FV_K, FV_K, 64, Stats16Policy::KT))) return rc;
    if ((rc = make_tmap_2d(&h.ql_map, (const __half*)pl.q + (size_t)rows * FV_K, true, rows, FV_K, FV_K, 64, Stats16Policy::KT))) return rc;
    if ((rc = make_tmap_2d(&h.y_map, y, false, rows, FV_D, FV_D, 32, Stats16Policy::KT))) return rc;
    if (!fallback_only && (rc = launch_tc<Stats16Policy>(h, (int)n_images, st))) return rc;
    h.b = pt;                                                  // the gated 3xTF32 kernel walks its own (shorter) segments
    return launch_tc<StatsGatedPolicy>(h, (int)n_images, st);
}

// statistics for any D (K = 256) on tensor cores; S rows have pitch 2d, s0 comes as partials
bool tc_fv_stats_generic_supported(const pvs_model* g, int64_t n_images)
{
    return tc_available() && g->k == FV_K && g->d >= 1 && n_images * ((2 * g->d + 127) / 128) < 2147483000LL;
}

int tc_fv_stats_generic(const float* q, const float* y, int d, const int64_t* offsets, int64_t n_images, float* S,
                        float* s0part, int* smax_dev, cudaStream_t st)
{
    if (n_images <= 0) return PVS_OK;
    PVS_CHECK((((uintptr_t)q) & 15) == 0, PVS_ERR_BAD_ARG, "posterior buffer must be 16-byte aligned");
    StatsGenParams p{};
    p.y = y; p.q = q; p.offsets = offsets; p.S = S; p.s0part = s0part; p.n_images = n_images;
    p.d = d; p.n_mt = (2 * d + 127) / 128;
    int seg = 2;                                               // 128-descriptor tiles per statistics segment
    if (const char* e = getenv("PVS_FV_SEG")) { const int v = atoi(e); if (v >= 1) seg = v; }
    p.segk = seg * (128 / ST_KT) > 1 ? seg * (128 / ST_KT) / 2 : 1;   // 3xTF32: K = 8 per MMA, half the descriptors per segment
    p.smax = smax_dev;
    PVS_CUDA(cudaMemsetAsync(s0part, 0, (size_t)n_images * TC_FV_S0_PARTS * FV_K * 4, st));
    fv_smax_kernel<<<1, 1024, 0, st>>>(offsets, n_images, p.segk, p.segk, smax_dev);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return launch_tc<StatsGenPolicy>(p, (int)(n_images * p.n_mt), st);
}

}  // namespace pvs
