// pvs_tc_fvfused.cu -- Fisher vector, K = 256 / D = 64: GMM posterior AND per-image statistics
// in ONE kernel, so that the posteriors (1 KB per descriptor, 60 % of the HBM bytes of the
// unfused path) never leave the SM.  fp16x2 operands (pvs_tc.cuh), one CTA per SM,
// persistent over images; tile = 128 descriptors of one image.
//
//   A1   [128 t x 128]  interleaved (y'^2, y') fp16 hi + lo, built once per tile by four converter
//        warps.  The same bytes serve TWO MMAs: as the K-major A operand of the logit MMA
//        (rows = descriptors) and, through an MN-major descriptor, as the A operand of the
//        statistics MMA (rows = contraction index t): the canonical 128-byte-swizzled layouts of
//        the two are transposes of each other.
//   MMA1 L[128 t x 256 k] = A1 . W'^T in two chunks of 128 components (N = 64 chunks ran the tensor
//        pipe at half rate: ~64 cycles per tcgen05.mma whatever N <= 128); W' (128 KB) does not fit
//        beside the rest, so it streams from L2 through a three-stage TMA ring (never waited for).
//   softmax: eight warps (two per TMEM lane quarter, column halves); max, exp stashed back
//        into the accumulator, sum; then q = e / sum scaled by 2^14 is split into fp16 hi + lo
//        and written as the MN-major B operand of
//   MMA2 S[128 (y'^2, y') x 256 k] += A1^T . Q, chunk by chunk (one Q buffer, used twice per tile), accumulated in
//        TMEM over all tiles of the image; the zeroth-order sums are column sums of the Q
//        chunks, taken by the converter warps while the MMAs run.
//   image end: S / T (operand scales undone) and the zeroth-order partials go to the same
//        buffers the unfused kernels fill; fv_finalize is unchanged.
// TMEM: 256 columns of logits + 256 of statistics = all 512; shared memory: A1 64 KB, W' ring
// 96 KB, Q buffer 64 KB.  Everything is single-buffered per tile (TMEM and shared memory are
// full), so the phases of a tile run back to back and the kernel is bound by that chain, not
// by HBM (it reads 256 B per descriptor).  Measured (role timing, -DPVS_TIMING): per 128-descriptor
// tile the tensor pipe is busy 6.9 k cycles (96 MMAs of N = 128 at their 64-cycle floor) but waits
// 8.9 k cycles for the softmax warps and 1.8 k for the converters, 17.8 k in all = 9.4 ms for the
// C2 batch against 8.2 ms for the two unfused HBM-bound kernels.  With the logits and the
// statistics filling TMEM there is no second accumulator to overlap softmax(i) with MMA(i+1);
// the way forward is a cluster that splits the 256 components.  Kept as
// an opt-in path (PVS_FV_FUSED=1) with its own parity test.
// Replaces predict_proba + the two statistics GEMMs of pyvisim/encoders/fisher_vector.py:99-104.
#include "pvs_tc.cuh"
#include "pvs_kernels.cuh"
#include <string.h>

namespace pvs {
namespace tc {

namespace fused {
#ifdef PVS_TIMING
// cycles summed over CTAs: MMA warp [0] wait a1_full [1] wait l_free [2] wait w_full [3] wait q_full [4] total;
// softmax warp 2 [5] wait l_full [6] pass 1 [7] pass 2 [8] pass 3 (incl. q_empty waits) [9] q_empty waits [10] image end [11] total
__device__ unsigned long long g_ft[16];
#define FT0(v) const long long v = clock64()
#define FTA(slot, v) ft[slot] += clock64() - (v)
#else
#define FT0(v)
#define FTA(slot, v)
#endif
// N = 128 per MMA: tcgen05.mma has a floor of ~64 cycles per instruction, so N = 64 chunks ran the tensor pipe at half rate
constexpr int K = 256, D = 64, AUG = 128, TT = 128, NCH = 2, CH = 128;
// W' stage = one k-block (64 operand columns) of one 128-component chunk, hi + lo: three stages cover the L2 latency
constexpr int A1_BYTES = 65536, W_STAGE = 32768, W_STAGES = 3, Q_BUF = 65536;
constexpr int OFF_A1 = 0, OFF_W = OFF_A1 + A1_BYTES, OFF_Q = OFF_W + W_STAGES * W_STAGE, OFF_MISC = OFF_Q + Q_BUF;
constexpr int OFF_XCH = OFF_MISC + 256, SMEM_BYTES = OFF_XCH + 2048;      // dynamic shared memory is declared 1024-aligned
constexpr int THREADS = 448;               // warp 0 TMA, warp 1 MMA, warps 2-9 softmax, warps 10-13 converters
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");

struct Params {
    CUtensorMap w_hi, w_lo;                // W' [256, 128] fp16 (interleaved (-P/2 4^e, mu P 2^e)), box 64 cols x 128 rows
    float cst[K];                          // per-component constants, read through the constant bank (uniform index)
    const float* y;                        // [rows, 64]
    const int64_t* offsets;
    float* S;                              // [n_images, 256, 128]: [k][ s1 (64) | s2 (64) ], already / T
    float* s0part;                         // [n_images, 16, 256] raw zeroth-order partial sums (slots 0-3 used)
    int64_t n_images;
    const int* flag;                       // != 0: fp16x2 range exceeded -> this kernel does nothing
    float sc_y, un1, un2;                  // 2^-e, 2^(e-14), 2^(2e-14)
};

__global__ void __launch_bounds__(THREADS, 1) kernel(const __grid_constant__ Params p)
{
    if (*p.flag != 0) return;                                  // uniform over the grid
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();               // the swizzled tiles need 1024-byte alignment
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
    uint64_t *a1_full = bars, *a1_free = bars + 1, *l_full = bars + 2, *l_free = bars + 3, *s_full = bars + 4;
    uint64_t *w_full = bars + 5, *w_empty = bars + 8, *q_full = bars + 11, *q_empty = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        mbar_init(a1_full, 4);
        mbar_init(a1_free, 1);
        mbar_init(l_full, 1);
        mbar_init(l_free, 8);
        mbar_init(s_full, 1);
        for (int i = 0; i < W_STAGES; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
        }
        mbar_init(q_full, 4);
        mbar_init(q_empty, 1 + 4);
        fence_barrier_init();
        tma_prefetch_desc(&p.w_hi);
        tma_prefetch_desc(&p.w_lo);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_L = tmem, tmem_S = tmem + 256;

    auto n_tiles_of = [&](int64_t img, int64_t& r0, int& T) {
        r0 = p.offsets[img];
        T = (int)(p.offsets[img + 1] - r0);
        return (T + TT - 1) / TT;
    };

    if (warp == 0) {
        // ---- TMA: W' chunks, in the order the logit MMAs consume them ----
        if (lane == 0) {
            int ws = 0;
            uint32_t wph = 0;
            for (int64_t img = blockIdx.x; img < p.n_images; img += gridDim.x) {
                int64_t r0; int T;
                const int nt = n_tiles_of(img, r0, T);
                for (int tile = 0; tile < nt; ++tile)
                    for (int n = 0; n < NCH; ++n)
                        for (int kb = 0; kb < 2; ++kb) {
                            mbar_wait(&w_empty[ws], wph ^ 1);
                            mbar_expect_tx(&w_full[ws], W_STAGE);
                            uint8_t* st = smem + OFF_W + ws * W_STAGE;
                            tma_load_2d(st, &p.w_hi, &w_full[ws], kb * 64, n * CH);
                            tma_load_2d(st + 16384, &p.w_lo, &w_full[ws], kb * 64, n * CH);
                            if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        constexpr uint32_t idesc1 = make_idesc(false, false, false, 128, CH, true);     // K-major x K-major
        constexpr uint32_t idesc2 = make_idesc(false, true, true, 128, CH, true);       // MN-major x MN-major
        const uint32_t a1 = smem_u32(smem + OFF_A1);
        int ws = 0;
        uint32_t wph = 0, g = 0, imgs = 0;
#ifdef PVS_TIMING
        long long ft[16] = {0};
#endif
        FT0(t_all);
        for (int64_t img = blockIdx.x; img < p.n_images; img += gridDim.x) {
            int64_t r0; int T;
            const int nt = n_tiles_of(img, r0, T);
            for (int tile = 0; tile < nt; ++tile, ++g) {
                FT0(t0);
                mbar_wait(a1_full, g & 1);
                FTA(0, t0);
                FT0(t1);
                mbar_wait(l_free, (g & 1) ^ 1);                // the softmax warps have left the previous logits
                FTA(1, t1);
                tcgen05_fence_after();
                for (int n = 0; n < NCH; ++n)
                    for (int kb = 0; kb < 2; ++kb) {
                        FT0(t2);
                        mbar_wait(&w_full[ws], wph);
                        FTA(2, t2);
                        tcgen05_fence_after();
                        const uint32_t wst = smem_u32(smem + OFF_W + ws * W_STAGE);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t a_hi = make_smem_desc(a1 + kb * 32768 + ks * 32, 16, 1024, LAYOUT_SW128);
                                const uint64_t a_lo = make_smem_desc(a1 + kb * 32768 + 16384 + ks * 32, 16, 1024, LAYOUT_SW128);
                                const uint64_t b_hi = make_smem_desc(wst + ks * 32, 16, 1024, LAYOUT_SW128);
                                const uint64_t b_lo = make_smem_desc(wst + 16384 + ks * 32, 16, 1024, LAYOUT_SW128);
                                const uint32_t d = tmem_L + (uint32_t)(n * CH);
                                umma<true>(d, a_hi, b_lo, idesc1, (kb | ks) ? 1u : 0u);
                                umma<true>(d, a_lo, b_hi, idesc1, 1u);
                                umma<true>(d, a_hi, b_hi, idesc1, 1u);
                            }
                            umma_commit(&w_empty[ws]);
                        }
                        __syncwarp();
                        if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
                    }
                if (elect_one()) umma_commit(l_full);
                __syncwarp();
                for (int n = 0; n < NCH; ++n) {
                    const uint32_t use = 2 * g + (uint32_t)n;
                    FT0(t3);
                    mbar_wait(q_full, use & 1);
                    FTA(3, t3);
                    tcgen05_fence_after();
                    const uint32_t qb = smem_u32(smem + OFF_Q);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < TT / 16; ++ks) {
                            // A1 read as MN-major: 64-column blocks 32 KB apart (the two k-blocks of MMA1),
                            // 8-row groups 1 KB apart, a k-step = 16 descriptors = 2 KB; Q: two 64-component blocks 16 KB apart
                            const uint64_t a_hi = make_smem_desc(a1 + ks * 2048, 32768, 1024, LAYOUT_SW128);
                            const uint64_t a_lo = make_smem_desc(a1 + 16384 + ks * 2048, 32768, 1024, LAYOUT_SW128);
                            const uint64_t b_hi = make_smem_desc(qb + ks * 2048, 16384, 1024, LAYOUT_SW128);
                            const uint64_t b_lo = make_smem_desc(qb + 32768 + ks * 2048, 16384, 1024, LAYOUT_SW128);
                            const uint32_t d = tmem_S + (uint32_t)(n * CH);
                            umma<true>(d, a_hi, b_lo, idesc2, (tile | ks) ? 1u : 0u);
                            umma<true>(d, a_lo, b_hi, idesc2, 1u);
                            umma<true>(d, a_hi, b_hi, idesc2, 1u);
                        }
                        umma_commit(q_empty);
                    }
                    __syncwarp();
                }
                if (elect_one()) {
                    umma_commit(a1_free);
                    if (tile == nt - 1) umma_commit(s_full);
                }
                __syncwarp();
            }
            if (nt > 0) ++imgs;
        }
        (void)imgs;
#ifdef PVS_TIMING
        FTA(4, t_all);
        if (lane == 0) for (int i = 0; i < 5; ++i) atomicAdd(&g_ft[i], (unsigned long long)ft[i]);
#endif
    } else if (warp < 10) {
        // ---- softmax / epilogue: thread = descriptor row, warp pair (quarter, half) splits the 256 components ----
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int c0 = half * 128;
        const float* cstv = p.cst;
        float* xch = reinterpret_cast<float*>(smem + OFF_XCH) + quarter * 128;     // [2 halves][max, sum][32 rows]
        auto pair_barrier = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory"); };
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const int trow = quarter * 32 + lane;                  // row of the tile = TMEM lane
        constexpr float LOG2E = 1.4426950408889634f;
        uint32_t g = 0, imgs = 0;
#ifdef PVS_TIMING
        long long ft[16] = {0};
#endif
        FT0(t_all);
        for (int64_t img = blockIdx.x; img < p.n_images; img += gridDim.x) {
            int64_t r0; int T;
            const int nt = n_tiles_of(img, r0, T);
            for (int tile = 0; tile < nt; ++tile, ++g) {
                const bool valid = tile * TT + trow < T;
                FT0(t5);
                mbar_wait(l_full, g & 1);
                FTA(5, t5);
                FT0(t6);
                tcgen05_fence_after();
                const uint32_t tl = tmem_L + lane_off;
                float va[32], vb[32];
                auto addc = [&](float (&v)[32], int c) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += cstv[c + j];
                };
                // pass 1: maximum of this warp's 128 columns
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                tmem_ld32(tl + c0, va);
#pragma unroll 1
                for (int c = c0; c < c0 + 128; c += 64) {
                    tmem_ld_wait();
                    tmem_ld32(tl + c + 32, vb);
                    addc(va, c);
#pragma unroll
                    for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], va[j]);
                    tmem_ld_wait();
                    if (c + 64 < c0 + 128) tmem_ld32(tl + c + 64, va);
                    addc(vb, c + 32);
#pragma unroll
                    for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], vb[j]);
                }
                float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                xch[half * 64 + lane] = mx;
                pair_barrier();
                mx = fmaxf(mx, xch[(half ^ 1) * 64 + lane]);
                const float base = (mx > -INFINITY && mx < INFINITY) ? mx : 0.f;
                const float nb = -base * LOG2E;
                FTA(6, t6);
                FT0(t7);
                // pass 2: e = exp(l - max), stashed back into the accumulator columns, and its sum
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
                auto exp_chunk = [&](float (&v)[32]) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float e;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v[j], LOG2E, nb)));
                        v[j] = e;
                        s4[j & 3] += e;
                    }
                };
                tmem_ld32(tl + c0, va);
#pragma unroll 1
                for (int c = c0; c < c0 + 128; c += 64) {
                    tmem_ld_wait();
                    tmem_ld32(tl + c + 32, vb);
                    addc(va, c);
                    exp_chunk(va);
                    tmem_st32(tl + c, va);
                    tmem_ld_wait();
                    if (c + 64 < c0 + 128) tmem_ld32(tl + c + 64, va);
                    addc(vb, c + 32);
                    exp_chunk(vb);
                    tmem_st32(tl + c + 32, vb);
                }
                const float spart = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                xch[half * 64 + 32 + lane] = spart;
                tmem_st_wait();
                pair_barrier();
                const float s_lo = half == 0 ? spart : xch[32 + lane], s_hi = half == 0 ? xch[64 + 32 + lane] : spart;
                // rows past the end of the image contribute nothing (their A1 rows are zero as well)
                const float sc = valid ? 16384.f / (s_lo + s_hi) : 0.f;
                FTA(7, t7);
                FT0(t8);
                // pass 3: q 2^14 as fp16 hi + lo rows of the MN-major operand of the statistics MMA
                {
                    // this warp's 128 components = chunk `half` of the statistics MMA (the Q buffer is used twice per tile)
                    const uint32_t use = 2 * g + (uint32_t)half;
                    uint8_t* qh = smem + OFF_Q;
                    uint8_t* ql = qh + 32768;
                    auto conv8 = [&](const float (&v)[32], int j8, uint4& h, uint4& l) {
                        float x[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) x[u] = v[8 * j8 + u] * sc;
                        split8_h(x, h, l);
                    };
                    auto put8 = [&](int blk, int j, const uint4& h, const uint4& l) {
                        const uint32_t off = (uint32_t)(blk * 16384 + trow * 128 + ((j ^ (trow & 7)) << 4));
                        *reinterpret_cast<uint4*>(qh + off) = h;
                        *reinterpret_cast<uint4*>(ql + off) = l;
                    };
                    tmem_ld32(tl + c0, va);
                    tmem_ld32(tl + c0 + 32, vb);
                    tmem_ld_wait();
                    // the first 32 components are converted before the wait for the buffer
                    uint4 ha[4], la[4];
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) conv8(va, j8, ha[j8], la[j8]);
                    tmem_ld32(tl + c0 + 64, va);
                    FT0(t9);
                    mbar_wait(q_empty, (use & 1) ^ 1);
                    FTA(9, t9);
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) put8(0, j8, ha[j8], la[j8]);
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        uint4 h, l;
                        conv8(vb, j8, h, l);
                        put8(0, 4 + j8, h, l);
                    }
                    tmem_ld_wait();
                    tmem_ld32(tl + c0 + 96, vb);
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        uint4 h, l;
                        conv8(va, j8, h, l);
                        put8(1, j8, h, l);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        uint4 h, l;
                        conv8(vb, j8, h, l);
                        put8(1, 4 + j8, h, l);
                    }
                    fence_proxy_async();
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(q_full);
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(l_free);
                FTA(8, t8);
            }
            FT0(t10);
            // image end: statistics / T, operand scales undone, in the [k][ s1 | s2 ] layout of fv_finalize.
            // TMEM lane m = augmented column: block kb = m / 64 holds dims [32 kb, 32 kb + 32) as (y'^2, y') pairs.
            float* Simg = p.S + img * (int64_t)(K * AUG);
            const int m = trow, dd = 32 * (m >> 6) + ((m & 63) >> 1);
            const bool lin = m & 1;                             // odd: y' -> s1, even: y'^2 -> s2
            const int col = lin ? dd : D + dd;
            if (nt > 0) {
                mbar_wait(s_full, imgs & 1);
                ++imgs;
                tcgen05_fence_after();
                const float scale = (lin ? p.un1 : p.un2) / (float)T;
                const uint32_t ts = tmem_S + lane_off;
#pragma unroll 1
                for (int c = c0; c < c0 + 128; c += 32) {
                    float v[32];
                    tmem_ld32(ts + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) Simg[(int64_t)(c + j) * AUG + col] = v[j] * scale;
                }
                tcgen05_fence_before();
            } else {
                const float nanv = __int_as_float(0x7fc00000);  // T == 0 -> NaN, like the reference
                for (int c = c0; c < c0 + 128; ++c) Simg[(int64_t)c * AUG + col] = nanv;
            }
            FTA(10, t10);
        }
#ifdef PVS_TIMING
        FTA(11, t_all);
        if (warp == 2 && lane == 0) for (int i = 5; i < 12; ++i) atomicAdd(&g_ft[i], (unsigned long long)ft[i]);
#endif
    } else {
        // ---- converters: Y rows -> A1 (interleaved (y'^2, y') fp16 hi + lo); zeroth-order sums of the Q chunks ----
        const int cw = warp - 10;                              // rows [32 cw, 32 cw + 32) of the tile
        const int c = lane & 7;
        float4 yv[16];
        auto fetch = [&](int64_t r0, int T, int tile) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int t = tile * TT + cw * 32 + i * 4 + (lane >> 3);
                    yv[kb * 8 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t < T) yv[kb * 8 + i] = __ldg(reinterpret_cast<const float4*>(p.y + (r0 + t) * D + kb * 32 + c * 4));
                }
        };
        uint32_t g = 0;
        int64_t img = blockIdx.x;
        int64_t r0 = 0; int T = 0, nt = 0, tile = 0;
        // first tile to fetch
        while (img < p.n_images && (nt = n_tiles_of(img, r0, T)) == 0) img += gridDim.x;
        if (img < p.n_images) fetch(r0, T, 0);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        while (img < p.n_images) {
            mbar_wait(a1_free, (g & 1) ^ 1);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = cw * 32 + i * 4 + (lane >> 3);
                    const float4 v = yv[kb * 8 + i];
                    const float a = v.x * p.sc_y, b = v.y * p.sc_y, cc = v.z * p.sc_y, d = v.w * p.sc_y;
                    const float x[8] = {a * a, a, b * b, b, cc * cc, cc, d * d, d};
                    uint4 h, l;
                    split8_h(x, h, l);
                    const uint32_t off = (uint32_t)(kb * 32768 + row * 128 + ((c ^ (row & 7)) << 4));
                    *reinterpret_cast<uint4*>(smem + OFF_A1 + off) = h;
                    *reinterpret_cast<uint4*>(smem + OFF_A1 + 16384 + off) = l;
                }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a1_full);
            // loads of the next tile fly while this one goes through its MMAs and softmax
            const int64_t cur_img = img;
            const bool last = tile == nt - 1;
            int64_t nimg = img, nr0 = r0; int nT = T, nnt = nt, ntile = tile + 1;
            if (last) {
                nimg = img + gridDim.x;
                ntile = 0;
                while (nimg < p.n_images && (nnt = n_tiles_of(nimg, nr0, nT)) == 0) nimg += gridDim.x;
            }
            if (nimg < p.n_images) fetch(nr0, nT, ntile);
            for (int n = 0; n < NCH; ++n) {
                const uint32_t use = 2 * g + (uint32_t)n;
                mbar_wait(q_full, use & 1);
                // warp cw: column block cw & 1 (64 components = 32 pairs, one per lane), rows [64 (cw / 2), +64)
                const uint8_t* qh = smem + OFF_Q + (cw & 1) * 16384;
                const uint8_t* ql = qh + 32768;
                float ax = 0.f, ay = 0.f;
#pragma unroll 8
                for (int rr = 0; rr < 64; ++rr) {
                    const int r = (cw >> 1) * 64 + rr;
                    const uint32_t off = (uint32_t)(r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
                    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(qh + off));
                    const float2 l = __half22float2(*reinterpret_cast<const __half2*>(ql + off));
                    ax += h.x + l.x;
                    ay += h.y + l.y;
                }
                acc[2 * n] += ax;
                acc[2 * n + 1] += ay;
                __syncwarp();
                if (lane == 0) mbar_arrive(q_empty);
            }
            ++g;
            if (last) {
                // partial slot = row half; components 128 n + 64 (cw & 1) + 2 lane
                float* dst = p.s0part + (cur_img * TC_FV_S0_PARTS + (cw >> 1)) * (int64_t)K + (cw & 1) * 64;
#pragma unroll
                for (int n = 0; n < NCH; ++n) {
                    *reinterpret_cast<float2*>(dst + n * CH + 2 * lane) = make_float2(acc[2 * n] * (1.f / 16384.f), acc[2 * n + 1] * (1.f / 16384.f));
                    acc[2 * n] = acc[2 * n + 1] = 0.f;
                }
            }
            img = nimg; r0 = nr0; T = nT; nt = nnt; tile = ntile;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem);
}
}  // namespace fused
}  // namespace tc

using namespace tc;

// PVS_FV_FUSED selects how posterior + statistics run.  Default 2: the 2-CTA cluster kernel of pvs_tc_fvfused2.cu -- the
// posteriors never leave the SM (2.3 MB of DRAM traffic per image instead of 7.1), the statistics are folded in segments of
// 256 descriptors with fire-and-forget RED.ADDs: 746 k images/s on the full C2 batch, worst image 9.9e-5 from the CUDA-core
// path (8.4e-5 from fp64).  0: the posterior kernel and the statistics kernel of pvs_tc_fv.cu with the same segments (693 k
// images/s on the same box, worst image 8.9e-5); it also serves the calls that want the arg-max of the posteriors.
// 1: this file's single-CTA kernel (statistics of a whole image accumulated in the tensor core: 712 k images/s but 58 of
// the 8 189 images 1e-4 .. 2.8e-4 off; kept as a reference for the tests only).  Table in DESIGN.md.
int tc_fv_fused_mode()
{
    const char* e = getenv("PVS_FV_FUSED");
    return !e ? 2 : (e[0] == '1' ? 1 : e[0] == '0' ? 0 : 2);
}

int tc_fv_poststats_fused(const TcFvPlan& pl, const pvs_model* g, const float* y, const int64_t* offsets, int64_t n_images,
                          cudaStream_t st)
{
    if (n_images <= 0) return PVS_OK;
    static PerDeviceOnce configured;
    if (configured.need()) {
        PVS_CUDA(cudaFuncSetAttribute(fused::kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::SMEM_BYTES));
        configured.mark();
    }
    fused::Params p{};
    int rc;
    if ((rc = make_tmap_2d(&p.w_hi, g->th0, true, fused::K, fused::AUG, fused::AUG, 64, 128))) return rc;
    if ((rc = make_tmap_2d(&p.w_lo, g->th1, true, fused::K, fused::AUG, fused::AUG, 64, 128))) return rc;
    PVS_CHECK((int)g->cst_host.size() == fused::K, PVS_ERR_BAD_ARG, "GMM model lacks the host copy of its constants");
    memcpy(p.cst, g->cst_host.data(), sizeof(p.cst));
    p.y = y; p.offsets = offsets; p.S = pl.S; p.s0part = pl.s0part; p.n_images = n_images;
    p.flag = pl.flag;
    p.sc_y = ldexpf(1.f, -g->h_exp); p.un1 = ldexpf(1.f, g->h_exp - 14); p.un2 = ldexpf(1.f, 2 * g->h_exp - 14);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_images < sms ? (int)n_images : sms;
    fused::kernel<<<grid, fused::THREADS, fused::SMEM_BYTES, st>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PVS_ERR_CUDA, "fused posterior + statistics kernel launch failed: %s", cudaGetErrorString(e));
#ifdef PVS_TIMING
    if (getenv("PVS_TIMING_PRINT")) {
        cudaStreamSynchronize(st);
        unsigned long long h[16], z[16] = {0};
        cudaMemcpyFromSymbol(h, fused::g_ft, sizeof(h));
        cudaMemcpyToSymbol(fused::g_ft, z, sizeof(z));
        const double np = grid * 1e3;
        fprintf(stderr, "[fused timing] per CTA (kcycles): mma total %.0f wait a1_full %.0f l_free %.0f w_full %.0f q_full %.0f | softmax total %.0f wait l_full %.0f p1 %.0f p2 %.0f p3 %.0f (q_empty %.0f) image end %.0f\n",
                h[4] / np, h[0] / np, h[1] / np, h[2] / np, h[3] / np, h[11] / np, h[5] / np, h[6] / np, h[7] / np, h[8] / np, h[9] / np, h[10] / np);
    }
#endif
    return PVS_OK;
}

}  // namespace pvs
