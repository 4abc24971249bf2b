// pvs_tc_fvfused2.cu -- Fisher vector, K = 256 / D = 64: posterior + per-image statistics in one
// kernel, CLUSTER version.  pvs_tc_fvfused.cu showed that one CTA cannot overlap anything: the
// logits (256 columns) and the statistics (256 columns) fill TMEM.  Here a 2-CTA cluster splits
// the 256 mixture components: both CTAs walk the same 128-descriptor tiles of the same images,
// CTA r owns components [128 r, 128 r + 128).  Per CTA that leaves room for
//   TMEM   two logit accumulators (2 x 128 columns) + two statistics buffers (2 x 128 columns: segments of four tiles
//          alternate between them and are folded into S in global memory, see the MMA warp)
//   smem   W' slice RESIDENT (64 KB, no streaming), two A1 tiles (2 x 64 KB), one Q chunk buffer (32 KB)
// so the logit MMA of tile i+1 runs while the softmax warps work on tile i.  All MMAs are
// cta_group::1; what crosses the CTA boundary is two floats per descriptor (row maximum, row
// sum of exponentials), exchanged through distributed shared memory with cluster-scope mbarriers.
//
// Per CTA: warp 0 loads W' once; warp 1 issues MMA1(g) then MMA2(g-1) (software pipeline);
// warps 2-9 softmax (two per TMEM lane quarter, 64 components each): max -> exchange -> exp,
// stash, sum -> exchange -> q 2^14 as fp16 hi + lo rows of the MN-major operand of MMA2 (chunk of
// 64 components, one buffer used twice per tile); warps 10-13 convert Y rows into A1(g+1)
// (interleaved (y'^2, y') fp16 hi + lo; K-major operand of MMA1 and, read through an MN-major
// descriptor, operand of MMA2) and then take the zeroth-order sums of the Q chunks of tile g.
// Image end: S / T and the zeroth-order partials go where the unfused kernels put them.
//
// Status (B200, C2 batch): parity-green, stress-tested on the full batch, 780 k images/s for the whole FV step against
// 705-720 k with the two unfused kernels; opt-in (PVS_FV_FUSED=2) until it has seen more boxes.  Role timing per
// 128-descriptor tile and cluster: tensor pipe busy 4.8 k cycles (MMA1 N = 128: 1.5 k; MMA2 as two N = 64 chunks at the
// 64-cycle-per-instruction floor: 3.1 k), softmax chain ~7 k (accumulator -> max -> exp -> sums -> peer exchange ->
// operand rows; the two warps of a lane quarter share a scheduler), converters 0.5 k on the critical path.
// Two synchronisation hazards found on the way (both only bit in uninstrumented builds on the full batch):
//  * one Q-buffer barrier pair polled by parity: with softmax(g) overlapping MMA2(g - 1) a producer can be two uses
//    ahead of the barrier -> ring of four barriers (use % 4);
//  * one exchange barrier for all four lane quarters: a fast warp's next-tile arrivals complete the current phase
//    -> one barrier per lane quarter.
// Variants measured and dropped: two softmax teams on alternate tiles (704 threads = 80 registers: spills and the
// converter bubble dominate, 10-13 ms for the kernel); one softmax warp per lane quarter with the whole row in registers
// (a single warp per scheduler runs at ~0.3 IPC, 12 ms).  Next: shared memory for an N = 128 Q buffer and a third A1
// tile (4-CTA cluster, 32 KB W' slice per CTA).
#include "pvs_tc.cuh"
#include "pvs_kernels.cuh"
#include <string.h>
#include <type_traits>

namespace pvs {
namespace tc {
namespace fusedc {

#ifdef PVS_TIMING
// MMA warp: [0] wait a1_full [1] wait l_free [2] wait q_full [3] total; softmax warp 2: [4] wait l_full
// [5] pass 1 [6] exchange max [7] pass 2 [8] exchange sum [9] pass 3 [10] wait q_empty [11] image end [12] total
__device__ unsigned long long g_ft[16];
#define FT0(v) const long long v = clock64()
#define FTA(slot, v) ft[slot] += clock64() - (v)
#else
#define FT0(v)
#define FTA(slot, v)
#endif

constexpr int K = 256, CK = 128, D = 64, AUG = 128, TT = 128, QC = 64;
#ifndef PVS_FOLD_HALVES
#define PVS_FOLD_HALVES 1
#endif
// 1: the first warp of every lane quarter folds (it is ahead of its partner); 2: both fold 64 of the 128 columns each --
// measured slower (0.949 against 0.929 of the two-kernel path's time): the second warp is the late one of the pair
constexpr int FOLD_HALVES = PVS_FOLD_HALVES;
constexpr int A1_BYTES = 65536, W_BYTES = 65536, Q_BYTES = 32768;
constexpr int OFF_A1 = 0, OFF_W = 2 * A1_BYTES, OFF_Q = OFF_W + W_BYTES, OFF_BAR = OFF_Q + Q_BYTES;
constexpr int OFF_XI = OFF_BAR + 256;          // intra-CTA exchange: [4 quarters][32 rows], one slot per warp pair
constexpr int OFF_XO = OFF_XI + 512;           // inter-CTA exchange (written by the peer): [tile parity][shift, sum][128 rows]
constexpr int SMEM_BYTES = OFF_XO + 2048;
constexpr int THREADS = 448;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");

struct Params {
    CUtensorMap w_hi, w_lo;                // W' [256, 128] fp16, box 64 cols x 128 rows
    float cst[K];
    const float* y;                        // [rows, 64]
    const int64_t* offsets;
    float* S;                              // [n_images, 256, 128]
    float* s0part;                         // [n_images, 16, 256] (slots 0-3 used)
    int64_t n_images;
    const int* flag;
    float sc_y, un1, un2;
    int seg;                               // tiles per statistics segment
};

__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_remote_f32(uint32_t cluster_addr, float v)
{
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
__device__ __forceinline__ void arrive_remote_release(uint32_t cluster_bar_addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void wait_acquire_cluster(uint64_t* bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 6000000000LL) __trap();
    }
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) kernel(const __grid_constant__ Params p)
{
    const int SEG = p.seg;
    if (*p.flag != 0) return;                                  // uniform over the grid
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t *a1_full = bars, *a1_free = bars + 4, *l_full = bars + 6, *l_free = bars + 8;           // a1_full[buffer * 2 + k-block]
    // The one Q buffer is used twice per tile, and the softmax of tile g overlaps the statistics MMAs of tile g - 1: a
    // producer can be TWO uses ahead of the barrier it polls, and a parity wait is only unambiguous one phase ahead.
    // So use u signals / polls barrier u % 4 (with a single barrier the uninstrumented build overwrote a chunk that was
    // still being read and hung on the full C2 batch).
    uint64_t *s_full = bars + 10, *q_full = bars + 11, *q_empty = bars + 15, *w_res = bars + 19, *x_sum = bars + 20;   // q_full[4], q_empty[4], x_sum[4]
    uint64_t* s_free = bars + 24;                              // s_free[2]: the segment buffer has been folded into S; the second s_full
    uint64_t* s_full2 = bars + 26;                             // sits at bars + 26
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank(), peer = rank ^ 1;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&a1_full[i], 4);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a1_free[i], 1);
            mbar_init(&l_full[i], 1);
            mbar_init(&l_free[i], 8);
        }
        mbar_init(s_full, 1);
        mbar_init(s_full2, 1);
        mbar_init(&s_free[0], FOLD_HALVES * 4);                // the folding warps (FOLD_HALVES per lane quarter)
        mbar_init(&s_free[1], FOLD_HALVES * 4);
        for (int i = 0; i < 4; ++i) {
            mbar_init(&q_full[i], 4);
            mbar_init(&q_empty[i], 1 + 4);
        }
        mbar_init(w_res, 1);
        // One exchange barrier PER LANE QUARTER, signalled by the 32 lanes of the peer's half-0 warp of the same quarter.
        // With one barrier for all four, a fast warp's arrivals for tile g + 1 were counted into the phase of tile g while a
        // slower sibling had not delivered its rows yet (the siblings only meet again at the Q buffer): stale rows, then
        // misaligned phases and a hang at full size.
        for (int i = 0; i < 4; ++i) mbar_init(&x_sum[i], 32);
        fence_barrier_init();
        tma_prefetch_desc(&p.w_hi);
        tma_prefetch_desc(&p.w_lo);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    cluster_sync();                                            // the peer's barriers exist before anybody signals them
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    // logits: columns [0,128) and [128,256); statistics of the even / odd segments: [256,384), [384,512)
    const uint32_t tmem_S = tmem + 256;
    auto s_full_of = [&](uint32_t sb) { return sb ? s_full2 : s_full; };

    auto n_tiles_of = [&](int64_t img, int64_t& r0, int& T) {
        r0 = p.offsets[img];
        T = (int)(p.offsets[img + 1] - r0);
        return (T + TT - 1) / TT;
    };

    if (warp == 0) {
        if (lane == 0) {                                       // this CTA's 128 rows of W', both k-blocks, hi and lo
            mbar_expect_tx(w_res, W_BYTES);
            uint8_t* w = smem + OFF_W;
            tma_load_2d(w, &p.w_hi, w_res, 0, (int)rank * CK);
            tma_load_2d(w + 16384, &p.w_hi, w_res, 64, (int)rank * CK);
            tma_load_2d(w + 32768, &p.w_lo, w_res, 0, (int)rank * CK);
            tma_load_2d(w + 49152, &p.w_lo, w_res, 64, (int)rank * CK);
        }
    } else if (warp == 1) {
        // ---- MMA issuer: MMA1(g), then MMA2(g - 1) ----
        constexpr uint32_t idesc1 = make_idesc(false, false, false, 128, CK, true);     // K-major x K-major, N = 128
        constexpr uint32_t idesc2 = make_idesc(false, true, true, 128, QC, true);       // MN-major x MN-major, N = 64
        const uint32_t a1b = smem_u32(smem + OFF_A1), wb = smem_u32(smem + OFF_W), qb = smem_u32(smem + OFF_Q);
#ifdef PVS_TIMING
        long long ft[16] = {0};
#endif
        FT0(t_all);
        mbar_wait(w_res, 0);
        // Statistics accumulate in SEGMENTS of SEG tiles, alternating between two TMEM buffers; a finished segment is added to
        // the image's S rows in global memory (L2-resident) by CUDA cores with round-to-nearest fp32 adds (`fold` in the
        // softmax warps) while the next segment is multiplied into the other buffer.  tcgen05.mma adds every K = 16 slice to
        // the accumulator with truncation, a bias that grows with the number of accumulation steps (384 for a 2 000-descriptor
        // image) and that d_sigma = (S2 - 2 mu S1 + mu^2 S0) / sigma^2 - S0 amplifies: accumulated over the whole image in one
        // buffer, the worst images of the full C2 batch were 1.2e-4 (this kernel) to 2.8e-4 (the other paths) off the fp64
        // result; folded every four tiles they are at 6e-5.
        uint32_t g = 0, n_seg = 0;
        struct Prev { int tile, nt; uint32_t sg; } prev{0, 0, 0};
        bool have_prev = false;
        auto mma2 = [&](uint32_t gp, const Prev& t) {
            const uint32_t a1 = a1b + (gp & 1) * A1_BYTES;
            const uint32_t sb = t.sg & 1;
            const bool fresh = t.tile % SEG == 0, seg_last = t.tile % SEG == SEG - 1 || t.tile == t.nt - 1;
            if (fresh && t.sg >= 2) {                          // the segment that used this buffer has been folded
                mbar_wait(&s_free[sb], ((t.sg >> 1) & 1) ^ 1);
                tcgen05_fence_after();
            }
            for (int n = 0; n < 2; ++n) {
                const uint32_t use = 2 * gp + (uint32_t)n;
                FT0(t2);
                mbar_wait(&q_full[use & 3], (use >> 2) & 1);
                FTA(2, t2);
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < TT / 16; ++ks) {
                        const uint64_t a_hi = make_smem_desc(a1 + ks * 2048, 32768, 1024, LAYOUT_SW128);
                        const uint64_t a_lo = make_smem_desc(a1 + 16384 + ks * 2048, 32768, 1024, LAYOUT_SW128);
                        const uint64_t b_hi = make_smem_desc(qb + ks * 2048, 16384, 1024, LAYOUT_SW128);
                        const uint64_t b_lo = make_smem_desc(qb + 16384 + ks * 2048, 16384, 1024, LAYOUT_SW128);
                        const uint32_t d = tmem_S + sb * CK + (uint32_t)(n * QC);
                        umma<true>(d, a_hi, b_lo, idesc2, (!fresh || ks) ? 1u : 0u);
                        umma<true>(d, a_lo, b_hi, idesc2, 1u);
                        umma<true>(d, a_hi, b_hi, idesc2, 1u);
                    }
                    umma_commit(&q_empty[use & 3]);
                }
                __syncwarp();
            }
            if (elect_one()) {
                umma_commit(&a1_free[gp & 1]);
                if (seg_last) umma_commit(s_full_of(sb));
            }
            __syncwarp();
        };
        for (int64_t img = cluster_id; img < p.n_images; img += n_clusters) {
            int64_t r0; int T;
            const int nt = n_tiles_of(img, r0, T);
            for (int tile = 0; tile < nt; ++tile, ++g) {
                const uint32_t b = g & 1, ph = (g >> 1) & 1;
                FT0(t1);
                mbar_wait(&l_free[b], ph ^ 1);
                FTA(1, t1);
                for (int kb = 0; kb < 2; ++kb) {
                    FT0(t0);
                    mbar_wait(&a1_full[b * 2 + kb], ph);       // the converters publish the two k-blocks separately
                    FTA(0, t0);
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint32_t a1 = a1b + b * A1_BYTES;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t a_hi = make_smem_desc(a1 + kb * 32768 + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t a_lo = make_smem_desc(a1 + kb * 32768 + 16384 + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t b_hi = make_smem_desc(wb + kb * 16384 + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t b_lo = make_smem_desc(wb + 32768 + kb * 16384 + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint32_t d = tmem + b * CK;
                            umma<true>(d, a_hi, b_lo, idesc1, (kb | ks) ? 1u : 0u);
                            umma<true>(d, a_lo, b_hi, idesc1, 1u);
                            umma<true>(d, a_hi, b_hi, idesc1, 1u);
                        }
                        if (kb == 1) umma_commit(&l_full[b]);
                    }
                    __syncwarp();
                }
                if (have_prev) mma2(g - 1, prev);
                have_prev = true;
                if (tile % SEG == 0) ++n_seg;
                prev = Prev{tile, nt, n_seg - 1};
            }
        }
        if (have_prev) mma2(g - 1, prev);
#ifdef PVS_TIMING
        FTA(3, t_all);
        if (lane == 0 && rank == 0) for (int i = 0; i < 4; ++i) atomicAdd(&g_ft[i], (unsigned long long)ft[i]);
#endif
    } else if (warp < 10) {
        // ---- softmax / epilogue: thread = descriptor row; warp (quarter, half) takes 64 of this CTA's 128 components ----
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const int c0 = half * 64;                              // first column of this warp inside the CTA's 128
        const float* cstv = p.cst + rank * CK;
        float* xi = reinterpret_cast<float*>(smem + OFF_XI) + quarter * 32;
        float* xo = reinterpret_cast<float*>(smem + OFF_XO);
        const uint32_t xo_remote = map_to_cta(smem_u32(xo), peer);
        const uint32_t xsum_remote = map_to_cta(smem_u32(&x_sum[quarter]), peer);
        auto pair_barrier = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory"); };
        // value shared by the two warps of a lane quarter: half 1 deposits, half 0 combines and puts the result back
        auto pair_combine = [&](float v, bool is_max) {
            if (half == 1) xi[lane] = v;
            pair_barrier();
            if (half == 0) {
                const float o = xi[lane];
                v = is_max ? fmaxf(v, o) : v + o;
                xi[lane] = v;
            }
            pair_barrier();
            if (half == 1) v = xi[lane];
            return v;
        };
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const int trow = quarter * 32 + lane;
        constexpr float LOG2E = 1.4426950408889634f;
        uint32_t g = 0, n_seg = 0;
        // this thread's place in fv_finalize's [k][ s1 | s2 ] layout: accumulator lane m = row of the (y'^2, y') operand
        const int m = trow, dd = 32 * (m >> 6) + ((m & 63) >> 1);
        const bool lin = m & 1;
        const int col = lin ? dd : D + dd;
        // Fold the finished statistics segment (tile t was its last) into S: the first segment of an image stores its partial
        // sums, the others add them (fp32, round to nearest; every address belongs to one thread).  S stays in raw operand
        // units; fv_finalize applies the operand scales and 1 / T.  Done by the first warp of every lane quarter right after
        // it has published its Q chunk of the FOLLOWING tile: by then the segment's last MMAs have long finished, and the warp
        // would otherwise wait for its partner (which publishes the second chunk of a tile and runs about 2 k cycles behind).
        struct Rec { int64_t img; int tile, nt; uint32_t sg; bool valid; };
        Rec h1{0, 0, 0, 0, false}, h2{0, 0, 0, 0, false};     // tiles g - 1 and g - 2
        const int fold_c0 = FOLD_HALVES == 2 ? 64 * half : 0, fold_c1 = FOLD_HALVES == 2 ? fold_c0 + 64 : CK;
        auto fold = [&](const Rec& t) {
            if (!t.valid || !(t.tile % SEG == SEG - 1 || t.tile == t.nt - 1)) return;
            const uint32_t sb = t.sg & 1;
            mbar_wait(s_full_of(sb), (t.sg >> 1) & 1);
            tcgen05_fence_after();
            const bool first = t.tile < SEG;
            float* Sp = p.S + t.img * (int64_t)(K * AUG) + (int64_t)(rank * CK) * AUG + col;
            const uint32_t ts = tmem_S + sb * CK + lane_off;
#pragma unroll 1
            for (int c = fold_c0; c < fold_c1; c += 32) {
                float v[32];
                tmem_ld32(ts + c, v);
                tmem_ld_wait();
                if (first) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) __stcg(Sp + (c + j) * AUG, v[j]);
                } else {                                       // fire-and-forget RED.ADD at L2: nothing to wait for
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(Sp + (c + j) * AUG, v[j]);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[sb]);
        };
#ifdef PVS_TIMING
        long long ft[16] = {0};
#endif
        FT0(t_all);
        for (int64_t img = cluster_id; img < p.n_images; img += n_clusters) {
            int64_t r0; int T;
            const int nt = n_tiles_of(img, r0, T);
            for (int tile = 0; tile < nt; ++tile, ++g) {
                const uint32_t b = g & 1, ph = (g >> 1) & 1;
                const bool valid = tile * TT + trow < T;
                if (tile % SEG == 0) ++n_seg;
                const Rec cur{img, tile, nt, n_seg - 1, true};
                FT0(t4);
                mbar_wait(&l_full[b], ph);
                FTA(4, t4);
                // the segment that ended with tile g - 2 is complete (l_full(g) was committed after the statistics MMAs of tile
                // g - 2): fold it while no logits are held in registers
                FT0(t11f);
                if (half < FOLD_HALVES) fold(h2);
                FTA(11, t11f);

                FT0(t5);
                tcgen05_fence_after();
                const uint32_t tl = tmem + b * CK + lane_off + c0;
                float va[32], vb[32];
                auto addc = [&](float (&v)[32], int c) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += cstv[c + j];
                };
                // pass 1: maximum of this warp's 64 columns
                tmem_ld32(tl, va);
                tmem_ld32(tl + 32, vb);
                tmem_ld_wait();
                addc(va, c0);
                addc(vb, c0 + 32);
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], fmaxf(va[j], vb[j]));
                float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                // the logits now live in registers: hand the accumulator back to the MMA warp at once
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&l_free[b]);
                FTA(5, t5);
                FT0(t6);
                // this CTA's maximum over its 128 components (partner warp = the other 64 columns).  The peer CTA is
                // NOT consulted for the maximum: each CTA shifts by its own and the two are reconciled with the sums.
                mx = pair_combine(mx, true);
                FTA(6, t6);
                FT0(t7);
                const float base = (mx > -INFINITY && mx < INFINITY) ? mx : 0.f;
                const float nb = -base * LOG2E;
                // pass 2: e = exp(l - base) in registers (nothing goes back to TMEM), and its sum
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float e0, e1;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(va[j], LOG2E, nb)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(vb[j], LOG2E, nb)));
                    va[j] = e0;
                    vb[j] = e1;
                    s4[j & 3] += e0 + e1;
                }
                float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                FTA(7, t7);
                FT0(t8);
                sum = pair_combine(sum, false);
                // one exchange with the peer CTA per tile: (shift, sum of exponentials) of its 128 components
                const uint32_t xoff = (g & 1) * 1024u;
                if (half == 0) {
                    st_remote_f32(xo_remote + xoff + (uint32_t)trow * 4, base);
                    st_remote_f32(xo_remote + xoff + 512u + (uint32_t)trow * 4, sum);
                    arrive_remote_release(xsum_remote);
                }
                wait_acquire_cluster(&x_sum[quarter], g & 1);
                const float pbase = xo[(g & 1) * 256 + trow], psum = xo[(g & 1) * 256 + 128 + trow];
                const float big = fmaxf(base, pbase);
                float f_me, f_peer;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f_me) : "f"((base - big) * LOG2E));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f_peer) : "f"((pbase - big) * LOG2E));
                const float tot = sum * f_me + psum * f_peer;
                const float sc = valid ? 16384.f * f_me / tot : 0.f;
                FTA(8, t8);
                FT0(t9);
                // pass 3: q 2^14 as fp16 hi + lo rows of the MN-major operand of the statistics MMA (chunk `half`)
                {
                    const uint32_t use = 2 * g + (uint32_t)half;
                    uint8_t* qh = smem + OFF_Q;
                    uint8_t* ql = qh + 16384;
                    FT0(t10);
                    if (use > 0) mbar_wait(&q_empty[(use - 1) & 3], ((use - 1) >> 2) & 1);   // the previous use has been consumed
                    FTA(10, t10);
#pragma unroll
                    for (int j8 = 0; j8 < 8; ++j8) {
                        float x[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) x[u] = (j8 < 4 ? va[8 * j8 + u] : vb[8 * (j8 - 4) + u]) * sc;
                        uint4 h, l;
                        split8_h(x, h, l);
                        const uint32_t off = (uint32_t)(trow * 128 + ((j8 ^ (trow & 7)) << 4));
                        *reinterpret_cast<uint4*>(qh + off) = h;
                        *reinterpret_cast<uint4*>(ql + off) = l;
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&q_full[use & 3]);
                }
                FTA(9, t9);
                h2 = h1;
                h1 = cur;
            }
            FT0(t11);
            if (nt == 0 && half == 0) {                        // T = 0: NaN encoding, like the reference's division by zero
                float* Simg = p.S + img * (int64_t)(K * AUG) + (int64_t)(rank * CK) * AUG;
                const float nanv = __int_as_float(0x7fc00000);
                for (int c = 0; c < CK; ++c) Simg[(int64_t)c * AUG + col] = nanv;
            }
            FTA(11, t11);
        }
        if (half < FOLD_HALVES) { fold(h2); fold(h1); }
#ifdef PVS_TIMING
        FTA(12, t_all);
        if (warp == 2 && lane == 0 && rank == 0) for (int i = 4; i < 13; ++i) atomicAdd(&g_ft[i], (unsigned long long)ft[i]);
#endif
    } else {
        // ---- converters: A1(g + 1) first, then the zeroth-order sums of tile g ----
        const int cw = warp - 10;
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
        auto convert = [&](uint32_t gg) {                      // registers -> A1[gg & 1]
            mbar_wait(&a1_free[gg & 1], ((gg >> 1) & 1) ^ 1);
            uint8_t* a1 = smem + OFF_A1 + (gg & 1) * A1_BYTES;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {                   // published separately: MMA1 starts on the first k-block
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = cw * 32 + i * 4 + (lane >> 3);
                    const float4 v = yv[kb * 8 + i];
                    const float a = v.x * p.sc_y, b = v.y * p.sc_y, cc = v.z * p.sc_y, d = v.w * p.sc_y;
                    const float x[8] = {a * a, a, b * b, b, cc * cc, cc, d * d, d};
                    uint4 h, l;
                    split8_h(x, h, l);
                    const uint32_t off = (uint32_t)(kb * 32768 + row * 128 + ((c ^ (row & 7)) << 4));
                    *reinterpret_cast<uint4*>(a1 + off) = h;
                    *reinterpret_cast<uint4*>(a1 + 16384 + off) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a1_full[(gg & 1) * 2 + kb]);
            }
        };
        // walk the (image, tile) sequence one tile ahead of the tile whose Q chunks are being summed
        int64_t f_img = cluster_id, f_r0 = 0;
        int f_T = 0, f_nt = 0, f_tile = 0;
        auto f_settle = [&]() {                                // skip empty images
            while (f_img < p.n_images && (f_nt = n_tiles_of(f_img, f_r0, f_T)) == 0) f_img += n_clusters;
        };
        auto f_next = [&]() {
            if (++f_tile >= f_nt) { f_tile = 0; f_img += n_clusters; f_settle(); }
        };
        f_settle();
        uint32_t g = 0;                                        // tile whose Q chunks are summed next
        if (f_img < p.n_images) {
            fetch(f_r0, f_T, 0);
            convert(0);
        }
        int64_t s_img = f_img;
        bool s_last = f_img < p.n_images && f_nt == 1;
        bool more = f_img < p.n_images;
        if (more) {
            f_next();
            if (f_img < p.n_images) fetch(f_r0, f_T, f_tile);
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        while (more) {
            const bool have_next = f_img < p.n_images;
            if (have_next) convert(g + 1);
            const int64_t n_img = f_img;
            const bool n_last = have_next && f_tile == f_nt - 1;
            if (have_next) {
                f_next();
                if (f_img < p.n_images) fetch(f_r0, f_T, f_tile);
            }
            for (int n = 0; n < 2; ++n) {
                const uint32_t use = 2 * g + (uint32_t)n;
                mbar_wait(&q_full[use & 3], (use >> 2) & 1);
                const uint8_t* qh = smem + OFF_Q;
                const uint8_t* ql = qh + 16384;
                float ax = 0.f, ay = 0.f;
#pragma unroll 8
                for (int rr = 0; rr < 32; ++rr) {
                    const int r = cw * 32 + rr;
                    const uint32_t off = (uint32_t)(r * 128 + (((lane >> 2) ^ (r & 7)) << 4) + (lane & 3) * 4);
                    const float2 h = __half22float2(*reinterpret_cast<const __half2*>(qh + off));
                    const float2 l = __half22float2(*reinterpret_cast<const __half2*>(ql + off));
                    ax += h.x + l.x;
                    ay += h.y + l.y;
                }
                acc[2 * n] += ax;
                acc[2 * n + 1] += ay;
                __syncwarp();
                if (lane == 0) mbar_arrive(&q_empty[use & 3]);
            }
            if (s_last) {
                float* dst = p.s0part + (s_img * TC_FV_S0_PARTS + cw) * (int64_t)K + rank * CK;
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    *reinterpret_cast<float2*>(dst + n * QC + 2 * lane) = make_float2(acc[2 * n] * (1.f / 16384.f), acc[2 * n + 1] * (1.f / 16384.f));
                    acc[2 * n] = acc[2 * n + 1] = 0.f;
                }
            }
            ++g;
            more = have_next;
            s_img = n_img;
            s_last = n_last;
        }
    }
    tcgen05_fence_before();
    cluster_sync();                                            // nobody exits while the peer may still write into its shared memory
    if (warp == 1) tmem_dealloc<512>(tmem);
}
}  // namespace fusedc
}  // namespace tc

using namespace tc;

int tc_fv_poststats_fused_cluster(const TcFvPlan& pl, const pvs_model* g, const float* y, const int64_t* offsets, int64_t n_images,
                                  cudaStream_t st)
{
    if (n_images <= 0) return PVS_OK;
    static PerDeviceOnce configured;
    if (configured.need()) {
        PVS_CUDA(cudaFuncSetAttribute(fusedc::kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fusedc::SMEM_BYTES));
        configured.mark();
    }
    fusedc::Params p{};
    int rc;
    if ((rc = make_tmap_2d(&p.w_hi, g->th0, true, fusedc::K, fusedc::AUG, fusedc::AUG, 64, 128))) return rc;
    if ((rc = make_tmap_2d(&p.w_lo, g->th1, true, fusedc::K, fusedc::AUG, fusedc::AUG, 64, 128))) return rc;
    PVS_CHECK((int)g->cst_host.size() == fusedc::K, PVS_ERR_BAD_ARG, "GMM model lacks the host copy of its constants");
    memcpy(p.cst, g->cst_host.data(), sizeof(p.cst));
    p.y = y; p.offsets = offsets; p.S = pl.S; p.s0part = pl.s0part; p.n_images = n_images; p.flag = pl.flag;
    p.sc_y = ldexpf(1.f, -g->h_exp); p.un1 = ldexpf(1.f, g->h_exp - 14); p.un2 = ldexpf(1.f, 2 * g->h_exp - 14);
    p.seg = 2;                                                 // see the table in DESIGN.md: 1 -> 7.6e-5 worst image / 506 k images/s, 2 -> 9.9e-5 / 618 k, 4 -> 1.3e-4 / 664 k
    if (const char* e = getenv("PVS_FV_SEG")) { const int v = atoi(e); if (v >= 1) p.seg = v; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int clusters = n_images < sms / 2 ? (int)n_images : sms / 2;
    fusedc::kernel<<<2 * clusters, fusedc::THREADS, fusedc::SMEM_BYTES, st>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PVS_ERR_CUDA, "fused (cluster) posterior + statistics kernel launch failed: %s", cudaGetErrorString(e));
#ifdef PVS_TIMING
    if (getenv("PVS_TIMING_PRINT")) {
        cudaStreamSynchronize(st);
        unsigned long long h[16], z[16] = {0};
        cudaMemcpyFromSymbol(h, fusedc::g_ft, sizeof(h));
        cudaMemcpyToSymbol(fusedc::g_ft, z, sizeof(z));
        const double np = clusters * 1e3;
        fprintf(stderr, "[fusedc timing] per cluster (kcycles): mma total %.0f wait a1_full %.0f l_free %.0f q_full %.0f | softmax total %.0f wait l_full %.0f p1 %.0f xmax %.0f p2 %.0f xsum %.0f p3 %.0f (q_empty %.0f) image end %.0f\n",
                h[3] / np, h[0] / np, h[1] / np, h[2] / np, h[12] / np, h[4] / np, h[5] / np, h[6] / np, h[7] / np, h[8] / np, h[9] / np, h[10] / np, h[11] / np);
    }
#endif
    return PVS_OK;
}

}  // namespace pvs
