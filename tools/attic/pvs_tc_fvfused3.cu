// NOT PART OF THE LIBRARY (tools/attic): third cut of the fused posterior + statistics kernel, kept for its findings.
// Parity-green (it passed every FV test as PVS_FV_FUSED=3) but slower than pvs_tc_fvfused2.cu: 577 k images/s for the whole
// FV step against 664 k with the same segment length.  What the role timers / the event trace in here showed:
//  * two softmax teams on alternate tiles do NOT overlap as hoped: the two Q slots serialise the chunk stream of both
//    teams, a tile holds its A1 buffer from the conversion until its last statistics MMA (so conversion + logit MMAs +
//    softmax + statistics MMAs of a tile form one chain of ~21 k cycles for two tiles in flight), and the SFU is shared
//    (a warp-wide ex2 costs ~10 cycles, ~20 when the other team is in its exponential pass too);
//  * one issuing warp for MMA1(g) / MMA2(g - 1) in a fixed order couples the teams; two issuing warps fix that;
//  * a warp that polls two barriers must use mbarrier.test_wait: try_wait suspends the thread for a while when the phase is
//    not complete; a suspend-time hint on the ordinary spin loops changed nothing;
//  * stashing the exponentials over the logits (tcgen05.st) is free, recomputing them in pass 3 is not (SFU-bound);
//  * __launch_bounds__(448) gives 128 registers (4-warp granularity); __maxnreg__(144) does not launch (too many resources);
//  * the segment fold (statistics of SEG tiles -> fp32 adds) is what pvs_tc_fvfused2.cu took over from here.
// pvs_tc_fvfused3.cu -- Fisher vector, K = 256 / D = 64: posterior + per-image statistics in one kernel, third cut.
// Same decomposition as pvs_tc_fvfused2.cu (a 2-CTA cluster splits the 256 mixture components, both CTAs walk the
// same 128-descriptor tiles, W' slice resident, Q never leaves the SM), re-organised around what the role timers of
// that kernel showed (per tile 7.7 k cycles, of which 3.8 k were the two warps of a lane quarter and the two CTAs
// waiting for each other, and the statistics MMAs ran as N = 64 chunks at the 64-cycle-per-instruction floor):
//
//  * TWO SOFTMAX TEAMS.  Warps 2-5 take the even tiles, warps 6-9 the odd ones; a thread owns a descriptor row with
//    all 128 components of its CTA, so nothing is exchanged inside the CTA, and the two warps that share a scheduler
//    work on DIFFERENT tiles: one warp's latencies (accumulator loads, the exchange with the peer CTA, the wait for
//    a Q slot) are the other warp's issue slots.  The row is never held in registers: pass 1 (maximum), pass 2
//    (e = exp(l - max), stashed over the logits with tcgen05.st, and its sum) and pass 3 (q 2^14 as fp16 hi + lo)
//    each stream the 128 accumulator columns through 64 registers.
//  * Q CHUNKS ARE 32 DESCRIPTORS x 128 COMPONENTS (one per lane quarter, 16 KB with hi + lo, two slots): the
//    statistics MMAs run with N = 128 at full rate, a chunk is produced by ONE warp (no cross-warp barrier), and the
//    MMA warp consumes the four chunks of a tile in order.
//  * SEGMENTED STATISTICS.  tcgen05.mma adds every K = 16 slice to the accumulator with truncation, a bias that grows
//    with the number of accumulation steps (384 per 2 000-descriptor image) and is amplified by the cancellation in
//    d_sigma = (S2 - 2 mu S1 + mu^2 S0) / sigma^2 - S0: on the full C2 batch the worst image of the previous kernels
//    sat at 1.0e-4 (unfused) / 1.8e-4 (fused) of the fp64 result.  Here the statistics of SEG = 4 tiles accumulate in
//    one of two TMEM buffers; a finished segment is folded into the image's S rows in global memory (L2-resident,
//    fp32 round-to-nearest adds by the softmax warps, each team half of the columns) while the next segment is being
//    multiplied into the other buffer.  Bias / 4, and no drain at the end of an image on the critical path.
//
// Per CTA: warp 0 loads W' once; warp 1 issues MMA1(g) then MMA2(g - 1); warps 2-9 softmax teams; warps 10-13 convert
// Y rows into A1(g + 1) (interleaved (y'^2, y') fp16 hi + lo; K-major operand of MMA1 and, through an MN-major
// descriptor on the same bytes, operand of MMA2) and take the zeroth-order sums of the Q chunks of tile g.
// TMEM: logits of the even / odd tile (2 x 128 columns) + two statistics buffers (2 x 128 columns).
#include "pvs_tc.cuh"
#include "pvs_kernels.cuh"
#include <string.h>

namespace pvs {
namespace tc {
namespace fused3 {

#ifdef PVS_TIMING
// MMA warp: [0] wait a1_full [1] wait l_free [2] wait q_full [3] total [13] wait s_free; softmax warp 2 (team 0):
// [4] wait l_full [5] pass 1 [6] pass 2 [7] safety fold [8] exchange wait [9] pass 3 (incl. [10] wait q_empty) [11] fold [12] total
__device__ unsigned long long g_ft[16];
#define FT0(v) const long long v = clock64()
#define FTA(slot, v) ft[slot] += clock64() - (v)
// event trace of cluster 0 / CTA 0, tiles [TR0, TR0 + 48): [tile][event] = clock64
constexpr int TR0 = 64, TRN = 48;
__device__ long long g_trace[TRN][32];
#define TRACE(g, ev) do { if (blockIdx.x == 0 && (g) >= TR0 && (g) < TR0 + TRN && (threadIdx.x & 31) == 0) g_trace[(g) - TR0][ev] = clock64(); } while (0)
#else
#define FT0(v)
#define FTA(slot, v)
#define TRACE(g, ev)
#endif

constexpr int K = 256, CK = 128, D = 64, AUG = 128, TT = 128, SEG = 4;
constexpr int A1_BYTES = 65536, W_BYTES = 65536, QCH_BYTES = 16384, QPLANE = 8192, QBLK = 4096;
constexpr int OFF_A1 = 0, OFF_W = 2 * A1_BYTES, OFF_Q = OFF_W + W_BYTES, OFF_BAR = OFF_Q + 2 * QCH_BYTES;
constexpr int OFF_XO = OFF_BAR + 512;          // inter-CTA exchange (written by the peer): [team][shift, sum][128 rows]
constexpr int OFF_CST = OFF_XO + 2048;         // this CTA's 128 per-component constants
constexpr int SMEM_BYTES = OFF_CST + 512;
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
// non-blocking test (mbarrier.try_wait may suspend the thread for a while when the phase is not complete -- wrong for a
// warp that polls two barriers in turn: measured ~2.5 k cycles between a chunk's publication and its summation)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Every role walks the same sequence of (image, tile) of its cluster; g counts tiles, sg segments (a segment is up
// to SEG consecutive tiles of one image).
struct Walk {
    int64_t r0;
    int img, T, nt, tile;
    uint32_t g, sg;
    __device__ bool seg_first() const { return tile % SEG == 0; }
    __device__ bool seg_last() const { return tile % SEG == SEG - 1 || tile == nt - 1; }
    __device__ bool img_first() const { return tile < SEG; }             // first segment of its image
    __device__ bool img_last() const { return tile / SEG == (nt - 1) / SEG; }   // last segment of its image
};
template <class OnEmpty>
__device__ __forceinline__ bool walk_settle(Walk& w, const Params& p, int n_clusters, OnEmpty&& on_empty)
{
    while (w.img < p.n_images) {
        w.r0 = p.offsets[w.img];
        w.T = (int)(p.offsets[w.img + 1] - w.r0);
        w.nt = (w.T + TT - 1) / TT;
        if (w.nt > 0) {
            w.tile = 0;
            ++w.sg;
            return true;
        }
        on_empty(w.img);
        w.img += n_clusters;
    }
    return false;
}
template <class OnEmpty>
__device__ __forceinline__ bool walk_start(Walk& w, const Params& p, int cluster_id, int n_clusters, OnEmpty&& on_empty)
{
    w.img = cluster_id;
    w.g = 0;
    w.sg = 0xffffffffu;
    w.tile = 0; w.nt = 0; w.T = 0; w.r0 = 0;
    return walk_settle(w, p, n_clusters, on_empty);
}
template <class OnEmpty>
__device__ __forceinline__ bool walk_next(Walk& w, const Params& p, int n_clusters, OnEmpty&& on_empty)
{
    ++w.g;
    if (++w.tile < w.nt) {
        if (w.tile % SEG == 0) ++w.sg;
        return true;
    }
    w.img += n_clusters;
    return walk_settle(w, p, n_clusters, on_empty);
}
struct NoEmpty { __device__ void operator()(int) const {} };

// one base pointer (the register file is tight: 14 warps leave 128 registers per thread); x_sum[16] ends at 47, the TMEM
// slot sits at 48
struct Bars {
    uint64_t* b;
    __device__ explicit Bars(uint64_t* base) : b(base) {}
    __device__ uint64_t* a1_full(uint32_t i) const { return b + i; }
    __device__ uint64_t* a1_free(uint32_t i) const { return b + 4 + i; }
    __device__ uint64_t* l_full(uint32_t i) const { return b + 6 + i; }
    __device__ uint64_t* l_free(uint32_t i) const { return b + 8 + i; }
    __device__ uint64_t* q_full(uint32_t i) const { return b + 10 + i; }
    __device__ uint64_t* q_empty(uint32_t i) const { return b + 18 + i; }
    __device__ uint64_t* s_full(uint32_t i) const { return b + 26 + i; }
    __device__ uint64_t* s_free(uint32_t i) const { return b + 28 + i; }
    __device__ uint64_t* w_res() const { return b + 30; }
    __device__ uint64_t* x_sum(uint32_t i) const { return b + 31 + i; }
};

// volatile: the constants must be re-read where they are used -- hoisted out of the tile loop (which is what happens to
// plain loads, from shared memory or from the kernel parameters) 128 of them cannot stay in registers and get spilled
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ void softmax_role(const Params& p, uint8_t* smem, const Bars& B, uint32_t tmem, int warp, int lane,
                                             int cluster_id, int n_clusters, const int RANK)
{
    const int quarter = warp & 3, team = (warp - 2) >> 2;
    const int trow = quarter * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t peer = (uint32_t)RANK ^ 1u;
    float* xo = reinterpret_cast<float*>(smem + OFF_XO) + team * 256;
    const uint32_t xo_remote = map_to_cta(smem_u32(xo), peer);
    constexpr float LOG2E = 1.4426950408889634f;
    // this thread's place in fv_finalize's [k][ s1 | s2 ] layout: accumulator lane m = row of the (y'^2, y') operand
    const int m = trow, dd = 32 * (m >> 6) + ((m & 63) >> 1);
    const bool lin = m & 1;
    const int col = lin ? dd : D + dd;
    const int fc0 = team * 64;                                 // the statistics columns this team folds
    const uint32_t cst_s = smem_u32(smem + OFF_CST);
    // v[j] += constant of component c0 + j
    auto add_cst = [&](float (&v)[32], int c0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 k4 = lds_f4(cst_s + (uint32_t)(c0 + j) * 4);
            v[j] += k4.x; v[j + 1] += k4.y; v[j + 2] += k4.z; v[j + 3] += k4.w;
        }
    };
#ifdef PVS_TIMING
    long long ft[16] = {0};
#endif
    FT0(t_all);

    // fold the statistics of the segment that ends with tile f into S (global, L2-resident): the partial sums of SEG
    // tiles come out of TMEM; the first segment of an image stores them, the others add them with fire-and-forget
    // fp32 reductions (round-to-nearest; every address belongs to exactly one thread, so the order is the program
    // order and the result is deterministic).  S stays in raw operand units, fv_finalize applies un / T.
    auto fold = [&](const Walk& f) {
        const uint32_t sb = f.sg & 1;
        mbar_wait(B.s_full(sb), (f.sg >> 1) & 1);
        tcgen05_fence_after();
        float* Sp = p.S + f.img * (int64_t)(K * AUG) + (int64_t)(RANK * CK + fc0) * AUG + col;
        const bool first = f.img_first();
        const uint32_t ts = tmem + 256 + sb * CK + lane_off + fc0;
        float v0[32], v1[32];
        tmem_ld32(ts, v0);
        tmem_ld32(ts + 32, v1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.s_free(sb));             // the buffer is free as soon as its values are in registers
        if (first) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { __stcg(Sp + j * AUG, v0[j]); __stcg(Sp + (32 + j) * AUG, v1[j]); }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(Sp + j * AUG), "f"(v0[j]) : "memory");
                asm volatile("red.global.add.f32 [%0], %1;" ::"l"(Sp + (32 + j) * AUG), "f"(v1[j]) : "memory");
            }
        }
    };
    auto nan_fill = [&](int img) {                          // T = 0: NaN encoding, like the reference's division by zero
        float* Sp = p.S + img * (int64_t)(K * AUG) + (int64_t)(RANK * CK + fc0) * AUG + col;
        const float nanv = __int_as_float(0x7fc00000);
        for (int c = 0; c < 64; ++c) Sp[c * AUG] = nanv;
    };

    Walk w, f;                                                 // w: tile being processed, f: fold cursor (lags)
    bool w_ok = walk_start(w, p, cluster_id, n_clusters, NoEmpty());
    bool f_ok = walk_start(f, p, cluster_id, n_clusters, nan_fill);
    // advance the fold cursor over every tile < limit, folding the segments that end there
    auto fold_until = [&](uint32_t limit) {
        while (f_ok && f.g < limit) {
            if (f.seg_last()) fold(f);
            f_ok = walk_next(f, p, n_clusters, nan_fill);
        }
    };

    for (; w_ok; w_ok = walk_next(w, p, n_clusters, NoEmpty())) {
        const uint32_t g = w.g;
        if ((g & 1) != (uint32_t)team) continue;
        const uint32_t b = team, n = g >> 1;                   // logit buffer of this team, n-th tile of the team
        const bool valid = w.tile * TT + trow < w.T;
        FT0(t4);
        mbar_wait(B.l_full(b), n & 1);
        FTA(4, t4);
        TRACE(g, 2 + quarter);
        tcgen05_fence_after();
        const uint32_t tl = tmem + b * CK + lane_off;
        float va[32], vb[32];
        // All three passes stream the 128 accumulator columns in chunks of 32 through two register buffers: the load of
        // the next chunk is issued right after the wait for the current one and travels while the current one is
        // processed (tcgen05.wait::ld waits for every outstanding load, so a load issued BEFORE the wait would not overlap).
        // pass 1: maximum of the 128 logits of this row (constant added here: pre-loading it costs accuracy)
        FT0(t5);
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        auto max32 = [&](float (&v)[32], int c0) {
            add_cst(v, c0);
#pragma unroll
            for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], v[j]);
        };
        tmem_ld32(tl, va);
        tmem_ld32(tl + 32, vb);
        tmem_ld_wait();
        max32(va, 0);
        tmem_ld32(tl + 64, va);
        max32(vb, 32);
        tmem_ld_wait();
        tmem_ld32(tl + 96, vb);
        max32(va, 64);
        tmem_ld_wait();
        tmem_ld32(tl, va);                                     // first chunk of pass 2
        max32(vb, 96);
        const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float base = (mx > -INFINITY && mx < INFINITY) ? mx : 0.f;
        const float nb = -base * LOG2E;
        FTA(5, t5);
        // pass 2: e = exp(l - base), stashed over the logits, and its sum.  (Measured: the stash is free, the pass is bound by
        // the SFU -- ~10 cycles per warp-wide ex2, twice that while the other team is in its pass 2 as well; computing the
        // exponentials a second time in pass 3 instead of stashing them made pass 3 four times longer.)
        FT0(t6);
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        auto exp32 = [&](float (&v)[32], int c0) {
            add_cst(v, c0);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v[j], LOG2E, nb)));
                v[j] = e;
                s4[j & 3] += e;
            }
            tmem_st32(tl + c0, v);
        };
        tmem_ld_wait();
        tmem_ld32(tl + 32, vb);
        exp32(va, 0);
        tmem_ld_wait();
        tmem_ld32(tl + 64, va);
        exp32(vb, 32);
        tmem_ld_wait();
        tmem_ld32(tl + 96, vb);
        exp32(va, 64);
        tmem_ld_wait();
        exp32(vb, 96);
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        // one exchange with the peer CTA per tile: (shift, sum of exponentials) of its 128 components
        st_remote_f32(xo_remote + (uint32_t)trow * 4, base);
        st_remote_f32(xo_remote + 512u + (uint32_t)trow * 4, sum);
        uint64_t* xs = B.x_sum((quarter * 2 + team) * 2 + (n & 1));
        arrive_remote_release(map_to_cta(smem_u32(xs), peer));
        tmem_st_wait();
        FTA(6, t6);
        TRACE(g, 6 + quarter);
        // Safety fold: statistics segments that ended with tile g - 2 or earlier must be folded BEFORE the wait for a Q slot
        // below -- the statistics MMAs of tile g may be held until the buffer they want to overwrite has been folded, and
        // this warp's slot only becomes free through them.  (Those segments only depend on chunks produced before tile
        // g - 1, so waiting for them here cannot deadlock.)  Normally there is nothing left to do: the fold at the end of
        // the previous tile, off the critical path, has taken care of it.
        FT0(t7);
        if (g >= 2) fold_until(g - 1);
        FTA(7, t7);
        tmem_ld32(tl, va);                                     // first chunk of pass 3 travels during the exchange
        FT0(t8);
        wait_acquire_cluster(xs, (n >> 1) & 1);
        const float pbase = xo[trow], psum = xo[128 + trow];
        const float big = fmaxf(base, pbase);
        float f_me, f_peer;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f_me) : "f"((base - big) * LOG2E));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f_peer) : "f"((pbase - big) * LOG2E));
        const float tot = sum * f_me + psum * f_peer;
        const float sc = valid ? 16384.f * f_me / tot : 0.f;
        FTA(8, t8);
        // pass 3: q 2^14 as fp16 hi + lo rows of the MN-major operand of the statistics MMA: chunk = this quarter's 32
        // descriptors x 128 components.  The two Q slots serialise the chunk stream of both teams (produce -> statistics MMA
        // -> produce ...), so as little as possible happens between the slot becoming free and the chunk being published:
        // the first 64 components are converted into registers BEFORE the wait, the next 32 are already on their way.
        const uint32_t use = 4 * g + (uint32_t)quarter;
        uint8_t* qh = smem + OFF_Q + (use & 1) * QCH_BYTES;
        FT0(t9);
        uint4 H[8], L[8];
        auto split32 = [&](const float (&v)[32], int, uint4* h, uint4* l) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                float x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = v[8 * j8 + u] * sc;
                split8_h(x, h[j8], l[j8]);
            }
        };
        // 16-byte units u0 .. u0 + 3 of row `lane` in 64-component block blk
        auto store4 = [&](const uint4* h, const uint4* l, int blk, int u0) {
#pragma unroll
            for (int j8 = 0; j8 < 4; ++j8) {
                const uint32_t off = (uint32_t)(blk * QBLK + lane * 128 + (((u0 + j8) ^ (lane & 7)) << 4));
                *reinterpret_cast<uint4*>(qh + off) = h[j8];
                *reinterpret_cast<uint4*>(qh + QPLANE + off) = l[j8];
            }
        };
        tmem_ld_wait();
        tmem_ld32(tl + 32, vb);
        split32(va, 0, H, L);
        tmem_ld_wait();
        tmem_ld32(tl + 64, va);
        split32(vb, 32, H + 4, L + 4);
        FT0(t10);
        if (use >= 2) mbar_wait(B.q_empty((use - 2) & 7), ((use - 2) >> 3) & 1);   // the slot's previous chunk has been consumed
        FTA(10, t10);
        TRACE(g, 10 + quarter);
        store4(H, L, 0, 0);
        store4(H + 4, L + 4, 0, 4);
        tmem_ld_wait();
        tmem_ld32(tl + 96, vb);
        split32(va, 64, H, L);
        store4(H, L, 1, 0);
        tmem_ld_wait();
        tcgen05_fence_before();                                // the accumulator goes back to the MMA warp
        __syncwarp();
        if (lane == 0) mbar_arrive(B.l_free(b));
        split32(vb, 96, H, L);
        store4(H, L, 1, 4);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(B.q_full(use & 7));
        FTA(9, t9);
        TRACE(g, 14 + quarter);
        // fold what has finished in the meantime (segments that ended with tile g - 1 or earlier; may wait a moment for the
        // last statistics MMAs of tile g - 1): this team now has nothing to do until the logits of tile g + 2 arrive
        FT0(t11);
        fold_until(g);
        FTA(11, t11);
    }
    fold_until(0xffffffffu);                                   // the segments that ended after this team's last tile
#ifdef PVS_TIMING
    FTA(12, t_all);
    if (warp == 2 && lane == 0 && RANK == 0) for (int i = 4; i < 13; ++i) atomicAdd(&g_ft[i], (unsigned long long)ft[i]);
#endif
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) kernel(const __grid_constant__ Params p)
{
    if (*p.flag != 0) return;                                  // uniform over the grid
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    const Bars B(bars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(B.a1_full(i), 4);
        for (int i = 0; i < 2; ++i) {
            mbar_init(B.a1_free(i), 1);
            mbar_init(B.l_full(i), 1);
            mbar_init(B.l_free(i), 4);                        // the four warps of the team that owns the buffer
            mbar_init(B.s_full(i), 1);
            mbar_init(B.s_free(i), 8);                        // both teams fold half of the columns each
        }
        // Q chunk use u (= 4 g + quarter) lives in slot u & 1 and signals barrier u & 7: softmax(g) overlaps softmax(g - 1)
        // and the statistics MMAs of tile g - 1, so a producer polls the barrier of use u - 2 while uses of tile g - 1 may
        // still be pending; with a ring of eight the previous phase of that barrier belongs to tile <= g - 2, which is
        // complete before l_full(g) (a parity wait is only unambiguous one phase ahead).
        for (int i = 0; i < 8; ++i) {
            mbar_init(B.q_full(i), 1);
            mbar_init(B.q_empty(i), 1 + 1);                   // MMA commit + the converter warp that summed the chunk
        }
        mbar_init(B.w_res(), 1);
        // exchange barriers: per (lane quarter, team), ring of two (the peer cannot be two tiles of a team ahead: its
        // next arrival needs this CTA's arrival for the tile in between, which comes after this CTA's wait)
        for (int i = 0; i < 16; ++i) mbar_init(B.x_sum(i), 32);
        fence_barrier_init();
        tma_prefetch_desc(&p.w_hi);
        tma_prefetch_desc(&p.w_lo);
    }
    if (warp == 0) {
        float* cs = reinterpret_cast<float*>(smem + OFF_CST);
        for (int i = lane; i < CK; i += 32) cs[i] = p.cst[rank * CK + i];
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    cluster_sync();                                            // the peer's barriers exist before anybody signals them
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;

    const uint32_t a1b = smem_u32(smem + OFF_A1), wb = smem_u32(smem + OFF_W), qb = smem_u32(smem + OFF_Q);
    if (warp == 0) {
        // ---- W' load, then the logit MMAs: MMA1(g) as soon as A1(g) is converted and the team's accumulator is free ----
        // Two issuing warps: with one warp issuing MMA1(g) and the statistics MMAs in a fixed order, the logits of a team's
        // next tile waited behind the other team's chunks (or the chunks behind the next conversion) and the whole kernel
        // ran as one chain.  tcgen05.commit only tracks the MMAs of the committing thread, which is what is wanted here.
        constexpr uint32_t idesc1 = make_idesc(false, false, false, 128, CK, true);     // K-major x K-major, N = 128
        if (lane == 0) {                                       // this CTA's 128 rows of W', both k-blocks, hi and lo
            mbar_expect_tx(B.w_res(), W_BYTES);
            uint8_t* w = smem + OFF_W;
            tma_load_2d(w, &p.w_hi, B.w_res(), 0, (int)rank * CK);
            tma_load_2d(w + 16384, &p.w_hi, B.w_res(), 64, (int)rank * CK);
            tma_load_2d(w + 32768, &p.w_lo, B.w_res(), 0, (int)rank * CK);
            tma_load_2d(w + 49152, &p.w_lo, B.w_res(), 64, (int)rank * CK);
        }
        __syncwarp();
#ifdef PVS_TIMING
        long long ft[16] = {0};
#endif
        mbar_wait(B.w_res(), 0);
        Walk w;
        for (bool ok = walk_start(w, p, cluster_id, n_clusters, NoEmpty()); ok; ok = walk_next(w, p, n_clusters, NoEmpty())) {
            const uint32_t g = w.g, b = g & 1, ph = (g >> 1) & 1;
            FT0(t1);
            mbar_wait(B.l_free(b), ph ^ 1);
            FTA(1, t1);
            for (int kb = 0; kb < 2; ++kb) {
                FT0(t0);
                mbar_wait(B.a1_full(b * 2 + kb), ph);          // the converters publish the two k-blocks separately
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
                    if (kb == 1) umma_commit(B.l_full(b));
                }
                __syncwarp();
            }
            TRACE(g, 0);
        }
#ifdef PVS_TIMING
        if (lane == 0 && rank == 0) for (int i = 0; i < 2; ++i) atomicAdd(&g_ft[i], (unsigned long long)ft[i]);
#endif
    } else if (warp == 1) {
        // ---- statistics MMAs: the four Q chunks of every tile, in order ----
        constexpr uint32_t idesc2 = make_idesc(false, true, true, 128, CK, true);       // MN-major x MN-major, N = 128
#ifdef PVS_TIMING
        long long ft[16] = {0};
#endif
        FT0(t_all);
        Walk t;
        for (bool ok = walk_start(t, p, cluster_id, n_clusters, NoEmpty()); ok; ok = walk_next(t, p, n_clusters, NoEmpty())) {
            const uint32_t gp = t.g, sb = t.sg & 1;
            const uint32_t a1 = a1b + (gp & 1) * A1_BYTES;
            const uint32_t d = tmem + 256 + sb * CK;
            const bool fresh = t.seg_first();
            if (fresh && t.sg >= 2) {                          // the segment that used this buffer has been folded by both teams
                FT0(t13);
                mbar_wait(B.s_free(sb), ((t.sg >> 1) & 1) ^ 1);
                FTA(13, t13);
                tcgen05_fence_after();
            }
            for (int q = 0; q < 4; ++q) {
                const uint32_t use = 4 * gp + (uint32_t)q;
                FT0(t2);
                mbar_wait(B.q_full(use & 7), (use >> 3) & 1);
                FTA(2, t2);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t qs = qb + (use & 1) * QCH_BYTES;
#pragma unroll
                    for (int k2 = 0; k2 < 2; ++k2) {
                        const int ks = 2 * q + k2;             // 16 descriptors per k-step
                        const uint64_t a_hi = make_smem_desc(a1 + ks * 2048, 32768, 1024, LAYOUT_SW128);
                        const uint64_t a_lo = make_smem_desc(a1 + 16384 + ks * 2048, 32768, 1024, LAYOUT_SW128);
                        const uint64_t b_hi = make_smem_desc(qs + k2 * 2048, QBLK, 1024, LAYOUT_SW128);
                        const uint64_t b_lo = make_smem_desc(qs + QPLANE + k2 * 2048, QBLK, 1024, LAYOUT_SW128);
                        umma<true>(d, a_hi, b_lo, idesc2, (fresh && q == 0 && k2 == 0) ? 0u : 1u);
                        umma<true>(d, a_lo, b_hi, idesc2, 1u);
                        umma<true>(d, a_hi, b_hi, idesc2, 1u);
                    }
                    umma_commit(B.q_empty(use & 7));
                }
                __syncwarp();
                TRACE(gp, 18 + q);
            }
            if (elect_one()) {
                umma_commit(B.a1_free(gp & 1));
                if (t.seg_last()) umma_commit(B.s_full(sb));
            }
            __syncwarp();
        }
#ifdef PVS_TIMING
        FTA(3, t_all);
        if (lane == 0 && rank == 0) {
            for (int i = 2; i < 4; ++i) atomicAdd(&g_ft[i], (unsigned long long)ft[i]);
            atomicAdd(&g_ft[13], (unsigned long long)ft[13]);
        }
#endif
    } else if (warp < 10) {
        softmax_role(p, smem, B, tmem, warp, lane, cluster_id, n_clusters, (int)rank);
    } else {
        // ---- converters: A1(g + 1) first, then the zeroth-order sums of this warp's Q chunk of tile g ----
        const int cw = warp - 10;                              // = the lane quarter whose chunk this warp sums
        const int c = lane & 7;
        float4 yv[16];
        auto fetch = [&](const Walk& t) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = t.tile * TT + cw * 32 + i * 4 + (lane >> 3);
                    yv[kb * 8 + i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < t.T) yv[kb * 8 + i] = __ldg(reinterpret_cast<const float4*>(p.y + (t.r0 + r) * D + kb * 32 + c * 4));
                }
        };
        auto convert = [&](uint32_t gg) {                      // registers -> A1[gg & 1] (the caller saw a1_free complete)
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
                if (lane == 0) mbar_arrive(B.a1_full((gg & 1) * 2 + kb));
            }
        };
        // lanes 0-15 take the even rows of a Q chunk, lanes 16-31 the odd ones; a lane owns four components of each
        // 64-component block
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int jl = (lane & 15) >> 1, sub = (lane & 1) * 8, rpar = lane >> 4;
        auto sum_chunk = [&](const Walk& t) {                  // zeroth-order sums of chunk (t.g, cw) (the caller saw q_full complete)
            const uint32_t use = 4 * t.g + (uint32_t)cw;
            const uint8_t* qh = smem + OFF_Q + (use & 1) * QCH_BYTES;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
                for (int rr = 0; rr < 16; ++rr) {
                    const int r = 2 * rr + rpar;
                    const uint32_t off = (uint32_t)(blk * QBLK + r * 128 + ((jl ^ (r & 7)) << 4) + sub);
                    const uint2 hv = *reinterpret_cast<const uint2*>(qh + off);
                    const uint2 lv = *reinterpret_cast<const uint2*>(qh + QPLANE + off);
                    const float2 h0 = __half22float2(*reinterpret_cast<const __half2*>(&hv.x));
                    const float2 h1 = __half22float2(*reinterpret_cast<const __half2*>(&hv.y));
                    const float2 l0 = __half22float2(*reinterpret_cast<const __half2*>(&lv.x));
                    const float2 l1 = __half22float2(*reinterpret_cast<const __half2*>(&lv.y));
                    a0 += h0.x + l0.x;
                    a1 += h0.y + l0.y;
                    a2 += h1.x + l1.x;
                    a3 += h1.y + l1.y;
                }
                acc[4 * blk] += a0; acc[4 * blk + 1] += a1; acc[4 * blk + 2] += a2; acc[4 * blk + 3] += a3;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(B.q_empty(use & 7));
            if (t.tile == t.nt - 1) {                          // image end: raw sums / 2^14 into this warp's partial slot
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
                if (lane < 16) {
                    float* dst = p.s0part + ((int64_t)t.img * TC_FV_S0_PARTS + cw) * (int64_t)K + rank * CK + (lane >> 1) * 8 + (lane & 1) * 4;
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk)
                        *reinterpret_cast<float4*>(dst + blk * 64) =
                            make_float4(acc[4 * blk] * (1.f / 16384.f), acc[4 * blk + 1] * (1.f / 16384.f), acc[4 * blk + 2] * (1.f / 16384.f),
                                        acc[4 * blk + 3] * (1.f / 16384.f));
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            }
        };
        // Two duties, whichever is ready first: convert the tile whose rows are in the registers (needs its A1 buffer back
        // from the statistics MMAs two tiles earlier), sum the next Q chunk (its arrival releases the chunk's slot, which
        // the softmax warps of the following chunks are waiting for).  In a fixed order each duty sat behind the other's wait.
        Walk f, s;                                             // f: tile to convert (rows fetched), s: tile to sum
        bool f_ok = walk_start(f, p, cluster_id, n_clusters, NoEmpty());
        bool s_ok = walk_start(s, p, cluster_id, n_clusters, NoEmpty());
        if (f_ok) fetch(f);
        const long long t_start = clock64();
        while (f_ok || s_ok) {
            bool did = false;
            if (f_ok && mbar_test(B.a1_free(f.g & 1), ((f.g >> 1) & 1) ^ 1)) {
                convert(f.g);
                if (cw == 0) TRACE(f.g, 1);
                f_ok = walk_next(f, p, n_clusters, NoEmpty());
                if (f_ok) fetch(f);
                did = true;
            }
            if (s_ok && s.g < f.g) {                           // only tiles that have been converted can have chunks
                const uint32_t use = 4 * s.g + (uint32_t)cw;
                if (mbar_test(B.q_full(use & 7), (use >> 3) & 1)) {
                    sum_chunk(s);
                    TRACE(s.g, 22 + cw);
                    s_ok = walk_next(s, p, n_clusters, NoEmpty());
                    did = true;
                }
            }
            if (!did && clock64() - t_start > 20000000000LL) __trap();
        }
    }
    tcgen05_fence_before();
    cluster_sync();                                            // nobody exits while the peer may still write into its shared memory
    if (warp == 1) tmem_dealloc<512>(tmem);
}
}  // namespace fused3
}  // namespace tc

using namespace tc;

int tc_fv_poststats_fused3(const TcFvPlan& pl, const pvs_model* g, const float* y, const int64_t* offsets, int64_t n_images,
                           cudaStream_t st)
{
    if (n_images <= 0) return PVS_OK;
    static PerDeviceOnce configured;
    if (configured.need()) {
        PVS_CUDA(cudaFuncSetAttribute(fused3::kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fused3::SMEM_BYTES));
        configured.mark();
    }
    fused3::Params p{};
    int rc;
    if ((rc = make_tmap_2d(&p.w_hi, g->th0, true, fused3::K, fused3::AUG, fused3::AUG, 64, 128))) return rc;
    if ((rc = make_tmap_2d(&p.w_lo, g->th1, true, fused3::K, fused3::AUG, fused3::AUG, 64, 128))) return rc;
    PVS_CHECK((int)g->cst_host.size() == fused3::K, PVS_ERR_BAD_ARG, "GMM model lacks the host copy of its constants");
    memcpy(p.cst, g->cst_host.data(), sizeof(p.cst));
    p.y = y; p.offsets = offsets; p.S = pl.S; p.s0part = pl.s0part; p.n_images = n_images; p.flag = pl.flag;
    p.sc_y = ldexpf(1.f, -g->h_exp); p.un1 = ldexpf(1.f, g->h_exp - 14); p.un2 = ldexpf(1.f, 2 * g->h_exp - 14);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int clusters = n_images < sms / 2 ? (int)n_images : sms / 2;
    fused3::kernel<<<2 * clusters, fused3::THREADS, fused3::SMEM_BYTES, st>>>(p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PVS_ERR_CUDA, "fused (teams) posterior + statistics kernel launch failed: %s", cudaGetErrorString(e));
#ifdef PVS_TIMING
    if (getenv("PVS_TIMING_PRINT")) {
        cudaStreamSynchronize(st);
        unsigned long long h[16], z[16] = {0};
        cudaMemcpyFromSymbol(h, fused3::g_ft, sizeof(h));
        cudaMemcpyToSymbol(fused3::g_ft, z, sizeof(z));
        const double np = clusters * 1e3;
        fprintf(stderr, "[fused3 timing] per cluster (kcycles): mma total %.0f wait a1_full %.0f l_free %.0f q_full %.0f s_free %.0f | team-0 warp total %.0f wait l_full %.0f p1 %.0f p2 %.0f safety fold %.0f xwait %.0f p3 %.0f (q_empty %.0f) fold %.0f\n",
                h[3] / np, h[0] / np, h[1] / np, h[2] / np, h[13] / np, h[12] / np, h[4] / np, h[5] / np, h[6] / np, h[7] / np, h[8] / np, h[9] / np, h[10] / np, h[11] / np);
        if (getenv("PVS_TRACE")) {
            static long long tr[fused3::TRN][32];
            cudaMemcpyFromSymbol(tr, fused3::g_trace, sizeof(tr));
            const long long t0 = tr[0][1] ? tr[0][1] : tr[0][0];
            fprintf(stderr, "tile | conv mma1 | lfull q0-3 | p2done q0-3 | slot q0-3 | pub q0-3 | mma2 c0-3 | sum c0-3   (cycles / 100 since the conversion of tile %d)\n", fused3::TR0);
            for (int i = 0; i < fused3::TRN; ++i) {
                fprintf(stderr, "%3d |", fused3::TR0 + i);
                const int order[] = {1, 0, -1, 2, 3, 4, 5, -1, 6, 7, 8, 9, -1, 10, 11, 12, 13, -1, 14, 15, 16, 17, -1, 18, 19, 20, 21, -1, 22, 23, 24, 25};
                for (int e : order) {
                    if (e < 0) fprintf(stderr, " |");
                    else fprintf(stderr, " %5lld", (tr[i][e] - t0) / 100);
                }
                fprintf(stderr, "\n");
            }
        }
    }
#endif
    return PVS_OK;
}

}  // namespace pvs
