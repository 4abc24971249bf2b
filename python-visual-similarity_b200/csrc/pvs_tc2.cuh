// pvs_tc2.cuh -- CTA-pair (tcgen05 cta_group::2) variant of the warp-specialised skeleton.
//
// Two CTAs of a 2-CTA cluster (the two SMs of a TPC) execute ONE tcgen05.mma together:
// the pair computes a [256 x BLOCK_N] tile, CTA r holds rows [128 r, 128 r + 128) of A and
// rows [BLOCK_N/2 r, ...) of B in its own shared memory and the [128 x BLOCK_N] slice of the
// accumulator in its own TMEM.  Compared with two independent 128-row CTAs each B byte is
// fetched from L2 and written to shared memory once per pair instead of once per CTA, which
// is what lifts the L2->SMEM feed limit of the single-CTA kernels; and a weight operand of
// K=256 rows fits resident (half per CTA) where it had to be re-streamed per tile.
//
// Roles per CTA (same warp ids in both CTAs):
//   warp 0      TMA producer (one lane): loads this CTA's operand halves; the transaction
//               bytes of BOTH CTAs complete on the LEADER's `full` mbarrier
//   warp 1      leader CTA only: single-thread MMA issuer; both CTAs: TMEM alloc / dealloc
//   warps 2-5   epilogue, one TMEM lane quarter each, rows of this CTA
//   warps 6..   (P::MANUAL) operand producers: fp32 global -> tf32 hi/lo swizzled tiles in
//               this CTA's shared memory, then a remote arrive on the leader's `full`.
//               P::PGROUPS groups of four warps; group g fills the k-blocks with running
//               index == g (mod PGROUPS), so PGROUPS stages are being loaded at any time
//               (the loads are latency-bound: this is what buys memory-level parallelism)
// Barriers: full[s] lives in the leader (peer arrives remotely); empty[s] and tfull[a] exist
// in both CTAs and are signalled by tcgen05.commit with a 2-CTA multicast; tempty[a] lives
// in the leader and collects the epilogue warps of both CTAs.
//
// Policy P (see pvs_tc.cuh for the single-CTA contract) additionally provides
//   B_RESIDENT : the whole B operand (all k-blocks) is loaded once per kernel into a
//                dedicated region and reused by every tile (weights); else B is staged
//   NKB_RES    : number of resident k-blocks (B_RESIDENT only)
//   ACC_INIT   : the epilogue warps pre-load every accumulator buffer (P::acc_init, e.g. a
//                per-column constant) before the MMA warp may use it, and the first MMA of a
//                tile accumulates onto that instead of overwriting
// and its callbacks receive the CTA rank and pair index instead of reading blockIdx.
#pragma once
#include <stdlib.h>
#include "pvs_tc.cuh"

namespace pvs {
namespace tc2 {
using namespace tc;

constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;      // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in the LEADER CTA.  Default (.release.cta)
// semantics on purpose: a cluster-scope release / acquire makes ptxas emit MEMBAR.ALL.GPU
// and an L1 invalidate (CCTL.IVALL) around every arrive / wait, which tripled the fill
// latency of the operand producers.  What crosses the CTA boundary here is shared memory
// read by the tensor core (async proxy), ordered by fence.proxy.async + the mbarrier itself.
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
#ifdef PVS_MBAR_HINT
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)PVS_MBAR_HINT)
        : "memory");
#else
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait_cl(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cl(bar, parity)) {
        if (clock64() - t0 > 6000000000LL) __trap();
    }
}
// TMA load into THIS CTA's shared memory, bytes complete on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot_in_smem)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
template <bool BF16>
__device__ __forceinline__ void umma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    if constexpr (BF16) {
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// completion of all prior MMAs of this thread -> arrive on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

#ifdef PVS_TIMING
// role wait-time counters (cycles, summed over CTAs): [0] MMA wait full, [1] MMA wait tempty,
// [2] MMA total, [3] epilogue(warp 2) wait tfull, [4] epilogue total, [5] producer(warp 6) wait empty,
// [6] producer total, [7] epilogue body
__device__ unsigned long long g_tc2_timing[16];          // [8..15]: policy-defined phases (PVS_TPHASE)
#define PVS_T0(var) const long long var = clock64()
#define PVS_TACC(slot, t0) timing_acc[slot] += clock64() - (t0)
#define PVS_TPHASE(slot, t0, cond) do { if (cond) atomicAdd(&g_tc2_timing[slot], (unsigned long long)(clock64() - (t0))); } while (0)
#else
#define PVS_T0(var)
#define PVS_TACC(slot, t0)
#define PVS_TPHASE(slot, t0, cond)
#endif

// Optional: static constexpr int EPI_WARPS = 8 -- two epilogue warps per TMEM lane quarter (warps 2..9; the
// policy splits the accumulator columns between the two and reconciles row reductions through shared memory);
// the operand producers then start at warp 10.
template <class P, class = void> struct policy_epi_warps { static constexpr int value = 4; };
template <class P> struct policy_epi_warps<P, decltype((void)P::EPI_WARPS)> { static constexpr int value = P::EPI_WARPS; };

// Optional: static constexpr bool PREFETCH2 = true -- producers keep two k-blocks of global loads in flight
template <class P, class = void> struct policy_prefetch2 { static constexpr bool value = false; };
template <class P> struct policy_prefetch2<P, decltype((void)P::PREFETCH2)> { static constexpr bool value = P::PREFETCH2; };

template <class P>
struct Layout2 {
    static constexpr int EW = policy_epi_warps<P>::value;
    static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
    static constexpr int PARTS = P::PASSES == 3 ? 2 : 1;
    static constexpr int B_STAGE = P::B_RESIDENT ? 0 : PARTS * P::B_BYTES;
    static constexpr int STAGE_BYTES = PARTS * P::A_BYTES + B_STAGE;
    static constexpr int RING_BYTES = P::STAGES * STAGE_BYTES;
    static constexpr int RES_BYTES = P::B_RESIDENT ? P::NKB_RES * PARTS * P::B_BYTES : 0;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = 1024 + RING_BYTES + RES_BYTES + BAR_BYTES + P::SCRATCH_BYTES;
    static constexpr int TMEM_COLS = 2 * P::BLOCK_N <= 32 ? 32 : 2 * P::BLOCK_N <= 64 ? 64 : 2 * P::BLOCK_N <= 128 ? 128
                                     : 2 * P::BLOCK_N <= 256 ? 256 : 512;
    static constexpr int THREADS = 64 + 32 * EW + (P::MANUAL ? 128 * P::PGROUPS : 0);
    // arrivals per phase on the leader's full barrier
    static constexpr uint32_t FULL_COUNT = (P::TMA_BYTES > 0 ? 1 : 0) + (P::MANUAL ? 8 : 0);
    static_assert(2 * P::BLOCK_N <= 512, "accumulator does not fit TMEM twice");
    static_assert(P::BLOCK_N % 16 == 0 && P::BLOCK_N <= 256, "pair MMA needs N % 16 == 0, N <= 256");
    static_assert(P::A_BYTES % 1024 == 0 && P::B_BYTES % 1024 == 0, "operand tiles must keep 1024-B alignment");
    static_assert(P::STAGES + 8 <= 30, "barrier block too small");
    // every stage must be refilled by one producer group only: a group that ran two uses ahead of
    // another one on the same stage would pass the parity wait of the empty barrier too early
    static_assert(!P::MANUAL || (P::PGROUPS >= 1 && P::STAGES % P::PGROUPS == 0), "PGROUPS must divide STAGES");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
};

template <class P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Layout2<P>::THREADS, 1)
tc2_kernel(const __grid_constant__ typename P::Params prm)
{
    using L = Layout2<P>;
    if constexpr (policy_gated<P>::value) {
        if (!P::enabled(prm)) return;                        // uniform over the grid (both CTAs of every pair)
    }
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET on the __shared__ array: going through an integer cast would
    // turn every later access into a generic LD / ST instead of LDS / STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* res = smem + L::RING_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::RING_BYTES + L::RES_BYTES);
    uint64_t* empty = full + P::STAGES;
    uint64_t* tfull = empty + P::STAGES;
    uint64_t* tempty = tfull + 2;
    uint64_t* bres = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);
    uint8_t* scratch = smem + L::RING_BYTES + L::RES_BYTES + L::BAR_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P::STAGES; ++s) {
            mbar_init(&full[s], L::FULL_COUNT);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 2 * L::EW);
        }
        mbar_init(bres, 1);
        fence_barrier_init();
        P::prefetch(prm);
    }
    if (warp == 1) tmem_alloc2<L::TMEM_COLS>(tmem_slot);
    tcgen05_fence_before();
    cluster_sync_all();                                     // peer barriers initialised, TMEM allocated
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = P::num_tiles(prm);

    if (warp == 0) {
        if (lane == 0) {
            if constexpr (P::B_RESIDENT) {
                if (rank == 0) mbar_expect_tx(bres, 2u * (uint32_t)L::RES_BYTES);
                P::load_resident(prm, rank, res, bres);
            }
            if constexpr (P::TMA_BYTES > 0) {
                int stage = 0;
                uint32_t phase = 0;
                int it = 0;
                for (int t; (t = P::tile_at(prm, it, pair, n_pairs, n_tiles)) >= 0; ++it) {
                    const typename P::Tile tl = P::tile(prm, t);
                    if constexpr (P::TILE_SYNC) P::tile_sync(prm, it, rank, n_pairs);   // keep the pairs in step (L2 reuse)
                    for (int kb = 0; kb < tl.nkb; ++kb) {
                        mbar_wait_cl(&empty[stage], phase ^ 1);
                        if (rank == 0) mbar_expect_tx(&full[stage], 2u * (uint32_t)P::TMA_BYTES);
                        uint8_t* sp = smem + stage * L::STAGE_BYTES;
                        P::load(prm, tl, kb, rank, sp, sp + P::A_BYTES, sp + L::PARTS * P::A_BYTES,
                                sp + L::PARTS * P::A_BYTES + P::B_BYTES, &full[stage]);
                        if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
                if constexpr (P::TILE_SYNC) P::tile_sync_done(prm, it, rank);
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {   // the whole warp walks the pipeline; one elected lane issues the MMAs and commits
            constexpr bool H = P::BF16 || policy_f16<P>::value;   // kind::f16
            constexpr uint32_t idesc = make_idesc(P::BF16, false, false, 256, P::BLOCK_N, policy_f16<P>::value);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
#ifdef PVS_TIMING
            long long timing_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
            PVS_T0(t_role);
            if constexpr (P::B_RESIDENT) mbar_wait_cl(bres, 0);
            for (int it = 0, t; (t = P::tile_at(prm, it, pair, n_pairs, n_tiles)) >= 0; ++it) {
                const typename P::Tile tl = P::tile(prm, t);
                // ACC_INIT: the buffer must have been initialised (phase n) rather than merely be free
                PVS_T0(t_te);
                mbar_wait_cl(&tempty[acc], P::ACC_INIT ? acc_phase : acc_phase ^ 1);
                PVS_TACC(1, t_te);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P::BLOCK_N);
                for (int kb = 0; kb < tl.nkb; ++kb) {
                    PVS_T0(t_fu);
                    mbar_wait_cl(&full[stage], phase);
                    PVS_TACC(0, t_fu);
                    tcgen05_fence_after();
                    const uint32_t sp = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint32_t a_hi = sp, a_lo = sp + P::A_BYTES;
                    const uint32_t b_hi = P::B_RESIDENT ? smem_u32(res) + (uint32_t)(kb * L::PARTS * P::B_BYTES)
                                                        : sp + L::PARTS * P::A_BYTES;
                    const uint32_t b_lo = b_hi + P::B_BYTES;
                    if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < P::KSTEPS; ++ks) {
                        const uint64_t da_hi = make_smem_desc(a_hi + ks * 32, 16, 1024, LAYOUT_SW128);
                        const uint64_t db_hi = make_smem_desc(b_hi + ks * 32, 16, 1024, LAYOUT_SW128);
                        const uint32_t first = (P::ACC_INIT || kb > 0 || ks > 0) ? 1u : 0u;
                        if constexpr (P::PASSES == 3) {
                            const uint64_t da_lo = make_smem_desc(a_lo + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t db_lo = make_smem_desc(b_lo + ks * 32, 16, 1024, LAYOUT_SW128);
                            umma2<H>(d_tmem, da_hi, db_lo, idesc, first);
                            umma2<H>(d_tmem, da_lo, db_hi, idesc, 1u);
                            umma2<H>(d_tmem, da_hi, db_hi, idesc, 1u);
                        } else {
                            umma2<H>(d_tmem, da_hi, db_hi, idesc, first);
                        }
                    }
                    umma2_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma2_commit(&tfull[acc]);
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
#ifdef PVS_TIMING
            PVS_TACC(2, t_role);
            if (lane == 0) for (int i = 0; i < 3; ++i) atomicAdd(&g_tc2_timing[i], (unsigned long long)timing_acc[i]);
#endif
        }
    } else if (warp >= 2 + L::EW) {
        if constexpr (P::MANUAL) {
            // Operand producers.  fetch() only issues the global loads of a k-block into
            // registers, store() splits and writes the swizzled tiles.  The loads of the NEXT
            // k-block of this group are issued before waiting for its stage to drain, so the
            // global-memory latency overlaps the wait instead of following it.
            const int pw = (warp - 2 - L::EW) & 3, grp = (warp - 2 - L::EW) >> 2;
            // position of a k-block in this pair's stream of (tile, k-block) pairs; idx = running index
            struct Pos { long long idx; int it, kb, t; typename P::Tile tl; };
            auto seek = [&](Pos& q) {                        // settle on the next k-block owned by this group
                while (q.t >= 0) {
                    if (q.kb >= q.tl.nkb) {
                        ++q.it; q.kb = 0;
                        q.t = P::tile_at(prm, q.it, pair, n_pairs, n_tiles);
                        if (q.t >= 0) q.tl = P::tile(prm, q.t);
                        continue;
                    }
                    if ((int)(q.idx % P::PGROUPS) == grp) return;
                    ++q.kb; ++q.idx;
                }
            };
            auto next = [&](Pos& q) { ++q.kb; ++q.idx; seek(q); };
            Pos s{0, 0, 0, P::tile_at(prm, 0, pair, n_pairs, n_tiles), {}};
            if (s.t >= 0) s.tl = P::tile(prm, s.t);
            seek(s);
#ifdef PVS_TIMING
            long long timing_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
            PVS_T0(t_role);
            auto publish = [&](const Pos& q, const typename P::Regs& r) {
                const int stage = (int)(q.idx % P::STAGES);
                const uint32_t phase = (uint32_t)((q.idx / P::STAGES) & 1);
                PVS_T0(t_em);
                mbar_wait_cl(&empty[stage], phase ^ 1);
                PVS_TACC(5, t_em);
                uint8_t* sp = smem + stage * L::STAGE_BYTES;
                P::store(prm, q.tl, q.kb, r, sp, sp + P::A_BYTES, pw, lane);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&full[stage]);
            };
            if constexpr (policy_prefetch2<P>::value) {
                // two k-blocks of global loads in flight per group (register double buffer): with a single
                // group and short tiles one block in flight leaves the loads latency-bound
                typename P::Regs r0, r1;
                Pos f = s;
                if (s.t >= 0) {
                    P::fetch(prm, f.tl, f.kb, rank, pw, lane, r0);
                    next(f);
                    if (f.t >= 0) P::fetch(prm, f.tl, f.kb, rank, pw, lane, r1);
                }
                while (s.t >= 0) {
                    publish(s, r0);
                    next(s);
                    if (s.t < 0) break;
                    next(f);
                    if (f.t >= 0) P::fetch(prm, f.tl, f.kb, rank, pw, lane, r0);
                    publish(s, r1);
                    next(s);
                    if (s.t < 0) break;
                    next(f);
                    if (f.t >= 0) P::fetch(prm, f.tl, f.kb, rank, pw, lane, r1);
                }
            } else {
                // Operand producers.  fetch() only issues the global loads of a k-block into
                // registers, store() splits and writes the swizzled tiles.  The loads of the NEXT
                // k-block of this group are issued before waiting for its stage to drain, so the
                // global-memory latency overlaps the wait instead of following it.
                typename P::Regs cur;
                if (s.t >= 0) P::fetch(prm, s.tl, s.kb, rank, pw, lane, cur);
                while (s.t >= 0) {
                    publish(s, cur);
                    next(s);
                    if (s.t >= 0) P::fetch(prm, s.tl, s.kb, rank, pw, lane, cur);
                }
            }
#ifdef PVS_TIMING
            PVS_TACC(6, t_role);
            if (warp == 2 + L::EW && lane == 0 && rank == 0) for (int i = 5; i < 7; ++i) atomicAdd(&g_tc2_timing[i], (unsigned long long)timing_acc[i]);
#endif
        }
    } else {
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        typename P::EpiState st;
        P::epi_init(prm, scratch, (int)threadIdx.x - 64);
        if constexpr (P::ACC_INIT) {
            for (int a = 0; a < 2; ++a) {
                P::acc_init(prm, tmem_base + (uint32_t)(a * P::BLOCK_N) + ((uint32_t)(quarter * 32) << 16), lane, scratch);
                tmem_st_wait();
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&tempty[a]);
            }
        }
#ifdef PVS_TIMING
        long long timing_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        PVS_T0(t_role);
        for (int it = 0, t; (t = P::tile_at(prm, it, pair, n_pairs, n_tiles)) >= 0; ++it) {
            const typename P::Tile tl = P::tile(prm, t);
            P::epi_begin(prm, tl, st, rank, quarter, lane);
            PVS_T0(t_tf);
            mbar_wait_cl(&tfull[acc], acc_phase);
            PVS_TACC(3, t_tf);
            tcgen05_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(acc * P::BLOCK_N) + ((uint32_t)(quarter * 32) << 16);
            PVS_T0(t_body);
            P::epilogue(prm, tl, rank, tacc, quarter, lane, scratch, st);
            PVS_TACC(7, t_body);
            if constexpr (P::ACC_INIT) {
                P::acc_init(prm, tacc, lane, scratch);
                tmem_st_wait();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        tma_store_wait_read<0>();                            // TMA stores of the last tile have read their smem tiles
#ifdef PVS_TIMING
        PVS_TACC(4, t_role);
        if (warp == 2 && lane == 0 && rank == 0) { atomicAdd(&g_tc2_timing[3], (unsigned long long)timing_acc[3]); atomicAdd(&g_tc2_timing[4], (unsigned long long)timing_acc[4]); atomicAdd(&g_tc2_timing[7], (unsigned long long)timing_acc[7]); }
#endif
    }
    tcgen05_fence_before();
    cluster_sync_all();                                     // nobody exits while the peer may still signal it
    if (warp == 1) tmem_dealloc2<L::TMEM_COLS>(tmem_base);
}

template <class P>
int launch_tc2(const typename P::Params& prm, int n_tiles, cudaStream_t st, int pairs_override = 0)
{
    using L = Layout2<P>;
    static PerDeviceOnce configured;
    if (configured.need()) {
        PVS_CUDA(cudaFuncSetAttribute(tc2_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM_BYTES));
        configured.mark();
    }
    if (n_tiles <= 0) return PVS_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int pairs = pairs_override > 0 ? pairs_override : (n_tiles < sms / 2 ? n_tiles : sms / 2);
    tc2_kernel<P><<<2 * pairs, L::THREADS, L::SMEM_BYTES, st>>>(prm);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PVS_ERR_CUDA, "tcgen05 pair kernel launch failed: %s", cudaGetErrorString(e));
#ifdef PVS_TIMING
    if (getenv("PVS_TIMING_PRINT")) {
        cudaStreamSynchronize(st);
        unsigned long long h[16], z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyFromSymbol(h, g_tc2_timing, sizeof(h));
        cudaMemcpyToSymbol(g_tc2_timing, z, sizeof(z));
        const double np = pairs;
        fprintf(stderr, "[tc2 timing %s] per pair (kcycles): mma total %.0f wait_full %.0f wait_tempty %.0f | epi total %.0f wait_tfull %.0f body %.0f | prod total %.0f wait_empty %.0f\n",
                __PRETTY_FUNCTION__, h[2] / np / 1e3, h[0] / np / 1e3, h[1] / np / 1e3, h[4] / np / 1e3, h[3] / np / 1e3, h[7] / np / 1e3, h[6] / np / 1e3, h[5] / np / 1e3);
        fprintf(stderr, "    phases (kcycles per pair): %.0f %.0f %.0f %.0f %.0f %.0f\n", h[8] / np / 1e3, h[9] / np / 1e3, h[10] / np / 1e3,
                h[11] / np / 1e3, h[12] / np / 1e3, h[13] / np / 1e3);
    }
#endif
    return PVS_OK;
}

}  // namespace tc2
}  // namespace pvs
