// pvs_tc_sim.cu -- all-pairs cosine similarity with fused per-row top-k on tcgen05 (bf16),
// CTA-pair (cta_group::2) kernel.
//
// Replaces `cosine_similarity(q, all)` + `np.argsort(-s)[:k]` per query
// (pyvisim/eval.py:37-43, 76-80, 131-132) by one sweep: scores = Qn . DBn^T as a bf16
// tensor-core contraction with fp32 accumulation in TMEM; the score matrix is never
// written.  A pair of CTAs computes a [256 queries x 256 database rows] tile per
// accumulator; every epilogue thread owns one query row (= one TMEM lane).
//
// Fused top-k.  Per row: a threshold tau (the current k-th best key) in registers and an
// append buffer of CAP >= 2k keys in global memory (L2 resident, private to the row).
// A score enters the buffer only if its key beats tau -- one 8-byte store, no dependent
// loads.  When a row's buffer could overflow during the next 32 columns the warp prunes it
// cooperatively: the 32 lanes load the buffer into registers, find the k-th largest key by
// a radix search over the key bits (one warp reduction per two bits), compact the k survivors
// back and raise tau (k > 256: bitonic sort in shared memory instead).  A row is pruned
// O(log(n_db / k)) times, so the epilogue stays far below the MMA time of a tile.  The
// partial lists are emitted unsorted; the merge kernel sorts.  Ordering key = (score, lowest index first), the same total order as the
// CUDA-core path, so both paths return identical lists up to score rounding.
//
// Work decomposition: unit = (block of 256 query rows) x (stripe of database blocks); a
// pair sweeps the tiles of a unit back to back, keeping the rows' top-k state.  Units are
// dealt round-robin, query-block major, so the ~74 pairs running at any time cover about
// sqrt(74) query blocks x sqrt(74) stripes and walk the feature dimension in step: each
// operand k-slice is fetched from HBM once and then served from L2 to the other pairs.
// The per-stripe partial lists are merged by the bitonic merge kernel (pvs_simt.cu).
#include <stdlib.h>
#include "pvs_tc2.cuh"
#include "pvs_kernels.cuh"

namespace pvs {
namespace tc2 {

struct SimParams {
    CUtensorMap q_map, db_map;     // bf16 rows, or the fp16 hi planes of the split format
    CUtensorMap q_lo, db_lo;       // fp16 lo planes (PVS_F16X2 operands only)
    int seg, nseg;                 // split kernel: k-blocks per accumulation segment, segments per tile
    float* dense;                  // split kernel, dense mode: score matrix [n_q, ld_dense] instead of top-k lists
    int64_t ld_dense;
    unsigned long long* bufs;      // [grid CTAs, 128 rows, cap] append buffers
    float* part_scores;            // [stripes, n_q, k]
    int64_t* part_idx;             // [stripes, n_q, k]
    int64_t n_q, n_db, idx_offset;
    int k, cap, nkb, q_blocks, db_blocks, stripes, tiles_per_unit, n_units;
    unsigned* sync_ctr;            // zeroed before launch: arrivals of the pairs' tile steps (drift limiter)
    int sync_steps;                // tile steps of the busiest pair
};

__device__ __forceinline__ unsigned long long sim_key(float s, unsigned idx)
{
    unsigned u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ float sim_key_score(unsigned long long key)
{
    unsigned u = (unsigned)(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}

struct SimState {
    unsigned long long tau;   // k-th best key of the row so far (0 while fewer than k were seen)
    float tau_s;              // its score: cheap first-level reject
    int cnt;                  // entries in the row's append buffer
};

// descending bitonic sort of n (power of two) keys in shared memory by one warp
__device__ __forceinline__ void warp_bitonic_desc(unsigned long long* sm, int n, int lane)
{
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int p = lane; p < (n >> 1); p += 32) {
                const int i = ((p & ~(stride - 1)) << 1) | (p & (stride - 1));
                const int l = i + stride;
                const unsigned long long a = sm[i], b = sm[l];
                const bool desc = (i & size) == 0;
                if (desc ? (a < b) : (a > b)) { sm[i] = b; sm[l] = a; }
            }
            __syncwarp();
        }
    }
}

template <int CAP_>
struct SimPolicy {
    using Params = SimParams;
    using EpiState = SimState;
    struct Tile { int nkb, qb, dbb, stripe; bool first, last; };
    static constexpr bool BF16 = true, MANUAL = false, B_RESIDENT = false, ACC_INIT = false, TILE_SYNC = true;
    static constexpr int CAP = CAP_, PASSES = 1, BLOCK_N = 256, KSTEPS = 4, NKB_RES = 0, PGROUPS = 1;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = 128 * 128, TMA_BYTES = A_BYTES + B_BYTES;
    static constexpr int SCRATCH_BYTES = CAP <= 512 ? 0 : 4 * CAP * 8;
    static constexpr int STAGES = (226 * 1024 - 1024 - 256 - SCRATCH_BYTES) / TMA_BYTES;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.q_map); tma_prefetch_desc(&p.db_map); }
    __device__ static int num_tiles(const Params& p) { return p.n_units * p.tiles_per_unit; }
    // units are dealt round-robin to pairs; the tiles of a unit are consecutive
    __device__ static int tile_at(const Params& p, int it, int pair, int n_pairs, int)
    {
        const int j = it / p.tiles_per_unit, w = it - j * p.tiles_per_unit;
        const long long u = (long long)pair + (long long)j * n_pairs;
        return u < p.n_units ? (int)(u * p.tiles_per_unit + w) : -1;
    }
    __device__ static Tile tile(const Params& p, int i)
    {
        const int u = i / p.tiles_per_unit, w = i - u * p.tiles_per_unit;
        const int qb = u / p.stripes, s = u - qb * p.stripes;
        const int dbb = s * p.tiles_per_unit + w;
        Tile t;
        t.qb = qb; t.dbb = dbb; t.stripe = s;
        t.nkb = dbb < p.db_blocks ? p.nkb : 0;             // stripes may overhang the last block
        t.first = w == 0;
        t.last = w == p.tiles_per_unit - 1;
        return t;
    }
    __device__ static void load(const Params& p, const Tile& t, int kb, int rank, uint8_t* a, uint8_t*, uint8_t* b, uint8_t*,
                                uint64_t* bar)
    {
        tma_load_2d_pair(a, &p.q_map, bar, kb * 64, t.qb * 256 + rank * 128);
        tma_load_2d_pair(b, &p.db_map, bar, kb * 64, t.dbb * 256 + rank * 128);
    }
    // Drift limiter.  The pairs that run at the same time share query / database blocks through
    // L2 only while they walk the feature dimension of the same tile step together; nothing else
    // keeps them together over units of 100+ tiles, and a pair that falls behind finds its
    // operands evicted.  The leader's TMA thread therefore starts tile step s only after every
    // pair has reached step s (one atomic + a short spin per ~200 us tile).  All CTAs of the
    // grid are co-resident (grid <= SM count), so the spin cannot deadlock; it is a performance
    // hint only and times out harmlessly.
    __device__ static void tile_sync(const Params& p, int it, int rank, int n_pairs)
    {
        if (rank != 0 || !p.sync_ctr) return;
        atomicAdd(p.sync_ctr, 1u);
        volatile unsigned* ctr = p.sync_ctr;
        if (ctr[1]) return;                                    // a pair timed out before: run unsynchronised
        const unsigned want = (unsigned)n_pairs * (unsigned)(it + 1);
        const long long t0 = clock64();
        while (ctr[0] < want) {
            __nanosleep(200);
            if (clock64() - t0 > 4000000LL) {                  // ~2 ms: not all pairs are resident (another kernel
                ctr[1] = 1u;                                   // holds SMs); never hang, and stop waiting for good
                break;
            }
        }
    }
    // a pair with fewer tile steps than the busiest one donates its missing arrivals on exit
    __device__ static void tile_sync_done(const Params& p, int steps_done, int rank)
    {
        if (rank != 0 || !p.sync_ctr) return;
        if (steps_done < p.sync_steps) atomicAdd(p.sync_ctr, (unsigned)(p.sync_steps - steps_done));
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile& t, EpiState& st, int, int, int)
    {
        if (t.first) { st.tau = 0ull; st.tau_s = -INFINITY; st.cnt = 0; }
    }

    static constexpr int NPL = CAP / 32;                      // keys per lane when the buffer sits in registers
    static constexpr bool REG_SELECT = CAP <= 512;

    // The warp reduces the buffer of the row owned by lane `owner` to its best k keys
    // (unordered) and updates the owner's threshold.  With `emit` the survivors are also
    // written to the unit's partial result.
    __device__ static void prune_row(const Params& p, unsigned long long* warp_bufs, unsigned long long* sm, int owner,
                                     int lane, EpiState& st, bool emit, float* ps, int64_t* pi)
    {
        const int n = __shfl_sync(0xffffffffu, st.cnt, owner);
        unsigned long long* base = warp_bufs + (size_t)owner * CAP;
        __syncwarp();                                          // the owner's appends are visible to the warp
        int m = n;
        unsigned long long new_tau = 0ull;
        if constexpr (REG_SELECT) {
            unsigned long long e[NPL];
#pragma unroll
            for (int j = 0; j < NPL; ++j) e[j] = (lane + 32 * j < n) ? __ldcg(base + lane + 32 * j) : 0ull;   // L2: written by another lane
            if (n > p.k) {
                // k-th largest key: radix search from the top, two bits per step; the three
                // counts of a step travel in one warp reduction (each <= 512 < 2^10)
                unsigned long long T = 0ull;
#pragma unroll 1
                for (int bit = 62; bit >= 0; bit -= 2) {
                    const unsigned long long c1 = T | (1ull << bit), c2 = T | (2ull << bit), c3 = T | (3ull << bit);
                    unsigned cnt = 0;
#pragma unroll
                    for (int j = 0; j < NPL; ++j)
                        cnt += (e[j] >= c1 ? 1u : 0u) + (e[j] >= c2 ? 1024u : 0u) + (e[j] >= c3 ? 1048576u : 0u);
                    cnt = __reduce_add_sync(0xffffffffu, cnt);
                    const unsigned n1 = cnt & 1023u, n2 = (cnt >> 10) & 1023u, n3 = cnt >> 20;
                    T = n3 >= (unsigned)p.k ? c3 : n2 >= (unsigned)p.k ? c2 : n1 >= (unsigned)p.k ? c1 : T;
                }
                new_tau = T;                                   // keys are unique: exactly k keys are >= T
                m = p.k;
            } else if (n == p.k) {
                unsigned long long mn = ~0ull;
#pragma unroll
                for (int j = 0; j < NPL; ++j) if (lane + 32 * j < n && e[j] < mn) mn = e[j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, mn, o); mn = t < mn ? t : mn; }
                new_tau = mn;
            }
            // compact the survivors to the front of the buffer (and to the partial result)
            int off = 0;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const bool keep = lane + 32 * j < n && e[j] >= new_tau;
                const unsigned mk = __ballot_sync(0xffffffffu, keep);
                const int pos = off + __popc(mk & ((1u << lane) - 1u));
                if (keep) {
                    if (n > p.k) base[pos] = e[j];
                    if (emit) {
                        ps[pos] = sim_key_score(e[j]);
                        pi[pos] = (int64_t)(0xffffffffu - (unsigned)(e[j] & 0xffffffffu)) + p.idx_offset;
                    }
                }
                off += __popc(mk);
            }
            if (emit)
                for (int i = m + lane; i < p.k; i += 32) { ps[i] = -INFINITY; pi[i] = -1; }
        } else {
            int np2 = 32;
            while (np2 < n) np2 <<= 1;
            for (int i = lane; i < np2; i += 32) sm[i] = i < n ? __ldcg(base + i) : 0ull;
            __syncwarp();
            warp_bitonic_desc(sm, np2, lane);
            m = n < p.k ? n : p.k;
            for (int i = lane; i < m; i += 32) base[i] = sm[i];
            if (emit) {
                for (int i = lane; i < p.k; i += 32) {
                    const unsigned long long key = i < m ? sm[i] : 0ull;
                    ps[i] = key ? sim_key_score(key) : -INFINITY;
                    pi[i] = key ? (int64_t)(0xffffffffu - (unsigned)(key & 0xffffffffu)) + p.idx_offset : -1;
                }
            }
            new_tau = n >= p.k ? sm[p.k - 1] : 0ull;
        }
        __syncwarp();                                          // buffer / sm are reused by the next row
        if (lane == owner) {
            st.cnt = m;
            st.tau = new_tau;
            st.tau_s = new_tau ? sim_key_score(new_tau) : -INFINITY;
        }
    }

    __device__ static void epilogue(const Params& p, const Tile& t, int rank, uint32_t tmem, int quarter, int lane,
                                    uint8_t* scratch, EpiState& st)
    {
        const int r_in = quarter * 32 + lane;
        const int64_t row = (int64_t)t.qb * 256 + rank * 128 + r_in;
        const bool valid = row < p.n_q;
        unsigned long long* warp_bufs = p.bufs + ((size_t)blockIdx.x * 128 + quarter * 32) * CAP;
        unsigned long long* buf = warp_bufs + (size_t)lane * CAP;
        unsigned long long* sm = reinterpret_cast<unsigned long long*>(scratch) + quarter * CAP;
        if (t.nkb > 0) {
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                __syncwarp();                                  // appends diverge; everything below is warp-collective
                unsigned need = __ballot_sync(0xffffffffu, valid && st.cnt > CAP - 32);
                while (need) {
                    const int owner = __ffs(need) - 1;
                    need &= need - 1;
                    prune_row(p, warp_bufs, sm, owner, lane, st, false, nullptr, nullptr);
                }
                float v[32];
                tmem_ld32(tmem + c, v);
                tmem_ld_wait();
                const int64_t col0 = (int64_t)t.dbb * BLOCK_N + c;
                if (valid && col0 < p.n_db) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float s = v[j];
                        if (s >= st.tau_s && col0 + j < p.n_db) {   // rare once tau has settled
                            const unsigned long long key = sim_key(s, (unsigned)(col0 + j));
                            if (key > st.tau) buf[st.cnt++] = key;
                        }
                    }
                }
            }
        }
        if (t.last) {
            __syncwarp();
            for (int owner = 0; owner < 32; ++owner) {
                const int64_t orow = (int64_t)t.qb * 256 + rank * 128 + quarter * 32 + owner;
                if (orow >= p.n_q) break;                      // warp-uniform
                float* ps = p.part_scores + ((size_t)t.stripe * p.n_q + orow) * p.k;
                int64_t* pi = p.part_idx + ((size_t)t.stripe * p.n_q + orow) * p.k;
                prune_row(p, warp_bufs, sm, owner, lane, st, true, ps, pi);
            }
        }
    }
};


// ------------------------------------------------------------------------------------------------
// fp32-accurate variant: operands are PVS_F16X2 planes (v 2^15 = hi + lo, both fp16), every product
// is hi*lo + lo*hi + hi*hi (three kind::f16 MMAs, 22 mantissa bits) and the accumulation is SEGMENTED.
//
// Why segments: tcgen05.mma adds each K = 16 slice to the TMEM accumulator with truncation, so over a
// d = 32 768 contraction (6 144 accumulation steps) the sum drifts by up to ~4e-4 relative -- a bias that is
// useless for "top-k indices exact".  Here a tile's contraction is cut into segments of `seg` k-blocks
// (64 deep each); the tensor core accumulates one segment from zero, the eight epilogue warps (two per TMEM
// lane quarter, 128 columns each) add the finished segment into fp32 registers with round-to-nearest while
// the next segment is already being multiplied into the other TMEM buffer.  The drift is then bounded by
// (3 * 4 * seg) * 2^-23 of the segment's own partial sums, about 6e-6 of sum|q_i d_i| <= 1 for seg = 4.
// After the last segment the totals go back into the (now idle) accumulator columns with tcgen05.st and the
// first warp of each lane quarter runs the same fused top-k scan as the bf16 kernel.  DENSE = true writes the
// score matrix instead (pvs_cosine_matrix).
// ------------------------------------------------------------------------------------------------
struct SimSplitState : SimState {
    float run[128];                        // running sums of this warp's 128 columns; all zero between tiles
    __device__ SimSplitState()
    {
#pragma unroll
        for (int j = 0; j < 128; ++j) run[j] = 0.f;
    }
};

template <int CAP_, bool DENSE_>
struct SimSplitPolicy : SimPolicy<CAP_> {
    using Base = SimPolicy<CAP_>;
    using Params = SimParams;
    using EpiState = SimSplitState;
    struct Tile { int nkb, kb0, qb, dbb, stripe; bool first, last, first_seg, last_seg; };
    static constexpr bool BF16 = false, F16 = true, MANUAL = false, B_RESIDENT = false, ACC_INIT = false, TILE_SYNC = true, DENSE = DENSE_;
    static constexpr int CAP = CAP_, PASSES = 3, BLOCK_N = 256, KSTEPS = 4, NKB_RES = 0, PGROUPS = 1, EPI_WARPS = 8;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = 128 * 128, TMA_BYTES = 2 * (A_BYTES + B_BYTES);
    static constexpr int SCRATCH_BYTES = (DENSE || CAP <= 512) ? 0 : 4 * CAP * 8;
    static constexpr int STAGES = (226 * 1024 - 1024 - 256 - SCRATCH_BYTES) / TMA_BYTES;
    static_assert(STAGES >= 2, "the split kernel needs at least two operand stages");
    __device__ static void prefetch(const Params& p)
    {
        tma_prefetch_desc(&p.q_map); tma_prefetch_desc(&p.db_map); tma_prefetch_desc(&p.q_lo); tma_prefetch_desc(&p.db_lo);
    }
    __device__ static int num_tiles(const Params& p) { return p.n_units * p.tiles_per_unit * p.nseg; }
    // units are dealt round-robin to pairs; the tiles of a unit and the segments of a tile are consecutive
    __device__ static int tile_at(const Params& p, int it, int pair, int n_pairs, int)
    {
        const int per_unit = p.tiles_per_unit * p.nseg;
        const int j = it / per_unit, w = it - j * per_unit;
        const long long u = (long long)pair + (long long)j * n_pairs;
        return u < p.n_units ? (int)(u * per_unit + w) : -1;
    }
    __device__ static Tile tile(const Params& p, int i)
    {
        const int per_unit = p.tiles_per_unit * p.nseg;
        const int u = i / per_unit, w2 = i - u * per_unit;
        const int w = w2 / p.nseg, sg = w2 - w * p.nseg;
        const int qb = u / p.stripes, s = u - qb * p.stripes;
        const int dbb = s * p.tiles_per_unit + w;
        Tile t;
        t.qb = qb; t.dbb = dbb; t.stripe = s;
        t.kb0 = sg * p.seg;
        const int left = p.nkb - t.kb0;
        t.nkb = dbb < p.db_blocks ? (left < p.seg ? left : p.seg) : 0;     // stripes may overhang the last block
        t.first_seg = sg == 0;
        t.last_seg = sg == p.nseg - 1;
        t.first = w == 0 && t.first_seg;
        t.last = w == p.tiles_per_unit - 1 && t.last_seg;
        return t;
    }
    __device__ static void load(const Params& p, const Tile& t, int kb, int rank, uint8_t* a_hi, uint8_t* a_lo, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        const int c = (t.kb0 + kb) * 64, rq = t.qb * 256 + rank * 128, rd = t.dbb * 256 + rank * 128;
        tma_load_2d_pair(a_hi, &p.q_map, bar, c, rq);
        tma_load_2d_pair(a_lo, &p.q_lo, bar, c, rq);
        tma_load_2d_pair(b_hi, &p.db_map, bar, c, rd);
        tma_load_2d_pair(b_lo, &p.db_lo, bar, c, rd);
    }
    // the drift limiter works on database tiles, not on their segments
    __device__ static void tile_sync(const Params& p, int it, int rank, int n_pairs)
    {
        if (it % p.nseg == 0) Base::tile_sync(p, it / p.nseg, rank, n_pairs);
    }
    __device__ static void tile_sync_done(const Params& p, int tiles_done, int rank) { Base::tile_sync_done(p, tiles_done / p.nseg, rank); }
    __device__ static void epi_begin(const Params&, const Tile& t, EpiState& st, int, int, int)
    {
        if (t.first) { st.tau = 0ull; st.tau_s = -INFINITY; st.cnt = 0; }
    }
    __device__ static void quarter_barrier(int quarter) { asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory"); }

    __device__ __forceinline__ static void epilogue(const Params& p, const Tile& t, int rank, uint32_t tmem, int quarter, int lane,
                                                    uint8_t* scratch, EpiState& st)
    {
        const int half = (((int)threadIdx.x >> 5) - 2) >> 2;       // which 128 of the 256 columns this warp folds
        const int c0 = half * 128;
        const int r_in = quarter * 32 + lane;
        const int64_t row = (int64_t)t.qb * 256 + rank * 128 + r_in;
        const bool valid = row < p.n_q;
        constexpr float UNSCALE = 1.f / 1073741824.f;              // operands carry 2^15 each
        if (t.nkb > 0) {
            // 8 columns at a time: 128 running sums + one chunk must fit the 168 registers a 320-thread CTA gets
            if (!t.last_seg) {
#pragma unroll
                for (int c = 0; c < 128; c += 8) {
                    float v[8];
                    tmem_ld8(tmem + c0 + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) st.run[c + j] += v[j];
                }
                return;
            }
            // last segment: totals
            if constexpr (DENSE) {
                float* orow = p.dense + row * p.ld_dense + (int64_t)t.dbb * BLOCK_N + c0;
                const bool vec = (p.ld_dense & 3) == 0 && ((uintptr_t)p.dense & 15) == 0;
#pragma unroll
                for (int c = 0; c < 128; c += 8) {
                    float v[8];
                    tmem_ld8(tmem + c0 + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) { v[j] = (st.run[c + j] + v[j]) * UNSCALE; st.run[c + j] = 0.f; }
                    const int64_t col0 = (int64_t)t.dbb * BLOCK_N + c0 + c;
                    if (!valid || col0 >= p.n_db) continue;
                    if (vec && col0 + 8 <= p.n_db) {
#pragma unroll
                        for (int j = 0; j < 8; j += 4) *reinterpret_cast<float4*>(orow + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (col0 + j < p.n_db) orow[c + j] = v[j];
                    }
                }
                return;
            } else {
#pragma unroll
                for (int c = 0; c < 128; c += 8) {
                    float v[8];
                    tmem_ld8(tmem + c0 + c, v);
                    tmem_ld_wait();
#pragma unroll
                    // the sums are zero again from here on: nothing is live across the (out-of-line) scan below, and the
                    // first segment of the next tile needs no special case
                    for (int j = 0; j < 8; ++j) { v[j] = (st.run[c + j] + v[j]) * UNSCALE; st.run[c + j] = 0.f; }
                    tmem_st8(tmem + c0 + c, v);
                }
                tmem_st_wait();
                tcgen05_fence_before();
                quarter_barrier(quarter);                          // both column halves of these 32 rows are back in TMEM
                tcgen05_fence_after();
            }
        }
        if constexpr (!DENSE) {
            if (half != 0) return;                                 // the scan + top-k state belong to the first warp of the quarter
            // Out of line and on a COPY of the three top-k words: inlined, the scan's registers are added to the 128
            // running sums of the fold loop (spills in the per-segment path); by reference, the whole state object --
            // sums included -- is forced into local memory.
            SimState ts = static_cast<const SimState&>(st);
            scan_tile(p, t.nkb, t.qb, t.dbb, t.stripe, t.last, rank, tmem, quarter, lane, scratch, ts);
            static_cast<SimState&>(st) = ts;
        }
    }

    // fused top-k over the 256 totals of this warp's 32 rows (now back in TMEM), as in SimPolicy::epilogue
    __device__ __noinline__ static void scan_tile(const Params& p, int nkb, int qb, int dbb, int stripe, bool last, int rank,
                                                  uint32_t tmem, int quarter, int lane, uint8_t* scratch, SimState& ts)
    {
        const int64_t row = (int64_t)qb * 256 + rank * 128 + quarter * 32 + lane;
        const bool valid = row < p.n_q;
        unsigned long long* warp_bufs = p.bufs + ((size_t)blockIdx.x * 128 + quarter * 32) * CAP;
        unsigned long long* buf = warp_bufs + (size_t)lane * CAP;
        unsigned long long* sm = reinterpret_cast<unsigned long long*>(scratch) + quarter * CAP;
        if (nkb > 0) {
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                __syncwarp();
                unsigned need = __ballot_sync(0xffffffffu, valid && ts.cnt > CAP - 32);
                while (need) {
                    const int owner = __ffs(need) - 1;
                    need &= need - 1;
                    Base::prune_row(p, warp_bufs, sm, owner, lane, ts, false, nullptr, nullptr);
                }
                float v[32];
                tmem_ld32(tmem + c, v);
                tmem_ld_wait();
                const int64_t col0 = (int64_t)dbb * BLOCK_N + c;
                if (valid && col0 < p.n_db) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float sc = v[j];
                        if (sc >= ts.tau_s && col0 + j < p.n_db) {
                            const unsigned long long key = sim_key(sc, (unsigned)(col0 + j));
                            if (key > ts.tau) buf[ts.cnt++] = key;
                        }
                    }
                }
            }
        }
        if (last) {
            __syncwarp();
            for (int owner = 0; owner < 32; ++owner) {
                const int64_t orow = (int64_t)qb * 256 + rank * 128 + quarter * 32 + owner;
                if (orow >= p.n_q) break;
                float* ps = p.part_scores + ((size_t)stripe * p.n_q + orow) * p.k;
                int64_t* pi = p.part_idx + ((size_t)stripe * p.n_q + orow) * p.k;
                Base::prune_row(p, warp_bufs, sm, owner, lane, ts, true, ps, pi);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Exact re-evaluation of the candidates the tensor scores cannot order.  Input: per query the shortlist of
// k' = k + margin candidates sorted by tensor score (error <= band).  Two neighbours closer than 2 * band may
// be in the wrong order, and the k-th / (k+1)-th decide membership; every maximal run of such neighbours that
// touches the first k positions is re-scored EXACTLY -- sum over d of (hi + lo)(hi + lo) in fp64 from the same
// operand planes -- and re-ranked (score descending, lowest index first).  One warp per query row.  At d = 32 768
// the band (2e-5, a worst-case bound; the measured error is ~1.5e-6) is of the order of the spacing of the best
// hundred scores of a database of 10^4..10^5 high-dimensional rows, so most of a list may be re-scored: ~130 KB of
// operand reads per candidate, a few per cent of the tensor kernel's time.
// stats[0] += re-scored candidates, stats[1] += rows that could not be certified: the last shortlist entry is still
// within 2 * band of the k-th score (a row beyond the shortlist could belong to the top-k: margin too small -- only
// with degenerate inputs such as many identical rows), or a run of more than 32 chained candidates.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sim_rescue_kernel(const float* __restrict__ s_in, const int64_t* __restrict__ i_in, int kp, int k, int64_t n_q,
                  const __half* __restrict__ q_hi, const __half* __restrict__ q_lo, const __half* __restrict__ db_hi,
                  const __half* __restrict__ db_lo, int64_t d, int64_t idx_offset, float band,
                  float* __restrict__ s_out, int64_t* __restrict__ i_out, unsigned long long* __restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= n_q) return;
    const float* s = s_in + q * kp;
    const int64_t* ix = i_in + q * kp;
    for (int j = lane; j < k; j += 32) {
        s_out[q * k + j] = j < kp ? s[j] : -INFINITY;
        i_out[q * k + j] = j < kp ? ix[j] : -1;
    }
    __syncwarp();
    const float gap = 2.f * band;
    // Membership: a candidate ranked past position k by the tensor scores can belong to the true top-k only if its
    // tensor score is within 2 * band of the k-th one; anything below `cut` is certainly out, whatever chain it hangs on.
    // The shortlist is long enough iff its last entry is below `cut` (or the database had fewer than kp rows).
    const bool full = kp > k && ix[kp - 1] >= 0;
    const float cut = (k <= kp && ix[k - 1] >= 0) ? s[k - 1] - gap : -INFINITY;
    if (full && s[kp - 1] >= cut && lane == 0) atomicAdd(&stats[1], 1ull);
    int i = 0;
    while (i < k && i < kp) {
        int j = i;
        // warp-uniform scan (same loads in every lane); past position k - 1 only possible members extend a run
        while (j + 1 < kp && ix[j + 1] >= 0 && s[j] - s[j + 1] <= gap && (j + 1 < k || s[j + 1] >= cut)) ++j;
        if (j == i) { ++i; continue; }
        const int n = j - i + 1;
        if (n > 32) {                                              // e.g. many identical rows: bit-equal tensor scores are already in index order
            if (lane == 0) atomicAdd(&stats[1], 1ull);
            i = j + 1;
            continue;
        }
        double mine = 0.0;
        int64_t my_idx = -1;
        for (int m = 0; m < n; ++m) {
            const int64_t id = ix[i + m];
            const __half* dh = db_hi + (id - idx_offset) * d;
            const __half* dl = db_lo + (id - idx_offset) * d;
            const __half* qh = q_hi + q * d;
            const __half* ql = q_lo + q * d;
            double acc = 0.0;
            for (int64_t e = (int64_t)lane * 8; e < d; e += 256) {
                const uint4 a = *reinterpret_cast<const uint4*>(qh + e), b = *reinterpret_cast<const uint4*>(ql + e);
                const uint4 c = *reinterpret_cast<const uint4*>(dh + e), f = *reinterpret_cast<const uint4*>(dl + e);
                const __half2* a2 = reinterpret_cast<const __half2*>(&a);
                const __half2* b2 = reinterpret_cast<const __half2*>(&b);
                const __half2* c2 = reinterpret_cast<const __half2*>(&c);
                const __half2* f2 = reinterpret_cast<const __half2*>(&f);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 qa = __half22float2(a2[u]), qb = __half22float2(b2[u]);
                    const float2 da = __half22float2(c2[u]), db = __half22float2(f2[u]);
                    acc = fma((double)(qa.x + qb.x), (double)(da.x + db.x), acc);      // hi + lo is exact in fp32
                    acc = fma((double)(qa.y + qb.y), (double)(da.y + db.y), acc);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == m) { mine = acc; my_idx = id; }
        }
        // rank inside the run: score descending, lowest index first
        int rank = 0;
        for (int m = 0; m < n; ++m) {
            const double os = __shfl_sync(0xffffffffu, mine, m);
            const long long oi = __shfl_sync(0xffffffffu, (long long)my_idx, m);
            if (lane < n && (os > mine || (os == mine && oi < (long long)my_idx))) ++rank;
        }
        if (lane < n && i + rank < k) {
            s_out[q * k + i + rank] = (float)(mine * (1.0 / 1073741824.0));
            i_out[q * k + i + rank] = my_idx;
        }
        if (lane == 0) atomicAdd(&stats[0], (unsigned long long)n);
        i = j + 1;
    }
}
}  // namespace tc2

using namespace tc2;

struct SimPlan { int q_blocks, db_blocks, stripes, tiles_per_unit, n_units, pairs, cap; size_t bufs, ps, pi, ctr, total; };

static int sim_cap(int k)
{
    int c = 512;
    while (c < 2 * k) c <<= 1;
    return c;                                                 // 512 .. 2048 for k <= 1024
}

static SimPlan sim_plan(int64_t n_q, int64_t n_db, int k)
{
    SimPlan pl{};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_pairs = sms / 2;
    pl.cap = sim_cap(k);
    pl.q_blocks = (int)ceil_div(n_q, 256);
    pl.db_blocks = (int)ceil_div(n_db, 256);
    int s = 1;
    while (s * s < n_pairs) ++s;                              // ~sqrt(pairs) stripes: square working set in L2
    const int s_fill = (int)ceil_div(3 * n_pairs, pl.q_blocks);   // enough units to balance the machine
    if (s < s_fill) s = s_fill;
    if (const char* e = getenv("PVS_SIM_STRIPES")) { const int v = atoi(e); if (v > 0) s = v; }   // tuning hook
    const int s_max_merge = 8192 / k > 0 ? 8192 / k : 1;      // bitonic merge handles parts*k <= 8192
    if (s > s_max_merge) s = s_max_merge;
    if (s > pl.db_blocks) s = pl.db_blocks;
    if (s < 1) s = 1;
    pl.tiles_per_unit = (int)ceil_div(pl.db_blocks, s);
    pl.stripes = (int)ceil_div(pl.db_blocks, pl.tiles_per_unit);
    pl.n_units = pl.q_blocks * pl.stripes;
    pl.pairs = pl.n_units < n_pairs ? pl.n_units : n_pairs;
    size_t off = 0;
    pl.bufs = off; off += align_up((size_t)pl.pairs * 2 * 128 * pl.cap * 8, 1024);
    pl.ps = off;   off += align_up((size_t)pl.stripes * n_q * k * 4, 1024);
    pl.pi = off;   off += align_up((size_t)pl.stripes * n_q * k * 8, 1024);
    pl.ctr = off;  off += 1024;
    pl.total = off + 1024;
    return pl;
}

bool tc_sim_supported(int dtype, int64_t n_q, int64_t n_db, int64_t d, int k)
{
    return tc_available() && dtype == PVS_BF16 && d % 8 == 0 && d >= 64 && k <= PVS_TOPK_MAX && n_q > 0 && n_db > 0 &&
           n_q < 2147483000LL && n_db < 2147483000LL && d < 2147483000LL &&
           ceil_div(n_q, 256) * (ceil_div(n_db, 256) + 128) < 2000000000LL;
}

size_t tc_sim_workspace_bytes(int64_t n_q, int64_t n_db, int k) { return sim_plan(n_q, n_db, k).total; }

int tc_sim_topk(const void* q, const void* db, int64_t n_q, int64_t n_db, int64_t d, int k, int64_t idx_offset,
                float* scores_out, int64_t* idx_out, void* ws, size_t ws_bytes, cudaStream_t st)
{
    PVS_CHECK((((uintptr_t)q | (uintptr_t)db) & 15) == 0, PVS_ERR_BAD_ARG, "bf16 operands must be 16-byte aligned");
    const SimPlan pl = sim_plan(n_q, n_db, k);
    PVS_CHECK(ws && ws_bytes >= pl.total, PVS_ERR_WORKSPACE, "similarity workspace %zu < %zu", ws_bytes, pl.total);
    char* base = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    SimParams p{};
    int rc;
    if ((rc = make_tmap_2d(&p.q_map, q, true, n_q, d, d, 64, 128))) return rc;
    if ((rc = make_tmap_2d(&p.db_map, db, true, n_db, d, d, 64, 128))) return rc;
    p.bufs = (unsigned long long*)(base + pl.bufs);
    p.part_scores = (float*)(base + pl.ps);
    p.part_idx = (int64_t*)(base + pl.pi);
    p.n_q = n_q; p.n_db = n_db; p.idx_offset = idx_offset; p.k = k; p.cap = pl.cap;
    p.nkb = (int)ceil_div(d, 64);
    p.q_blocks = pl.q_blocks; p.db_blocks = pl.db_blocks; p.stripes = pl.stripes;
    p.tiles_per_unit = pl.tiles_per_unit; p.n_units = pl.n_units;
    const int n_tiles = pl.n_units * pl.tiles_per_unit;
    p.sync_ctr = getenv("PVS_SIM_NOSYNC") ? nullptr : (unsigned*)(base + pl.ctr);
    p.sync_steps = (int)ceil_div(pl.n_units, pl.pairs) * pl.tiles_per_unit;
    PVS_CUDA(cudaMemsetAsync(base + pl.ctr, 0, 1024, st));
    switch (pl.cap) {
        case 512: rc = launch_tc2<SimPolicy<512>>(p, n_tiles, st, pl.pairs); break;
        case 1024: rc = launch_tc2<SimPolicy<1024>>(p, n_tiles, st, pl.pairs); break;
        default: rc = launch_tc2<SimPolicy<2048>>(p, n_tiles, st, pl.pairs); break;
    }
    if (rc) return rc;
    return launch_topk_merge(p.part_scores, p.part_idx, pl.stripes, n_q, k, scores_out, idx_out, st);
}

}  // namespace pvs

// =====================================================================================
// fp32-accurate similarity on split operands (SimSplitPolicy): fused top-k with exact tie
// resolution, and the dense score matrix
// =====================================================================================
namespace pvs {
using namespace tc2;

constexpr int SIM3_MARGIN = 24;                                // shortlist = k + margin candidates per query

static int sim3_seg()
{
    int g = 4;
    if (const char* e = getenv("PVS_SIM_SEG")) { const int v = atoi(e); if (v > 0) g = v; }
    return g;
}

// Band on |tensor score - exact score| for unit-norm rows.  Model: every MMA (3 per k-step, 4 k-steps per 64-deep k-block,
// `seg` k-blocks per segment) adds its slice to the accumulator with an error of at most one ulp of the segment's partial
// sum, which is bounded by the segment's share of sum |q_i d_i| <= 1; every fold of a segment into the fp32 running sum
// rounds to nearest (half an ulp of a value <= 1); the hi + lo split and the dropped lo * lo term cost 2^-21.  The first two
// terms are worst cases that assume every step loses a full ulp in the same direction at full magnitude; measured on VLAD-
// shaped and Gaussian rows at d = 64 .. 32 768 the error stays below a tenth of their sum (1.5e-6 at d = 32 768), so half of
// it is used -- 7e-6 at d = 32 768 instead of 1.4e-5, which matters because every candidate pair closer than twice the band
// is re-scored exactly (131 KB of operand reads per candidate).  The tests compare the returned indices with the fp64
// ranking, so a band that is too small shows up as an index mismatch, not as a silent error.
static float sim3_band(int64_t d, int seg)
{
    const int nkb = (int)ceil_div(d, 64), nseg = (int)ceil_div(nkb, seg);
    return 0.5f * ((12.f * seg) * 1.1920929e-7f + (float)nseg * 5.9604645e-8f) + 4.7683716e-7f;
}

bool tc_sim3_supported(int64_t n_q, int64_t n_db, int64_t d, int k)
{
    if (!(tc_available() && d % 8 == 0 && d >= 64 && k >= 1 && k + SIM3_MARGIN <= PVS_TOPK_MAX && n_q > 0 && n_db > 0 &&
          n_q < 2147483000LL && n_db < 2147483000LL && d < 2147483000LL)) return false;
    const int64_t nseg = ceil_div(ceil_div(d, 64), sim3_seg());
    return ceil_div(n_q, 256) * (ceil_div(n_db, 256) + 128) * nseg < 2000000000LL;
}

struct Sim3Ws { SimPlan pl; int kp; size_t ms, mi, stats, total; };
static Sim3Ws sim3_ws(int64_t n_q, int64_t n_db, int k)
{
    Sim3Ws w{};
    w.kp = k + SIM3_MARGIN;
    w.pl = sim_plan(n_q, n_db, w.kp);
    size_t off = w.pl.total;
    w.ms = off;    off += align_up((size_t)n_q * w.kp * 4, 1024);
    w.mi = off;    off += align_up((size_t)n_q * w.kp * 8, 1024);
    w.stats = off; off += 1024;
    w.total = off + 1024;
    return w;
}
size_t tc_sim3_workspace_bytes(int64_t n_q, int64_t n_db, int k) { return sim3_ws(n_q, n_db, k).total; }

static int sim3_fill(SimParams& p, const void* q, const void* db, int64_t n_q, int64_t n_db, int64_t d)
{
    PVS_CHECK((((uintptr_t)q | (uintptr_t)db) & 15) == 0, PVS_ERR_BAD_ARG, "split operands must be 16-byte aligned");
    const __half* qh = (const __half*)q;
    const __half* dh = (const __half*)db;
    int rc;
    if ((rc = make_tmap_2d(&p.q_map, qh, true, n_q, d, d, 64, 128))) return rc;
    if ((rc = make_tmap_2d(&p.q_lo, qh + n_q * d, true, n_q, d, d, 64, 128))) return rc;
    if ((rc = make_tmap_2d(&p.db_map, dh, true, n_db, d, d, 64, 128))) return rc;
    if ((rc = make_tmap_2d(&p.db_lo, dh + n_db * d, true, n_db, d, d, 64, 128))) return rc;
    p.n_q = n_q; p.n_db = n_db;
    p.nkb = (int)ceil_div(d, 64);
    p.seg = sim3_seg();
    p.nseg = (int)ceil_div(p.nkb, p.seg);
    return PVS_OK;
}

// q / db: PVS_F16X2 planes ([2, n, d] fp16).  ws[stats] (two uint64 at the START of the aligned workspace's stats
// slot, returned through stats_out_dev when given): re-scored candidates, rows with an unresolved run.
int tc_sim3_topk(const void* q, const void* db, int64_t n_q, int64_t n_db, int64_t d, int k, int64_t idx_offset,
                 float* scores_out, int64_t* idx_out, void* ws, size_t ws_bytes, cudaStream_t st)
{
    const Sim3Ws w = sim3_ws(n_q, n_db, k);
    const SimPlan& pl = w.pl;
    PVS_CHECK(ws && ws_bytes >= w.total, PVS_ERR_WORKSPACE, "similarity workspace %zu < %zu", ws_bytes, w.total);
    char* base = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    SimParams p{};
    if (int rc = sim3_fill(p, q, db, n_q, n_db, d)) return rc;
    p.bufs = (unsigned long long*)(base + pl.bufs);
    p.part_scores = (float*)(base + pl.ps);
    p.part_idx = (int64_t*)(base + pl.pi);
    p.idx_offset = idx_offset; p.k = w.kp; p.cap = pl.cap;
    p.q_blocks = pl.q_blocks; p.db_blocks = pl.db_blocks; p.stripes = pl.stripes;
    p.tiles_per_unit = pl.tiles_per_unit; p.n_units = pl.n_units;
    const int n_tiles = pl.n_units * pl.tiles_per_unit * p.nseg;
    p.sync_ctr = getenv("PVS_SIM_NOSYNC") ? nullptr : (unsigned*)(base + pl.ctr);
    p.sync_steps = (int)ceil_div(pl.n_units, pl.pairs) * pl.tiles_per_unit;
    PVS_CUDA(cudaMemsetAsync(base + pl.ctr, 0, 1024, st));
    PVS_CUDA(cudaMemsetAsync(base + w.stats, 0, 1024, st));
    int rc;
    switch (pl.cap) {
        case 512: rc = launch_tc2<SimSplitPolicy<512, false>>(p, n_tiles, st, pl.pairs); break;
        case 1024: rc = launch_tc2<SimSplitPolicy<1024, false>>(p, n_tiles, st, pl.pairs); break;
        default: rc = launch_tc2<SimSplitPolicy<2048, false>>(p, n_tiles, st, pl.pairs); break;
    }
    if (rc) return rc;
    float* ms = (float*)(base + w.ms);
    int64_t* mi = (int64_t*)(base + w.mi);
    if ((rc = launch_topk_merge(p.part_scores, p.part_idx, pl.stripes, n_q, w.kp, ms, mi, st))) return rc;
    const __half* qh = (const __half*)q;
    const __half* dh = (const __half*)db;
    PVS_LAUNCH(sim_rescue_kernel, (unsigned)ceil_div(n_q, 8), 256, 0, st, ms, mi, w.kp, k, n_q, qh, qh + n_q * d, dh, dh + n_db * d, d,
               idx_offset, sim3_band(d, p.seg), scores_out, idx_out, (unsigned long long*)(base + w.stats));
    return PVS_OK;
}

// offset of the two uint64 counters inside a workspace of tc_sim3_workspace_bytes() (after 1024-byte alignment)
size_t tc_sim3_stats_offset(int64_t n_q, int64_t n_db, int k) { return sim3_ws(n_q, n_db, k).stats; }

// dense score matrix s[n_q, ld] = Q . DB^T from split planes
int tc_sim3_dense(const void* q, const void* db, int64_t n_q, int64_t n_db, int64_t d, float* s, int64_t ld, cudaStream_t st)
{
    SimParams p{};
    if (int rc = sim3_fill(p, q, db, n_q, n_db, d)) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    p.dense = s; p.ld_dense = ld; p.k = 1; p.cap = 512;
    p.q_blocks = (int)ceil_div(n_q, 256); p.db_blocks = (int)ceil_div(n_db, 256);
    p.stripes = p.db_blocks; p.tiles_per_unit = 1;             // one database block per unit: every (query block, block) pair is a tile
    p.n_units = p.q_blocks * p.stripes;
    PVS_CHECK((int64_t)p.q_blocks * p.stripes * p.nseg < 2000000000LL, PVS_ERR_BAD_SHAPE, "score matrix too large for one launch");
    p.sync_ctr = nullptr;
    const int pairs = p.n_units < sms / 2 ? p.n_units : sms / 2;
    return launch_tc2<SimSplitPolicy<512, true>>(p, p.n_units * p.nseg, st, pairs);
}

}  // namespace pvs
