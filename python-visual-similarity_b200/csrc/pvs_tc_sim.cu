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
    CUtensorMap q_map, db_map;
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
