// pvs_tc_sim.cu -- all-pairs cosine similarity with fused per-row top-k on tcgen05 (bf16).
//
// Replaces `cosine_similarity(q, all)` + `np.argsort(-s)[:k]` per query
// (pyvisim/eval.py:37-43, 76-80, 131-132) by one sweep: scores = Qn . DBn^T as a bf16
// tensor-core contraction with fp32 accumulation in TMEM; the score matrix is never
// written.  Each epilogue thread owns one query row (= one TMEM lane), compares the 256
// scores of every tile against the row's current k-th best and pushes the rare survivors
// into a k-entry min-heap.  Ordering key = (score, lowest index first), the same total
// order as the CUDA-core path, so both paths return identical lists up to score rounding.
//
// Work decomposition: unit = (block of 128 query rows) x (stripe of database blocks).  A CTA
// sweeps all tiles of a unit back to back so the heap state stays in registers / L2;
// units that run concurrently share query blocks and database blocks through L2.  The S
// partial lists per row are merged by the bitonic merge kernel (pvs_simt.cu).
#include "pvs_tc.cuh"
#include "pvs_kernels.cuh"

namespace pvs {
namespace tc {

struct SimParams {
    CUtensorMap q_map, db_map;
    unsigned long long* heaps;     // [grid, 128, k] scratch
    float* part_scores;            // [stripes, n_q, k]
    int64_t* part_idx;             // [stripes, n_q, k]
    int64_t n_q, n_db, idx_offset;
    int k, nkb, q_blocks, db_blocks, stripes, tiles_per_unit, n_units;
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
    unsigned long long tau;   // heap root = worst of the current top-k (0 while the heap is filling)
    float tau_s;              // its score, for the cheap first-level reject
    int count;
};

struct SimPolicy {
    using Params = SimParams;
    using EpiState = SimState;
    struct Tile { int nkb, qb, dbb, stripe; bool first, last; };
    static constexpr bool BF16 = true, A_MN = false, B_MN = false, EPI_READS_STAGES = false, MANUAL = false;
    static constexpr int PASSES = 1, BLOCK_N = 256, KSTEPS = 4, STAGES = 4;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = BLOCK_N * 128, A_LBO = 0, B_LBO = 0, SCRATCH_BYTES = 0;
    static constexpr int TMA_BYTES = A_BYTES + B_BYTES;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.q_map); tma_prefetch_desc(&p.db_map); }
    __device__ static int num_tiles(const Params& p) { return p.n_units * p.tiles_per_unit; }
    // units are dealt round-robin to CTAs; the tiles of a unit are consecutive
    __device__ static int tile_at(const Params& p, int it, int)
    {
        const int j = it / p.tiles_per_unit, w = it - j * p.tiles_per_unit;
        const long long u = (long long)blockIdx.x + (long long)j * gridDim.x;
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
    __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* a, uint8_t*, uint8_t* b, uint8_t*,
                                uint64_t* bar)
    {
        tma_load_2d(a, &p.q_map, bar, kb * 64, t.qb * 128);
        tma_load_2d(b, &p.db_map, bar, kb * 64, t.dbb * BLOCK_N);
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile& t, EpiState& st, int, int)
    {
        if (t.first) { st.tau = 0ull; st.tau_s = -INFINITY; st.count = 0; }
    }

    __device__ static void sift_down(unsigned long long* h, int n, int i)
    {
        const unsigned long long x = h[i];
        while (true) {
            int c = 2 * i + 1;
            if (c >= n) break;
            unsigned long long cv = h[c];
            if (c + 1 < n) { const unsigned long long r = h[c + 1]; if (r < cv) { cv = r; ++c; } }
            if (cv >= x) break;
            h[i] = cv;
            i = c;
        }
        h[i] = x;
    }

    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane, uint8_t*,
                                    EpiState& st)
    {
        const int r_in = quarter * 32 + lane;
        const int64_t row = (int64_t)t.qb * 128 + r_in;
        const bool valid = row < p.n_q;
        unsigned long long* heap = p.heaps + ((size_t)blockIdx.x * 128 + r_in) * p.k;
        if (t.nkb > 0) {
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                float v[32];
                __syncwarp();                                            // heap updates diverge; tcgen05.ld is warp-collective
                tmem_ld32(tmem + c, v);
                tmem_ld_wait();
                const int64_t col0 = (int64_t)t.dbb * BLOCK_N + c;
                if (!valid || col0 >= p.n_db) continue;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float s = v[j];
                    if (!(s >= st.tau_s)) continue;                      // common case: below the k-th best
                    if (col0 + j >= p.n_db) continue;
                    const unsigned long long key = sim_key(s, (unsigned)(col0 + j));
                    if (st.count < p.k) {
                        heap[st.count++] = key;
                        if (st.count == p.k) {
                            for (int i = p.k / 2 - 1; i >= 0; --i) sift_down(heap, p.k, i);
                            st.tau = heap[0];
                            st.tau_s = sim_key_score(st.tau);
                        }
                    } else if (key > st.tau) {
                        heap[0] = key;
                        sift_down(heap, p.k, 0);
                        st.tau = heap[0];
                        st.tau_s = sim_key_score(st.tau);
                    }
                }
            }
            __syncwarp();
        }
        if (t.last && valid) {
            float* ps = p.part_scores + ((size_t)t.stripe * p.n_q + row) * p.k;
            int64_t* pi = p.part_idx + ((size_t)t.stripe * p.n_q + row) * p.k;
            for (int i = 0; i < p.k; ++i) {
                if (i < st.count) {
                    const unsigned long long key = heap[i];
                    ps[i] = sim_key_score(key);
                    pi[i] = (int64_t)(0xffffffffu - (unsigned)(key & 0xffffffffu)) + p.idx_offset;
                } else {
                    ps[i] = -INFINITY;
                    pi[i] = -1;
                }
            }
        }
    }
};

}  // namespace tc

using namespace tc;

struct SimPlan { int q_blocks, db_blocks, stripes, tiles_per_unit, n_units, grid; size_t heaps, ps, pi, total; };

static SimPlan sim_plan(int64_t n_q, int64_t n_db, int k)
{
    SimPlan pl{};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    pl.q_blocks = (int)ceil_div(n_q, 128);
    pl.db_blocks = (int)ceil_div(n_db, 256);
    int s = (int)ceil_div(3 * sms, pl.q_blocks);              // enough units to balance the machine
    const int s_max_merge = 8192 / k > 0 ? 8192 / k : 1;      // bitonic merge handles parts*k <= 8192
    if (s > s_max_merge) s = s_max_merge;
    if (s > pl.db_blocks) s = pl.db_blocks;
    if (s < 1) s = 1;
    pl.tiles_per_unit = (int)ceil_div(pl.db_blocks, s);
    pl.stripes = (int)ceil_div(pl.db_blocks, pl.tiles_per_unit);
    pl.n_units = pl.q_blocks * pl.stripes;
    pl.grid = pl.n_units < sms ? pl.n_units : sms;
    size_t off = 0;
    pl.heaps = off; off += align_up((size_t)pl.grid * 128 * k * 8, 1024);
    pl.ps = off;    off += align_up((size_t)pl.stripes * n_q * k * 4, 1024);
    pl.pi = off;    off += align_up((size_t)pl.stripes * n_q * k * 8, 1024);
    pl.total = off + 1024;
    return pl;
}

bool tc_sim_supported(int dtype, int64_t n_q, int64_t n_db, int64_t d, int k)
{
    return tc_available() && dtype == PVS_BF16 && d % 8 == 0 && d >= 64 && k <= PVS_TOPK_MAX && n_q > 0 && n_db > 0 &&
           n_q < 2147483000LL && n_db < 2147483000LL && d < 2147483000LL &&
           ceil_div(n_q, 128) * ceil_div(n_db, 256) < 2000000000LL;
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
    if ((rc = make_tmap_2d(&p.db_map, db, true, n_db, d, d, 64, 256))) return rc;
    p.heaps = (unsigned long long*)(base + pl.heaps);
    p.part_scores = (float*)(base + pl.ps);
    p.part_idx = (int64_t*)(base + pl.pi);
    p.n_q = n_q; p.n_db = n_db; p.idx_offset = idx_offset; p.k = k;
    p.nkb = (int)ceil_div(d, 64);
    p.q_blocks = pl.q_blocks; p.db_blocks = pl.db_blocks; p.stripes = pl.stripes;
    p.tiles_per_unit = pl.tiles_per_unit; p.n_units = pl.n_units;
    if ((rc = launch_tc<SimPolicy>(p, pl.n_units * pl.tiles_per_unit, st, pl.grid))) return rc;
    return launch_topk_merge(p.part_scores, p.part_idx, pl.stripes, n_q, k, scores_out, idx_out, st);
}

}  // namespace pvs
