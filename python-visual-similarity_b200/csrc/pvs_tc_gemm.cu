// pvs_tc_gemm.cu -- host side of the tcgen05 machinery (tensor-map encoding) and the plain
// contraction policies used to validate it: C = A * B^T with K-major operands (tf32 single
// pass, 3xTF32, bf16) and C = A^T * B with MN-major operands (tf32).  Exposed through
// pvs_debug_tc_gemm for the GPU tests; the production kernels (pvs_tc_fv.cu,
// pvs_tc_sim.cu) instantiate the same skeleton with fused epilogues.
#include "pvs_tc.cuh"
#include "pvs_tc2.cuh"
#include "pvs_kernels.cuh"

namespace pvs {
namespace tc {

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, bool bf16, int64_t rows, int64_t cols, int64_t ld,
                 int box_cols, int box_rows, bool mn_major)
{
    EncodeTiledFn fn = encode_tiled_fn();
    PVS_CHECK(fn, PVS_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const int esz = bf16 ? 2 : 4;
    PVS_CHECK(box_cols * esz == 128, PVS_ERR_BAD_ARG, "TMA box must span exactly one 128-B swizzle row");
    PVS_CHECK(((uintptr_t)base & 15) == 0 && (ld * esz) % 16 == 0, PVS_ERR_BAD_SHAPE,
              "TMA needs a 16-B aligned base and row pitch (ld=%lld)", (long long)ld);
    PVS_CHECK(box_rows >= 1 && box_rows <= 256, PVS_ERR_BAD_ARG, "TMA box rows %d out of range", box_rows);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)(rows > 0 ? rows : 1)};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PVS_CHECK(r == CUDA_SUCCESS, PVS_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PVS_OK;
}

// ---------------------------------------------------------------------------------------
// validation policies
// ---------------------------------------------------------------------------------------
struct GemmParams {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;
    float* c;
    int64_t ldc;
    int m, n, k;
    int m_blocks, n_blocks;
};
struct NoState {};

// C[m,n] = sum_k A[m,k] B[n,k];  A [M,K], B [N,K] row-major (K-major operands)
template <bool BF16_, int PASSES_, int BLOCK_N_>
struct GemmKMajor {
    using Params = GemmParams;
    using EpiState = NoState;
    struct Tile { int nkb, mb, nb; };
    static constexpr bool BF16 = BF16_, A_MN = false, B_MN = false, EPI_READS_STAGES = false, MANUAL = false;
    static constexpr int PASSES = PASSES_, BLOCK_N = BLOCK_N_, KSTEPS = 4, PGROUPS = 1;
    static constexpr int BK = BF16 ? 64 : 32;                       // elements per 128-B span
    static constexpr int A_BYTES = 128 * 128, B_BYTES = BLOCK_N * 128, A_LBO = 0, B_LBO = 0, SCRATCH_BYTES = 0;
    static constexpr int STAGE_ = (PASSES == 3 ? 2 : 1) * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (200 * 1024) / STAGE_ >= 4 ? 4 : (200 * 1024) / STAGE_;
    static constexpr int TMA_BYTES = STAGE_;
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.a_hi); tma_prefetch_desc(&p.b_hi); }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks * p.n_blocks; }
    __device__ static int tile_at(const Params&, int it, int n) { return strided_tile(it, n); }
    __device__ static Tile tile(const Params& p, int i) { return {p.k / BK, i / p.n_blocks, i % p.n_blocks}; }
    __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* a_hi, uint8_t* a_lo, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        tma_load_2d(a_hi, &p.a_hi, bar, kb * BK, t.mb * 128);
        tma_load_2d(b_hi, &p.b_hi, bar, kb * BK, t.nb * BLOCK_N);
        if constexpr (PASSES == 3) {
            tma_load_2d(a_lo, &p.a_lo, bar, kb * BK, t.mb * 128);
            tma_load_2d(b_lo, &p.b_lo, bar, kb * BK, t.nb * BLOCK_N);
        }
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane, uint8_t*,
                                    EpiState&)
    {
        const int row = t.mb * 128 + quarter * 32 + lane;
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + c0, v);
            tmem_ld_wait();
            if (row < p.m)
                for (int j = 0; j < 32; ++j) {
                    const int col = t.nb * BLOCK_N + c0 + j;
                    if (col < p.n) p.c[(int64_t)row * p.ldc + col] = v[j];
                }
        }
    }
};

// C[m,n] = sum_k A[k,m] B[k,n];  A [K,M], B [K,N] row-major (MN-major operands), tf32 only
template <int PASSES_, int BLOCK_N_>
struct GemmMNMajor {
    using Params = GemmParams;
    using EpiState = NoState;
    struct Tile { int nkb, mb, nb; };
    static constexpr bool BF16 = false, A_MN = true, B_MN = true, EPI_READS_STAGES = false, MANUAL = false;
    static constexpr int KT = 32;                                   // contraction rows per stage
    static constexpr int PASSES = PASSES_, BLOCK_N = BLOCK_N_, STAGES = 2, KSTEPS = KT / 8, PGROUPS = 1;
    static constexpr int A_LBO = KT * 128, B_LBO = KT * 128;        // one [KT x 128 B] box per 32 columns
    static constexpr int A_BYTES = 4 * A_LBO, B_BYTES = (BLOCK_N / 32) * B_LBO, SCRATCH_BYTES = 0;
    static constexpr int TMA_BYTES = (PASSES == 3 ? 2 : 1) * (A_BYTES + B_BYTES);
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.a_hi); tma_prefetch_desc(&p.b_hi); }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks * p.n_blocks; }
    __device__ static int tile_at(const Params&, int it, int n) { return strided_tile(it, n); }
    __device__ static Tile tile(const Params& p, int i) { return {(p.k + KT - 1) / KT, i / p.n_blocks, i % p.n_blocks}; }
    __device__ static void load(const Params& p, const Tile& t, int kb, uint8_t* a_hi, uint8_t* a_lo, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        for (int b = 0; b < 4; ++b) {
            tma_load_2d(a_hi + b * A_LBO, &p.a_hi, bar, t.mb * 128 + 32 * b, kb * KT);
            if constexpr (PASSES == 3) tma_load_2d(a_lo + b * A_LBO, &p.a_lo, bar, t.mb * 128 + 32 * b, kb * KT);
        }
        for (int b = 0; b < BLOCK_N / 32; ++b) {
            tma_load_2d(b_hi + b * B_LBO, &p.b_hi, bar, t.nb * BLOCK_N + 32 * b, kb * KT);
            if constexpr (PASSES == 3) tma_load_2d(b_lo + b * B_LBO, &p.b_lo, bar, t.nb * BLOCK_N + 32 * b, kb * KT);
        }
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, uint32_t tmem, int quarter, int lane, uint8_t* s,
                                    EpiState& st)
    {
        GemmKMajor<false, PASSES_, BLOCK_N_>::epilogue(p, {t.nkb, t.mb, t.nb}, tmem, quarter, lane, s, st);
    }
};

// CTA-pair variant (cta_group::2): the pair computes a [256 x BLOCK_N] tile; CTA `rank` loads
// A rows [256 mb + 128 rank, +128) and B rows [BLOCK_N nb + BLOCK_N/2 rank, +BLOCK_N/2).
// RES_: keep all k-blocks of B resident (k <= 32 * NKB_RES) to validate the resident path.
template <bool BF16_, int PASSES_, int BLOCK_N_, bool RES_>
struct GemmPair {
    using Params = GemmParams;
    using EpiState = NoState;
    struct Tile { int nkb, mb, nb; };
    static constexpr bool BF16 = BF16_, MANUAL = false, B_RESIDENT = RES_, ACC_INIT = false, TILE_SYNC = false;
    static constexpr int PASSES = PASSES_, BLOCK_N = BLOCK_N_, KSTEPS = 4, NKB_RES = 4, PGROUPS = 1;
    static constexpr int BK = BF16 ? 64 : 32;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = (BLOCK_N / 2) * 128, SCRATCH_BYTES = 0;
    static constexpr int PARTS_ = PASSES == 3 ? 2 : 1;
    static constexpr int TMA_BYTES = PARTS_ * (A_BYTES + (RES_ ? 0 : B_BYTES));
    static constexpr int STAGES = RES_ ? 2 : ((190 * 1024) / TMA_BYTES >= 4 ? 4 : (190 * 1024) / TMA_BYTES);
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.a_hi); tma_prefetch_desc(&p.b_hi); }
    __device__ static int num_tiles(const Params& p) { return p.m_blocks * p.n_blocks; }
    __device__ static int tile_at(const Params&, int it, int pair, int n_pairs, int n)
    {
        const long long t = (long long)pair + (long long)it * n_pairs;
        return t < n ? (int)t : -1;
    }
    __device__ static Tile tile(const Params& p, int i) { return {p.k / BK, i / p.n_blocks, i % p.n_blocks}; }
    __device__ static void load_resident(const Params& p, int rank, uint8_t* res, uint64_t* bar)
    {
        for (int kb = 0; kb < NKB_RES; ++kb) {
            tc2::tma_load_2d_pair(res + (kb * PARTS_) * B_BYTES, &p.b_hi, bar, kb * BK, rank * (BLOCK_N / 2));
            if constexpr (PASSES == 3)
                tc2::tma_load_2d_pair(res + (kb * PARTS_ + 1) * B_BYTES, &p.b_lo, bar, kb * BK, rank * (BLOCK_N / 2));
        }
    }
    __device__ static void load(const Params& p, const Tile& t, int kb, int rank, uint8_t* a_hi, uint8_t* a_lo, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        tc2::tma_load_2d_pair(a_hi, &p.a_hi, bar, kb * BK, t.mb * 256 + rank * 128);
        if constexpr (PASSES == 3) tc2::tma_load_2d_pair(a_lo, &p.a_lo, bar, kb * BK, t.mb * 256 + rank * 128);
        if constexpr (!RES_) {
            tc2::tma_load_2d_pair(b_hi, &p.b_hi, bar, kb * BK, t.nb * BLOCK_N + rank * (BLOCK_N / 2));
            if constexpr (PASSES == 3) tc2::tma_load_2d_pair(b_lo, &p.b_lo, bar, kb * BK, t.nb * BLOCK_N + rank * (BLOCK_N / 2));
        }
    }
    __device__ static void epi_init(const Params&, uint8_t*, int) {}
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, int rank, uint32_t tmem, int quarter, int lane, uint8_t*,
                                    EpiState&)
    {
        const int row = t.mb * 256 + rank * 128 + quarter * 32 + lane;
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + c0, v);
            tmem_ld_wait();
            if (row < p.m)
                for (int j = 0; j < 32; ++j) {
                    const int col = t.nb * BLOCK_N + c0 + j;
                    if (col < p.n) p.c[(int64_t)row * p.ldc + col] = v[j];
                }
        }
    }
};

template <class P>
static int run_gemm(const GemmParams& prm, cudaStream_t st) { return launch_tc<P>(prm, prm.m_blocks * prm.n_blocks, st); }
template <class P>
static int run_gemm2(const GemmParams& prm, cudaStream_t st) { return tc2::launch_tc2<P>(prm, prm.m_blocks * prm.n_blocks, st); }

}  // namespace tc

bool tc_available()
{
    // cached per device ordinal (0 = unknown, 1 = yes, 2 = no): a process may drive several GPUs
    static std::atomic<signed char> ok[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
    signed char v = ok[dev & 63].load(std::memory_order_relaxed);
    if (v == 0) {
        int major = 0;
        v = (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess && major == 10 &&
             tc::encode_tiled_fn() != nullptr) ? 1 : 2;
        cudaGetLastError();
        ok[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v == 1;
}

}  // namespace pvs

using namespace pvs;

// mode 0: tf32 single pass (K-major)   a_hi/b_hi fp32 [m,k] / [n,k]
// mode 1: 3xTF32 (K-major)             hi + lo parts
// mode 2: bf16 (K-major)               a_hi/b_hi bf16
// mode 3: tf32 single pass, MN-major   a_hi fp32 [k,m], b_hi fp32 [k,n]
// mode 4: 3xTF32, MN-major
// mode 5: bf16, CTA pair (cta_group::2)     mode 6: 3xTF32, CTA pair
// mode 7: 3xTF32, CTA pair, B operand resident in shared memory (k == 128, n == block_n)
extern "C" int pvs_debug_tc_gemm(int mode, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                                 float* c, int m, int n, int k, int block_n, void* stream)
{
    PVS_CHECK(tc_available(), PVS_ERR_UNSUPPORTED, "tcgen05 path needs an sm_100 device");
    PVS_CHECK(mode >= 0 && mode <= 7, PVS_ERR_BAD_ARG, "unknown mode %d", mode);
    PVS_CHECK(block_n == 64 || block_n == 128 || block_n == 256, PVS_ERR_BAD_ARG, "block_n must be 64/128/256");
    const bool bf16 = mode == 2 || mode == 5, mn = mode == 3 || mode == 4, three = mode == 1 || mode == 4 || mode >= 6;
    const bool pair = mode >= 5;
    PVS_CHECK(a_hi && b_hi && c && (!three || (a_lo && b_lo)), PVS_ERR_BAD_ARG, "NULL operand");
    tc::GemmParams p{};
    p.c = c; p.ldc = n; p.m = m; p.n = n; p.k = k;
    p.m_blocks = (m + (pair ? 255 : 127)) / (pair ? 256 : 128);
    p.n_blocks = (n + block_n - 1) / block_n;
    int rc;
    if (!mn) {
        const int bk = bf16 ? 64 : 32;
        const int brows = pair ? block_n / 2 : block_n;          // a pair CTA loads half of the B tile
        PVS_CHECK(k % bk == 0, PVS_ERR_BAD_SHAPE, "k must be a multiple of %d", bk);
        PVS_CHECK(mode != 7 || (k == 128 && n <= block_n), PVS_ERR_BAD_SHAPE, "mode 7 needs k == 128 and n <= block_n");
        if ((rc = tc::make_tmap_2d(&p.a_hi, a_hi, bf16, m, k, k, bk, 128))) return rc;
        if ((rc = tc::make_tmap_2d(&p.b_hi, b_hi, bf16, n, k, k, bk, brows))) return rc;
        if (three) {
            if ((rc = tc::make_tmap_2d(&p.a_lo, a_lo, false, m, k, k, bk, 128))) return rc;
            if ((rc = tc::make_tmap_2d(&p.b_lo, b_lo, false, n, k, k, bk, brows))) return rc;
        }
    } else {
        if ((rc = tc::make_tmap_2d(&p.a_hi, a_hi, false, k, m, m, 32, 32, true))) return rc;
        if ((rc = tc::make_tmap_2d(&p.b_hi, b_hi, false, k, n, n, 32, 32, true))) return rc;
        if (three) {
            if ((rc = tc::make_tmap_2d(&p.a_lo, a_lo, false, k, m, m, 32, 32, true))) return rc;
            if ((rc = tc::make_tmap_2d(&p.b_lo, b_lo, false, k, n, n, 32, 32, true))) return rc;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
#define RUN(BN)                                                                                  \
    switch (mode) {                                                                              \
        case 5: return tc::run_gemm2<tc::GemmPair<true, 1, BN, false>>(p, st);                   \
        case 6: return tc::run_gemm2<tc::GemmPair<false, 3, BN, false>>(p, st);                  \
        case 7: return tc::run_gemm2<tc::GemmPair<false, 3, BN, true>>(p, st);                   \
        case 0: return tc::run_gemm<tc::GemmKMajor<false, 1, BN>>(p, st);                        \
        case 1: return tc::run_gemm<tc::GemmKMajor<false, 3, BN>>(p, st);                        \
        case 2: return tc::run_gemm<tc::GemmKMajor<true, 1, BN>>(p, st);                         \
        case 3: return tc::run_gemm<tc::GemmMNMajor<1, BN>>(p, st);                              \
        default: return tc::run_gemm<tc::GemmMNMajor<3, BN>>(p, st);                             \
    }
    if (block_n == 64) { RUN(64) }
    if (block_n == 128) { RUN(128) }
    RUN(256)
#undef RUN
}
