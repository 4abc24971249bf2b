// pvs_tc_vlad.cu -- VLAD hard assignment on tcgen05 tensor cores (3xTF32, CTA pairs).
//
// Replaces KMeans.predict at pyvisim/encoders/vlad.py:95 (sklearn _k_means_lloyd.pyx
// _update_chunk_dense: one sgemm per 256-row chunk, then a strict `<` arg-min scan over
// ||c||^2 - 2 x.c).  Here a pair of CTAs scores a [256 descriptors x 256 centres] tile:
//   A  = descriptor rows, loaded as fp32 by four producer warps per CTA, split into tf32
//        hi/lo parts in registers and stored as 128-byte-swizzled UMMA tiles
//   B  = centres (hi/lo split once at model creation, zero-padded to a multiple of 32
//        columns); resident in shared memory for D = 64 / 128 (half of the centres per CTA),
//        streamed by TMA otherwise (D = 514: 17 k-blocks)
//   D  = x.c in TMEM (fp32, error-compensated: hi*lo + lo*hi + hi*hi)
// Epilogue: each thread owns one descriptor (one TMEM lane), scans its 256 scores
// ||c||^2 - 2 acc in column order with a strict `<` (lowest index wins ties, like the
// reference) and writes the int32 label.  The score matrix never leaves the SM.
#include <string.h>
#include <type_traits>
#include "pvs_tc2.cuh"
#include "pvs_kernels.cuh"

namespace pvs {
namespace tc2 {

struct AssignParams {
    CUtensorMap c_hi, c_lo;            // centres [k, d_pad] fp32 (hi / lo parts), box 32 cols x 128 rows
    const float* x;                    // [rows, d]
    const float* c2;                   // [k] squared norms
    const float* centers;              // [k, d] fp32 (exact re-evaluation of near-ties)
    int32_t* labels;                   // [rows]
    int64_t rows;
    int d, k, nkb, m_blocks;
    int* flag;                         // fp16x2 range flag: raised by the fp16x2 kernel when some |x 2^-e| leaves fp16's
                                       // range; the 3xTF32 kernel behind it then redoes the call (NULL: no gating)
    float sc_x, neg2s, lim;            // fp16x2: 2^-e, -2 * 4^e, largest admissible |x 2^-e|
};
struct AssignState {};

__device__ __forceinline__ uint32_t swz128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// One warp owns rows [32 pw, 32 pw + 32) of a [128 rows x 32 floats] K-major tile (hi and lo
// parts).  fetch: loads from src[row * d + col0 ...]; columns >= d and rows >= rows_total
// read as zero.  VEC = floats per load the row alignment allows (4: d % 4 == 0,
// 2: d % 2 == 0, else 1).  store: tf32 split + swizzled 16-byte stores.
struct ARegs { float4 v[8]; };
template <int VEC>
__device__ __forceinline__ void fetch_tile_rows(const float* __restrict__ src, int d, int64_t row0, int64_t rows_total,
                                                int col0, int pw, int lane, ARegs& g)
{
    const int col = col0 + (lane & 7) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gr = row0 + pw * 32 + i * 4 + (lane >> 3);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < rows_total) {
            const float* p = src + gr * (int64_t)d + col;
            if constexpr (VEC == 4) {
                if (col < d) v = __ldg(reinterpret_cast<const float4*>(p));
            } else if constexpr (VEC == 2) {
                if (col < d) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v.x = t.x; v.y = t.y; }
                if (col + 2 < d) { const float2 t = __ldg(reinterpret_cast<const float2*>(p + 2)); v.z = t.x; v.w = t.y; }
            } else {
                if (col < d) v.x = __ldg(p);
                if (col + 1 < d) v.y = __ldg(p + 1);
                if (col + 2 < d) v.z = __ldg(p + 2);
                if (col + 3 < d) v.w = __ldg(p + 3);
            }
        }
        g.v[i] = v;
    }
}
__device__ __forceinline__ void store_tile_rows(const ARegs& g, uint8_t* hi, uint8_t* lo, int pw, int lane)
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
        const uint32_t off = swz128(r, c);
        *reinterpret_cast<float4*>(hi + off) = h;
        *reinterpret_cast<float4*>(lo + off) = l;
    }
}

// fp16x2 variant of the producer: a lane fills one 16-byte chunk = 8 columns of a [128 rows x 64 fp16] tile
struct ARegs16 { float4 v[16]; };
template <int VEC>
__device__ __forceinline__ void fetch_tile_rows16(const float* __restrict__ src, int d, int64_t row0, int64_t rows_total,
                                                  int col0, int pw, int lane, ARegs16& g)
{
    const int col = col0 + (lane & 7) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gr = row0 + pw * 32 + i * 4 + (lane >> 3);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (gr < rows_total) {
            const float* p = src + gr * (int64_t)d + col;
            if constexpr (VEC == 4) {
                if (col < d) a = __ldg(reinterpret_cast<const float4*>(p));
                if (col + 4 < d) b = __ldg(reinterpret_cast<const float4*>(p + 4));
            } else if constexpr (VEC == 2) {
                if (col < d) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); a.x = t.x; a.y = t.y; }
                if (col + 2 < d) { const float2 t = __ldg(reinterpret_cast<const float2*>(p + 2)); a.z = t.x; a.w = t.y; }
                if (col + 4 < d) { const float2 t = __ldg(reinterpret_cast<const float2*>(p + 4)); b.x = t.x; b.y = t.y; }
                if (col + 6 < d) { const float2 t = __ldg(reinterpret_cast<const float2*>(p + 6)); b.z = t.x; b.w = t.y; }
            } else {
                float t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = col + u < d ? __ldg(p + u) : 0.f;
                a = make_float4(t[0], t[1], t[2], t[3]);
                b = make_float4(t[4], t[5], t[6], t[7]);
            }
        }
        g.v[2 * i] = a;
        g.v[2 * i + 1] = b;
    }
}
// scale by 2^-e, split into fp16 hi + lo, swizzled 16-byte stores; returns the largest |x 2^-e| seen
__device__ __forceinline__ float store_tile_rows16(const ARegs16& g, float sc, uint8_t* hi, uint8_t* lo, int pw, int lane)
{
    const int c = lane & 7;
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = pw * 32 + i * 4 + (lane >> 3);
        const float4 a = g.v[2 * i], b = g.v[2 * i + 1];
        const float x[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, b.x * sc, b.y * sc, b.z * sc, b.w * sc};
#pragma unroll
        for (int u = 0; u < 8; ++u) amax = fmaxf(amax, fabsf(x[u]));
        uint4 h, l;
        split8_h(x, h, l);
        const uint32_t off = swz128(r, c);
        *reinterpret_cast<uint4*>(hi + off) = h;
        *reinterpret_cast<uint4*>(lo + off) = l;
    }
    return amax;
}

// H_ = false: 3xTF32 (k-block = 32 columns).  H_ = true: fp16x2 -- descriptors and centres scaled by 2^-e
// (e from the centres: max |c| 2^-e <= 128) and split into fp16 hi + lo; three kind::f16 MMAs give the same
// 22-bit products at twice the tensor rate, and a k-block is 64 columns.
template <bool RES_, int NKB_, int VEC_, bool H_ = false>
struct AssignPolicy {
    using Params = AssignParams;
    using EpiState = AssignState;
    struct Tile { int nkb, mb; };
    static constexpr bool BF16 = false, F16 = H_, MANUAL = true, B_RESIDENT = RES_, ACC_INIT = false, TILE_SYNC = false;
    static constexpr int KCOLS = H_ ? 64 : 32;                     // operand columns per k-block (one 128-B row)
    // fp16x2: a lane holds a k-block of 16 float4 in registers, so two producer groups (448 threads) instead of three
    static constexpr int PASSES = 3, BLOCK_N = 256, KSTEPS = 4, NKB_RES = NKB_;
    static constexpr int STAGES = H_ ? (RES_ ? 4 : 2) : 3, PGROUPS = H_ ? 2 : 3;
    static constexpr int A_BYTES = 128 * 128, B_BYTES = 128 * 128, SCRATCH_BYTES = 1024;
    static constexpr int TMA_BYTES = RES_ ? 0 : 2 * B_BYTES;
    __device__ static bool enabled(const Params& p) { return H_ || !p.flag || *p.flag != 0; }
    __device__ static void prefetch(const Params& p) { tma_prefetch_desc(&p.c_hi); tma_prefetch_desc(&p.c_lo); }
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
            tma_load_2d_pair(res + (2 * kb) * B_BYTES, &p.c_hi, bar, kb * KCOLS, rank * 128);
            tma_load_2d_pair(res + (2 * kb + 1) * B_BYTES, &p.c_lo, bar, kb * KCOLS, rank * 128);
        }
    }
    __device__ static void load(const Params& p, const Tile&, int kb, int rank, uint8_t*, uint8_t*, uint8_t* b_hi,
                                uint8_t* b_lo, uint64_t* bar)
    {
        tma_load_2d_pair(b_hi, &p.c_hi, bar, kb * KCOLS, rank * 128);
        tma_load_2d_pair(b_lo, &p.c_lo, bar, kb * KCOLS, rank * 128);
    }
    using Regs = typename std::conditional<H_, ARegs16, ARegs>::type;
    __device__ static void fetch(const Params& p, const Tile& t, int kb, int rank, int pw, int lane, Regs& g)
    {
        if constexpr (H_) fetch_tile_rows16<VEC_>(p.x, p.d, (int64_t)t.mb * 256 + rank * 128, p.rows, kb * 64, pw, lane, g);
        else fetch_tile_rows<VEC_>(p.x, p.d, (int64_t)t.mb * 256 + rank * 128, p.rows, kb * 32, pw, lane, g);
    }
    __device__ static void store(const Params& p, const Tile&, int, const Regs& g, uint8_t* a_hi, uint8_t* a_lo, int pw, int lane)
    {
        if constexpr (H_) {
            // inf raises the flag, NaN does not (it just propagates through either variant)
            if (store_tile_rows16(g, p.sc_x, a_hi, a_lo, pw, lane) > p.lim) *p.flag = 1;
        } else {
            store_tile_rows(g, a_hi, a_lo, pw, lane);
        }
    }
    __device__ static void epi_init(const Params& p, uint8_t* scratch, int tid)
    {
        float* c2 = reinterpret_cast<float*>(scratch);
        for (int j = tid; j < BLOCK_N; j += 128) c2[j] = j < p.k ? p.c2[j] : INFINITY;   // padded centres never win
        epi_barrier();
    }
    __device__ static void epi_begin(const Params&, const Tile&, EpiState&, int, int, int) {}
    __device__ static void epilogue(const Params& p, const Tile& t, int rank, uint32_t tmem, int quarter, int lane,
                                    uint8_t* scratch, EpiState&)
    {
        const float* c2 = reinterpret_cast<const float*>(scratch);
        const int64_t row = (int64_t)t.mb * 256 + rank * 128 + quarter * 32 + lane;
        const float m2 = H_ ? p.neg2s : -2.f;                  // fp16x2: the accumulator holds x.c / 4^e
        // best (strict <: the lowest index wins exact ties, like sklearn's scan) and the VALUE of the runner-up, branch-free:
        // min over the elements of max(score, best so far) -- two FMNMX per element; tracking the runner-up's index as well
        // made the epilogue the bottleneck of the 128-D kernel (VLAD C1 1.67 M -> 1.15 M images/s)
        float best = INFINITY, second = INFINITY;
        int bi = 0;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
            float v[32];
            tmem_ld32(tmem + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float s = fmaf(m2, v[j], c2[c + j]);
#ifndef PVS_ASSIGN_NO_TIE
                second = fminf(second, fmaxf(s, best));
#endif
                if (s < best) { best = s; bi = c + j; }
            }
        }
        // Near-tie: the split-operand scores carry up to ~1e-6 of relative error (22-bit products, 24 truncating accumulation
        // steps at D = 128), so when the runner-up is closer than 2e-6 of (||c||^2 + |score|) the two are re-evaluated EXACTLY
        // (||c||^2 - 2 x.c in fp64 from the fp32 inputs) and the strict-< rule is applied to the exact scores.  The reference
        // (vlad.py:95 -> sklearn's fp32 sgemm) resolves gaps below ~1e-6 by the rounding noise of its BLAS kernel's summation
        // order, which differs between CPUs; the exact order is the only reproducible choice, and every remaining mismatch
        // with an fp32 reference run is such a sub-1e-6 tie (tests: assert_labels).  A warp with a flagged row scans the
        // accumulator a second time for the runner-up's index (tcgen05.ld is warp-collective) and then works the flagged rows
        // off together: 32 lanes share the two 2 D-long fp64 dot products.  (A 1e-5 threshold flagged a row in nearly every
        // warp of the RootSIFT benchmark and doubled the kernel's time.)
        const bool tie = row < p.rows && second - best <= 2e-6f * (c2[bi] + fabsf(best));
        unsigned need = __ballot_sync(0xffffffffu, tie);
        if (need) {
            float sec = INFINITY;
            int si = -1;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N; c += 32) {
                float v[32];
                tmem_ld32(tmem + c, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float s = fmaf(m2, v[j], c2[c + j]);
                    if (c + j != bi && s < sec) { sec = s; si = c + j; }
                }
            }
            while (need) {
                const int owner = __ffs(need) - 1;
                need &= need - 1;
                const int a_i = __shfl_sync(0xffffffffu, bi, owner), b_i = __shfl_sync(0xffffffffu, si, owner);
                if (b_i < 0 || b_i >= p.k) continue;           // warp-uniform
                const float* xr = p.x + (row - lane + owner) * (int64_t)p.d;
                const float* ca = p.centers + (int64_t)a_i * p.d;
                const float* cb = p.centers + (int64_t)b_i * p.d;
                double sa = 0.0, sb = 0.0;
                for (int i = lane; i < p.d; i += 32) {
                    const double xv = (double)xr[i], a = (double)ca[i], b = (double)cb[i];
                    sa = fma(a, a - 2.0 * xv, sa);             // c^2 - 2 x c, term by term
                    sb = fma(b, b - 2.0 * xv, sb);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    sa += __shfl_xor_sync(0xffffffffu, sa, o);
                    sb += __shfl_xor_sync(0xffffffffu, sb, o);
                }
                if (lane == owner && (sb < sa || (sb == sa && b_i < a_i))) bi = b_i;
            }
        }
        if (row < p.rows) p.labels[row] = bi;
    }
};

}  // namespace tc2

using namespace tc2;

// K-Means model preparation: zero-padded [k, d_pad] hi / lo copies of the centres
int tc_prepare_kmeans(pvs_model* m)
{
    if (!tc_available() || m->kind != PVS_MODEL_KMEANS || m->k > 256) return PVS_OK;
    const int d_pad = (m->d + 31) / 32 * 32;
    const size_t n = (size_t)m->k * d_pad;
    std::vector<float> hc((size_t)m->k * m->d), hi(n, 0.f), lo(n, 0.f);
    PVS_CUDA(cudaMemcpy(hc.data(), m->centers, hc.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int j = 0; j < m->k; ++j)
        for (int i = 0; i < m->d; ++i) {
            // round-to-nearest-away at 13 dropped mantissa bits == cvt.rna.tf32.f32 for finite values
            const float x = hc[(size_t)j * m->d + i];
            uint32_t u;
            memcpy(&u, &x, 4);
            uint32_t uh = (u + 0x1000u) & 0xFFFFE000u;
            float h;
            memcpy(&h, &uh, 4);
            const float r = x - h;
            memcpy(&u, &r, 4);
            uint32_t ul = (u + 0x1000u) & 0xFFFFE000u;
            float l;
            memcpy(&l, &ul, 4);
            hi[(size_t)j * d_pad + i] = h;
            lo[(size_t)j * d_pad + i] = l;
        }
    float* buf = nullptr;
    PVS_CUDA(cudaMalloc((void**)&buf, 2 * n * sizeof(float)));
    cudaError_t e = cudaMemcpy(buf, hi.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(buf + n, lo.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(buf); return fail(PVS_ERR_CUDA, "centre upload failed: %s", cudaGetErrorString(e)); }
    m->tc0 = buf;
    m->tc1 = buf + n;
    m->tc_ld = d_pad;
    // fp16x2 copies: centres * 2^-e with max |c| 2^-e in (64, 128], fp16 hi + lo, padded to 64 columns
    float cmax = 0.f;
    for (float v : hc) cmax = fmaxf(cmax, fabsf(v));
    if (cmax > 0.f && std::isfinite(cmax)) {
        const int ex = (int)ceilf(log2f(cmax)) - 7;
        if (ex > -30 && ex < 30) {
            const int ld16 = (m->d + 63) / 64 * 64;
            const size_t n16 = (size_t)m->k * ld16;
            std::vector<__half> hl(2 * n16, __float2half_rn(0.f));
            const float sc = ldexpf(1.f, -ex);
            for (int j = 0; j < m->k; ++j)
                for (int i = 0; i < m->d; ++i) {
                    const float x = hc[(size_t)j * m->d + i] * sc;
                    const __half h = __float2half_rn(x);
                    hl[(size_t)j * ld16 + i] = h;
                    hl[n16 + (size_t)j * ld16 + i] = __float2half_rn(x - __half2float(h));
                }
            void* hb = nullptr;
            int* flags = nullptr;
            PVS_CUDA(cudaMalloc(&hb, 2 * n16 * sizeof(__half)));
            if (cudaMalloc((void**)&flags, 64 * sizeof(int)) != cudaSuccess ||
                cudaMemcpy(hb, hl.data(), 2 * n16 * sizeof(__half), cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFree(hb);
                if (flags) cudaFree(flags);
                return fail(PVS_ERR_CUDA, "fp16 centre upload failed");
            }
            m->th0 = hb;
            m->th1 = (const __half*)hb + n16;
            m->th_ld = ld16;
            m->h_flags = flags;
            m->h_exp = ex;
            m->h_ok = true;
        }
    }
    return PVS_OK;
}

bool tc_assign_supported(const pvs_model* km, int64_t rows)
{
    return tc_available() && km->tc0 && km->k <= 256 && km->k >= 8 && rows / 256 + 2 < 2147483647LL;
}

int tc_vlad_assign(const pvs_model* km, const float* x, int64_t rows, int32_t* labels, cudaStream_t st)
{
    if (rows <= 0) return PVS_OK;
    AssignParams p{};
    int rc;
    if ((rc = make_tmap_2d(&p.c_hi, km->tc0, false, km->k, km->tc_ld, km->tc_ld, 32, 128))) return rc;
    if ((rc = make_tmap_2d(&p.c_lo, km->tc1, false, km->k, km->tc_ld, km->tc_ld, 32, 128))) return rc;
    p.x = x; p.c2 = km->c2; p.centers = km->centers; p.labels = labels; p.rows = rows;
    p.d = km->d; p.k = km->k; p.nkb = km->tc_ld / 32;
    p.m_blocks = (int)ceil_div(rows, 256);
    const bool a16 = ((uintptr_t)x & 15) == 0 && km->d % 4 == 0;
    const bool a8 = ((uintptr_t)x & 7) == 0 && km->d % 2 == 0;
    if (km->h_ok && km->th0 && !getenv("PVS_VLAD_NO_FP16X2")) {
        // fp16x2 kernel first; it raises the flag when a descriptor leaves the fp16 range, and only then
        // does the 3xTF32 kernel behind it run (and overwrite the labels)
        AssignParams h = p;
        h.flag = km->h_flags + (km->h_flag_next.fetch_add(1, std::memory_order_relaxed) & 63u);
        h.sc_x = ldexpf(1.f, -km->h_exp); h.neg2s = -2.f * ldexpf(1.f, 2 * km->h_exp); h.lim = 60000.f;
        h.nkb = km->th_ld / 64;
        PVS_CUDA(cudaMemsetAsync(h.flag, 0, sizeof(int), st));
        if ((rc = make_tmap_2d(&h.c_hi, km->th0, true, km->k, km->th_ld, km->th_ld, 64, 128))) return rc;
        if ((rc = make_tmap_2d(&h.c_lo, km->th1, true, km->k, km->th_ld, km->th_ld, 64, 128))) return rc;
        if (km->d == 64 && a16) rc = launch_tc2<AssignPolicy<true, 1, 4, true>>(h, h.m_blocks, st);
        else if (km->d == 128 && a16) rc = launch_tc2<AssignPolicy<true, 2, 4, true>>(h, h.m_blocks, st);
        else if (a16) rc = launch_tc2<AssignPolicy<false, 0, 4, true>>(h, h.m_blocks, st);
        else if (a8) rc = launch_tc2<AssignPolicy<false, 0, 2, true>>(h, h.m_blocks, st);
        else rc = launch_tc2<AssignPolicy<false, 0, 1, true>>(h, h.m_blocks, st);
        if (rc) return rc;
        p.flag = h.flag;
    }
    if (km->d == 64 && a16) return launch_tc2<AssignPolicy<true, 2, 4>>(p, p.m_blocks, st);
    if (km->d == 128 && a16) return launch_tc2<AssignPolicy<true, 4, 4>>(p, p.m_blocks, st);
    if (a16) return launch_tc2<AssignPolicy<false, 0, 4>>(p, p.m_blocks, st);
    if (a8) return launch_tc2<AssignPolicy<false, 0, 2>>(p, p.m_blocks, st);
    return launch_tc2<AssignPolicy<false, 0, 1>>(p, p.m_blocks, st);
}

}  // namespace pvs
