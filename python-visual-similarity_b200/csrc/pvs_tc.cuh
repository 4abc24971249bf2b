// pvs_tc.cuh -- sm_100a tensor-core machinery shared by the tcgen05 kernels:
// mbarrier / TMA / tcgen05 (alloc, mma, commit, ld) PTX wrappers, UMMA shared-memory and
// instruction descriptors, host-side TMA tensor-map encoding, and one warp-specialised
// persistent kernel skeleton (TMA producer warp -> MMA issuer warp -> 4 epilogue warps,
// smem ring + double-buffered TMEM accumulators) that every contraction instantiates
// with a small policy struct.
//
// Precision modes of the contraction (all accumulate in fp32 in TMEM):
//   PASSES == 1 : one tcgen05.mma per k-step (bf16 x bf16 for similarity).
//   PASSES == 3 : "3xTF32" -- both operands are supplied as tf32-exact hi + lo parts and
//                 the product is hi*lo + lo*hi + hi*hi, which restores fp32-class accuracy
//                 (tools/tf32_split_study.py: any cheaper split misses the 1e-4 parity bar).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "pvs_common.cuh"

namespace pvs {
namespace tc {

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
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
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 6000000000LL) __trap();      // ~3 s at 2 GHz
    }
}

// one lane of a converged warp; ptxas then knows the guarded region is single-threaded and
// issues UTCHMMA / UTMALDG directly (a plain `lane == 0` test makes it wrap every such
// instruction in an ELECT loop, ~20 extra instructions per MMA)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled TMA load global -> shared, completion on an mbarrier (c0 = innermost coordinate)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// 1-D bulk copy global -> shared (no tensor map): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// 2-D tiled TMA store shared -> global (bulk async-group of the issuing thread); rows / columns
// outside the tensor are clipped by the hardware
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; one elected thread issues on behalf of the CTA
template <bool BF16>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    if constexpr (BF16) {
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// all previously issued MMAs of this thread arrive on `bar` when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 16-column variants (register-tight epilogues)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8])
{
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8])
{
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16])
{
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the mirror store: thread i of the warp writes 32 consecutive fp32 columns of lane (base + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32])
{
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// barrier among the 128 epilogue threads only (id 1; id 0 is __syncthreads)
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// tf32 split: x = hi + lo (+ O(2^-22 |x|)), both parts exactly representable in tf32
__device__ __forceinline__ float tf32_rn(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo)
{
    hi = tf32_rn(x);
    lo = tf32_rn(x - hi);
}

// fp16x2 split: x = hi + lo (+ O(2^-22 |x|)) with both parts fp16 -- for operands that were scaled into
// fp16's range by an exact power of two
__device__ __forceinline__ uint32_t pack_h2(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
// eight fp32 values -> 16-byte chunks of fp16 hi and lo parts (x = hi + lo + O(2^-22 |x|))
__device__ __forceinline__ void split8_h(const float (&x)[8], uint4& hi, uint4& lo)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
        const float2 back = __half22float2(hh);
        h[i] = pack_h2(hh);
        l[i] = pack_h2(__floats2half2_rn(x[2 * i] - back.x, x[2 * i + 1] - back.y));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle (bit layout: cute/arch/mma_sm100_desc.hpp).
//   K-major operand : SWIZZLE_128B (2): rows of 128 B (one swizzle span along K), 16-B chunks
//                     XOR-ed with (row & 7); 8-row groups SBO = 1024 B apart.
//   MN-major tf32   : SWIZZLE_128B_BASE32B (1) -- the only MN-major layout tf32 supports
//                     (cutlass sm100_common.inl:92): rows of 128 B run along M/N, 32-B chunks
//                     XOR-ed with (row & 3); 4 consecutive K rows form one 512-B atom;
//                     LBO = byte stride between 128-B column blocks, SBO = between 4-row groups.
//                     The matching TMA mode is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}
// Instruction descriptor: fp32 accumulate, A/B format tf32 (2) or bf16 (1), majors, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc(bool bf16, bool a_mn, bool b_mn, int m, int n, bool f16 = false)
{
    const uint32_t fmt = f16 ? 0u : bf16 ? 1u : 2u;          // kind::f16: 0 = fp16, 1 = bf16; kind::tf32: 2 = tf32
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((a_mn ? 1u : 0u) << 15) |
           ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Optional policy members (detected, so existing policies need not declare them):
//   static constexpr bool F16 = true   operands are fp16 (kind::f16, format 0); with PASSES == 3 this is the
//                                      "fp16x2" split x = hi + lo, hi*lo + lo*hi + hi*hi: 22 mantissa bits like
//                                      3xTF32 at twice the tensor rate and half the shared-memory bytes, for
//                                      operands that were scaled into fp16's range by an exact power of two
//   static bool enabled(const Params&) uniform early exit: the kernel returns at once when this is false
//                                      (two precision variants are launched back to back and a device-side
//                                      flag decides which one does the work -- no host synchronisation)
template <class P, class = void> struct policy_f16 { static constexpr bool value = false; };
template <class P> struct policy_f16<P, decltype((void)P::F16)> { static constexpr bool value = P::F16; };
//   static constexpr int STAGE_EXTRA   policy-owned bytes at the end of every stage (e.g. a TMA-filled staging tile
//                                      plus the policy's own mbarrier), with
//   static void init_stage(uint8_t*)   called once per stage (pointer to the extra region) before the barriers
//                                      are published
template <class P, class = void> struct policy_extra { static constexpr int value = 0; };
template <class P> struct policy_extra<P, decltype((void)P::STAGE_EXTRA)> { static constexpr int value = P::STAGE_EXTRA; };
//   static constexpr bool TMA_OWN_BARRIER = true   the policy's load() completes its TMA bytes on its own mbarrier
//                                      (in the extra region) instead of full[s]; the MANUAL warps wait for
//                                      it inside store(), post-process the tiles and only then arrive on full[s]
template <class P, class = void> struct policy_own_tx { static constexpr bool value = false; };
template <class P> struct policy_own_tx<P, decltype((void)P::TMA_OWN_BARRIER)> { static constexpr bool value = P::TMA_OWN_BARRIER; };
//   static constexpr bool FOLD_WARPS = true   (with MANUAL) four more warps behind the producers own the accumulator read-out:
//                                      they wait for tfull, run fold(prm, tile, tmem, quarter, lane, FoldState&) and
//                                      release the accumulator; the epilogue warps then only consume() stages and call
//                                      epilogue_host(prm, tile, EpiState&) (no TMEM access, no accumulator hand-shake)
template <class P, class = void> struct policy_foldwarps { static constexpr bool value = false; };
template <class P> struct policy_foldwarps<P, decltype((void)P::FOLD_WARPS)> { static constexpr bool value = P::FOLD_WARPS; };
template <class P, class = void> struct policy_gated { static constexpr bool value = false; };
template <class P> struct policy_gated<P, decltype((void)&P::enabled)> { static constexpr bool value = true; };

// ---------------------------------------------------------------------------------------
// host: TMA tensor maps (driver entry point resolved at run time, no libcuda link)
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
// 2-D row-major [rows, cols] tensor with a row pitch of ld elements; box = box_cols x box_rows,
// box_cols * elem_bytes must be 128 (one swizzle span).
int make_tmap_2d(CUtensorMap* out, const void* base, bool bf16, int64_t rows, int64_t cols, int64_t ld,
                 int box_cols, int box_rows, bool mn_major = false);

// ---------------------------------------------------------------------------------------
// kernel skeleton
// ---------------------------------------------------------------------------------------
// Policy P provides:
//   struct Params; struct Tile { int nkb; ... };
//   static constexpr: bool BF16; int PASSES, BLOCK_N, STAGES, KSTEPS (k-steps per k-block);
//                     bool A_MN, B_MN; int A_BYTES, B_BYTES (one precision part of one stage),
//                     A_LBO, B_LBO (MN-major block stride, bytes); bool EPI_READS_STAGES;
//                     int SCRATCH_BYTES (extra smem for the epilogue)
//   __device__ static int   num_tiles(const Params&);
//   __device__ static int   tile_at(const Params&, int it, int n_tiles);   it-th tile of this CTA or -1
//                                  (strided_tile() = blockIdx.x + it * gridDim.x is the default order)
//   __device__ static Tile  tile(const Params&, int idx);
//   __device__ static void  load(const Params&, const Tile&, int kb, uint8_t* a_hi, uint8_t* a_lo,
//                                uint8_t* b_hi, uint8_t* b_lo, uint64_t* bar);     (one thread)
//   __device__ static void  consume(...)   (only if EPI_READS_STAGES; 128 epilogue threads)
//   __device__ static void  epilogue(const Params&, const Tile&, uint32_t tmem_acc, int quarter, int lane,
//                                    uint8_t* scratch, EpiState&);
__device__ __forceinline__ int strided_tile(int it, int n_tiles)
{
    const long long t = (long long)blockIdx.x + (long long)it * gridDim.x;
    return t < n_tiles ? (int)t : -1;
}

#ifdef PVS_TIMING
// cycles summed over CTAs: [0] MMA wait full, [1] MMA wait tempty, [2] MMA total,
// [3] epilogue (warp 2) wait tfull, [4] epilogue total, [5] producer (warp 6) wait empty, [6] producer total
__device__ unsigned long long g_tc_timing[8];
#define PVS1_T0(var) const long long var = clock64()
#define PVS1_ADD(slot, t0, cond) do { if (cond) atomicAdd(&g_tc_timing[slot], (unsigned long long)(clock64() - (t0))); } while (0)
#else
#define PVS1_T0(var)
#define PVS1_ADD(slot, t0, cond)
#endif

template <class P>
struct Layout {
    static constexpr int PARTS = P::PASSES == 3 ? 2 : 1;
    static constexpr int EXTRA = policy_extra<P>::value;
    static constexpr int STAGE_BYTES = PARTS * (P::A_BYTES + P::B_BYTES) + EXTRA;
    static constexpr int RING_BYTES = P::STAGES * STAGE_BYTES;
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = 1024 /*alignment slack*/ + RING_BYTES + BAR_BYTES + P::SCRATCH_BYTES;
    static_assert(EXTRA % 1024 == 0, "stages must keep 1024-B alignment");
    static_assert(2 * P::STAGES + 4 <= 30, "barrier block too small");
    static constexpr int TMEM_COLS = 2 * P::BLOCK_N <= 32 ? 32 : 2 * P::BLOCK_N <= 64 ? 64 : 2 * P::BLOCK_N <= 128 ? 128
                                     : 2 * P::BLOCK_N <= 256 ? 256 : 512;
    // threads: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 epilogue,
    // warps 6-9 (only with P::MANUAL) operand producers that load fp32 from global memory,
    // split it into tf32 hi/lo parts and store the swizzled tiles themselves
    static constexpr bool FOLD = policy_foldwarps<P>::value;
    static_assert(!FOLD || P::MANUAL, "fold warps sit behind the producer warps");
    static constexpr int THREADS = (P::MANUAL ? 192 + 128 * P::PGROUPS : 192) + (FOLD ? 128 : 0);
    static constexpr bool OWN_TX = policy_own_tx<P>::value;
    static constexpr uint32_t FULL_COUNT = (P::TMA_BYTES > 0 && !OWN_TX ? 1 : 0) + (P::MANUAL ? 4 : 0);
    static_assert(2 * P::BLOCK_N <= 512, "accumulator does not fit TMEM twice");
    static constexpr bool F16 = policy_f16<P>::value;
    static_assert(!(P::BF16 && (P::A_MN || P::B_MN)), "MN-major operands are implemented for tf32 and fp16 only");
    static_assert(P::A_BYTES % 1024 == 0 && P::B_BYTES % 1024 == 0, "operand tiles must keep 1024-B alignment");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
    // every stage must be refilled by one producer group only: a group that ran two uses ahead of
    // another one on the same stage would pass the parity wait of the empty barrier too early
    static_assert(!P::MANUAL || P::STAGES % P::PGROUPS == 0, "PGROUPS must divide STAGES");
};

template <class P>
__global__ void __launch_bounds__(Layout<P>::THREADS, 1) tc_kernel(const __grid_constant__ typename P::Params prm)
{
    using L = Layout<P>;
    if constexpr (policy_gated<P>::value) {
        if (!P::enabled(prm)) return;                        // uniform over the grid
    }
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET on the __shared__ array: going through an integer cast would
    // turn every later access into a generic LD / ST instead of LDS / STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::RING_BYTES);
    uint64_t* empty = full + P::STAGES;
    uint64_t* tfull = empty + P::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    uint8_t* scratch = smem + L::RING_BYTES + L::BAR_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P::STAGES; ++s) {
            mbar_init(&full[s], L::FULL_COUNT);
            mbar_init(&empty[s], 1 + (P::EPI_READS_STAGES ? 4 : 0));
            if constexpr (L::EXTRA > 0) P::init_stage(smem + s * L::STAGE_BYTES + L::PARTS * (P::A_BYTES + P::B_BYTES));
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
        }
        fence_barrier_init();
        P::prefetch(prm);
    }
    if (warp == 1) tmem_alloc<L::TMEM_COLS>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = P::num_tiles(prm);

    if (warp == 0) {
        if (lane == 0 && P::TMA_BYTES > 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0, t; (t = P::tile_at(prm, it, n_tiles)) >= 0; ++it) {
                const typename P::Tile tl = P::tile(prm, t);
                for (int kb = 0; kb < tl.nkb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if constexpr (!L::OWN_TX) mbar_expect_tx(&full[stage], P::TMA_BYTES);
                    uint8_t* sp = smem + stage * L::STAGE_BYTES;
                    P::load(prm, tl, kb, sp, sp + P::A_BYTES, sp + L::PARTS * P::A_BYTES,
                            sp + L::PARTS * P::A_BYTES + P::B_BYTES, &full[stage]);
                    if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {   // the whole warp walks the pipeline; one elected lane issues the MMAs and commits
            constexpr bool F16 = L::F16;
            constexpr bool H = P::BF16 || F16;               // kind::f16
            constexpr uint32_t idesc = make_idesc(P::BF16, P::A_MN, P::B_MN, 128, P::BLOCK_N, F16);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int it = 0, t; (t = P::tile_at(prm, it, n_tiles)) >= 0; ++it) {
                const typename P::Tile tl = P::tile(prm, t);
                PVS1_T0(t_te);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                PVS1_ADD(1, t_te, lane == 0);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P::BLOCK_N);
                for (int kb = 0; kb < tl.nkb; ++kb) {
                    PVS1_T0(t_fu);
                    mbar_wait(&full[stage], phase);
                    PVS1_ADD(0, t_fu, lane == 0);
                    tcgen05_fence_after();
                    const uint32_t sp = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint32_t a_hi = sp, a_lo = sp + P::A_BYTES;
                    const uint32_t b_hi = sp + L::PARTS * P::A_BYTES, b_lo = b_hi + P::B_BYTES;
                    if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < P::KSTEPS; ++ks) {
                        // one k-step: tf32 K = 8, 16-bit K = 16.  K-major: 32 B along the 128-B row either way.
                        // MN-major: K rows of 128 B; tf32 uses 4-row atoms (512 B, SWIZZLE_128B_BASE32B), 16-bit
                        // types 8-row atoms (1024 B, SWIZZLE_128B) -- a k-step spans two atoms in both cases.
                        constexpr uint32_t mn_step = F16 ? 2048 : 1024, mn_sbo = F16 ? 1024 : 512;
                        constexpr uint32_t mn_layout = F16 ? LAYOUT_SW128 : LAYOUT_SW128_BASE32B;
                        const uint32_t a_off = P::A_MN ? ks * mn_step : ks * 32;
                        const uint32_t b_off = P::B_MN ? ks * mn_step : ks * 32;
                        const uint32_t a_l = P::A_MN ? P::A_LBO : 16, b_l = P::B_MN ? P::B_LBO : 16;
                        const uint32_t a_s = P::A_MN ? mn_sbo : 1024, b_s = P::B_MN ? mn_sbo : 1024;
                        const uint32_t a_t = P::A_MN ? mn_layout : LAYOUT_SW128;
                        const uint32_t b_t = P::B_MN ? mn_layout : LAYOUT_SW128;
                        const uint64_t da_hi = make_smem_desc(a_hi + a_off, a_l, a_s, a_t);
                        const uint64_t db_hi = make_smem_desc(b_hi + b_off, b_l, b_s, b_t);
                        const uint32_t first = (kb > 0 || ks > 0) ? 1u : 0u;
                        if constexpr (P::PASSES == 3) {
                            const uint64_t da_lo = make_smem_desc(a_lo + a_off, a_l, a_s, a_t);
                            const uint64_t db_lo = make_smem_desc(b_lo + b_off, b_l, b_s, b_t);
                            umma<H>(d_tmem, da_hi, db_lo, idesc, first);
                            umma<H>(d_tmem, da_lo, db_hi, idesc, 1u);
                            umma<H>(d_tmem, da_hi, db_hi, idesc, 1u);
                        } else {
                            umma<H>(d_tmem, da_hi, db_hi, idesc, first);
                        }
                    }
                    umma_commit(&empty[stage]);          // smem slot free once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit(&tfull[acc]);    // accumulator complete
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (L::FOLD && warp >= 6 + 4 * P::PGROUPS) {
        if constexpr (L::FOLD) {
            // accumulator read-out by dedicated warps (one per TMEM lane quarter)
            const int quarter = warp & 3;
            int acc = 0;
            uint32_t acc_phase = 0;
            typename P::FoldState fs;
            for (int it = 0, t; (t = P::tile_at(prm, it, n_tiles)) >= 0; ++it) {
                const typename P::Tile tl = P::tile(prm, t);
                mbar_wait(&tfull[acc], acc_phase);
                tcgen05_fence_after();
                P::fold(prm, tl, tmem_base + (uint32_t)(acc * P::BLOCK_N) + ((uint32_t)(quarter * 32) << 16), quarter, lane, fs);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 6) {
        if constexpr (P::MANUAL) {
            // P::PGROUPS groups of four producer warps; group g fills the k-blocks whose
            // running index is g (mod PGROUPS).  fetch() issues the global loads of a k-block
            // into registers, store() splits and writes the swizzled tiles; the loads of the
            // group's NEXT k-block are in flight while it waits for that stage to drain.
            const int pw = (warp - 6) & 3, grp = (warp - 6) >> 2;
            long long idx = 0;
            int it = 0, kb = 0;
            int t = P::tile_at(prm, 0, n_tiles);
            typename P::Tile tl{};
            if (t >= 0) tl = P::tile(prm, t);
            auto seek = [&]() {
                while (t >= 0) {
                    if (kb >= tl.nkb) {
                        ++it; kb = 0;
                        t = P::tile_at(prm, it, n_tiles);
                        if (t >= 0) tl = P::tile(prm, t);
                        continue;
                    }
                    if ((int)(idx % P::PGROUPS) == grp) return;   // a stage is always refilled by the same group
                                                                  // (PGROUPS divides STAGES): parity waits stay unambiguous
                    ++kb; ++idx;
                }
            };
            seek();
            typename P::Regs cur;
            typename P::PState ps{};                         // per-warp producer state that lives across k-blocks
            if (t >= 0) P::fetch(prm, tl, kb, pw, lane, cur);
            while (t >= 0) {
                const int stage = (int)(idx % P::STAGES);
                const uint32_t phase = (uint32_t)((idx / P::STAGES) & 1);
                PVS1_T0(t_em);
                mbar_wait(&empty[stage], phase ^ 1);
                PVS1_ADD(5, t_em, warp == 6 && lane == 0);
                uint8_t* sp = smem + stage * L::STAGE_BYTES;
                P::store(prm, tl, kb, cur, sp, sp + P::A_BYTES, sp + L::PARTS * P::A_BYTES,
                         sp + L::PARTS * P::A_BYTES + P::B_BYTES, pw, grp, lane, ps);
                fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
                ++kb; ++idx;
                seek();
                if (t >= 0) P::fetch(prm, tl, kb, pw, lane, cur);
            }
        }
    } else {
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        typename P::EpiState st;
        P::epi_init(prm, scratch, (int)threadIdx.x - 64);
        for (int it = 0, t; (t = P::tile_at(prm, it, n_tiles)) >= 0; ++it) {
            const typename P::Tile tl = P::tile(prm, t);
            P::epi_begin(prm, tl, st, quarter, lane);
            if constexpr (P::EPI_READS_STAGES) {
                for (int kb = 0; kb < tl.nkb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    uint8_t* sp = smem + stage * L::STAGE_BYTES;
                    P::consume(prm, tl, kb, sp, sp + P::A_BYTES, sp + L::PARTS * P::A_BYTES,
                               sp + L::PARTS * P::A_BYTES + P::B_BYTES, st, quarter, lane);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    if (++stage == P::STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if constexpr (L::FOLD) {
                P::epilogue_host(prm, tl, st, quarter, lane);    // the fold warps own the accumulator
            } else {
                PVS1_T0(t_tf);
                mbar_wait(&tfull[acc], acc_phase);
                PVS1_ADD(3, t_tf, warp == 2 && lane == 0);
                tcgen05_fence_after();
                PVS1_T0(t_body);
                P::epilogue(prm, tl, tmem_base + (uint32_t)(acc * P::BLOCK_N) + ((uint32_t)(quarter * 32) << 16), quarter,
                            lane, scratch, st);
                PVS1_ADD(4, t_body, warp == 2 && lane == 0);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<L::TMEM_COLS>(tmem_base);
}

template <class P>
int launch_tc(const typename P::Params& prm, int n_tiles, cudaStream_t st, int grid_override = 0)
{
    using L = Layout<P>;
    static PerDeviceOnce configured;
    if (configured.need()) {
        PVS_CUDA(cudaFuncSetAttribute(tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM_BYTES));
        configured.mark();
    }
    if (n_tiles <= 0) return PVS_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = grid_override > 0 ? grid_override : (n_tiles < sms ? n_tiles : sms);   // persistent: one CTA per SM
    tc_kernel<P><<<grid, L::THREADS, L::SMEM_BYTES, st>>>(prm);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PVS_ERR_CUDA, "tcgen05 kernel launch failed: %s", cudaGetErrorString(e));
#ifdef PVS_TIMING
    if (getenv("PVS_TIMING_PRINT")) {
        cudaStreamSynchronize(st);
        unsigned long long h[8], z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyFromSymbol(h, g_tc_timing, sizeof(h));
        cudaMemcpyToSymbol(g_tc_timing, z, sizeof(z));
        const double np = grid;
        fprintf(stderr, "[tc timing %s] per CTA (kcycles): mma wait_full %.0f wait_tempty %.0f | epi wait_tfull %.0f body %.0f | prod wait_empty %.0f\n",
                __PRETTY_FUNCTION__, h[0] / np / 1e3, h[1] / np / 1e3, h[3] / np / 1e3, h[4] / np / 1e3, h[5] / np / 1e3);
    }
#endif
    return PVS_OK;
}

}  // namespace tc
}  // namespace pvs
