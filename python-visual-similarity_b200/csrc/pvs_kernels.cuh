// pvs_kernels.cuh -- launcher prototypes shared between the C-ABI layer and the kernel TUs.
#pragma once
#include "pvs_common.cuh"

namespace pvs {

// ---- generic fp32 CUDA-core contraction (any shape) -----------------------------------
// C[m,n] = alpha * sum_k A'[m,k] * B[n,k] + bias[n]
//   square_cat == 0 : A' = A                      (kdim = a_cols)
//   square_cat == 1 : A' = [A*A | A] per row      (kdim = 2*a_cols; B is [n, 2*a_cols])
int launch_gemm_nt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                   int64_t M, int N, int a_cols, int square_cat, float alpha, const float* bias,
                   cudaStream_t st);

// in-place softmax over each row of L [rows, k]; optional arg-max per row
int launch_row_softmax(float* L, int64_t rows, int k, int32_t* argmax_out, cudaStream_t st);
// labels[r] = argmin_j S[r, j], lowest index on ties
int launch_row_argmin(const float* S, int64_t rows, int k, int32_t* labels, cudaStream_t st);

// ---- VLAD aggregation + normalisation ---------------------------------------------------
int launch_vlad_aggregate(const float* y, int d, const int32_t* labels, const int64_t* offsets,
                          int64_t n_images, int64_t total_rows, const float* centers, int k, float power,
                          float norm_order, float eps, float* out, cudaStream_t st);

// ---- Fisher vector statistics + gradients ------------------------------------------------
// S [n_images, k, 2d+1] = per image ( q^T [y | y*y] , sum_t q ) / T
int launch_fv_stats(const float* q, const float* y, int d, int k, const int64_t* offsets,
                    int64_t n_images, float* S, cudaStream_t st);
// S rows have pitch ld (2d+1 with s0 in the last column, or 2d with s0 given as `parts`
// raw partial sums per image in s0part [n_images, parts, k] that still need / T)
int launch_fv_finalize(const float* S, int ld, const float* s0part, int parts, const int64_t* offsets,
                       const pvs_model* gmm, int64_t n_images, float power, float norm_order, float eps,
                       float* out, cudaStream_t st, float raw1 = 0.f, float raw2 = 0.f, const int* raw_gate = nullptr, float rawg = 0.f);
// with *raw_gate != 0 the scales are (rawg, rawg) instead of (raw1, raw2); a scale of 0 means S is S / T already.
// raw1 != 0 (and *raw_gate == 0 when a gate is given): S holds raw sums that still need raw1 / T (first-order columns), raw2 / T

// ---- similarity / top-k --------------------------------------------------------------------
int launch_l2_normalize(const float* x, int64_t n, int64_t d, void* out, int out_dtype, cudaStream_t st);
// uint8 -> float32 (the uint8 transport of the host entry points); n = number of elements, both pointers 16-byte aligned
int launch_u8_to_f32(const uint8_t* in, float* out, size_t n, cudaStream_t st);
int launch_bf16_to_f32(const void* x, int64_t n, float* out, cudaStream_t st);
int launch_f32_to_split(const float* x, int64_t n, void* out, cudaStream_t st);
// s[n, m] = cosine of raw rows, for small n * m (one CTA per pair, no workspace)
int launch_cosine_small(const float* x, int64_t n, const float* y, int64_t m, int64_t d, float* s, cudaStream_t st);
// per-row top-k of a dense score block S [rows, n_db] (row stride lds)
// out_ld: row pitch of the outputs (0 = k); upper_in / last_out: multi-pass ranking for k > PVS_TOPK_MAX
int launch_topk_rows(const float* S, int64_t lds, int64_t rows, int64_t n_db, int k, int64_t idx_offset,
                     float* scores_out, int64_t* idx_out, cudaStream_t st, int64_t out_ld = 0,
                     const unsigned long long* upper_in = nullptr, unsigned long long* last_out = nullptr);
int launch_topk_merge(const float* scores, const int64_t* idx, int parts, int64_t n_q, int k,
                      float* scores_out, int64_t* idx_out, cudaStream_t st);
int launch_label_metrics(const int64_t* idx, const int32_t* db_labels, const int32_t* q_labels,
                         int64_t n_q, int k, int32_t* hits, float* ap, cudaStream_t st);

// ---- tcgen05 tensor-core paths (pvs_tc_*.cu); return PVS_ERR_UNSUPPORTED when the shape
//      is outside what the kernel handles so the caller can take the SIMT path -------------
bool tc_available();
int tc_prepare_model(pvs_model* m);

// VLAD hard assignment (pvs_tc_vlad.cu): 3xTF32 scores + fused arg-min, any d, k <= 256
int tc_prepare_kmeans(pvs_model* m);
bool tc_assign_supported(const pvs_model* km, int64_t rows);
int tc_vlad_assign(const pvs_model* km, const float* x, int64_t rows, int32_t* labels, cudaStream_t st);

// Fisher vector, K = 256 / D = 64: workspace carve-up and the three tensor-core stages
constexpr int TC_FV_S0_PARTS = 16;   // producer warps of the stats kernel, one zeroth-order partial each
struct TcFvPlan {
    float* y;                    // [rows, 64] projected descriptors (NULL without PCA: y = input)
    float* q;                    // [rows, 256] posteriors
    float* S;                    // [n_images, 256, 128] first/second-order sums / T
    float* rinv;                 // [rows + 16, 4] fp16x2 path: per-descriptor softmax normaliser 1 / sum_k e (16-byte slots)
    float* s0part;               // [n_images, TC_FV_S0_PARTS, 256] raw zeroth-order partial sums
    int* flag;                   // right behind s0part (one memset clears both): raised by the projection when
                                 // some |y| leaves the fp16x2 operand range -> the 3xTF32 kernels do the work
    bool fp16x2;                 // this call may use the fp16x2 kernels (eligible model, PCA present)
    int n_tiles;                 // 128-row tiles
    int64_t rows;
    size_t total;
};
bool tc_fv_supported(const pvs_model* g, const pvs_model* pca, int64_t rows, int64_t n_images);
int tc_fv_plan(const pvs_model* g, const pvs_model* pca, int64_t rows, int64_t n_images, void* ws, TcFvPlan* plan);
int tc_fv_begin(const TcFvPlan& pl, int64_t n_images, cudaStream_t st);   // clears s0part + flag
int tc_fv_project(const TcFvPlan& pl, const pvs_model* g, const pvs_model* pca, const float* desc, int64_t rows, cudaStream_t st);
// fallback_only: launch just the gated 3xTF32 kernel (the fp16x2 work was done by the fused kernel)
int tc_fv_posterior(const TcFvPlan& pl, const pvs_model* g, const float* y, int64_t rows, int32_t* argmax, cudaStream_t st,
                    bool fallback_only = false);
int tc_fv_poststats_fused_cluster(const TcFvPlan& pl, const pvs_model* g, const float* y, const int64_t* offsets, int64_t n_images,
                                  cudaStream_t st);
// PVS_FV_FUSED: 0 (default) = posterior kernel + statistics kernel, statistics folded in segments (pvs_tc_fv.cu);
// 2 = both in one kernel, 2-CTA clusters that split the components, same fold (pvs_tc_fvfused2.cu); 1 = one CTA per SM
// (pvs_tc_fvfused.cu, whole-image accumulation: test reference only)
int tc_fv_fused_mode();
inline bool tc_fv_fused_enabled() { return tc_fv_fused_mode() != 0; }
int tc_fv_poststats_fused(const TcFvPlan& pl, const pvs_model* g, const float* y, const int64_t* offsets, int64_t n_images,
                          cudaStream_t st);
int tc_fv_stats(const TcFvPlan& pl, const pvs_model* g, const float* y, const int64_t* offsets, int64_t n_images, cudaStream_t st,
                bool fallback_only = false);

// generic fp32-accurate contraction on tensor cores (pvs_tc_gemmnt.cu), any shape
int tc_prepare_generic(pvs_model* m);
bool tc_gemm_nt_supported(int64_t m, int n);
int tc_gemm_nt(const float* a, int64_t lda, int a_cols, bool square_cat, const float* b_hi, const float* b_lo, int b_ld,
               int n, float* c, int64_t ldc, int64_t m, float alpha, const float* bias, cudaStream_t st);

bool tc_fv_stats_generic_supported(const pvs_model* g, int64_t n_images);
int tc_fv_stats_generic(const float* q, const float* y, int d, const int64_t* offsets, int64_t n_images, float* S,
                        float* s0part, int* smax_dev, cudaStream_t st);   // smax_dev: one int of scratch (segments per image slot)

// similarity + fused top-k on bf16 tensor cores (pvs_tc_sim.cu)
bool tc_sim_supported(int dtype, int64_t n_q, int64_t n_db, int64_t d, int k);
size_t tc_sim_workspace_bytes(int64_t n_q, int64_t n_db, int k);
int tc_sim_topk(const void* q, const void* db, int64_t n_q, int64_t n_db, int64_t d, int k, int64_t idx_offset,
                float* scores_out, int64_t* idx_out, void* ws, size_t ws_bytes, cudaStream_t st);

// fp32-accurate similarity on PVS_F16X2 operand planes (three kind::f16 passes, segmented accumulation):
// fused top-(k + margin) shortlist + exact re-evaluation of the candidates closer than the error bound, and the
// dense score matrix
bool tc_sim3_supported(int64_t n_q, int64_t n_db, int64_t d, int k);
size_t tc_sim3_workspace_bytes(int64_t n_q, int64_t n_db, int k);
size_t tc_sim3_stats_offset(int64_t n_q, int64_t n_db, int k);
int tc_sim3_topk(const void* q, const void* db, int64_t n_q, int64_t n_db, int64_t d, int k, int64_t idx_offset,
                 float* scores_out, int64_t* idx_out, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_sim3_dense(const void* q, const void* db, int64_t n_q, int64_t n_db, int64_t d, float* s, int64_t ld, cudaStream_t st);

}  // namespace pvs
