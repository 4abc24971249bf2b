/*
 * pvs_b200.h  --  C ABI of the B200-native encode-and-compare path for pyvisim.
 *
 * The reference (MechaCritter/Python-Visual-Similarity, pyvisim 0.1.3) is pure Python and
 * has no FFI layer: its operator boundary is the Python class API
 * (VLADEncoder / FisherVectorEncoder / Pipeline  .encode() / .similarity_score()).  Every
 * export below replaces the *body* of one reference routine -- the scikit-learn / NumPy
 * calls under that API -- and is what a ctypes stub inside the reference would bind
 * (INTEGRATION.md shows that stub).  The reference routine each export replaces is cited
 * as path:line relative to the reference root.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++ / torch types.  Every call returns 0 on success or
 *     a negative pvs_status; pvs_last_error() gives the thread-local message.
 *     Nothing throws or exits across this boundary.
 *   - *_dev pointers are device pointers on the current CUDA device; the caller owns every
 *     buffer.  Calls are asynchronous and ordered on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream).  The only hidden device memory is the
 *     per-model constant block owned by a pvs_model handle; scratch comes from the
 *     caller through an explicit workspace whose size pvs_*_workspace_bytes() reports.
 *   - *_host entry points take host pointers, do their own staged H2D / D2H copies and
 *     synchronise before returning (they are what encode() on NumPy arrays calls).
 *   - descriptors are packed row-major fp32 [total_rows, d_in]; image i owns rows
 *     offsets[i] .. offsets[i+1]  (int64, n_images + 1 entries, offsets[0] == 0) -- the
 *     CSR form of the reference's Python loop over images (encoders/vlad.py:87,
 *     encoders/fisher_vector.py:89).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     PVS_ERR_CUDA.
 */
#ifndef PVS_B200_H
#define PVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVS_VERSION 100 /* 0.1.0 */

typedef enum pvs_status {
    PVS_OK = 0,
    PVS_ERR_BAD_ARG = -1,     /* NULL pointer, negative size, unknown enum            */
    PVS_ERR_BAD_SHAPE = -2,   /* dimension mismatch between model / pca / descriptors  */
    PVS_ERR_CUDA = -3,        /* CUDA runtime error (message has the cudaError string) */
    PVS_ERR_WORKSPACE = -4,   /* workspace too small                                   */
    PVS_ERR_UNSUPPORTED = -5, /* e.g. norm_order <= 0                                  */
    PVS_ERR_NCCL = -6         /* NCCL error (message has ncclGetErrorString)           */
} pvs_status;

typedef enum pvs_model_kind { PVS_MODEL_KMEANS = 1, PVS_MODEL_GMM_DIAG = 2, PVS_MODEL_PCA = 3 } pvs_model_kind;
/* PVS_F16X2: a normalised row v as two fp16 planes with v * 2^15 = hi + lo (22 mantissa bits): hi plane [n, d]
 * followed by the lo plane [n, d].  Operand format of the fp32-accurate tensor-core similarity. */
typedef enum pvs_dtype { PVS_F32 = 0, PVS_BF16 = 1, PVS_F16X2 = 2 } pvs_dtype;

/* compute path selector for the contractions (see DESIGN.md "kernels") */
typedef enum pvs_path {
    PVS_PATH_AUTO = 0,   /* tcgen05 tensor-core kernels when the shape allows, else SIMT */
    PVS_PATH_SIMT = 1,   /* fp32 CUDA-core kernels (any K, D)                            */
    PVS_PATH_TENSOR = 2  /* force the tcgen05 kernels; PVS_ERR_UNSUPPORTED if shape can't */
} pvs_path;

typedef struct pvs_model pvs_model; /* opaque; device-resident, immutable after create */
typedef struct pvs_comm pvs_comm;   /* opaque; one NCCL communicator of this process          */

/* ---- library ------------------------------------------------------------------------ */
int pvs_version(void);
const char* pvs_last_error(void);
/* sm count / compute capability of the current device; PVS_ERR_CUDA when there is none */
int pvs_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
/* number of kernel launches issued by this library on this thread since the last reset
 * (bench.py reports it as gpu_launches) */
int64_t pvs_launch_count(void);
void pvs_launch_count_reset(void);
int pvs_set_path(int path); /* pvs_path; process-wide default PVS_PATH_AUTO */
/* Optional per-stage device timing: while enabled every kernel stage is bracketed by CUDA
 * events on the stream it is launched on; pvs_profile_read() folds the finished records
 * and returns the accumulated milliseconds / launch count of one stage. */
int pvs_profile_enable(int on);
int pvs_profile_stage_count(void);
const char* pvs_profile_stage_name(int stage);
int pvs_profile_read(int stage, double* total_ms, int64_t* launches);

/* ---- model handles (weights uploaded once) ------------------------------------------ */
/* K-Means centres, fp32 [k, d].  Replaces the state read by KMeans.predict at
 * encoders/vlad.py:95-96 (cluster_centers_). */
int pvs_kmeans_create(const float* centers_host, int k, int d, pvs_model** out);
/* Diagonal GMM, fp64 [k], [k,d], [k,d], [k,d] exactly as scikit-learn stores them
 * (weights_, means_, covariances_, precisions_cholesky_), read at
 * encoders/fisher_vector.py:95-99.  Derived constants are folded in fp64 on the host. */
int pvs_gmm_create(const double* weights_host, const double* means_host,
                   const double* covariances_host, const double* precisions_cholesky_host,
                   int k, int d, pvs_model** out);
/* PCA (whiten=False): components fp32 [d_out, d_in], mean fp32 [d_in]
 * (encoders/vlad.py:89-90, encoders/fisher_vector.py:91-92). */
int pvs_pca_create(const float* components_host, const float* mean_host, int d_out, int d_in,
                   pvs_model** out);
int pvs_model_destroy(pvs_model* m);
int pvs_model_dims(const pvs_model* m, int* kind, int* k, int* d, int* d_in);

/* ---- a1: PCA projection  (sklearn PCA.transform called at vlad.py:90, fisher_vector.py:92)
 * y[rows, d_out] = x @ C^T - mean @ C^T, fp32 */
int pvs_pca_project(const pvs_model* pca, const float* x_dev, int64_t rows, float* y_dev, void* stream);

/* ---- a2-a4: VLAD  (VLADEncoder.encode body, encoders/vlad.py:88-111) ------------------
 * out_dev: fp32 [n_images, k * d]  (cluster-major; flatten=False is a host-side reshape)
 * labels_out_dev: optional int32 [total_rows] hard assignments (NULL to skip)
 * pca may be NULL.  power = power_norm_weight, norm_order = ord of the per-cluster norm
 * (>0, or INFINITY), eps = epsilon.  Images with zero descriptors produce zero rows here;
 * quirk Q1 (the reference aborts the batch, vlad.py:92-93) is reproduced by the host
 * wrapper, which knows the iteration order. */
size_t pvs_vlad_workspace_bytes(const pvs_model* kmeans, const pvs_model* pca, int64_t total_rows, int64_t n_images);
int pvs_vlad_encode(const pvs_model* kmeans, const pvs_model* pca, const float* desc_dev,
                    const int64_t* offsets_dev, int64_t n_images, int64_t total_rows, float power,
                    float norm_order, float eps, float* out_dev, int32_t* labels_out_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- a5-a8: Fisher vector  (FisherVectorEncoder.encode body, fisher_vector.py:90-133) ---
 * out_dev: fp32 [n_images, 2*k*d + k] = [d_pi | d_mu | d_sigma] per image, power- and
 * globally ord-normalised.  argmax_out_dev: optional int32 [total_rows], arg-max posterior. */
size_t pvs_fv_workspace_bytes(const pvs_model* gmm, const pvs_model* pca, int64_t total_rows, int64_t n_images);
int pvs_fv_encode(const pvs_model* gmm, const pvs_model* pca, const float* desc_dev,
                  const int64_t* offsets_dev, int64_t n_images, int64_t total_rows, float power,
                  float norm_order, float eps, float* out_dev, int32_t* argmax_out_dev,
                  void* workspace_dev, size_t workspace_bytes, void* stream);
/* posterior only: q fp32 [rows, k] (GaussianMixture.predict_proba, fisher_vector.py:99);
 * y_dev is already PCA-projected [rows, d] */
int pvs_gmm_posterior(const pvs_model* gmm, const float* y_dev, int64_t rows, float* q_dev, void* stream);
/* hard assignment only: labels int32 [rows] (KMeans.predict, vlad.py:95).  The score scratch of the
 * CUDA-core path (k > 256) comes from the caller like every other workspace; 0 bytes on the tcgen05 path. */
size_t pvs_kmeans_assign_workspace_bytes(const pvs_model* kmeans, int64_t rows);
int pvs_kmeans_assign(const pvs_model* kmeans, const float* y_dev, int64_t rows, int32_t* labels_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- a10: cosine similarity  (_utils.py:312-330 -> sklearn cosine_similarity) --------- */
/* row L2-normalise fp32 [n, d] -> fp32 or bf16 [n, d], or PVS_F16X2 planes [2, n, d]; zero rows stay zero */
int pvs_l2_normalize_rows(const float* x_dev, int64_t n, int64_t d, void* out_dev, int out_dtype, void* stream);
/* full matrix, fp32: s[n, m] = normalise(x) @ normalise(y)^T.  x/y are RAW (un-normalised)
 * fp32; workspace holds the two normalised copies. */
size_t pvs_cosine_matrix_workspace_bytes(int64_t n, int64_t m, int64_t d);
int pvs_cosine_matrix(const float* x_dev, int64_t n, const float* y_dev, int64_t m, int64_t d,
                      float* s_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- a10 + a11: similarity with fused top-k  (eval.py:37-43, 76-80, 131-132) ----------
 * q_dev [n_q, d], db_dev [n_db, d]: ALREADY row-normalised, dtype fp32, bf16 or PVS_F16X2 planes.
 * bf16: one tensor-core pass, scores within 1e-2.  PVS_F16X2 (and fp32 rows, converted on the fly): three
 * kind::f16 passes with segmented accumulation, a shortlist of k + 8 candidates per query and an exact
 * re-evaluation of every candidate the tensor scores cannot order: indices equal the fp64 ranking of the
 * operands, scores within 1e-4 (tests/test_gpu_sim_exact.py).
 * For every query row the k best database rows by score, ties broken by lowest index
 * (the reference's np.argsort order on exact ties is unspecified).  Scores descending.
 * idx_out = local database row + db_index_offset (global ids for a sharded database).
 * k <= PVS_TOPK_MAX on the fused tensor-core paths; with fp32 operands any k <= n_db is accepted (the
 * reference's default k=None ranks the whole database, eval.py:76-80): beyond PVS_TOPK_MAX the scores of a
 * block of query rows are materialised in the workspace and ranked in passes of PVS_TOPK_MAX. */
#define PVS_TOPK_MAX 1024
size_t pvs_cosine_topk_workspace_bytes(int64_t n_q, int64_t n_db, int64_t d, int k, int dtype);
int pvs_cosine_topk(const void* q_dev, const void* db_dev, int dtype, int64_t n_q, int64_t n_db,
                    int64_t d, int k, int64_t db_index_offset, float* scores_out_dev,
                    int64_t* idx_out_dev, void* workspace_dev, size_t workspace_bytes, void* stream);
/* After a pvs_cosine_topk call that took the fp32-accurate tensor-core path (PVS_F16X2 operands, or PVS_F32 rows
 * on a device with tcgen05): number of shortlisted candidates whose score was re-evaluated exactly because the
 * tensor-core scores could not order them (closer than the error bound), and number of query rows whose ambiguous
 * run reached the end of the shortlist (0 unless many database rows are near-identical).  Synchronises `stream`. */
int pvs_cosine_topk_exact_stats(const void* workspace_dev, int64_t n_q, int64_t n_db, int k, int64_t* rescored,
                                int64_t* unresolved, void* stream);
/* merge `parts` top-k lists per row (scores/idx laid out [parts, n_q, k]) into one
 * [n_q, k] list with the same ordering rule -- used when the DATABASE is sharded. */
int pvs_topk_merge(const float* scores_dev, const int64_t* idx_dev, int parts, int64_t n_q, int k,
                   float* scores_out_dev, int64_t* idx_out_dev, void* stream);

/* ---- C1: the one collective of the path -----------------------------------
 * Row-block sharded retrieval: rank r scores query rows shard(r) against the database, so its
 * top-k lists (eval.py:40-43 per query) are final for those rows; the lists of all ranks are
 * assembled with an NCCL all-gather.  NCCL is loaded at run time (the copy already in the
 * process, e.g. PyTorch's, is preferred; pvs_nccl_load(path) names another one).
 *   pvs_comm_unique_id : rank 0 creates the 128-byte id and ships it to the other ranks by
 *                        any means (torch.distributed broadcast in the Python wrapper)
 *   pvs_comm_create    : ncclCommInitRank on the current device (collective call)
 *   pvs_allgather_topk : scores fp32 / indices int64 [rows_per_rank, k] of every rank ->
 *                        [world * rows_per_rank, k] on every rank, rank-major, on `stream`;
 *                        rows_per_rank is the same on all ranks (the caller pads). */
int pvs_nccl_load(const char* library_path);
int pvs_comm_unique_id(void* id_out_128);
int pvs_comm_create(const void* id_128, int world, int rank, pvs_comm** out);
int pvs_comm_destroy(pvs_comm* comm);
int pvs_allgather_topk(pvs_comm* comm, const float* scores_dev, const int64_t* idx_dev, int64_t rows_per_rank,
                       int k, float* scores_all_dev, int64_t* idx_all_dev, void* stream);

/* ---- f1: label logic on top-k lists  (eval.py:82-98, 126-145) ------------------------- */
/* hits_out[q] = any(db_labels[idx[q, :k]] == query_labels[q]);  ap_out[q] = average precision
 * with the reference's quirk Q6 (R counted inside the truncated list). Either may be NULL. */
int pvs_topk_label_metrics(const int64_t* idx_dev, const int32_t* db_labels_dev,
                           const int32_t* query_labels_dev, int64_t n_q, int k,
                           int32_t* hits_out_dev, float* ap_out_dev, void* stream);

/* ---- f4: learn()  (ImageEncoderBase.learn, encoders/_base_encoder.py:311-342) ----------------
 * The reference fits KMeans (VLAD) or a diagonal GaussianMixture (FV) with scikit-learn.  These calls are
 * ONE pass of the respective iteration over the training descriptors x_dev [rows, d]: the E-step runs
 * the encode path's own kernels (hard assignment / posteriors), the M-step accumulators leave in fp64.
 * The parameter update (division, variance floor, convergence test) is a few hundred scalars and stays
 * with the caller, where sklearn does it (cluster/_kmeans.py:_kmeans_single_lloyd,
 * mixture/_base.py:fit_predict).
 *   pvs_kmeans_lloyd_step : labels[rows] = arg-min (lowest index on ties), sums[k, d] = sum of the members,
 *                           counts[k], inertia[1] = sum ||x - c_label||^2
 *   pvs_gmm_em_step       : s0[k] = sum_t q, s1[k, d] = q^T x, s2[k, d] = q^T x^2,
 *                           loglik[1] = sum_t logsumexp_k (weighted log prob)  (lower bound * rows)
 *   pvs_cluster_sums      : the M-step accumulator alone (labels NULL = every row in cluster 0: column sums)
 *   pvs_rows_sub          : x[r, :] -= v  (sklearn centres X on its mean before Lloyd, _kmeans.py:fit) */
int pvs_rows_sub(float* x_dev, int64_t rows, int d, const float* v_dev, void* stream);
int pvs_cluster_sums(const float* x_dev, const int32_t* labels_dev, int64_t rows, int d, int k,
                     const float* centers_dev, double* sums_dev, int64_t* counts_dev, double* inertia_dev, void* stream);
size_t pvs_kmeans_lloyd_workspace_bytes(const pvs_model* kmeans, int64_t rows);
int pvs_kmeans_lloyd_step(const pvs_model* kmeans, const float* x_dev, int64_t rows, int32_t* labels_dev,
                          double* sums_dev, int64_t* counts_dev, double* inertia_dev, void* workspace_dev,
                          size_t workspace_bytes, void* stream);
size_t pvs_gmm_em_workspace_bytes(const pvs_model* gmm, int64_t rows);
int pvs_gmm_em_step(const pvs_model* gmm, const float* x_dev, int64_t rows, double* s0_dev, double* s1_dev,
                    double* s2_dev, double* loglik_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* ---- host-buffer entry points (what encode()/similarity_func call with NumPy arrays) --
 * Pinned or pageable host pointers; inputs are staged in image chunks of at most
 * chunk_rows descriptor rows (0 = library default) on two streams so H2D, compute and
 * D2H overlap.  They return after the last D2H has completed. */
int pvs_vlad_encode_host(const pvs_model* kmeans, const pvs_model* pca, const float* desc_host,
                         const int64_t* offsets_host, int64_t n_images, float power, float norm_order,
                         float eps, float* out_host, int32_t* labels_out_host, int64_t chunk_rows);
int pvs_fv_encode_host(const pvs_model* gmm, const pvs_model* pca, const float* desc_host,
                       const int64_t* offsets_host, int64_t n_images, float power, float norm_order,
                       float eps, float* out_host, int64_t chunk_rows);
/* uint8 transport (a declared option, not the reference's dtype): descriptors whose values are integers in 0..255 --
 * what OpenCV's SIFT produces (features/_features.py:96-110 returns them as float32) and how they are usually stored --
 * may be passed as uint8 rows.  They cross PCIe in a quarter of the bytes, are widened to float32 on the device and then
 * take exactly the path of the float32 entry points: results are bit-identical to passing the same values as float32. */
int pvs_vlad_encode_host_u8(const pvs_model* kmeans, const pvs_model* pca, const uint8_t* desc_host,
                            const int64_t* offsets_host, int64_t n_images, float power, float norm_order,
                            float eps, float* out_host, int32_t* labels_out_host, int64_t chunk_rows);
int pvs_fv_encode_host_u8(const pvs_model* gmm, const pvs_model* pca, const uint8_t* desc_host,
                          const int64_t* offsets_host, int64_t n_images, float power, float norm_order,
                          float eps, float* out_host, int64_t chunk_rows);
int pvs_cosine_matrix_host(const float* x_host, int64_t n, const float* y_host, int64_t m, int64_t d,
                           float* s_host);
int pvs_cosine_topk_host(const float* q_host, int64_t n_q, const float* db_host, int64_t n_db, int64_t d,
                         int k, int use_bf16, float* scores_out_host, int64_t* idx_out_host);

/* ---- validation hook for the tcgen05 machinery (used by tests/test_gpu_tc.py) ----------
 * mode 0: tf32 single pass, 1: 3xTF32 (hi+lo parts), 2: bf16 -- C[m,n] = A[m,k] B[n,k]^T;
 * mode 3: tf32, 4: 3xTF32 with MN-major operands   -- C[m,n] = A[k,m]^T B[k,n]. */
int pvs_debug_tc_gemm(int mode, const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                      float* c_dev, int m, int n, int k, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVS_B200_H */
