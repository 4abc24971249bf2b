// pvs_api.cu -- the C ABI declared in include/pvs_b200.h: model handles, stage
// orchestration on caller-provided device buffers, and the host-buffer entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

#include "pvs_kernels.cuh"

namespace pvs {

std::string& last_error()
{
    static thread_local std::string s;
    return s;
}
int fail(int status, const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return status;
}
std::atomic<long long> g_launches{0};
std::atomic<int> g_path{PVS_PATH_AUTO};

static int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(PVS_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    return PVS_OK;
}

// rows of the score / logit scratch processed per contraction launch (VLAD assignment)
constexpr int64_t ASSIGN_CHUNK_ROWS = 1 << 18;

// ---- optional per-stage device timing (CUDA events on the launching stream) -------------
enum Stage { ST_PCA = 0, ST_KM_SCORES, ST_KM_ARGMIN, ST_VLAD_AGG, ST_GMM_LOGITS, ST_GMM_SOFTMAX, ST_FV_STATS,
             ST_FV_FINALIZE, ST_L2NORM, ST_SIM_GEMM, ST_TOPK_SELECT, ST_TC_VLAD_ASSIGN, ST_TC_FV_POSTERIOR,
             ST_TC_FV_STATS, ST_TC_SIM_TOPK, ST_TC_FV_PREP, ST_TC_FV_PROJECT, ST_TC_GEMM_PCA, ST_TC_GEMM_LOGITS, ST_TC_FV_FUSED, ST_TC_SIM_DENSE, ST_TC_SIM3_TOPK, ST_COUNT };
static const char* kStageNames[ST_COUNT] = {
    "pca_project", "kmeans_scores", "kmeans_argmin", "vlad_aggregate", "gmm_logits", "gmm_softmax", "fv_stats",
    "fv_finalize", "l2_normalize", "sim_gemm", "topk_select", "tc_vlad_assign", "tc_fv_posterior", "tc_fv_stats",
    "tc_sim_topk", "tc_fv_prep", "tc_fv_project", "tc_gemm_pca", "tc_gemm_logits", "tc_fv_poststats_fused", "tc_sim3_dense", "tc_sim3_topk"};
struct StageRec { int stage; cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::atomic<int> g_prof_on{0};
static std::vector<StageRec> g_prof_recs;
static double g_prof_ms[ST_COUNT];
static long long g_prof_n[ST_COUNT];

struct StageScope {
    int stage; cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; bool on;
    StageScope(int stage_, cudaStream_t st_) : stage(stage_), st(st_), on(g_prof_on.load() != 0)
    {
        if (!on) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { on = false; return; }
        cudaEventRecord(a, st);
    }
    ~StageScope()
    {
        if (!on) return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> l(g_prof_mu);
        g_prof_recs.push_back({stage, a, b});
    }
};
#define PVS_STAGE(id, st, expr) [&]() -> int { ::pvs::StageScope scope__(id, st); return (expr); }()

}  // namespace pvs

using namespace pvs;

// =====================================================================================
// library
// =====================================================================================
extern "C" int pvs_version(void) { return PVS_VERSION; }
extern "C" const char* pvs_last_error(void) { return last_error().c_str(); }
extern "C" int64_t pvs_launch_count(void) { return g_launches.load(); }
extern "C" void pvs_launch_count_reset(void) { g_launches.store(0); }
extern "C" int pvs_set_path(int path)
{
    PVS_CHECK(path >= PVS_PATH_AUTO && path <= PVS_PATH_TENSOR, PVS_ERR_BAD_ARG, "unknown path %d", path);
    g_path.store(path);
    return PVS_OK;
}

extern "C" int pvs_profile_enable(int on)
{
    std::lock_guard<std::mutex> l(g_prof_mu);
    for (auto& r : g_prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof_recs.clear();
    for (int i = 0; i < ST_COUNT; ++i) { g_prof_ms[i] = 0.0; g_prof_n[i] = 0; }
    g_prof_on.store(on ? 1 : 0);
    return PVS_OK;
}
extern "C" int pvs_profile_stage_count(void) { return ST_COUNT; }
extern "C" const char* pvs_profile_stage_name(int stage) { return stage >= 0 && stage < ST_COUNT ? kStageNames[stage] : ""; }
extern "C" int pvs_profile_read(int stage, double* total_ms, int64_t* launches)
{
    PVS_CHECK(stage >= 0 && stage < ST_COUNT, PVS_ERR_BAD_ARG, "unknown stage %d", stage);
    std::lock_guard<std::mutex> l(g_prof_mu);
    for (auto& r : g_prof_recs) {               // fold finished records into the totals
        float ms = 0.f;
        PVS_CUDA(cudaEventSynchronize(r.b));
        PVS_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        g_prof_ms[r.stage] += ms;
        g_prof_n[r.stage] += 1;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof_recs.clear();
    if (total_ms) *total_ms = g_prof_ms[stage];
    if (launches) *launches = g_prof_n[stage];
    return PVS_OK;
}

extern "C" int pvs_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem)
{
    if (int s = require_device()) return s;
    int dev = 0;
    PVS_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    PVS_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return PVS_OK;
}

// =====================================================================================
// models
// =====================================================================================
static int upload_block(pvs_model* m, const std::vector<float>& host)
{
    m->block_bytes = host.size() * sizeof(float);
    PVS_CUDA(cudaMalloc(&m->block, m->block_bytes));
    PVS_CUDA(cudaMemcpy(m->block, host.data(), m->block_bytes, cudaMemcpyHostToDevice));
    return PVS_OK;
}

extern "C" int pvs_kmeans_create(const float* centers, int k, int d, pvs_model** out)
{
    PVS_CHECK(centers && out, PVS_ERR_BAD_ARG, "pvs_kmeans_create: NULL argument");
    PVS_CHECK(k > 0 && d > 0, PVS_ERR_BAD_SHAPE, "pvs_kmeans_create: k=%d d=%d", k, d);
    if (int s = require_device()) return s;
    std::vector<float> h((size_t)k * d + k);
    memcpy(h.data(), centers, (size_t)k * d * sizeof(float));
    for (int j = 0; j < k; ++j) {          // fp32 squared norms, as sklearn's row_norms(squared=True)
        float s = 0.f;
        for (int i = 0; i < d; ++i) s += centers[(size_t)j * d + i] * centers[(size_t)j * d + i];
        h[(size_t)k * d + j] = s;
    }
    pvs_model* m = new pvs_model();
    m->kind = PVS_MODEL_KMEANS; m->k = k; m->d = d; m->d_in = d;
    if (int s = upload_block(m, h)) { delete m; return s; }
    const float* b = (const float*)m->block;
    m->centers = b;
    m->c2 = b + (size_t)k * d;
    if (int s = tc_prepare_kmeans(m)) { pvs_model_destroy(m); return s; }
    *out = m;
    return PVS_OK;
}

extern "C" int pvs_gmm_create(const double* w, const double* mu, const double* cov, const double* pc,
                              int k, int d, pvs_model** out)
{
    PVS_CHECK(w && mu && cov && pc && out, PVS_ERR_BAD_ARG, "pvs_gmm_create: NULL argument");
    PVS_CHECK(k > 0 && d > 0, PVS_ERR_BAD_SHAPE, "pvs_gmm_create: k=%d d=%d", k, d);
    if (int s = require_device()) return s;
    const size_t kd = (size_t)k * d;
    // layout: wcat[k,2d] cst[k] mu[kd] var[kd] pi[k] g_pi[k] g_mu[kd] g_sig[kd]
    std::vector<float> h(2 * kd + k + kd + kd + k + k + kd + kd);
    float* wcat = h.data();
    float* cst = wcat + 2 * kd;
    float* hmu = cst + k;
    float* hvar = hmu + kd;
    float* hpi = hvar + kd;
    float* gpi = hpi + k;
    float* gmu = gpi + k;
    float* gsig = gmu + kd;
    const double ln2pi = log(2.0 * M_PI);
    for (int j = 0; j < k; ++j) {
        double quad = 0.0, logdet = 0.0;
        const double sw = sqrt(w[j]);
        for (int i = 0; i < d; ++i) {
            const size_t e = (size_t)j * d + i;
            const double P = pc[e] * pc[e];                 // precisions = precisions_cholesky_**2
            wcat[(size_t)j * 2 * d + i] = (float)(-0.5 * P);
            wcat[(size_t)j * 2 * d + d + i] = (float)(mu[e] * P);
            quad += mu[e] * mu[e] * P;
            logdet += log(pc[e]);
            hmu[e] = (float)mu[e];
            hvar[e] = (float)cov[e];
            gmu[e] = (float)(1.0 / (sw * sqrt(cov[e])));
            gsig[e] = (float)(1.0 / (sqrt(2.0) * sw * cov[e]));
        }
        cst[j] = (float)(-0.5 * (d * ln2pi + quad) + logdet + log(w[j]));
        hpi[j] = (float)w[j];
        gpi[j] = (float)(1.0 / sw);
    }
    pvs_model* m = new pvs_model();
    m->kind = PVS_MODEL_GMM_DIAG; m->k = k; m->d = d; m->d_in = d;
    {   // power-of-two operand scale of the fp16x2 path: |mu| + 6 sigma <= 128 * 2^h_exp, and the scaled
        // weights must sit well inside fp16's range (degenerate variances, e.g. sklearn's reg_covar
        // floor, rule the model out -> it stays on 3xTF32)
        double r = 0.0;
        for (size_t e = 0; e < kd; ++e) r = fmax(r, fabs(mu[e]) + 6.0 * sqrt(cov[e]));
        if (r > 0.0 && isfinite(r)) {
            const int ex = (int)ceil(log2(r)) - 7;
            double wmax = 0.0;
            for (size_t e = 0; e < kd; ++e) {
                const double P = pc[e] * pc[e];
                wmax = fmax(wmax, fmax(0.5 * P * ldexp(1.0, 2 * ex), fabs(mu[e]) * P * ldexp(1.0, ex)));
            }
            m->h_exp = ex;
            m->h_ok = ex > -40 && ex < 40 && wmax < 16384.0;
        }
    }
    if (int s = upload_block(m, h)) { delete m; return s; }
    m->cst_host.assign(cst, cst + k);
    const float* b = (const float*)m->block;
    m->wcat = b;
    m->cst = b + 2 * kd;
    m->mu = m->cst + k;
    m->var = m->mu + kd;
    m->pi = m->var + kd;
    m->g_pi = m->pi + k;
    m->g_mu = m->g_pi + k;
    m->g_sig = m->g_mu + kd;
    if (int s = tc_prepare_model(m)) { pvs_model_destroy(m); return s; }
    if (int s = tc_prepare_generic(m)) { pvs_model_destroy(m); return s; }
    *out = m;
    return PVS_OK;
}

extern "C" int pvs_pca_create(const float* comp, const float* mean, int d_out, int d_in, pvs_model** out)
{
    PVS_CHECK(comp && mean && out, PVS_ERR_BAD_ARG, "pvs_pca_create: NULL argument");
    PVS_CHECK(d_out > 0 && d_in > 0, PVS_ERR_BAD_SHAPE, "pvs_pca_create: d_out=%d d_in=%d", d_out, d_in);
    if (int s = require_device()) return s;
    std::vector<float> h((size_t)d_out * d_in + d_out);
    memcpy(h.data(), comp, (size_t)d_out * d_in * sizeof(float));
    for (int j = 0; j < d_out; ++j) {      // bias = -(mean @ C^T), fp32 like sklearn's _transform
        float s = 0.f;
        for (int i = 0; i < d_in; ++i) s += mean[i] * comp[(size_t)j * d_in + i];
        h[(size_t)d_out * d_in + j] = -s;
    }
    pvs_model* m = new pvs_model();
    m->kind = PVS_MODEL_PCA; m->k = 0; m->d = d_out; m->d_in = d_in;
    if (int s = upload_block(m, h)) { delete m; return s; }
    m->comp = (const float*)m->block;
    m->bias = m->comp + (size_t)d_out * d_in;
    if (int s = tc_prepare_model(m)) { pvs_model_destroy(m); return s; }
    if (int s = tc_prepare_generic(m)) { pvs_model_destroy(m); return s; }
    *out = m;
    return PVS_OK;
}

extern "C" int pvs_model_destroy(pvs_model* m)
{
    if (!m) return PVS_OK;
    if (m->block) cudaFree(m->block);
    if (m->tc0) cudaFree((void*)m->tc0);
    if (m->tcg0) cudaFree((void*)m->tcg0);
    if (m->th0) cudaFree((void*)m->th0);
    if (m->h_flags) cudaFree(m->h_flags);
    delete m;
    return PVS_OK;
}

extern "C" int pvs_model_dims(const pvs_model* m, int* kind, int* k, int* d, int* d_in)
{
    PVS_CHECK(m, PVS_ERR_BAD_ARG, "pvs_model_dims: NULL model");
    if (kind) *kind = m->kind;
    if (k) *k = m->k;
    if (d) *d = m->d;
    if (d_in) *d_in = m->d_in;
    return PVS_OK;
}

// =====================================================================================
// stages on device buffers
// =====================================================================================
static int check_chain(const pvs_model* cl, int want_kind, const pvs_model* pca, const char* who)
{
    PVS_CHECK(cl && cl->kind == want_kind, PVS_ERR_BAD_ARG, "%s: wrong or NULL clustering model", who);
    if (pca) {
        PVS_CHECK(pca->kind == PVS_MODEL_PCA, PVS_ERR_BAD_ARG, "%s: pca handle is not a PCA model", who);
        PVS_CHECK(pca->d == cl->d, PVS_ERR_BAD_SHAPE, "%s: PCA outputs %d dims but the clustering model takes %d",
                  who, pca->d, cl->d);
    }
    return PVS_OK;
}

extern "C" int pvs_pca_project(const pvs_model* pca, const float* x, int64_t rows, float* y, void* stream)
{
    PVS_CHECK(pca && pca->kind == PVS_MODEL_PCA, PVS_ERR_BAD_ARG, "pvs_pca_project: not a PCA model");
    PVS_CHECK(rows >= 0 && (rows == 0 || (x && y)), PVS_ERR_BAD_ARG, "pvs_pca_project: bad buffers");
    // The tensor core truncates when it accumulates (error grows linearly with the number of MMA
    // steps); for the 514-long VGG16 projection that alone costs 5.6e-5 of the 1e-4 parity budget
    // of the Fisher vector (measured on the golden case), so long contractions stay on the
    // round-to-nearest CUDA-core kernel.
    if (g_path.load() != PVS_PATH_SIMT && pca->tcg0 && pca->d_in <= 256 && tc_gemm_nt_supported(rows, pca->d))
        return PVS_STAGE(ST_TC_GEMM_PCA, (cudaStream_t)stream,
                         tc_gemm_nt(x, pca->d_in, pca->d_in, false, pca->tcg0, pca->tcg1, pca->tcg_ld, pca->d, y, pca->d, rows,
                                    1.f, pca->bias, (cudaStream_t)stream));
    return PVS_STAGE(ST_PCA, (cudaStream_t)stream,
                     launch_gemm_nt(x, pca->d_in, pca->comp, pca->d_in, y, pca->d, rows, pca->d, pca->d_in, 0, 1.f,
                                    pca->bias, (cudaStream_t)stream));
}

static bool assign_use_tc(const pvs_model* km, int64_t rows)
{
    return g_path.load() != PVS_PATH_SIMT && tc_assign_supported(km, rows);
}

extern "C" size_t pvs_kmeans_assign_workspace_bytes(const pvs_model* km, int64_t rows)
{
    if (!km || km->kind != PVS_MODEL_KMEANS || rows <= 0) return 0;
    if (assign_use_tc(km, rows)) return 0;                   // the tcgen05 kernel keeps the scores in TMEM
    const int64_t chunk = rows < ASSIGN_CHUNK_ROWS ? rows : ASSIGN_CHUNK_ROWS;
    return align_up((size_t)chunk * km->k * sizeof(float), 256) + 256;
}

extern "C" int pvs_kmeans_assign(const pvs_model* km, const float* y, int64_t rows, int32_t* labels, void* workspace,
                                 size_t workspace_bytes, void* stream)
{
    PVS_CHECK(km && km->kind == PVS_MODEL_KMEANS, PVS_ERR_BAD_ARG, "pvs_kmeans_assign: not a K-Means model");
    PVS_CHECK(rows >= 0 && (rows == 0 || (y && labels)), PVS_ERR_BAD_ARG, "pvs_kmeans_assign: bad buffers");
    if (rows == 0) return PVS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (assign_use_tc(km, rows)) return PVS_STAGE(ST_TC_VLAD_ASSIGN, st, tc_vlad_assign(km, y, rows, labels, st));
    PVS_CHECK(g_path.load() != PVS_PATH_TENSOR, PVS_ERR_UNSUPPORTED, "pvs_kmeans_assign: the tensor-core path handles k <= 256 only");
    const size_t need = pvs_kmeans_assign_workspace_bytes(km, rows);
    PVS_CHECK(workspace && workspace_bytes >= need, PVS_ERR_WORKSPACE, "pvs_kmeans_assign: workspace %zu < required %zu",
              workspace_bytes, need);
    const int64_t chunk = rows < ASSIGN_CHUNK_ROWS ? rows : ASSIGN_CHUNK_ROWS;
    float* scores = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    for (int64_t r = 0; r < rows; r += chunk) {
        const int64_t n = rows - r < chunk ? rows - r : chunk;
        if (int rc = PVS_STAGE(ST_KM_SCORES, st, launch_gemm_nt(y + r * km->d, km->d, km->centers, km->d, scores, km->k, n, km->k, km->d, 0, -2.f, km->c2, st))) return rc;
        if (int rc = PVS_STAGE(ST_KM_ARGMIN, st, launch_row_argmin(scores, n, km->k, labels + r, st))) return rc;
    }
    return PVS_OK;
}

extern "C" int pvs_gmm_posterior(const pvs_model* g, const float* y, int64_t rows, float* q, void* stream)
{
    PVS_CHECK(g && g->kind == PVS_MODEL_GMM_DIAG, PVS_ERR_BAD_ARG, "pvs_gmm_posterior: not a GMM model");
    PVS_CHECK(rows >= 0 && (rows == 0 || (y && q)), PVS_ERR_BAD_ARG, "pvs_gmm_posterior: bad buffers");
    cudaStream_t st = (cudaStream_t)stream;
    if (g_path.load() != PVS_PATH_SIMT && g->tcg0 && tc_gemm_nt_supported(rows, g->k)) {
        if (int rc = PVS_STAGE(ST_TC_GEMM_LOGITS, st, tc_gemm_nt(y, g->d, g->d, true, g->tcg0, g->tcg1, g->tcg_ld, g->k, q, g->k, rows,
                                                                 1.f, g->cst, st))) return rc;
    } else if (int rc = PVS_STAGE(ST_GMM_LOGITS, st, launch_gemm_nt(y, g->d, g->wcat, 2 * g->d, q, g->k, rows, g->k, g->d, 1, 1.f, g->cst, st))) return rc;
    return PVS_STAGE(ST_GMM_SOFTMAX, st, launch_row_softmax(q, rows, g->k, nullptr, st));
}

// ---- VLAD ------------------------------------------------------------------------------
struct VladWs { size_t y, scores, labels, total; };
static VladWs vlad_ws(const pvs_model* km, const pvs_model* pca, int64_t rows)
{
    VladWs w{};
    const int64_t chunk = rows < ASSIGN_CHUNK_ROWS ? rows : ASSIGN_CHUNK_ROWS;
    size_t off = 0;
    w.y = off;      off += pca ? align_up((size_t)rows * km->d * 4, 256) : 0;
    w.scores = off; off += align_up((size_t)chunk * km->k * 4, 256);
    w.labels = off; off += align_up((size_t)rows * 4, 256);
    w.total = off + 256;
    return w;
}

extern "C" size_t pvs_vlad_workspace_bytes(const pvs_model* km, const pvs_model* pca, int64_t total_rows, int64_t)
{
    if (!km || total_rows < 0) return 0;
    return vlad_ws(km, pca, total_rows).total;
}

extern "C" int pvs_vlad_encode(const pvs_model* km, const pvs_model* pca, const float* desc,
                               const int64_t* offsets, int64_t n_images, int64_t total_rows, float power,
                               float norm_order, float eps, float* out, int32_t* labels_out, void* workspace,
                               size_t workspace_bytes, void* stream)
{
    if (int s = check_chain(km, PVS_MODEL_KMEANS, pca, "pvs_vlad_encode")) return s;
    PVS_CHECK(n_images >= 0 && total_rows >= 0, PVS_ERR_BAD_ARG, "pvs_vlad_encode: negative size");
    PVS_CHECK(norm_order > 0.f, PVS_ERR_UNSUPPORTED, "norm_order must be > 0 (or inf), got %g", (double)norm_order);
    if (n_images == 0) return PVS_OK;
    PVS_CHECK(offsets && out && (total_rows == 0 || desc), PVS_ERR_BAD_ARG, "pvs_vlad_encode: NULL buffer");
    const VladWs w = vlad_ws(km, pca, total_rows);
    PVS_CHECK(workspace && workspace_bytes >= w.total, PVS_ERR_WORKSPACE,
              "pvs_vlad_encode: workspace %zu < required %zu", workspace_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const float* y = desc;
    if (pca) {
        float* yp = (float*)(ws + w.y);
        if (int rc = pvs_pca_project(pca, desc, total_rows, yp, stream)) return rc;
        y = yp;
    }
    float* scores = (float*)(ws + w.scores);
    int32_t* labels = labels_out ? labels_out : (int32_t*)(ws + w.labels);
    const int64_t chunk = total_rows < ASSIGN_CHUNK_ROWS ? total_rows : ASSIGN_CHUNK_ROWS;
    const bool tc = assign_use_tc(km, total_rows);
    PVS_CHECK(tc || g_path.load() != PVS_PATH_TENSOR, PVS_ERR_UNSUPPORTED,
              "pvs_vlad_encode: the tensor-core assignment handles k <= 256 only");
    if (tc)
        if (int rc = PVS_STAGE(ST_TC_VLAD_ASSIGN, st, tc_vlad_assign(km, y, total_rows, labels, st))) return rc;
    for (int64_t r = 0; r < total_rows && !tc; r += chunk) {
        const int64_t n = total_rows - r < chunk ? total_rows - r : chunk;
        if (int rc = PVS_STAGE(ST_KM_SCORES, st, launch_gemm_nt(y + r * km->d, km->d, km->centers, km->d, scores, km->k, n,
                                                                km->k, km->d, 0, -2.f, km->c2, st))) return rc;
        if (int rc = PVS_STAGE(ST_KM_ARGMIN, st, launch_row_argmin(scores, n, km->k, labels + r, st))) return rc;
    }
    return PVS_STAGE(ST_VLAD_AGG, st, launch_vlad_aggregate(y, km->d, labels, offsets, n_images, total_rows, km->centers, km->k,
                                                            power, norm_order, eps, out, st));
}

// ---- Fisher vector ---------------------------------------------------------------------
struct FvWs { size_t y, q, s, s0, smax, total; };
static FvWs fv_ws(const pvs_model* g, const pvs_model* pca, int64_t rows, int64_t n_images)
{
    FvWs w{};
    size_t off = 0;
    w.y = off; off += pca ? align_up((size_t)rows * g->d * 4, 256) : 0;
    w.q = off; off += align_up((size_t)rows * g->k * 4, 256);
    w.s = off; off += align_up((size_t)n_images * g->k * (2 * g->d + 1) * 4, 256);
    w.s0 = off; off += align_up((size_t)n_images * TC_FV_S0_PARTS * g->k * 4, 256);
    w.smax = off; off += 256;                              // two ints: statistics segments per image slot
    w.total = off + 256;
    return w;
}

static bool fv_use_tc(const pvs_model* g, const pvs_model* pca, int64_t rows, int64_t n_images)
{
    return g_path.load() != PVS_PATH_SIMT && tc_fv_supported(g, pca, rows, n_images) && !getenv("PVS_FV_FORCE_GENERIC");
}

extern "C" size_t pvs_fv_workspace_bytes(const pvs_model* g, const pvs_model* pca, int64_t total_rows, int64_t n_images)
{
    if (!g || total_rows < 0 || n_images < 0) return 0;
    if (g->kind == PVS_MODEL_GMM_DIAG && fv_use_tc(g, pca, total_rows, n_images)) {
        TcFvPlan pl;
        tc_fv_plan(g, pca, total_rows, n_images, nullptr, &pl);
        return pl.total;
    }
    return fv_ws(g, pca, total_rows, n_images).total;
}

extern "C" int pvs_fv_encode(const pvs_model* g, const pvs_model* pca, const float* desc, const int64_t* offsets,
                             int64_t n_images, int64_t total_rows, float power, float norm_order, float eps,
                             float* out, int32_t* argmax_out, void* workspace, size_t workspace_bytes, void* stream)
{
    if (int s = check_chain(g, PVS_MODEL_GMM_DIAG, pca, "pvs_fv_encode")) return s;
    PVS_CHECK(n_images >= 0 && total_rows >= 0, PVS_ERR_BAD_ARG, "pvs_fv_encode: negative size");
    PVS_CHECK(norm_order > 0.f, PVS_ERR_UNSUPPORTED, "norm_order must be > 0 (or inf), got %g", (double)norm_order);
    if (n_images == 0) return PVS_OK;
    PVS_CHECK(offsets && out && (total_rows == 0 || desc), PVS_ERR_BAD_ARG, "pvs_fv_encode: NULL buffer");
    if (fv_use_tc(g, pca, total_rows, n_images)) {
        // tcgen05 path: 3xTF32 contractions, see pvs_tc_fv.cu
        cudaStream_t st = (cudaStream_t)stream;
        TcFvPlan pl;
        tc_fv_plan(g, pca, total_rows, n_images, nullptr, &pl);
        PVS_CHECK(workspace && workspace_bytes >= pl.total, PVS_ERR_WORKSPACE,
                  "pvs_fv_encode: workspace %zu < required %zu", workspace_bytes, pl.total);
        char* ws = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
        tc_fv_plan(g, pca, total_rows, n_images, ws, &pl);
        const float* y = pca ? pl.y : desc;
        if (int rc = tc_fv_begin(pl, n_images, st)) return rc;
        if (pca)
            if (int rc = PVS_STAGE(ST_TC_FV_PROJECT, st, tc_fv_project(pl, g, pca, desc, total_rows, st))) return rc;
        float raw1 = 0.f, raw2 = 0.f;
        if (pl.fp16x2 && !argmax_out && tc_fv_fused_enabled()) {
            // posterior + statistics in one kernel (the posteriors never leave the SM); behind it the two
            // 3xTF32 kernels that only run when the projection raised the range flag
            const int mode = tc_fv_fused_mode();
            if (mode == 2) { raw1 = ldexpf(1.f, g->h_exp - 14); raw2 = ldexpf(1.f, 2 * g->h_exp - 14); }   // raw sums, folded in global memory
            if (int rc = PVS_STAGE(ST_TC_FV_FUSED, st, mode == 2 ? tc_fv_poststats_fused_cluster(pl, g, y, offsets, n_images, st)
                                                                 : tc_fv_poststats_fused(pl, g, y, offsets, n_images, st))) return rc;
            if (int rc = PVS_STAGE(ST_TC_FV_POSTERIOR, st, tc_fv_posterior(pl, g, y, total_rows, argmax_out, st, true))) return rc;
            if (int rc = PVS_STAGE(ST_TC_FV_STATS, st, tc_fv_stats(pl, g, y, offsets, n_images, st, true))) return rc;
        } else {
            // the fp16x2 statistics kernel folds raw segment sums into S as well
            if (pl.fp16x2) { raw1 = ldexpf(1.f, g->h_exp - 14); raw2 = ldexpf(1.f, 2 * g->h_exp - 14); }
            else raw1 = raw2 = 1.f;                          // 3xTF32 kernels only: raw sums in plain units
            if (int rc = PVS_STAGE(ST_TC_FV_POSTERIOR, st, tc_fv_posterior(pl, g, y, total_rows, argmax_out, st))) return rc;
            if (int rc = PVS_STAGE(ST_TC_FV_STATS, st, tc_fv_stats(pl, g, y, offsets, n_images, st))) return rc;
        }
        // every tensor statistics kernel leaves raw segment-folded sums in S, the fp16x2 ones in operand units (raw1, raw2),
        // the 3xTF32 fallback behind the range flag in plain units (rawg = 1): fv_finalize reads the flag and applies the
        // matching scale / T.  Only the single-CTA fused kernel (PVS_FV_FUSED=1) writes S / T (raw1 = 0).
        return PVS_STAGE(ST_FV_FINALIZE, st, launch_fv_finalize(pl.S, 2 * g->d, pl.s0part, TC_FV_S0_PARTS, offsets, g, n_images, power,
                                                                 norm_order, eps, out, st, raw1, raw2, pl.flag, 1.f));
    }
    PVS_CHECK(g_path.load() != PVS_PATH_TENSOR, PVS_ERR_UNSUPPORTED,
              "pvs_fv_encode: the tensor-core path handles K=256, D=64 (d_in %% 32 == 0) only");
    const FvWs w = fv_ws(g, pca, total_rows, n_images);
    PVS_CHECK(workspace && workspace_bytes >= w.total, PVS_ERR_WORKSPACE,
              "pvs_fv_encode: workspace %zu < required %zu", workspace_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const float* y = desc;
    if (pca) {
        float* yp = (float*)(ws + w.y);
        if (int rc = pvs_pca_project(pca, desc, total_rows, yp, stream)) return rc;
        y = yp;
    }
    float* q = (float*)(ws + w.q);
    float* S = (float*)(ws + w.s);
    if (g_path.load() != PVS_PATH_SIMT && g->tcg0 && tc_gemm_nt_supported(total_rows, g->k)) {
        if (int rc = PVS_STAGE(ST_TC_GEMM_LOGITS, st, tc_gemm_nt(y, g->d, g->d, true, g->tcg0, g->tcg1, g->tcg_ld, g->k, q, g->k,
                                                                 total_rows, 1.f, g->cst, st))) return rc;
    } else if (int rc = PVS_STAGE(ST_GMM_LOGITS, st, launch_gemm_nt(y, g->d, g->wcat, 2 * g->d, q, g->k, total_rows, g->k, g->d, 1,
                                                                    1.f, g->cst, st))) return rc;
    if (int rc = PVS_STAGE(ST_GMM_SOFTMAX, st, launch_row_softmax(q, total_rows, g->k, argmax_out, st))) return rc;
    if (g_path.load() != PVS_PATH_SIMT && tc_fv_stats_generic_supported(g, n_images)) {
        float* s0part = (float*)(ws + w.s0);
        if (int rc = PVS_STAGE(ST_TC_FV_STATS, st, tc_fv_stats_generic(q, y, g->d, offsets, n_images, S, s0part, (int*)(ws + w.smax), st))) return rc;
        return PVS_STAGE(ST_FV_FINALIZE, st, launch_fv_finalize(S, 2 * g->d, s0part, TC_FV_S0_PARTS, offsets, g, n_images, power,
                                                                 norm_order, eps, out, st, 1.f, 1.f));   // raw sums, plain units
    }
    if (int rc = PVS_STAGE(ST_FV_STATS, st, launch_fv_stats(q, y, g->d, g->k, offsets, n_images, S, st))) return rc;
    return PVS_STAGE(ST_FV_FINALIZE, st, launch_fv_finalize(S, 2 * g->d + 1, nullptr, 0, offsets, g, n_images, power,
                                                             norm_order, eps, out, st));
}

// ---- similarity ----------------------------------------------------------------------------
extern "C" int pvs_l2_normalize_rows(const float* x, int64_t n, int64_t d, void* out, int out_dtype, void* stream)
{
    PVS_CHECK(n >= 0 && d > 0 && (n == 0 || (x && out)), PVS_ERR_BAD_ARG, "pvs_l2_normalize_rows: bad arguments");
    return PVS_STAGE(ST_L2NORM, (cudaStream_t)stream, launch_l2_normalize(x, n, d, out, out_dtype, (cudaStream_t)stream));
}

extern "C" size_t pvs_cosine_matrix_workspace_bytes(int64_t n, int64_t m, int64_t d)
{
    if (n < 0 || m < 0 || d <= 0) return 0;
    return align_up((size_t)n * d * 4, 256) + align_up((size_t)m * d * 4, 256) + 256;
}

static bool matrix_use_tc(int64_t n, int64_t m, int64_t d)
{
    return g_path.load() != PVS_PATH_SIMT && tc_sim3_supported(n, m, d, 1);
}

extern "C" int pvs_cosine_matrix(const float* x, int64_t n, const float* y, int64_t m, int64_t d, float* s,
                                 void* workspace, size_t workspace_bytes, void* stream)
{
    PVS_CHECK(n >= 0 && m >= 0, PVS_ERR_BAD_ARG, "pvs_cosine_matrix: negative size");
    PVS_CHECK(d >= 2, PVS_ERR_BAD_SHAPE, "Cosine similarity requires at least 2 features. Got %lld", (long long)d);
    if (n == 0 || m == 0) return PVS_OK;
    PVS_CHECK(x && y && s, PVS_ERR_BAD_ARG, "pvs_cosine_matrix: NULL buffer");
    PVS_CHECK(d < 2147483647LL && m < 2147483647LL, PVS_ERR_BAD_SHAPE, "pvs_cosine_matrix: dimension too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (n * m <= 1024) return launch_cosine_small(x, n, y, m, d, s, st);     // e.g. similarity_score of two images
    const size_t need = pvs_cosine_matrix_workspace_bytes(n, m, d);
    PVS_CHECK(workspace && workspace_bytes >= need, PVS_ERR_WORKSPACE, "pvs_cosine_matrix: workspace %zu < %zu",
              workspace_bytes, need);
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    if (matrix_use_tc(n, m, d)) {
        // tcgen05: rows normalised straight into fp16 hi + lo planes (the same 4 bytes per element as an fp32 copy),
        // three kind::f16 passes with segmented accumulation (pvs_tc_sim.cu: SimSplitPolicy<.., DENSE>)
        void* xp = ws;
        void* yp = ws + align_up((size_t)n * d * 4, 256);
        if (int rc = PVS_STAGE(ST_L2NORM, st, launch_l2_normalize(x, n, d, xp, PVS_F16X2, st))) return rc;
        if (int rc = PVS_STAGE(ST_L2NORM, st, launch_l2_normalize(y, m, d, yp, PVS_F16X2, st))) return rc;
        return PVS_STAGE(ST_TC_SIM_DENSE, st, tc_sim3_dense(xp, yp, n, m, d, s, m, st));
    }
    PVS_CHECK(g_path.load() != PVS_PATH_TENSOR, PVS_ERR_UNSUPPORTED,
              "pvs_cosine_matrix: the tensor-core path needs d %% 8 == 0 and d >= 64");
    float* xn = (float*)ws;
    float* yn = (float*)(ws + align_up((size_t)n * d * 4, 256));
    if (int rc = launch_l2_normalize(x, n, d, xn, PVS_F32, st)) return rc;
    if (int rc = launch_l2_normalize(y, m, d, yn, PVS_F32, st)) return rc;
    return PVS_STAGE(ST_SIM_GEMM, st, launch_gemm_nt(xn, d, yn, d, s, m, n, (int)m, (int)d, 0, 1.f, nullptr, st));
}

// S[n, n_db] = Q . DB^T for ALREADY normalised fp32 rows (CUDA-core path)
static int dense_scores(const float* qf, const float* dbf, int64_t n, int64_t n_db, int64_t d, float* block, cudaStream_t st)
{
    return launch_gemm_nt(qf, d, dbf, d, block, n_db, n, (int)n_db, (int)d, 0, 1.f, nullptr, st);
}

// query rows scored per pass by the dense-block top-k path
static int64_t topk_block_rows(int64_t n_q, int64_t n_db)
{
    int64_t qb = (int64_t)(256ll << 20) / (n_db * 4 > 0 ? n_db * 4 : 1);
    if (qb < 1) qb = 1;
    if (qb > n_q) qb = n_q;
    return qb;
}

static bool sim_use_tc(int dtype, int64_t n_q, int64_t n_db, int64_t d, int k)
{
    return g_path.load() != PVS_PATH_SIMT && tc_sim_supported(dtype, n_q, n_db, d, k);
}
// fp32-accurate tensor path (split operands).  Planes always take it; fp32 rows are converted first, which only
// pays off above a few million multiply-adds (below that the dense CUDA-core block is as fast).
static bool sim3_use_tc(int dtype, int64_t n_q, int64_t n_db, int64_t d, int k)
{
    if (g_path.load() == PVS_PATH_SIMT || !tc_sim3_supported(n_q, n_db, d, k)) return false;
    if (dtype == PVS_F16X2) return true;
    return dtype == PVS_F32 && (g_path.load() == PVS_PATH_TENSOR || (double)n_q * (double)n_db * (double)d >= 16777216.0);
}

extern "C" size_t pvs_cosine_topk_workspace_bytes(int64_t n_q, int64_t n_db, int64_t d, int k, int dtype)
{
    if (n_q < 0 || n_db <= 0 || d <= 0 || k <= 0) return 0;
    if (sim_use_tc(dtype, n_q, n_db, d, k)) return tc_sim_workspace_bytes(n_q, n_db, k);
    if (sim3_use_tc(dtype, n_q, n_db, d, k))
        return tc_sim3_workspace_bytes(n_q, n_db, k) +
               (dtype == PVS_F32 ? align_up((size_t)n_q * d * 4, 1024) + align_up((size_t)n_db * d * 4, 1024) + 1024 : 0);
    const int64_t qb = topk_block_rows(n_q, n_db);
    size_t b = align_up((size_t)qb * n_db * 4, 256) + 2 * align_up((size_t)qb * 8, 256) + 256;
    if (dtype == PVS_BF16) b += align_up((size_t)n_q * d * 4, 256) + align_up((size_t)n_db * d * 4, 256);
    return b;
}

extern "C" int pvs_cosine_topk(const void* q, const void* db, int dtype, int64_t n_q, int64_t n_db, int64_t d, int k,
                               int64_t db_index_offset, float* scores_out, int64_t* idx_out, void* workspace,
                               size_t workspace_bytes, void* stream)
{
    PVS_CHECK(n_q >= 0 && n_db > 0 && d >= 2, PVS_ERR_BAD_SHAPE, "pvs_cosine_topk: bad shape");
    PVS_CHECK(k >= 1 && (k <= PVS_TOPK_MAX || k <= n_db), PVS_ERR_BAD_ARG,
              "k must be in [1, max(%d, n_db = %lld)] (got %d)", PVS_TOPK_MAX, (long long)n_db, k);
    PVS_CHECK(dtype == PVS_F32 || dtype == PVS_BF16 || dtype == PVS_F16X2, PVS_ERR_BAD_ARG, "unknown dtype %d", dtype);
    if (n_q == 0) return PVS_OK;
    PVS_CHECK(q && db && scores_out && idx_out, PVS_ERR_BAD_ARG, "pvs_cosine_topk: NULL buffer");
    PVS_CHECK(d < 2147483647LL && n_db < 2147483647LL, PVS_ERR_BAD_SHAPE, "pvs_cosine_topk: dimension too large");
    if (sim_use_tc(dtype, n_q, n_db, d, k))
        return PVS_STAGE(ST_TC_SIM_TOPK, (cudaStream_t)stream,
                         tc_sim_topk(q, db, n_q, n_db, d, k, db_index_offset, scores_out, idx_out, workspace,
                                     workspace_bytes, (cudaStream_t)stream));
    if (sim3_use_tc(dtype, n_q, n_db, d, k)) {
        cudaStream_t st = (cudaStream_t)stream;
        const size_t need = pvs_cosine_topk_workspace_bytes(n_q, n_db, d, k, dtype);
        PVS_CHECK(workspace && workspace_bytes >= need, PVS_ERR_WORKSPACE, "pvs_cosine_topk: workspace %zu < %zu",
                  workspace_bytes, need);
        const void* qp = q;
        const void* dbp = db;
        const size_t core = tc_sim3_workspace_bytes(n_q, n_db, k);
        if (dtype == PVS_F32) {                                 // normalised fp32 rows -> fp16 hi + lo planes behind the core workspace
            char* base = (char*)(((uintptr_t)workspace + core + 1023) & ~(uintptr_t)1023);
            void* q2 = base;
            void* db2 = base + align_up((size_t)n_q * d * 4, 1024);
            if (int rc = launch_f32_to_split((const float*)q, n_q * d, q2, st)) return rc;
            if (int rc = launch_f32_to_split((const float*)db, n_db * d, db2, st)) return rc;
            qp = q2;
            dbp = db2;
        }
        return PVS_STAGE(ST_TC_SIM3_TOPK, st, tc_sim3_topk(qp, dbp, n_q, n_db, d, k, db_index_offset, scores_out, idx_out,
                                                          workspace, core, st));
    }
    PVS_CHECK(dtype != PVS_F16X2, PVS_ERR_UNSUPPORTED,
              "pvs_cosine_topk: split (PVS_F16X2) operands need the tensor-core path: d %% 8 == 0, d >= 64, k <= %d",
              PVS_TOPK_MAX - 8);
    PVS_CHECK(g_path.load() != PVS_PATH_TENSOR, PVS_ERR_UNSUPPORTED,
              "pvs_cosine_topk: the tensor-core paths handle d %% 8 == 0, d >= 64, k <= %d only", PVS_TOPK_MAX - 8);
    const size_t need = pvs_cosine_topk_workspace_bytes(n_q, n_db, d, k, dtype);
    PVS_CHECK(workspace && workspace_bytes >= need, PVS_ERR_WORKSPACE, "pvs_cosine_topk: workspace %zu < %zu",
              workspace_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const int64_t qb = topk_block_rows(n_q, n_db);
    float* block = (float*)ws;
    char* p = ws + align_up((size_t)qb * n_db * 4, 256);
    unsigned long long* bound[2] = {(unsigned long long*)p, (unsigned long long*)(p + align_up((size_t)qb * 8, 256))};
    p += 2 * align_up((size_t)qb * 8, 256);
    const float* qf = (const float*)q;
    const float* dbf = (const float*)db;
    if (dtype == PVS_BF16) {
        float* q32 = (float*)p;
        float* db32 = (float*)(p + align_up((size_t)n_q * d * 4, 256));
        if (int rc = launch_bf16_to_f32(q, n_q * d, q32, st)) return rc;
        if (int rc = launch_bf16_to_f32(db, n_db * d, db32, st)) return rc;
        qf = q32;
        dbf = db32;
    }
    for (int64_t r = 0; r < n_q; r += qb) {
        const int64_t n = n_q - r < qb ? n_q - r : qb;
        if (int rc = PVS_STAGE(ST_SIM_GEMM, st, dense_scores(qf + r * d, dbf, n, n_db, d, block, st))) return rc;
        // k <= PVS_TOPK_MAX: one selection pass.  Larger k (the reference's default k=None ranks the whole
        // database, eval.py:76-80): pass p selects the next PVS_TOPK_MAX keys below the last key of pass p - 1.
        for (int done = 0, pass = 0; done < k; done += PVS_TOPK_MAX, ++pass) {
            const int kk = k - done < PVS_TOPK_MAX ? k - done : PVS_TOPK_MAX;
            if (int rc = PVS_STAGE(ST_TOPK_SELECT, st, launch_topk_rows(block, n_db, n, n_db, kk, db_index_offset, scores_out + r * k + done,
                                                                        idx_out + r * k + done, st, k, pass ? bound[(pass - 1) & 1] : nullptr,
                                                                        done + kk < k ? bound[pass & 1] : nullptr))) return rc;
        }
    }
    return PVS_OK;
}

extern "C" int pvs_cosine_topk_exact_stats(const void* workspace, int64_t n_q, int64_t n_db, int k, int64_t* rescored,
                                           int64_t* unresolved, void* stream)
{
    PVS_CHECK(workspace && n_q > 0 && n_db > 0 && k >= 1, PVS_ERR_BAD_ARG, "pvs_cosine_topk_exact_stats: bad arguments");
    const char* base = (const char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    unsigned long long h[2] = {0, 0};
    PVS_CUDA(cudaMemcpyAsync(h, base + tc_sim3_stats_offset(n_q, n_db, k), sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PVS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (rescored) *rescored = (int64_t)h[0];
    if (unresolved) *unresolved = (int64_t)h[1];
    return PVS_OK;
}

extern "C" int pvs_topk_merge(const float* scores, const int64_t* idx, int parts, int64_t n_q, int k,
                              float* scores_out, int64_t* idx_out, void* stream)
{
    PVS_CHECK(n_q >= 0, PVS_ERR_BAD_ARG, "pvs_topk_merge: negative size");
    if (n_q == 0) return PVS_OK;
    PVS_CHECK(scores && idx && scores_out && idx_out, PVS_ERR_BAD_ARG, "pvs_topk_merge: NULL buffer");
    return launch_topk_merge(scores, idx, parts, n_q, k, scores_out, idx_out, (cudaStream_t)stream);
}

extern "C" int pvs_topk_label_metrics(const int64_t* idx, const int32_t* db_labels, const int32_t* q_labels,
                                      int64_t n_q, int k, int32_t* hits, float* ap, void* stream)
{
    PVS_CHECK(n_q >= 0 && k >= 1, PVS_ERR_BAD_ARG, "pvs_topk_label_metrics: bad size");
    if (n_q == 0) return PVS_OK;
    PVS_CHECK(idx && db_labels && q_labels, PVS_ERR_BAD_ARG, "pvs_topk_label_metrics: NULL buffer");
    return launch_label_metrics(idx, db_labels, q_labels, n_q, k, hits, ap, (cudaStream_t)stream);
}

// =====================================================================================
// host-buffer entry points
// =====================================================================================
namespace {
struct Slot {
    cudaStream_t stream = nullptr;
    void* buf = nullptr;
    size_t bytes = 0;
    int64_t* offs_pinned = nullptr;
    size_t offs_cap = 0;
};
struct HostCtx {
    std::mutex mu;
    Slot slot[2];
};
// one context per device ordinal: the streams and the staging arena belong to a device
HostCtx& host_ctx()
{
    static HostCtx c[64];
    int dev = 0;
    cudaGetDevice(&dev);
    return c[dev & 63];
}
int slot_reserve(Slot& s, size_t bytes, size_t n_offs)
{
    if (!s.stream) PVS_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (s.bytes < bytes) {
        if (s.buf) PVS_CUDA(cudaFree(s.buf));
        s.buf = nullptr; s.bytes = 0;
        PVS_CUDA(cudaMalloc(&s.buf, bytes));
        s.bytes = bytes;
    }
    if (s.offs_cap < n_offs) {
        if (s.offs_pinned) PVS_CUDA(cudaFreeHost(s.offs_pinned));
        s.offs_pinned = nullptr; s.offs_cap = 0;
        PVS_CUDA(cudaMallocHost((void**)&s.offs_pinned, n_offs * sizeof(int64_t)));
        s.offs_cap = n_offs;
    }
    return PVS_OK;
}

// shared driver for the two encoders: chunk images so a chunk has <= chunk_rows rows
// desc_u8 != NULL: uint8 transport -- the chunk crosses PCIe as bytes into a staging area and is widened on the device
template <typename EncodeFn, typename WsFn>
int encode_host(const float* desc_host, const int64_t* offsets_host, int64_t n_images, int d_in, int64_t out_dim,
                float* out_host, int32_t* rows_i32_out_host, int64_t chunk_rows, WsFn ws_bytes, EncodeFn encode,
                const uint8_t* desc_u8 = nullptr)
{
    PVS_CHECK(n_images >= 0, PVS_ERR_BAD_ARG, "negative n_images");
    if (n_images == 0) return PVS_OK;
    PVS_CHECK(offsets_host && out_host, PVS_ERR_BAD_ARG, "NULL host buffer");
    PVS_CHECK(offsets_host[0] == 0, PVS_ERR_BAD_ARG, "offsets[0] must be 0");
    for (int64_t i = 0; i < n_images; ++i)
        PVS_CHECK(offsets_host[i + 1] >= offsets_host[i], PVS_ERR_BAD_ARG, "offsets must be non-decreasing");
    PVS_CHECK(offsets_host[n_images] == 0 || desc_host || desc_u8, PVS_ERR_BAD_ARG, "NULL descriptor buffer");
    if (int s = require_device()) return s;
    if (chunk_rows <= 0) chunk_rows = 1 << 19;
    HostCtx& ctx = host_ctx();
    std::lock_guard<std::mutex> lock(ctx.mu);

    int64_t i0 = 0;
    int which = 0;
    int rc = PVS_OK;
    while (i0 < n_images && rc == PVS_OK) {
        // greedy chunk: at least one image, at most chunk_rows rows, at most 65535*8 images
        int64_t i1 = i0 + 1;
        while (i1 < n_images && offsets_host[i1 + 1] - offsets_host[i0] <= chunk_rows && i1 - i0 < (1 << 19)) ++i1;
        const int64_t n = i1 - i0, rows = offsets_host[i1] - offsets_host[i0];
        Slot& s = ctx.slot[which];
        which ^= 1;
        const size_t b_desc = align_up((size_t)rows * d_in * 4, 256);
        const size_t b_offs = align_up((size_t)(n + 1) * 8, 256);
        const size_t b_out = align_up((size_t)n * out_dim * 4, 256);
        const size_t b_lab = rows_i32_out_host ? align_up((size_t)rows * 4, 256) : 0;
        const size_t b_ws = ws_bytes(rows, n);
        const size_t b_u8 = desc_u8 ? align_up((size_t)rows * d_in, 256) : 0;
        if (s.stream) PVS_CUDA(cudaStreamSynchronize(s.stream));      // previous use of this slot finished
        if ((rc = slot_reserve(s, b_desc + b_offs + b_out + b_lab + b_ws + b_u8 + 256, (size_t)n + 1))) break;
        char* p = (char*)s.buf;
        float* d_desc = (float*)p;            p += b_desc;
        int64_t* d_offs = (int64_t*)p;        p += b_offs;
        float* d_out = (float*)p;             p += b_out;
        int32_t* d_lab = rows_i32_out_host ? (int32_t*)p : nullptr; p += b_lab;
        void* d_ws = p;                       p += b_ws;
        uint8_t* d_u8 = (uint8_t*)p;
        for (int64_t i = 0; i <= n; ++i) s.offs_pinned[i] = offsets_host[i0 + i] - offsets_host[i0];
        if (rows > 0 && desc_u8) {
            PVS_CUDA(cudaMemcpyAsync(d_u8, desc_u8 + offsets_host[i0] * (int64_t)d_in, (size_t)rows * d_in, cudaMemcpyHostToDevice, s.stream));
            if ((rc = launch_u8_to_f32(d_u8, d_desc, (size_t)rows * d_in, s.stream))) break;
        } else if (rows > 0)
            PVS_CUDA(cudaMemcpyAsync(d_desc, desc_host + offsets_host[i0] * (int64_t)d_in, (size_t)rows * d_in * 4,
                                     cudaMemcpyHostToDevice, s.stream));
        PVS_CUDA(cudaMemcpyAsync(d_offs, s.offs_pinned, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s.stream));
        rc = encode(d_desc, d_offs, n, rows, d_out, d_lab, d_ws, b_ws, s.stream);
        if (rc != PVS_OK) break;
        PVS_CUDA(cudaMemcpyAsync(out_host + i0 * out_dim, d_out, (size_t)n * out_dim * 4, cudaMemcpyDeviceToHost, s.stream));
        if (d_lab && rows > 0)
            PVS_CUDA(cudaMemcpyAsync(rows_i32_out_host + offsets_host[i0], d_lab, (size_t)rows * 4,
                                     cudaMemcpyDeviceToHost, s.stream));
        i0 = i1;
    }
    for (int w = 0; w < 2; ++w)
        if (ctx.slot[w].stream) {
            cudaError_t e = cudaStreamSynchronize(ctx.slot[w].stream);
            if (e != cudaSuccess && rc == PVS_OK)
                rc = fail(PVS_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
        }
    return rc;
}
}  // namespace

extern "C" int pvs_vlad_encode_host(const pvs_model* km, const pvs_model* pca, const float* desc_host,
                                    const int64_t* offsets_host, int64_t n_images, float power, float norm_order,
                                    float eps, float* out_host, int32_t* labels_out_host, int64_t chunk_rows)
{
    if (int s = check_chain(km, PVS_MODEL_KMEANS, pca, "pvs_vlad_encode_host")) return s;
    PVS_CHECK(norm_order > 0.f, PVS_ERR_UNSUPPORTED, "norm_order must be > 0 (or inf), got %g", (double)norm_order);
    const int d_in = pca ? pca->d_in : km->d;
    return encode_host(
        desc_host, offsets_host, n_images, d_in, (int64_t)km->k * km->d, out_host, labels_out_host, chunk_rows,
        [&](int64_t rows, int64_t n) { return pvs_vlad_workspace_bytes(km, pca, rows, n); },
        [&](const float* dd, const int64_t* doff, int64_t n, int64_t rows, float* dout, int32_t* dlab, void* ws,
            size_t wsb, cudaStream_t st) {
            return pvs_vlad_encode(km, pca, dd, doff, n, rows, power, norm_order, eps, dout, dlab, ws, wsb, st);
        });
}

extern "C" int pvs_fv_encode_host(const pvs_model* g, const pvs_model* pca, const float* desc_host,
                                  const int64_t* offsets_host, int64_t n_images, float power, float norm_order,
                                  float eps, float* out_host, int64_t chunk_rows)
{
    if (int s = check_chain(g, PVS_MODEL_GMM_DIAG, pca, "pvs_fv_encode_host")) return s;
    PVS_CHECK(norm_order > 0.f, PVS_ERR_UNSUPPORTED, "norm_order must be > 0 (or inf), got %g", (double)norm_order);
    const int d_in = pca ? pca->d_in : g->d;
    return encode_host(
        desc_host, offsets_host, n_images, d_in, (int64_t)2 * g->k * g->d + g->k, out_host, nullptr, chunk_rows,
        [&](int64_t rows, int64_t n) { return pvs_fv_workspace_bytes(g, pca, rows, n); },
        [&](const float* dd, const int64_t* doff, int64_t n, int64_t rows, float* dout, int32_t*, void* ws, size_t wsb,
            cudaStream_t st) {
            return pvs_fv_encode(g, pca, dd, doff, n, rows, power, norm_order, eps, dout, nullptr, ws, wsb, st);
        });
}

extern "C" int pvs_vlad_encode_host_u8(const pvs_model* km, const pvs_model* pca, const uint8_t* desc_host,
                                       const int64_t* offsets_host, int64_t n_images, float power, float norm_order,
                                       float eps, float* out_host, int32_t* labels_out_host, int64_t chunk_rows)
{
    if (int s = check_chain(km, PVS_MODEL_KMEANS, pca, "pvs_vlad_encode_host_u8")) return s;
    PVS_CHECK(norm_order > 0.f, PVS_ERR_UNSUPPORTED, "norm_order must be > 0 (or inf), got %g", (double)norm_order);
    const int d_in = pca ? pca->d_in : km->d;
    return encode_host(
        nullptr, offsets_host, n_images, d_in, (int64_t)km->k * km->d, out_host, labels_out_host, chunk_rows,
        [&](int64_t rows, int64_t n) { return pvs_vlad_workspace_bytes(km, pca, rows, n); },
        [&](const float* dd, const int64_t* doff, int64_t n, int64_t rows, float* dout, int32_t* dlab, void* ws, size_t wsb,
            cudaStream_t st) {
            return pvs_vlad_encode(km, pca, dd, doff, n, rows, power, norm_order, eps, dout, dlab, ws, wsb, st);
        },
        desc_host);
}

extern "C" int pvs_fv_encode_host_u8(const pvs_model* g, const pvs_model* pca, const uint8_t* desc_host,
                                     const int64_t* offsets_host, int64_t n_images, float power, float norm_order,
                                     float eps, float* out_host, int64_t chunk_rows)
{
    if (int s = check_chain(g, PVS_MODEL_GMM_DIAG, pca, "pvs_fv_encode_host_u8")) return s;
    PVS_CHECK(norm_order > 0.f, PVS_ERR_UNSUPPORTED, "norm_order must be > 0 (or inf), got %g", (double)norm_order);
    const int d_in = pca ? pca->d_in : g->d;
    return encode_host(
        nullptr, offsets_host, n_images, d_in, (int64_t)2 * g->k * g->d + g->k, out_host, nullptr, chunk_rows,
        [&](int64_t rows, int64_t n) { return pvs_fv_workspace_bytes(g, pca, rows, n); },
        [&](const float* dd, const int64_t* doff, int64_t n, int64_t rows, float* dout, int32_t*, void* ws, size_t wsb,
            cudaStream_t st) {
            return pvs_fv_encode(g, pca, dd, doff, n, rows, power, norm_order, eps, dout, nullptr, ws, wsb, st);
        },
        desc_host);
}

extern "C" int pvs_cosine_matrix_host(const float* x, int64_t n, const float* y, int64_t m, int64_t d, float* s_host)
{
    PVS_CHECK(n >= 0 && m >= 0, PVS_ERR_BAD_ARG, "negative size");
    PVS_CHECK(d >= 2, PVS_ERR_BAD_SHAPE, "Cosine similarity requires at least 2 features. Got %lld", (long long)d);
    if (n == 0 || m == 0) return PVS_OK;
    PVS_CHECK(x && y && s_host, PVS_ERR_BAD_ARG, "NULL host buffer");
    if (int s = require_device()) return s;
    HostCtx& ctx = host_ctx();
    std::lock_guard<std::mutex> lock(ctx.mu);
    Slot& sl = ctx.slot[0];
    const size_t bx = align_up((size_t)n * d * 4, 256), by = align_up((size_t)m * d * 4, 256);
    const size_t bs = align_up((size_t)n * m * 4, 256), bw = pvs_cosine_matrix_workspace_bytes(n, m, d);
    if (int rc = slot_reserve(sl, bx + by + bs + bw + 256, 1)) return rc;
    char* p = (char*)sl.buf;
    float* dx = (float*)p; float* dy = (float*)(p + bx); float* ds = (float*)(p + bx + by); void* dw = p + bx + by + bs;
    PVS_CUDA(cudaMemcpyAsync(dx, x, (size_t)n * d * 4, cudaMemcpyHostToDevice, sl.stream));
    PVS_CUDA(cudaMemcpyAsync(dy, y, (size_t)m * d * 4, cudaMemcpyHostToDevice, sl.stream));
    if (int rc = pvs_cosine_matrix(dx, n, dy, m, d, ds, dw, bw, sl.stream)) return rc;
    PVS_CUDA(cudaMemcpyAsync(s_host, ds, (size_t)n * m * 4, cudaMemcpyDeviceToHost, sl.stream));
    PVS_CUDA(cudaStreamSynchronize(sl.stream));
    return PVS_OK;
}

extern "C" int pvs_cosine_topk_host(const float* q, int64_t n_q, const float* db, int64_t n_db, int64_t d, int k,
                                    int use_bf16, float* scores_out, int64_t* idx_out)
{
    PVS_CHECK(n_q >= 0 && n_db > 0 && d >= 2, PVS_ERR_BAD_SHAPE, "pvs_cosine_topk_host: bad shape");
    PVS_CHECK(k >= 1 && (k <= PVS_TOPK_MAX || (k <= n_db && !use_bf16)), PVS_ERR_BAD_ARG,
              "k must be in [1, %d] (fp32: up to n_db) (got %d)", PVS_TOPK_MAX, k);
    if (n_q == 0) return PVS_OK;
    PVS_CHECK(q && db && scores_out && idx_out, PVS_ERR_BAD_ARG, "NULL host buffer");
    if (int s = require_device()) return s;
    HostCtx& ctx = host_ctx();
    std::lock_guard<std::mutex> lock(ctx.mu);
    Slot& sl = ctx.slot[0];
    // fp32 request: rows are normalised straight into the split operand planes when the tensor path will take them
    const int dt = use_bf16 ? PVS_BF16 : (sim3_use_tc(PVS_F16X2, n_q, n_db, d, k) && sim3_use_tc(PVS_F32, n_q, n_db, d, k)) ? PVS_F16X2 : PVS_F32;
    const size_t esz = use_bf16 ? 2 : 4;
    const size_t bq = align_up((size_t)n_q * d * 4, 256), bdb = align_up((size_t)n_db * d * 4, 256);
    const size_t bqn = align_up((size_t)n_q * d * esz, 256), bdbn = align_up((size_t)n_db * d * esz, 256);
    const size_t bsc = align_up((size_t)n_q * k * 4, 256), bid = align_up((size_t)n_q * k * 8, 256);
    const size_t bw = pvs_cosine_topk_workspace_bytes(n_q, n_db, d, k, dt);
    if (int rc = slot_reserve(sl, bq + bdb + bqn + bdbn + bsc + bid + bw + 256, 1)) return rc;
    char* p = (char*)sl.buf;
    float* dq = (float*)p;            p += bq;
    float* ddb = (float*)p;           p += bdb;
    void* dqn = p;                    p += bqn;
    void* ddbn = p;                   p += bdbn;
    float* dsc = (float*)p;           p += bsc;
    int64_t* did = (int64_t*)p;       p += bid;
    void* dw = p;
    PVS_CUDA(cudaMemcpyAsync(dq, q, (size_t)n_q * d * 4, cudaMemcpyHostToDevice, sl.stream));
    PVS_CUDA(cudaMemcpyAsync(ddb, db, (size_t)n_db * d * 4, cudaMemcpyHostToDevice, sl.stream));
    if (int rc = launch_l2_normalize(dq, n_q, d, dqn, dt, sl.stream)) return rc;
    if (int rc = launch_l2_normalize(ddb, n_db, d, ddbn, dt, sl.stream)) return rc;
    if (int rc = pvs_cosine_topk(dqn, ddbn, dt, n_q, n_db, d, k, 0, dsc, did, dw, bw, sl.stream)) return rc;
    PVS_CUDA(cudaMemcpyAsync(scores_out, dsc, (size_t)n_q * k * 4, cudaMemcpyDeviceToHost, sl.stream));
    PVS_CUDA(cudaMemcpyAsync(idx_out, did, (size_t)n_q * k * 8, cudaMemcpyDeviceToHost, sl.stream));
    PVS_CUDA(cudaStreamSynchronize(sl.stream));
    return PVS_OK;
}
