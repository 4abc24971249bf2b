// pvs_comm.cu -- the one collective of the path: all-gather of the per-rank top-k lists
//  over NCCL.  Row-block sharded retrieval leaves every rank with the final
// (rows / W, k) score and index lists of its own query rows; pvs_allgather_topk assembles
// the (rows, k) result on every rank with two ncclAllGather calls in one group on the
// caller's stream.  NCCL is resolved at run time (dlopen) so the library has no link-time
// dependency on it: the copy already loaded by the host process (PyTorch's) is preferred.
#include <dlfcn.h>
#include <string.h>
#include "pvs_kernels.cuh"

namespace {
typedef struct { char internal[128]; } nccl_uid;                  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_destroy)(nccl_comm);
typedef int (*fn_allgather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
typedef int (*fn_group)(void);
typedef const char* (*fn_errstr)(int);
constexpr int NCCL_FLOAT32 = 7, NCCL_INT64 = 4;                   // ncclDataType_t

struct NcclApi {
    void* handle = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allgather allgather = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_errstr errstr = nullptr;
    bool ok = false;
};
NcclApi& api()
{
    static NcclApi a;
    return a;
}
int load_nccl(const char* path)
{
    NcclApi& a = api();
    if (a.ok) return PVS_OK;
    void* h = nullptr;
    if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // already in the process (PyTorch)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return pvs::fail(PVS_ERR_CUDA, "NCCL could not be loaded: %s", dlerror());
    a.handle = h;
    a.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
    a.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    a.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    a.allgather = (fn_allgather)dlsym(h, "ncclAllGather");
    a.group_start = (fn_group)dlsym(h, "ncclGroupStart");
    a.group_end = (fn_group)dlsym(h, "ncclGroupEnd");
    a.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    if (!a.get_uid || !a.init_rank || !a.destroy || !a.allgather || !a.group_start || !a.group_end)
        return pvs::fail(PVS_ERR_CUDA, "NCCL library lacks a required symbol");
    a.ok = true;
    return PVS_OK;
}
int nccl_fail(const char* what, int rc)
{
    const NcclApi& a = api();
    return pvs::fail(PVS_ERR_NCCL, "%s failed: %s", what, a.errstr ? a.errstr(rc) : "nccl error");
}
}  // namespace

struct pvs_comm {
    nccl_comm comm = nullptr;
    int world = 0, rank = 0;
};

extern "C" int pvs_nccl_load(const char* library_path) { return load_nccl(library_path); }

extern "C" int pvs_comm_unique_id(void* id_out_128)
{
    PVS_CHECK(id_out_128, PVS_ERR_BAD_ARG, "pvs_comm_unique_id: NULL buffer");
    if (int s = load_nccl(nullptr)) return s;
    nccl_uid id;
    if (int rc = api().get_uid(&id)) return nccl_fail("ncclGetUniqueId", rc);
    memcpy(id_out_128, &id, sizeof(id));
    return PVS_OK;
}

extern "C" int pvs_comm_create(const void* id_128, int world, int rank, pvs_comm** out)
{
    PVS_CHECK(id_128 && out && world >= 1 && rank >= 0 && rank < world, PVS_ERR_BAD_ARG, "pvs_comm_create: bad arguments");
    if (int s = load_nccl(nullptr)) return s;
    nccl_uid id;
    memcpy(&id, id_128, sizeof(id));
    pvs_comm* c = new pvs_comm();
    c->world = world;
    c->rank = rank;
    if (int rc = api().init_rank(&c->comm, world, id, rank)) { delete c; return nccl_fail("ncclCommInitRank", rc); }
    *out = c;
    return PVS_OK;
}

extern "C" int pvs_comm_destroy(pvs_comm* c)
{
    if (!c) return PVS_OK;
    if (c->comm && api().ok) api().destroy(c->comm);
    delete c;
    return PVS_OK;
}

extern "C" int pvs_allgather_topk(pvs_comm* c, const float* scores_dev, const int64_t* idx_dev, int64_t rows_per_rank,
                                  int k, float* scores_all_dev, int64_t* idx_all_dev, void* stream)
{
    PVS_CHECK(c && c->comm, PVS_ERR_BAD_ARG, "pvs_allgather_topk: NULL communicator");
    PVS_CHECK(rows_per_rank >= 0 && k >= 1, PVS_ERR_BAD_ARG, "pvs_allgather_topk: bad size");
    if (rows_per_rank == 0) return PVS_OK;
    PVS_CHECK(scores_dev && idx_dev && scores_all_dev && idx_all_dev, PVS_ERR_BAD_ARG, "pvs_allgather_topk: NULL buffer");
    const size_t count = (size_t)rows_per_rank * k;
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = api().group_start()) return nccl_fail("ncclGroupStart", rc);
    int r1 = api().allgather(scores_dev, scores_all_dev, count, NCCL_FLOAT32, c->comm, st);
    int r2 = api().allgather(idx_dev, idx_all_dev, count, NCCL_INT64, c->comm, st);
    int r3 = api().group_end();
    if (r1) return nccl_fail("ncclAllGather(scores)", r1);
    if (r2) return nccl_fail("ncclAllGather(indices)", r2);
    if (r3) return nccl_fail("ncclGroupEnd", r3);
    return PVS_OK;
}
