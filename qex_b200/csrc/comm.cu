// The one exchange step of the path: all-reduce(sum) of the packed outputs
// [vmat | excsum | nelec] and [dm_bar | theta_bar] across the grid shards, plus the broadcast that
// replicates (dm, theta, cotangents) from the rank that holds the host buffers.  NCCL over
// NVLink/NVSwitch; the reference has no multi-device code at all.
//
// libnccl is bound at run time (dlopen of the copy already loaded into the process by the host
// framework, else libnccl.so.2 from the loader path), so libqexxc.so itself has no link-time NCCL
// dependency and still loads on a box without it; every entry point fails with QEXXC_ERR_STATE then.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

struct qexxc_comm {
    void* nccl = nullptr;  // ncclComm_t
    int world = 1, rank = 0, device = 0;
    bool owned = false;
    long calls = 0;
};

namespace qexxc {
namespace {

struct NcclId {
    char internal[128];
};
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(void**, int, NcclId, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_bcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_version)(int*);

constexpr int kNcclFloat64 = 8;  // ncclDataType_t::ncclFloat64
constexpr int kNcclSum = 0;      // ncclRedOp_t::ncclSum

struct NcclApi {
    void* handle = nullptr;
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_bcast bcast = nullptr;
    fn_errstr errstr = nullptr;
    fn_version version = nullptr;
    bool ok = false;
};

NcclApi g_api;
std::once_flag g_once;

void load_nccl() {
    const char* names[] = {getenv("QEXXC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    // prefer the copy the process already has (torch bundles its own libnccl.so.2)
    for (const char* n : names)
        if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    for (const char* n : names)
        if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_api.handle = h;
    g_api.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
    g_api.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    g_api.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_api.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
    g_api.bcast = (fn_bcast)dlsym(h, "ncclBroadcast");
    g_api.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    g_api.version = (fn_version)dlsym(h, "ncclGetVersion");
    g_api.ok = g_api.get_id && g_api.init_rank && g_api.destroy && g_api.allreduce && g_api.bcast;
}

int need_nccl() {
    std::call_once(g_once, load_nccl);
    if (!g_api.ok) {
        set_error("NCCL is not available in this process (dlopen libnccl.so.2 failed): %s",
                  dlerror() ? dlerror() : "symbols missing");
        return QEXXC_ERR_STATE;
    }
    return QEXXC_OK;
}

int nccl_check(int rc, const char* what) {
    if (rc == 0) return QEXXC_OK;
    set_error("%s failed: NCCL error %d (%s)", what, rc, g_api.errstr ? g_api.errstr(rc) : "?");
    return QEXXC_ERR_CUDA;
}

}  // namespace
}  // namespace qexxc

using namespace qexxc;

extern "C" {

int qexxc_comm_nccl_version(int* version) {
    QX_ARG(version != nullptr, "null pointer");
    QX_TRY(need_nccl());
    *version = 0;
    if (g_api.version) return nccl_check(g_api.version(version), "ncclGetVersion");
    return QEXXC_OK;
}

int qexxc_comm_unique_id(unsigned char id[128]) {
    QX_ARG(id != nullptr, "null pointer");
    QX_TRY(need_nccl());
    NcclId u;
    QX_TRY(nccl_check(g_api.get_id(&u), "ncclGetUniqueId"));
    memcpy(id, u.internal, 128);
    return QEXXC_OK;
}

int qexxc_comm_create(qexxc_comm** out, int device, int world, int rank, const unsigned char id[128]) {
    QX_ARG(out != nullptr && id != nullptr, "null pointer");
    *out = nullptr;
    QX_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    QX_TRY(need_nccl());
    QX_CUDA(cudaSetDevice(device));
    NcclId u;
    memcpy(u.internal, id, 128);
    void* comm = nullptr;
    QX_TRY(nccl_check(g_api.init_rank(&comm, world, u, rank), "ncclCommInitRank"));
    qexxc_comm* c = new qexxc_comm();
    c->nccl = comm;
    c->world = world;
    c->rank = rank;
    c->device = device;
    c->owned = true;
    *out = c;
    return QEXXC_OK;
}

int qexxc_comm_wrap(qexxc_comm** out, void* nccl_comm, int device, int world, int rank) {
    QX_ARG(out != nullptr && nccl_comm != nullptr, "null pointer");
    QX_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    QX_TRY(need_nccl());
    qexxc_comm* c = new qexxc_comm();
    c->nccl = nccl_comm;
    c->world = world;
    c->rank = rank;
    c->device = device;
    c->owned = false;
    *out = c;
    return QEXXC_OK;
}

int qexxc_comm_destroy(qexxc_comm* c) {
    if (!c) return QEXXC_OK;
    int rc = QEXXC_OK;
    if (c->owned && c->nccl && g_api.ok) {
        cudaSetDevice(c->device);
        rc = nccl_check(g_api.destroy(c->nccl), "ncclCommDestroy");
    }
    delete c;
    return rc;
}

int qexxc_comm_rank(const qexxc_comm* c) { return c ? c->rank : -1; }
int qexxc_comm_world(const qexxc_comm* c) { return c ? c->world : 0; }
long qexxc_comm_calls(const qexxc_comm* c) { return c ? c->calls : 0; }

int qexxc_allreduce(qexxc_comm* c, double* buf_dev, long count, void* stream) {
    QX_ARG(c != nullptr && (buf_dev != nullptr || count == 0), "null pointer");
    QX_ARG(count >= 0, "negative count");
    if (c->world == 1 || count == 0) return QEXXC_OK;
    QX_CUDA(cudaSetDevice(c->device));
    c->calls++;
    return nccl_check(g_api.allreduce(buf_dev, buf_dev, (size_t)count, kNcclFloat64, kNcclSum, c->nccl,
                                      (cudaStream_t)stream),
                      "ncclAllReduce");
}

int qexxc_bcast(qexxc_comm* c, double* buf_dev, long count, int root, void* stream) {
    QX_ARG(c != nullptr && (buf_dev != nullptr || count == 0), "null pointer");
    QX_ARG(count >= 0 && root >= 0 && root < c->world, "bad count / root");
    if (c->world == 1 || count == 0) return QEXXC_OK;
    QX_CUDA(cudaSetDevice(c->device));
    c->calls++;
    return nccl_check(g_api.bcast(buf_dev, buf_dev, (size_t)count, kNcclFloat64, root, c->nccl, (cudaStream_t)stream),
                      "ncclBroadcast");
}

}  // extern "C"
