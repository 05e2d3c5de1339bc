// Collectives of the data-parallel training step: a thin layer over NCCL, resolved with dlopen.
//
// The reference trains on one device (no distributed backend); the batch-sharded step of this
// path sums the flat gradient buffer over the ranks of one NVSwitch box.  libuocr does not link libnccl: the
// image carries two builds (system 2.27.3, torch-bundled 2.28.9, both SONAME libnccl.so.2) and a process may
// already have mapped one of them -- the entry points below bind whichever is there (RTLD_NOLOAD first), so
// two NCCL copies never meet in one process.  Only the long-stable core ABI is used (ncclGetUniqueId,
// ncclCommInitRank, ncclAllReduce, ncclBroadcast, ncclCommDestroy), declared here instead of including nccl.h.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace {

typedef struct { char internal[UOCR_NCCL_UNIQUE_ID_BYTES]; } nccl_unique_id;
typedef void* nccl_comm_t;
enum { kNcclSuccess = 0 };
enum { kNcclSum = 0, kNcclMax = 2, kNcclMin = 3 };          // ncclRedOp_t
enum { kNcclFloat32 = 7, kNcclFloat64 = 8 };                // ncclDataType_t

struct Nccl {
    void* handle = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    const char* (*GetLastError)(nccl_comm_t) = nullptr;
};

std::mutex g_mutex;
Nccl g_nccl;
nccl_comm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;

template <typename F>
bool bind(F& fn, const char* name) {
    fn = reinterpret_cast<F>(dlsym(g_nccl.handle, name));
    return fn != nullptr;
}

int load_locked(const char* path) {
    if (g_nccl.handle) return UOCR_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // a copy the process already uses
    if (!h && path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        uocr::set_error("NCCL not found: %s", dlerror());
        return UOCR_ERR_COMM;
    }
    g_nccl.handle = h;
    bool ok = bind(g_nccl.GetVersion, "ncclGetVersion") && bind(g_nccl.GetUniqueId, "ncclGetUniqueId") &&
              bind(g_nccl.CommInitRank, "ncclCommInitRank") && bind(g_nccl.CommDestroy, "ncclCommDestroy") &&
              bind(g_nccl.AllReduce, "ncclAllReduce") && bind(g_nccl.Broadcast, "ncclBroadcast") &&
              bind(g_nccl.GetErrorString, "ncclGetErrorString");
    bind(g_nccl.GetLastError, "ncclGetLastError");                    // optional (NCCL >= 2.13)
    if (!ok) {
        uocr::set_error("NCCL library misses a core symbol: %s", dlerror());
        g_nccl = Nccl();
        return UOCR_ERR_COMM;
    }
    return UOCR_OK;
}

int nccl_fail(const char* what, int rc) {
    const char* detail = (g_nccl.GetLastError && g_comm) ? g_nccl.GetLastError(g_comm) : "";
    uocr::set_error("%s: %s %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error",
                    detail ? detail : "");
    return UOCR_ERR_COMM;
}

#define UOCR_NCCL(expr)                              \
    do {                                             \
        int rc__ = (expr);                           \
        if (rc__ != kNcclSuccess) return nccl_fail(#expr, rc__); \
    } while (0)

int require_comm() {
    if (!g_comm) {
        uocr::set_error("no communicator: call uocr_nccl_init first");
        return UOCR_ERR_COMM;
    }
    return UOCR_OK;
}

}  // namespace

extern "C" {

int uocr_nccl_load(const char* path) {
    std::lock_guard<std::mutex> lock(g_mutex);
    return load_locked(path);
}

int uocr_nccl_version(int* version) {
    UOCR_REQUIRE(version, "version is NULL");
    std::lock_guard<std::mutex> lock(g_mutex);
    int rc = load_locked(nullptr);
    if (rc != UOCR_OK) return rc;
    UOCR_NCCL(g_nccl.GetVersion(version));
    return UOCR_OK;
}

int uocr_nccl_unique_id(void* id_out) {
    UOCR_REQUIRE(id_out, "id_out is NULL");
    std::lock_guard<std::mutex> lock(g_mutex);
    int rc = load_locked(nullptr);
    if (rc != UOCR_OK) return rc;
    nccl_unique_id id;
    UOCR_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, sizeof(id.internal));
    return UOCR_OK;
}

int uocr_nccl_init(int rank, int world, const void* unique_id) {
    UOCR_REQUIRE(unique_id, "unique_id is NULL");
    UOCR_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
    std::lock_guard<std::mutex> lock(g_mutex);
    UOCR_REQUIRE(!g_comm, "communicator already initialised (one per process)");
    int rc = load_locked(nullptr);
    if (rc != UOCR_OK) return rc;
    nccl_unique_id id;
    memcpy(id.internal, unique_id, sizeof(id.internal));
    nccl_comm_t comm = nullptr;
    UOCR_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    g_comm = comm;
    g_rank = rank;
    g_world = world;
    return UOCR_OK;
}

int uocr_nccl_rank(int* rank, int* world) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (rank) *rank = g_rank;
    if (world) *world = g_comm ? g_world : 1;
    return UOCR_OK;
}

int uocr_nccl_finalize(void) {
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_comm) return UOCR_OK;
    nccl_comm_t comm = g_comm;
    g_comm = nullptr;
    g_rank = 0;
    g_world = 1;
    UOCR_NCCL(g_nccl.CommDestroy(comm));
    return UOCR_OK;
}

int uocr_allreduce_sum_f32(float* data, int64_t count, void* stream) {
    if (count == 0) return UOCR_OK;
    UOCR_REQUIRE(data && count > 0, "bad buffer");
    int rc = require_comm();
    if (rc != UOCR_OK) return rc;
    UOCR_NCCL(g_nccl.AllReduce(data, data, (size_t)count, kNcclFloat32, kNcclSum, g_comm, uocr::as_stream(stream)));
    return UOCR_OK;
}

int uocr_allreduce_f64(double* data, int64_t count, int op, void* stream) {
    if (count == 0) return UOCR_OK;
    UOCR_REQUIRE(data && count > 0, "bad buffer");
    UOCR_REQUIRE(op >= 0 && op <= 2, "op must be 0 (sum), 1 (max) or 2 (min)");
    int rc = require_comm();
    if (rc != UOCR_OK) return rc;
    const int red = op == 0 ? kNcclSum : (op == 1 ? kNcclMax : kNcclMin);
    UOCR_NCCL(g_nccl.AllReduce(data, data, (size_t)count, kNcclFloat64, red, g_comm, uocr::as_stream(stream)));
    return UOCR_OK;
}

int uocr_broadcast_f32(float* data, int64_t count, int root, void* stream) {
    if (count == 0) return UOCR_OK;
    UOCR_REQUIRE(data && count > 0, "bad buffer");
    int rc = require_comm();
    if (rc != UOCR_OK) return rc;
    UOCR_REQUIRE(root >= 0 && root < g_world, "root %d outside world %d", root, g_world);
    UOCR_NCCL(g_nccl.Broadcast(data, data, (size_t)count, kNcclFloat32, root, g_comm, uocr::as_stream(stream)));
    return UOCR_OK;
}

}  // extern "C"
