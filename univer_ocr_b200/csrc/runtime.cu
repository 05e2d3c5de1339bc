// libuocr runtime plumbing: errors, device info, stream-ordered pooled memory, streams,
// events, CUDA-graph capture.  Replaces what the reference gets from CuPy / numba.cuda
// (nn/gpu.py:5-29, cuda.synchronize() in nn/layers/convolutional.py:192).
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>
#include <unordered_map>

#include "common.cuh"

namespace uocr {

static thread_local char tl_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------- caching device allocator
// The layer stack allocates the same few dozen tensor sizes every step.  cudaMallocAsync's pool
// showed multi-millisecond slow paths in steady state (fresh physical mappings), so libuocr keeps
// its own size-keyed free lists per (device, stream): uocr_free() parks the block, uocr_malloc()
// returns a parked block of the same rounded size (or up to 1/8 larger) without touching the
// driver.  Reuse is safe because a parked block is only handed out to work queued LATER on the
// SAME stream it was freed on (stream order = reuse order).  Misses fall through to cudaMalloc.
struct PoolKey {
    int dev;
    cudaStream_t st;
    bool operator<(const PoolKey& o) const { return dev != o.dev ? dev < o.dev : st < o.st; }
};
static std::mutex g_pool_mutex;
static std::map<PoolKey, std::multimap<size_t, void*>> g_free;      // parked blocks by size
static std::unordered_map<void*, std::pair<size_t, int>> g_live;    // ptr -> (rounded size, device)
static size_t g_reserved = 0;
// While a launch sequence is being captured into a CUDA graph, blocks that are freed must not be handed out again:
// the graph's kernels carry their addresses and will use them on every replay.  They are parked in g_held and
// belong to the captured graph until it is destroyed.
struct HeldBlock { void* ptr; size_t size; int dev; cudaStream_t st; };
static bool g_hold = false;
static std::vector<HeldBlock> g_held;

static size_t round_size(size_t bytes) {
    if (bytes < 512) return 512;
    if (bytes < (1u << 20)) return (bytes + 511) & ~(size_t)511;
    return (bytes + ((1u << 16) - 1)) & ~(size_t)((1u << 16) - 1);       // 64 KiB granules above 1 MiB
}

int pool_alloc(void** ptr, size_t bytes, cudaStream_t st) {
    *ptr = nullptr;
    if (bytes == 0) return UOCR_OK;
    int dev;
    UOCR_CUDA(cudaGetDevice(&dev));
    const size_t want = round_size(bytes);
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        auto& lst = g_free[PoolKey{dev, st}];
        auto it = lst.lower_bound(want);
        if (it != lst.end() && it->first <= want + want / 8) {
            *ptr = it->second;
            g_live[*ptr] = {it->first, dev};
            lst.erase(it);
            return UOCR_OK;
        }
    }
    cudaError_t e = cudaMalloc(ptr, want);
    if (e != cudaSuccess) {
        // out of memory: drop every parked block and retry once
        cudaGetLastError();
        pool_trim();
        e = cudaMalloc(ptr, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            *ptr = nullptr;
            return UOCR_ERR_CUDA;
        }
    }
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_live[*ptr] = {want, dev};
    g_reserved += want;
    return UOCR_OK;
}

int pool_free(void* ptr, cudaStream_t st) {
    if (!ptr) return UOCR_OK;
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    auto it = g_live.find(ptr);
    if (it == g_live.end()) {
        set_error("uocr_free: pointer %p was not allocated by uocr_malloc", ptr);
        return UOCR_ERR_INVALID;
    }
    if (g_hold) g_held.push_back(HeldBlock{ptr, it->second.first, it->second.second, st});
    else g_free[PoolKey{it->second.second, st}].emplace(it->second.first, ptr);
    g_live.erase(it);
    return UOCR_OK;
}

struct CapturedGraph {
    cudaGraphExec_t exec = nullptr;
    std::vector<HeldBlock> held;
    uint64_t launches = 0;
};
static uint64_t g_capture_launch0 = 0;

int pool_trim() {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    cudaDeviceSynchronize();
    for (auto& kv : g_free) {
        for (auto& blk : kv.second) {
            cudaFree(blk.second);
            g_reserved -= blk.first;
        }
        kv.second.clear();
    }
    return UOCR_OK;
}

}  // namespace uocr

using namespace uocr;

extern "C" {

int uocr_version(void) { return UOCR_VERSION; }
const char* uocr_last_error(void) { return tl_error; }

int uocr_device_count(int* count) {
    UOCR_REQUIRE(count, "count is NULL");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return UOCR_ERR_CUDA;
    }
    return UOCR_OK;
}

int uocr_set_device(int device) {
    UOCR_CUDA(cudaSetDevice(device));
    return UOCR_OK;
}

int uocr_get_device(int* device) {
    UOCR_REQUIRE(device, "device is NULL");
    UOCR_CUDA(cudaGetDevice(device));
    return UOCR_OK;
}

int uocr_device_info(int device, char* name, size_t name_len, int* sm_count, int* cc_major,
                     int* cc_minor, size_t* total_mem, size_t* free_mem) {
    cudaDeviceProp prop;
    UOCR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name && name_len) {
        strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem || free_mem) {
        int cur;
        UOCR_CUDA(cudaGetDevice(&cur));
        if (cur != device) UOCR_CUDA(cudaSetDevice(device));
        size_t f = 0, t = 0;
        UOCR_CUDA(cudaMemGetInfo(&f, &t));
        if (cur != device) UOCR_CUDA(cudaSetDevice(cur));
        if (total_mem) *total_mem = t;
        if (free_mem) *free_mem = f;
    }
    return UOCR_OK;
}

int uocr_malloc(void** ptr, size_t bytes, void* stream) {
    UOCR_REQUIRE(ptr, "ptr is NULL");
    return pool_alloc(ptr, bytes, as_stream(stream));
}

int uocr_free(void* ptr, void* stream) { return pool_free(ptr, as_stream(stream)); }

int uocr_mempool_trim(void) { return pool_trim(); }

int uocr_mempool_reserved(size_t* bytes) {
    UOCR_REQUIRE(bytes, "bytes is NULL");
    *bytes = g_reserved;
    return UOCR_OK;
}

int uocr_host_alloc(void** ptr, size_t bytes) {
    UOCR_REQUIRE(ptr, "ptr is NULL");
    UOCR_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
    return UOCR_OK;
}

int uocr_host_free(void* ptr) {
    if (!ptr) return UOCR_OK;
    UOCR_CUDA(cudaFreeHost(ptr));
    return UOCR_OK;
}

int uocr_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return UOCR_OK;
    UOCR_REQUIRE(dst && src, "NULL pointer");
    UOCR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return UOCR_OK;
}

int uocr_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return UOCR_OK;
    UOCR_REQUIRE(dst && src, "NULL pointer");
    UOCR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return UOCR_OK;
}

int uocr_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
    if (bytes == 0) return UOCR_OK;
    UOCR_REQUIRE(dst && src, "NULL pointer");
    UOCR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return UOCR_OK;
}

int uocr_memset(void* dst, int byte_value, size_t bytes, void* stream) {
    if (bytes == 0) return UOCR_OK;
    UOCR_REQUIRE(dst, "NULL pointer");
    UOCR_CUDA(cudaMemsetAsync(dst, byte_value, bytes, as_stream(stream)));
    return UOCR_OK;
}

int uocr_stream_create(void** stream) {
    UOCR_REQUIRE(stream, "stream is NULL");
    cudaStream_t s;
    UOCR_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = s;
    return UOCR_OK;
}

int uocr_stream_destroy(void* stream) {
    if (!stream) return UOCR_OK;
    UOCR_CUDA(cudaStreamDestroy(as_stream(stream)));
    return UOCR_OK;
}

int uocr_stream_sync(void* stream) {
    UOCR_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return UOCR_OK;
}

int uocr_device_sync(void) {
    UOCR_CUDA(cudaDeviceSynchronize());
    return UOCR_OK;
}

int uocr_event_create(void** event) {
    UOCR_REQUIRE(event, "event is NULL");
    cudaEvent_t e;
    UOCR_CUDA(cudaEventCreate(&e));
    *event = e;
    return UOCR_OK;
}

int uocr_event_destroy(void* event) {
    if (!event) return UOCR_OK;
    UOCR_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)));
    return UOCR_OK;
}

int uocr_event_record(void* event, void* stream) {
    UOCR_REQUIRE(event, "event is NULL");
    UOCR_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), as_stream(stream)));
    return UOCR_OK;
}

int uocr_event_sync(void* event) {
    UOCR_REQUIRE(event, "event is NULL");
    UOCR_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)));
    return UOCR_OK;
}

int uocr_event_elapsed_ms(void* start, void* stop, float* ms) {
    UOCR_REQUIRE(start && stop && ms, "NULL argument");
    UOCR_CUDA(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start),
                                   reinterpret_cast<cudaEvent_t>(stop)));
    return UOCR_OK;
}

int uocr_stream_wait_event(void* stream, void* event) {
    UOCR_REQUIRE(event, "event is NULL");
    UOCR_CUDA(cudaStreamWaitEvent(as_stream(stream), reinterpret_cast<cudaEvent_t>(event), 0));
    return UOCR_OK;
}

int uocr_graph_begin(void* stream) {
    UOCR_REQUIRE(stream, "graph capture needs a non-default stream");
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        UOCR_REQUIRE(!g_hold, "a graph capture is already in progress");
        g_hold = true;
        g_held.clear();
    }
    g_capture_launch0 = g_launches.load(std::memory_order_relaxed);
    // relaxed: a pool miss during capture may call cudaMalloc (no work is enqueued by it)
    cudaError_t e = cudaStreamBeginCapture(as_stream(stream), cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        g_hold = false;
        set_error("cudaStreamBeginCapture: %s", cudaGetErrorString(e));
        return UOCR_ERR_CUDA;
    }
    return UOCR_OK;
}

static void release_held(std::vector<HeldBlock>& held) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    for (const HeldBlock& b : held) g_free[PoolKey{b.dev, b.st}].emplace(b.size, b.ptr);
    held.clear();
}

int uocr_graph_end(void* stream, void** graph_exec) {
    UOCR_REQUIRE(stream && graph_exec, "NULL argument");
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(as_stream(stream), &graph);
    CapturedGraph* cg = new CapturedGraph();
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        cg->held.swap(g_held);
        g_hold = false;
    }
    cg->launches = g_launches.load(std::memory_order_relaxed) - g_capture_launch0;
    if (e == cudaSuccess) {
        e = cudaGraphInstantiate(&cg->exec, graph, 0);
        cudaGraphDestroy(graph);
    }
    if (e != cudaSuccess) {
        set_error("graph capture / instantiate: %s", cudaGetErrorString(e));
        cudaGetLastError();
        release_held(cg->held);
        delete cg;
        return UOCR_ERR_CUDA;
    }
    *graph_exec = cg;
    return UOCR_OK;
}

int uocr_graph_launch(void* graph_exec, void* stream) {
    UOCR_REQUIRE(graph_exec, "graph_exec is NULL");
    CapturedGraph* cg = static_cast<CapturedGraph*>(graph_exec);
    UOCR_CUDA(cudaGraphLaunch(cg->exec, as_stream(stream)));
    g_launches.fetch_add(cg->launches, std::memory_order_relaxed);     // the kernels the replay runs
    return UOCR_OK;
}

int uocr_graph_destroy(void* graph_exec) {
    if (!graph_exec) return UOCR_OK;
    CapturedGraph* cg = static_cast<CapturedGraph*>(graph_exec);
    cudaError_t e = cudaGraphExecDestroy(cg->exec);
    release_held(cg->held);                 // the caller has synchronised: nothing replays any more
    delete cg;
    if (e != cudaSuccess) { set_error("cudaGraphExecDestroy: %s", cudaGetErrorString(e)); return UOCR_ERR_CUDA; }
    return UOCR_OK;
}

int uocr_launch_count(uint64_t* count) {
    UOCR_REQUIRE(count, "count is NULL");
    *count = g_launches.load(std::memory_order_relaxed);
    return UOCR_OK;
}

}  // extern "C"
