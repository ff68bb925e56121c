// smj_runtime.cu -- device selection, stream, memory and timing entry points of smj_device.h.
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <stdlib.h>
#include <string.h>

static SmjCtx              g_ctx = {-1, 0, nullptr, nullptr, 0, false, false, 0, 0};

// ---- optional per-kernel timing (bench.py's roofline numbers): CUDA events on the launch stream
struct ProfRec {
    const char *name;
    cudaEvent_t start, stop;
    u64         items;
};
static std::vector<ProfRec> g_prof;

void smj_prof_before(SmjCtx *ctx, const char *kernel_name) {
    ProfRec r;
    r.name  = kernel_name;
    r.items = ctx->next_items;
    cudaEventCreate(&r.start);
    cudaEventCreate(&r.stop);
    cudaEventRecord(r.start, ctx->stream);
    g_prof.push_back(r);
}
void smj_prof_after(SmjCtx *ctx) {
    if (!g_prof.empty()) cudaEventRecord(g_prof.back().stop, ctx->stream);
}
static std::mutex          g_mu;
static thread_local char   g_err[512] = "";

int smj_fail(cudaError_t e, const char *what, const char *file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d in `%s`", (int)e,
             cudaGetErrorString(e), file, line, what);
    if (getenv("SMJ_VERBOSE")) fprintf(stderr, "[smj_b200] %s\n", g_err);
    return -(int)e - 1000;
}
int smj_fail_msg(const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    if (getenv("SMJ_VERBOSE")) fprintf(stderr, "[smj_b200] %s\n", g_err);
    return -1;
}

static int ctx_init_locked(int device) {
    if (g_ctx.ready && (device < 0 || device == g_ctx.device)) return 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return smj_fail_msg("no CUDA device available: the sort-merge join kernels need a GPU "
                            "(there is no CPU fallback)");
    }
    if (device < 0) {
        const char *env = getenv("SMJ_DEVICE");
        if (!env) env = getenv("LOCAL_RANK");
        device = env ? atoi(env) : 0;
        if (device < 0 || device >= ndev) device = device % ndev;
    }
    if (device >= ndev) return smj_fail_msg("smjdev_init: device index out of range");
    SMJ_CHECK(cudaSetDevice(device));
    if (g_ctx.ready && g_ctx.own_stream) cudaStreamDestroy(g_ctx.own_stream);
    if (g_ctx.ready && g_ctx.copy_stream) cudaStreamDestroy(g_ctx.copy_stream);
    cudaDeviceProp prop;
    SMJ_CHECK(cudaGetDeviceProperties(&prop, device));
    g_ctx.device   = device;
    g_ctx.generation++;
    g_ctx.sm_count = prop.multiProcessorCount;
    SMJ_CHECK(cudaStreamCreateWithFlags(&g_ctx.own_stream, cudaStreamNonBlocking));
    SMJ_CHECK(cudaStreamCreateWithFlags(&g_ctx.copy_stream, cudaStreamNonBlocking));
    g_ctx.stream = g_ctx.own_stream;
    // keep freed scratch in the pool: cudaMallocAsync then costs microseconds, not a driver call
    cudaMemPool_t pool;
    SMJ_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long thresh = ~0ull;
    SMJ_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    g_ctx.ready = true;
    return 0;
}

bool smj_is_host_pointer(const void *p) {
    cudaPointerAttributes attr;
    if (!p || cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

SmjCtx *smj_ctx() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx.ready && ctx_init_locked(-1) != 0) return nullptr;
    // the caller may have switched devices (torch does); kernels must go to ours
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != g_ctx.device) cudaSetDevice(g_ctx.device);
    return &g_ctx;
}

extern "C" {

int smjdev_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int smjdev_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    return ctx_init_locked(device);
}
int smjdev_current_device(void) {
    SmjCtx *c = smj_ctx();
    return c ? c->device : -1;
}
int smjdev_set_stream(void *s) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    c->stream = (cudaStream_t)s;  // NULL is CUDA's legacy default stream (what torch uses by default)
    return 0;
}
int smjdev_use_own_stream(void) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    c->stream = c->own_stream;
    return 0;
}
void *smjdev_get_stream(void) {
    SmjCtx *c = smj_ctx();
    return c ? (void *)c->stream : nullptr;
}
int smjdev_sync(void) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    SMJ_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}
const char *smjdev_last_error(void) { return g_err; }
uint64_t    smjdev_launch_count(void) { return g_ctx.launches; }
int         smjdev_sm_count(void) {
    SmjCtx *c = smj_ctx();
    return c ? c->sm_count : 0;
}

void *smjdev_malloc(size_t bytes) {
    SmjCtx *c = smj_ctx();
    if (!c) return nullptr;
    void *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(&p, bytes, c->stream);
    if (e != cudaSuccess) {
        smj_fail(e, "cudaMallocAsync", __FILE__, __LINE__);
        return nullptr;
    }
    return p;
}
int smjdev_free(void *p) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (p) SMJ_CHECK(cudaFreeAsync(p, c->stream));
    return 0;
}
void *smjdev_host_alloc(size_t bytes) {
    if (!smj_ctx()) return nullptr;
    void *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) {
        smj_fail(e, "cudaMallocHost", __FILE__, __LINE__);
        return nullptr;
    }
    return p;
}
int smjdev_host_free(void *p) {
    if (p) SMJ_CHECK(cudaFreeHost(p));
    return 0;
}
int smjdev_memcpy_h2d(void *dst, const void *src, size_t bytes) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (bytes) SMJ_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}
int smjdev_memcpy_d2h(void *dst, const void *src, size_t bytes) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (bytes) SMJ_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}
int smjdev_memcpy_d2d(void *dst, const void *src, size_t bytes) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (bytes) SMJ_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
}
int smjdev_memset(void *dst, int value, size_t bytes) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (bytes) SMJ_CHECK(cudaMemsetAsync(dst, value, bytes, c->stream));
    return 0;
}
// Bulk exchange with the copy engines: copies enqueued between smjdev_exchange_begin() and
// smjdev_exchange_end() run on the library's copy stream behind everything enqueued on the compute
// stream so far, and the compute stream continues only when they are done.  dst / src may be peer
// GPUs' buffers (smjdev_ipc_open): a contiguous copy crosses NVLink at ~735 GB/s on B200, about
// twice what scattered stores from the SMs reach.
static cudaEvent_t g_xev[2] = {nullptr, nullptr};
static int exchange_events(void) {
    for (auto &e : g_xev)
        if (!e) SMJ_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return 0;
}
int smjdev_exchange_begin(void) {
    SmjCtx *c = smj_ctx();
    if (!c || exchange_events()) return -1;
    SMJ_CHECK(cudaEventRecord(g_xev[0], c->stream));
    SMJ_CHECK(cudaStreamWaitEvent(c->copy_stream, g_xev[0], 0));
    return 0;
}
int smjdev_exchange_copy(void *dst, const void *src, size_t bytes) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (bytes) SMJ_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->copy_stream));
    return 0;
}
int smjdev_exchange_end(void) {
    SmjCtx *c = smj_ctx();
    if (!c || exchange_events()) return -1;
    SMJ_CHECK(cudaEventRecord(g_xev[1], c->copy_stream));
    SMJ_CHECK(cudaStreamWaitEvent(c->stream, g_xev[1], 0));
    return 0;
}

int smjdev_pointer_kind(const void *p) {
    if (!smj_ctx()) return -1;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    switch (attr.type) {
        case cudaMemoryTypeHost: return 1;
        case cudaMemoryTypeDevice: return 2;
        case cudaMemoryTypeManaged: return 3;
        default: return 0;
    }
}

int smjdev_profile_begin(void) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    for (auto &r : g_prof) {
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    g_prof.clear();
    c->profiling = true;
    return 0;
}
// writes one line per kernel: name;launches;total_ms;total_items
int smjdev_profile_end(char *out, size_t cap) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    c->profiling = false;
    SMJ_CHECK(cudaStreamSynchronize(c->stream));
    struct Agg {
        u64    n, items;
        double ms;
    };
    std::map<std::string, Agg> agg;
    std::vector<std::string>   order;
    for (auto &r : g_prof) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.start, r.stop);
        if (!agg.count(r.name)) {
            order.push_back(r.name);
            agg[r.name] = {0, 0, 0.0};
        }
        Agg &a = agg[r.name];
        a.n++;
        a.items += r.items;
        a.ms += ms;
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    g_prof.clear();
    std::string txt;
    for (auto &k : order) {
        char line[512];
        snprintf(line, sizeof(line), "%s;%llu;%.6f;%llu\n", k.c_str(), agg[k].n, agg[k].ms,
                 agg[k].items);
        txt += line;
    }
    if (out && cap) {
        snprintf(out, cap, "%s", txt.c_str());
    }
    return 0;
}

int smjdev_event_create(void **ev) {
    if (!smj_ctx()) return -1;
    cudaEvent_t e;
    SMJ_CHECK(cudaEventCreate(&e));
    *ev = (void *)e;
    return 0;
}
int smjdev_event_record(void *ev) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    SMJ_CHECK(cudaEventRecord((cudaEvent_t)ev, c->stream));
    return 0;
}
int smjdev_event_elapsed_ms(void *start, void *stop, float *ms) {
    SMJ_CHECK(cudaEventSynchronize((cudaEvent_t)stop));
    SMJ_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return 0;
}
int smjdev_event_destroy(void *ev) {
    if (ev) SMJ_CHECK(cudaEventDestroy((cudaEvent_t)ev));
    return 0;
}

}  // extern "C"

// ---- buffers other processes (one per GPU) can map: CUDA IPC over NVLink / NVSwitch
extern "C" {
int smjdev_ipc_alloc(size_t bytes, void **dptr) {
    if (!smj_ctx()) return -1;
    if (bytes == 0) bytes = 16;
    SMJ_CHECK(cudaMalloc(dptr, bytes));  // pool (cudaMallocAsync) memory cannot be exported
    return 0;
}
int smjdev_ipc_free(void *dptr) {
    if (dptr) SMJ_CHECK(cudaFree(dptr));
    return 0;
}
int smjdev_ipc_export(void *dptr, unsigned char *handle64) {
    if (!smj_ctx()) return -1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    SMJ_CHECK(cudaIpcGetMemHandle(&h, dptr));
    memcpy(handle64, &h, 64);
    return 0;
}
int smjdev_ipc_open(const unsigned char *handle64, void **peer_ptr) {
    if (!smj_ctx()) return -1;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SMJ_CHECK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int smjdev_ipc_close(void *peer_ptr) {
    if (peer_ptr) SMJ_CHECK(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}
}

// ---------------------------------------------------------------------------------------
// Staged transfers for ORDINARY (pageable) host memory.  The reference's callers malloc their
// relations (src/main.c:516-528); a plain cudaMemcpy from such memory runs at ~11 GB/s on the
// B200 host because the driver stages it single-threaded.  Here several host threads copy
// 8 MiB chunks into a ring of pinned bounce buffers and each chunk is DMA-ed as soon as it is
// staged, so memcpy and PCIe traffic overlap and the host side is parallel.
// ---------------------------------------------------------------------------------------
#include <atomic>
#include <thread>

namespace {
constexpr int    STG_SLOTS = 12;
constexpr size_t STG_CHUNK = 8u << 20;
struct Stager {
    bool                   ready = false;
    int                    device = -1;
    void                  *bounce[STG_SLOTS];
    cudaEvent_t            ev[STG_SLOTS];
    std::atomic<long long> done[STG_SLOTS];  // index of the last chunk whose event was recorded
    cudaStream_t           stream = nullptr;
    cudaEvent_t            fence  = nullptr;
    std::mutex             mu;  // one staged transfer at a time (the ring is shared)
} g_stg;

int stager_init(SmjCtx *c) {
    if (g_stg.ready && g_stg.device == c->device) return 0;
    if (g_stg.ready) {  // device changed: rebuild
        for (int i = 0; i < STG_SLOTS; i++) {
            cudaFreeHost(g_stg.bounce[i]);
            cudaEventDestroy(g_stg.ev[i]);
        }
        cudaStreamDestroy(g_stg.stream);
        cudaEventDestroy(g_stg.fence);
        g_stg.ready = false;
    }
    for (int i = 0; i < STG_SLOTS; i++) {
        SMJ_CHECK(cudaMallocHost(&g_stg.bounce[i], STG_CHUNK));
        SMJ_CHECK(cudaEventCreateWithFlags(&g_stg.ev[i], cudaEventDisableTiming));
    }
    SMJ_CHECK(cudaStreamCreateWithFlags(&g_stg.stream, cudaStreamNonBlocking));
    SMJ_CHECK(cudaEventCreateWithFlags(&g_stg.fence, cudaEventDisableTiming));
    g_stg.device = c->device;
    g_stg.ready  = true;
    return 0;
}

int stager_threads(size_t nchunks) {
    unsigned hw = std::thread::hardware_concurrency();
    const char *e = getenv("SMJ_STAGE_THREADS");
    int t = e ? atoi(e) : (int)(hw ? hw / 2 : 4);
    t     = std::max(1, std::min(t, 8));
    return (int)std::min<size_t>((size_t)t, nchunks);
}

// dir 0: host -> device, dir 1: device -> host
int staged_copy(SmjCtx *c, char *dev, char *host, size_t bytes, int dir) {
    std::lock_guard<std::mutex> lk(g_stg.mu);
    int rc = stager_init(c);
    if (rc) return rc;
    // the device buffer may have been produced (or allocated) on the compute stream
    SMJ_CHECK(cudaEventRecord(g_stg.fence, c->stream));
    SMJ_CHECK(cudaStreamWaitEvent(g_stg.stream, g_stg.fence, 0));
    const size_t nchunks = (bytes + STG_CHUNK - 1) / STG_CHUNK;
    for (int i = 0; i < STG_SLOTS; i++) g_stg.done[i].store((long long)i - STG_SLOTS);
    std::atomic<size_t> next{0};
    std::atomic<int>    err{0};
    const int           device = c->device;
    auto worker = [&]() {
        if (cudaSetDevice(device) != cudaSuccess) {
            err.store((int)cudaErrorInvalidDevice);
            return;
        }
        while (!err.load()) {
            const size_t i = next.fetch_add(1);
            if (i >= nchunks) break;
            const int    slot = (int)(i % STG_SLOTS);
            const size_t off = i * STG_CHUNK, len = std::min(STG_CHUNK, bytes - off);
            while (g_stg.done[slot].load() != (long long)i - STG_SLOTS) std::this_thread::yield();
            cudaError_t e = cudaEventSynchronize(g_stg.ev[slot]);  // the slot's previous DMA is over
            if (e == cudaSuccess && dir == 0) {
                memcpy(g_stg.bounce[slot], host + off, len);
                e = cudaMemcpyAsync(dev + off, g_stg.bounce[slot], len, cudaMemcpyHostToDevice, g_stg.stream);
                if (e == cudaSuccess) e = cudaEventRecord(g_stg.ev[slot], g_stg.stream);
            } else if (e == cudaSuccess) {
                e = cudaMemcpyAsync(g_stg.bounce[slot], dev + off, len, cudaMemcpyDeviceToHost, g_stg.stream);
                if (e == cudaSuccess) e = cudaEventRecord(g_stg.ev[slot], g_stg.stream);
                if (e == cudaSuccess) e = cudaEventSynchronize(g_stg.ev[slot]);
                if (e == cudaSuccess) memcpy(host + off, g_stg.bounce[slot], len);
            }
            if (e != cudaSuccess) err.store((int)e);
            g_stg.done[slot].store((long long)i);
        }
    };
    const int nthreads = stager_threads(nchunks);
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    if (err.load()) return smj_fail((cudaError_t)err.load(), "staged_copy", __FILE__, __LINE__);
    // later work on the compute stream must see the uploaded data
    SMJ_CHECK(cudaEventRecord(g_stg.fence, g_stg.stream));
    SMJ_CHECK(cudaStreamWaitEvent(c->stream, g_stg.fence, 0));
    if (dir == 1) SMJ_CHECK(cudaStreamSynchronize(g_stg.stream));
    return 0;
}
}  // namespace

extern "C" {
// host -> device for any kind of host memory: one DMA for pinned memory, the threaded staging
// pipeline for pageable memory.  Asynchronous with respect to the library's stream.
int smjdev_upload(void *dst, const void *src, size_t bytes) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (bytes == 0) return 0;
    if (smjdev_pointer_kind(src) != 0 || bytes < 4 * STG_CHUNK || getenv("SMJ_NO_STAGING")) {
        SMJ_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        return 0;
    }
    return staged_copy(c, (char *)dst, (char *)const_cast<void *>(src), bytes, 0);
}
// device -> host, same idea; the data is in dst when the call returns
int smjdev_download(void *dst, const void *src, size_t bytes) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (bytes == 0) return 0;
    if (smjdev_pointer_kind(dst) != 0 || bytes < 4 * STG_CHUNK || getenv("SMJ_NO_STAGING")) {
        SMJ_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        SMJ_CHECK(cudaStreamSynchronize(c->stream));
        return 0;
    }
    return staged_copy(c, (char *)const_cast<void *>(src), (char *)dst, bytes, 1);
}
}

// ---------------------------------------------------------------------------------------
// Measurement helper (tools/p2p_store_bw.py): copy n 64-bit items with SM stores of `width` bytes
// per lane (8 or 16), run length `run` items: after every run the destination skips `gap` items,
// which mimics the (tile, bin) runs the scatter kernels write into peer memory.
// ---------------------------------------------------------------------------------------
namespace {
template <int WIDTH>
__global__ void __launch_bounds__(256) k_debug_copy(const u64 *__restrict__ src, u64 *__restrict__ dst, u64 n,
                                                    u32 run, u32 gap) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    if (WIDTH == 16) {
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; 2 * i + 1 < n; i += stride) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + 2 * i);
            const u64        o = 2 * i + ((2 * i) / run) * gap;
            *reinterpret_cast<ulonglong2 *>(dst + o) = v;
        }
    } else {
        for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            dst[i + (i / run) * gap] = src[i];
    }
}
}  // namespace
extern "C" int smjdev_debug_copy(const uint64_t *src, uint64_t *dst, uint64_t n, int width, uint32_t run,
                                 uint32_t gap) {
    SmjCtx *c = smj_ctx();
    if (!c) return -1;
    if (run == 0 || (run & 1) || (gap & 1)) return smj_fail_msg("smjdev_debug_copy: run and gap must be even");
    if (width == 16)
        SMJ_LAUNCH(c, k_debug_copy<16>, c->sm_count * 8, 256, 0, (const u64 *)src, (u64 *)dst, n, run, gap);
    else
        SMJ_LAUNCH(c, k_debug_copy<8>, c->sm_count * 8, 256, 0, (const u64 *)src, (u64 *)dst, n, run, gap);
    return 0;
}
