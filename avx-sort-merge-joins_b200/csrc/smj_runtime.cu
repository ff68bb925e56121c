// smj_runtime.cu -- device selection, stream, memory and timing entry points of smj_device.h.
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <stdlib.h>
#include <string.h>

// ---- contexts: one per device, created on demand; a thread is bound to one of them
constexpr int SMJ_MAX_DEVICES = 64;
static std::mutex               g_mu;
static SmjCtx                  *g_ctxs[SMJ_MAX_DEVICES] = {nullptr};
static SmjCtx                  *g_default = nullptr;  // context of the first successful init
static thread_local SmjCtx     *t_ctx = nullptr;
static std::atomic<u64>         g_launches{0};        // all devices, for smjdev_launch_count()
static thread_local char        g_err[512] = "";

void smj_count_launch(SmjCtx *ctx) {
    ctx->launches++;
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

// ---- optional per-kernel timing (bench.py's roofline numbers): CUDA events on the launch stream
void smj_prof_before(SmjCtx *ctx, const char *kernel_name) {
    SmjProfRec r;
    r.name  = kernel_name;
    r.items = ctx->next_items;
    cudaEventCreate(&r.start);
    cudaEventCreate(&r.stop);
    cudaEventRecord(r.start, ctx->stream);
    ctx->prof.push_back(r);
}
void smj_prof_after(SmjCtx *ctx) {
    if (!ctx->prof.empty()) cudaEventRecord(ctx->prof.back().stop, ctx->stream);
}

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

// creates (once) the context of `device`; g_mu held.  device < 0: $SMJ_DEVICE, $LOCAL_RANK or 0
static int ctx_create_locked(int device, SmjCtx **out) {
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
    if (device >= ndev || device >= SMJ_MAX_DEVICES) return smj_fail_msg("smjdev_init: device index out of range");
    if (g_ctxs[device]) {
        *out = g_ctxs[device];
        return 0;
    }
    SMJ_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    SMJ_CHECK(cudaGetDeviceProperties(&prop, device));
    SmjCtx *c   = new SmjCtx();
    c->device   = device;
    c->sm_count = prop.multiProcessorCount;
    c->launches = 0;
    c->ready = c->profiling = false;
    c->next_items = 0;
    SMJ_CHECK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    SMJ_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    SMJ_CHECK(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    // keep freed scratch in the pool: cudaMallocAsync then costs microseconds, not a driver call
    cudaMemPool_t pool;
    SMJ_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long thresh = ~0ull;
    SMJ_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    c->ready       = true;
    g_ctxs[device] = c;
    if (!g_default) g_default = c;
    *out = c;
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

// The calling thread's context.  The library makes its device the thread's current CUDA device
// and leaves it so (like cudaSetDevice itself, this is visible to the caller).
SmjCtx *smj_ctx() {
    if (!t_ctx) {
        std::lock_guard<std::mutex> lk(g_mu);
        t_ctx = g_default;
        if (!t_ctx && ctx_create_locked(-1, &t_ctx) != 0) return nullptr;
    }
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != t_ctx->device) cudaSetDevice(t_ctx->device);
    return t_ctx;
}
SmjCtx *smj_ctx_of(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    SmjCtx *c = nullptr;
    int     cur = -1;
    cudaGetDevice(&cur);
    if (ctx_create_locked(device, &c) != 0) c = nullptr;
    if (cur >= 0) cudaSetDevice(cur);
    return c;
}

int smj_func_smem(SmjCtx *ctx, const void *kernel, size_t bytes) {
    auto it = ctx->func_smem.find(kernel);
    if (it != ctx->func_smem.end() && it->second >= bytes) return 0;
    SMJ_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    ctx->func_smem[kernel] = bytes;
    return 0;
}

cudaEvent_t smj_pool_event(SmjCtx *ctx, size_t idx) {
    while (ctx->event_pool.size() <= idx) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ctx->event_pool.push_back(e);
    }
    return ctx->event_pool[idx];
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
// binds the calling thread to `device` (device < 0: keep the thread's binding if it has one, else
// $SMJ_DEVICE, $LOCAL_RANK or 0)
int smjdev_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (device < 0 && t_ctx) return 0;
    if (device < 0 && g_default) {
        t_ctx = g_default;
        return 0;
    }
    SmjCtx *c = nullptr;
    int     rc = ctx_create_locked(device, &c);
    if (rc) return rc;
    t_ctx = c;
    SMJ_CHECK(cudaSetDevice(c->device));
    return 0;
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
uint64_t    smjdev_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
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
// (the two ordering events are entries 0 and 1 of the context's event cache; the chunk events of
// an upload start at entry 2, see smj_join.cu)
int smjdev_exchange_begin(void) {
    SmjCtx *c = smj_ctx();
    cudaEvent_t ev = c ? smj_pool_event(c, 0) : nullptr;
    if (!ev) return -1;
    SMJ_CHECK(cudaEventRecord(ev, c->stream));
    SMJ_CHECK(cudaStreamWaitEvent(c->copy_stream, ev, 0));
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
    cudaEvent_t ev = c ? smj_pool_event(c, 1) : nullptr;
    if (!ev) return -1;
    SMJ_CHECK(cudaEventRecord(ev, c->copy_stream));
    SMJ_CHECK(cudaStreamWaitEvent(c->stream, ev, 0));
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
    for (auto &r : c->prof) {
        cudaEventDestroy(r.start);
        cudaEventDestroy(r.stop);
    }
    c->prof.clear();
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
    for (auto &r : c->prof) {
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
    c->prof.clear();
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
// 4 MiB chunks into a ring of pinned bounce buffers and each chunk is DMA-ed as soon as it is
// staged, so memcpy and PCIe traffic overlap and the host side is parallel.
// ---------------------------------------------------------------------------------------
#include <atomic>
#include <thread>
extern "C" void smj_stage_copy(void *dst, const void *src, size_t n);  // host/smj_copy.c

namespace {
constexpr int    STG_SLOTS = 16;
constexpr size_t STG_MIN_BYTES = 32u << 20;  // smaller transfers: one plain cudaMemcpyAsync
// bytes per bounce buffer; $SMJ_STAGE_CHUNK_KB (read once, before the first staged transfer).  16 x 4 MiB:
// the ring about fits the host's last-level cache (measured, 2 x 1 GB uploads + join: 8 MiB 48.8 ms,
// 4 MiB 44.6, 2 MiB 45.9, 1 MiB 49.6, 512 KiB 68 ms; 8 or 14 staging threads make no difference)
size_t stg_chunk() {
    static const size_t v = [] {
        const char *e = getenv("SMJ_STAGE_CHUNK_KB");
        const long  kb = e ? atol(e) : 0;
        return kb >= 64 && kb <= (64 << 10) ? (size_t)kb << 10 : (size_t)4 << 20;
    }();
    return v;
}
struct Stager {
    bool                   ready = false, ready_down = false;
    void                  *bounce[STG_SLOTS];       // uploads
    void                  *bounce_down[STG_SLOTS];  // downloads (the host READS these: never write-combined)
    cudaEvent_t            ev[STG_SLOTS];
    std::atomic<long long> done[STG_SLOTS];  // index of the last chunk whose event was recorded
    cudaStream_t           stream = nullptr;
    cudaEvent_t            fence  = nullptr;
    std::mutex             mu;  // one staged transfer at a time per device (the ring is shared)
};
Stager    *g_stgs[SMJ_MAX_DEVICES] = {nullptr};  // one ring per device, built on first use
std::mutex g_stgs_mu;

Stager *stager_of(SmjCtx *c) {
    std::lock_guard<std::mutex> lk(g_stgs_mu);
    if (!g_stgs[c->device]) g_stgs[c->device] = new Stager();
    return g_stgs[c->device];
}

// st->mu held
int stager_init(Stager *st) {
    if (st->ready) return 0;
    // $SMJ_STAGE_WC=1: write-combined bounce buffers for uploads (the host only writes them, front to
    // back; the DMA engine reads them without snooping the caches)
    const char *wc = getenv("SMJ_STAGE_WC");
    const unsigned flags = cudaHostAllocPortable | ((wc && atoi(wc)) ? cudaHostAllocWriteCombined : 0u);
    for (int i = 0; i < STG_SLOTS; i++) {
        SMJ_CHECK(cudaHostAlloc(&st->bounce[i], stg_chunk(), flags));
        st->bounce_down[i] = (flags & cudaHostAllocWriteCombined) ? nullptr : st->bounce[i];
        SMJ_CHECK(cudaEventCreateWithFlags(&st->ev[i], cudaEventDisableTiming));
    }
    st->ready_down = !(flags & cudaHostAllocWriteCombined);
    SMJ_CHECK(cudaStreamCreateWithFlags(&st->stream, cudaStreamNonBlocking));
    SMJ_CHECK(cudaEventCreateWithFlags(&st->fence, cudaEventDisableTiming));
    st->ready = true;
    return 0;
}

int stager_threads(size_t nchunks) {
    unsigned hw = std::thread::hardware_concurrency();
    const char *e = getenv("SMJ_STAGE_THREADS");
    int t = e ? atoi(e) : std::min((int)(hw ? hw / 2 : 4), 8);
    t     = std::max(1, std::min(t, 32));
    return (int)std::min<size_t>((size_t)t, nchunks);
}

// dir 0: host -> device, dir 1: device -> host
int staged_copy(SmjCtx *c, char *dev, char *host, size_t bytes, int dir) {
    Stager *st = stager_of(c);
    std::lock_guard<std::mutex> lk(st->mu);
    int rc = stager_init(st);
    if (rc) return rc;
    if (dir == 1 && !st->ready_down) {  // (write-combined upload ring: downloads get a ring of their own)
        for (int i = 0; i < STG_SLOTS; i++) SMJ_CHECK(cudaHostAlloc(&st->bounce_down[i], stg_chunk(), cudaHostAllocPortable));
        st->ready_down = true;
    }
    // the device buffer may have been produced (or allocated) on the compute stream
    SMJ_CHECK(cudaEventRecord(st->fence, c->stream));
    SMJ_CHECK(cudaStreamWaitEvent(st->stream, st->fence, 0));
    const size_t STG_CHUNK = stg_chunk();
    const size_t nchunks = (bytes + STG_CHUNK - 1) / STG_CHUNK;
    // slots in use: $SMJ_STAGE_SLOTS (2 .. 16)
    static const int NSLOT = [] {
        const char *e = getenv("SMJ_STAGE_SLOTS");
        const int   v = e ? atoi(e) : STG_SLOTS;
        return v >= 2 && v <= STG_SLOTS ? v : STG_SLOTS;
    }();
    for (int i = 0; i < NSLOT; i++) st->done[i].store((long long)i - NSLOT);
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
            const int    slot = (int)(i % NSLOT);
            const size_t off = i * STG_CHUNK, len = std::min(STG_CHUNK, bytes - off);
            while (st->done[slot].load() != (long long)i - NSLOT) std::this_thread::yield();
            cudaError_t e = cudaEventSynchronize(st->ev[slot]);  // the slot's previous DMA is over
            if (e == cudaSuccess && dir == 0) {
                smj_stage_copy(st->bounce[slot], host + off, len);
                e = cudaMemcpyAsync(dev + off, st->bounce[slot], len, cudaMemcpyHostToDevice, st->stream);
                if (e == cudaSuccess) e = cudaEventRecord(st->ev[slot], st->stream);
            } else if (e == cudaSuccess) {
                e = cudaMemcpyAsync(st->bounce_down[slot], dev + off, len, cudaMemcpyDeviceToHost, st->stream);
                if (e == cudaSuccess) e = cudaEventRecord(st->ev[slot], st->stream);
                if (e == cudaSuccess) e = cudaEventSynchronize(st->ev[slot]);
                if (e == cudaSuccess) memcpy(host + off, st->bounce_down[slot], len);
            }
            if (e != cudaSuccess) err.store((int)e);
            st->done[slot].store((long long)i);
        }
    };
    const int nthreads = stager_threads(nchunks);
    const bool dbg = getenv("SMJ_STAGE_DEBUG") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    if (dbg) {
        const auto t_enq = std::chrono::steady_clock::now();
        cudaStreamSynchronize(st->stream);
        const auto t_done = std::chrono::steady_clock::now();
        fprintf(stderr, "[smj_b200] staged copy of %.0f MB: all chunks enqueued after %.2f ms, DMA done after %.2f ms (%d threads)\n",
                bytes / 1e6, std::chrono::duration<double, std::milli>(t_enq - t_begin).count(),
                std::chrono::duration<double, std::milli>(t_done - t_begin).count(), nthreads);
    }
    if (err.load()) return smj_fail((cudaError_t)err.load(), "staged_copy", __FILE__, __LINE__);
    // later work on the compute stream must see the uploaded data
    SMJ_CHECK(cudaEventRecord(st->fence, st->stream));
    SMJ_CHECK(cudaStreamWaitEvent(c->stream, st->fence, 0));
    if (dir == 1) SMJ_CHECK(cudaStreamSynchronize(st->stream));
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
    if (smjdev_pointer_kind(src) != 0 || bytes < STG_MIN_BYTES || getenv("SMJ_NO_STAGING")) {
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
    if (smjdev_pointer_kind(dst) != 0 || bytes < STG_MIN_BYTES || getenv("SMJ_NO_STAGING")) {
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
