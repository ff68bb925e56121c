// smj_common.cuh -- shared helpers for the sm_100a kernels (internal, not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <unordered_map>
#include <vector>

typedef unsigned long long u64;
typedef unsigned int       u32;

// ---------------------------------------------------------------------------------------
// runtime context (smj_runtime.cu)
// ---------------------------------------------------------------------------------------
struct SmjProfRec {
    const char *name;
    cudaEvent_t start, stop;
    u64         items;
};
// One context per CUDA device.  A host thread works on the context it bound itself to with
// smjdev_init(device) (threads that never called it share the first context created), so several
// threads can drive several GPUs of one process concurrently (one thread per GPU: a context is
// not meant to be used by two threads at the same time).
struct SmjCtx {
    int          device;
    int          sm_count;
    cudaStream_t stream;      // stream all kernels are launched on
    cudaStream_t own_stream;  // the library's own stream (used when the caller sets none)
    cudaStream_t copy_stream; // host->device copies that overlap kernels on `stream`
    cudaStream_t aux_stream;  // second compute stream (multi-GPU join: pass B of R under the exchange of S)
    u64          launches;    // kernels launched so far on this device
    bool         ready;
    bool         profiling;   // smjdev_profile_begin(): CUDA events around every launch
    u64          next_items;  // items the next launch processes (annotation for the profile)
    // per-device caches (were function-local statics while the library knew one device)
    std::unordered_map<const void *, size_t> func_smem;   // dynamic shared memory limit set per kernel
    std::unordered_map<const void *, int>    func_occ;    // cached occupancy queries
    std::vector<cudaEvent_t>                 event_pool;  // untimed events (uploads, exchange)
    std::vector<SmjProfRec>                  prof;
};
void smj_prof_before(SmjCtx *ctx, const char *kernel_name);
void smj_prof_after(SmjCtx *ctx);
void smj_count_launch(SmjCtx *ctx);
#define SMJ_ITEMS(ctx, n) ((ctx)->next_items = (u64)(n))
SmjCtx *smj_ctx();                 // the calling thread's context; initialises lazily; nullptr on failure
SmjCtx *smj_ctx_of(int device);    // context of a given device (created on demand), does not bind
int     smj_fail(cudaError_t e, const char *what, const char *file, int line);
int     smj_fail_msg(const char *msg);
bool    smj_is_host_pointer(const void *p);  // mapped (pinned / registered) host memory
// raises the kernel's dynamic shared memory limit on ctx's device once (per device and kernel)
int     smj_func_smem(SmjCtx *ctx, const void *kernel, size_t bytes);
cudaEvent_t smj_pool_event(SmjCtx *ctx, size_t idx);  // idx-th cached untimed event of this device

#define SMJ_CHECK(expr)                                                        \
    do {                                                                       \
        cudaError_t e__ = (expr);                                              \
        if (e__ != cudaSuccess) return smj_fail(e__, #expr, __FILE__, __LINE__); \
    } while (0)

// every kernel launch goes through this so that gpu_launches is a real count
#define SMJ_LAUNCH(ctx, kernel, grid, block, smem, ...)                        \
    do {                                                                       \
        if ((ctx)->profiling) smj_prof_before((ctx), #kernel);                 \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);       \
        if ((ctx)->profiling) smj_prof_after((ctx));                           \
        (ctx)->next_items = 0;                                                 \
        smj_count_launch(ctx);                                                 \
        cudaError_t e__ = cudaGetLastError();                                  \
        if (e__ != cudaSuccess) return smj_fail(e__, #kernel, __FILE__, __LINE__); \
    } while (0)

// stream-ordered scratch allocation with automatic release at scope exit
struct SmjScratch {
    SmjCtx             *ctx;
    std::vector<void *> ptrs;
    u64                 bytes;
    explicit SmjScratch(SmjCtx *c) : ctx(c), bytes(0) {}
    SmjScratch(const SmjScratch &) = delete;
    SmjScratch &operator=(const SmjScratch &) = delete;
    ~SmjScratch() { release(); }
    void *get(size_t nbytes) {
        void *p = nullptr;
        if (nbytes == 0) nbytes = 16;
        if (cudaMallocAsync(&p, nbytes, ctx->stream) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        ptrs.push_back(p);
        bytes += nbytes;
        return p;
    }
    void release() {
        for (void *p : ptrs) cudaFreeAsync(p, ctx->stream);
        ptrs.clear();
    }
};

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#define SMJ_FULL_MASK 0xffffffffu

// bucket(x) = ((x - sub) >> shift) & mask.  sub = 1<<32 turns key into key-1 without touching
// the payload half (reference HASH_BIT_MODULO, src/partition/partition.c:29).
struct DigitFn {
    u64 sub;
    u32 shift;
    u32 mask;
    __host__ __device__ __forceinline__ u32 operator()(u64 x) const {
        return (u32)((x - sub) >> shift) & mask;
    }
    __host__ __device__ __forceinline__ u32 last_bin() const { return mask; }
};

// plain bit-field digit (LSD sort passes): no key-1 adjustment, saves the 64-bit subtract
struct FieldFn {
    u32 shift;
    u32 mask;
    __host__ __device__ __forceinline__ u32 operator()(u64 x) const {
        return (u32)(x >> shift) & mask;
    }
    __host__ __device__ __forceinline__ u32 last_bin() const { return mask; }
};

// destination of an item under range partitioning: number of splitters <= key (at most 32 parts)
struct SplitFn {
    u32 nparts;
    u32 sp[31];
    __host__ __device__ __forceinline__ u32 operator()(u64 x) const {
        const u32 k = (u32)(x >> 32);
        u32       d = 0;
        for (u32 i = 0; i + 1 < nparts; i++) d += (sp[i] <= k) ? 1u : 0u;  // nparts - 1 compares
        return d;
    }
    __host__ __device__ __forceinline__ u32 last_bin() const { return nparts - 1; }
};

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming (read-once) loads/stores: keep L1 clean, let L2 do the write combining
__device__ __forceinline__ u64 ld_stream(const u64 *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(u64 *p, u64 v) { __stcs(p, v); }

// pull one 128-byte line into L2 (no register, no dependency): used to fetch the tile a later CTA
// will start with, so that its loads find L2 instead of DRAM
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Bulk copies shared -> global (TMA, UBLKCP): one instruction of ONE thread writes a whole run of a
// bin; the copy engine reads shared memory and issues the stores, so the run costs the LSU pipe
// nothing.  Addresses and size are multiples of 16 bytes.  The writes of the generic proxy (STS)
// have to be made visible to the asynchronous proxy first (fence_proxy_async + barrier).
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, u32 bytes) {
    const u32 s = (u32)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(s), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk copies of this thread have READ their source (the shared-memory buffer may be reused)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stores the contiguous shared-memory range src[0 .. n) (8-byte items; src and dst have the same
// 16-byte phase) to global memory: one bulk copy for the 16-byte aligned middle, plain stores for an
// odd first / last item.  ONE thread, after every writer of the range has executed
// fence_proxy_async() and a CTA barrier; it must bulk_wait_read() before the range is overwritten
// or the CTA exits.
__device__ __forceinline__ void bulk_store_items(u64 *dst, const u64 *src, u32 n) {
    if (n && ((uintptr_t)dst & 8)) {
        __stcs(dst, *src);
        dst++, src++, n--;
    }
    if (n & 1) {
        n--;
        __stcs(dst + n, src[n]);
    }
    if (n) bulk_s2g(dst, src, n * 8u);
    bulk_commit();
}

// ---- bulk copies global -> shared (TMA, UBLKCP) completing on an mbarrier ------------------------
// One thread arms the barrier with the number of bytes in flight and issues the copies; the copy
// engine writes shared memory and counts the bytes off; everybody waits on the barrier's phase.
// The copy costs the LSU pipe and the issue slots nothing (a tile staged with LDG + STS costs two
// instructions per item and thread).  Source, destination and size are multiples of 16 bytes.
__device__ __forceinline__ void mbar_init(u64 *bar, u32 arrivals) {
    const u32 b = (u32)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    const u32 b = (u32)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, u32 bytes, u64 *bar) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst), b = (u32)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                 "l"(gmem_src), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    const u32 b = (u32)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}
// Staging of a contiguous global range src[0 .. n) (8-byte items, any 8-byte alignment) so that item i
// lands at dst_even[mis + i], mis = (src is 16-byte aligned ? 0 : 1); dst_even is 16-byte aligned.
// The 16-byte granules of the copy may include the item before src[0] (same 16-byte word, inside the
// array); an odd last item is not part of the bulk copy (its granule would end past the array) and is
// copied with plain instructions.  ONE thread, after mbar_init + a CTA barrier:
//     bytes = a.bytes() + b.bytes();  a.plain(); b.plain();          // plain stores before the arrive
//     mbar_arrive_expect_tx(bar, bytes);  a.issue(bar); b.issue(bar);
// everybody: mbar_wait(bar, phase).
struct BulkStage {
    u64       *dst_even;
    const u64 *src;
    u32        n, mis;
    __device__ __forceinline__ BulkStage(u64 *d, const u64 *s_, u32 n_)
        : dst_even(d), src(s_), n(n_), mis((u32)(((uintptr_t)s_ >> 3) & 1)) {}
    __device__ __forceinline__ u32 bytes() const { return n ? ((n + mis) & ~1u) * 8u : 0u; }
    __device__ __forceinline__ void plain() const {
        if (n && ((n + mis) & 1u)) dst_even[n + mis - 1] = src[n - 1];
    }
    __device__ __forceinline__ void issue(u64 *bar) const {
        if (bytes()) bulk_g2s(dst_even, src - mis, bytes(), bar);
    }
};

__device__ __forceinline__ u64 ld_relaxed(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u64 warp_reduce_add(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SMJ_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ u64 warp_reduce_or(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(SMJ_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ u64 warp_reduce_and(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v &= __shfl_xor_sync(SMJ_FULL_MASK, v, o);
    return v;
}

static inline u64 div_up(u64 a, u64 b) { return (a + b - 1) / b; }

// Key digits of the m-way join's two-level MSD partition, from the OR / AND of the items of R and
// S (bits = {OR_R, AND_R, OR_S, AND_S}): key bits [lo, hi] vary; level 1 = 9 bits at shift1, level 2
// = 9 bits at shift2, lbits key bits (at lo_shift) are left inside a bucket.  false when the key
// domain does not suit the path (fewer than 19 or more than 31 varying bits).
constexpr int SMJ_MSD_SAMPLE_WORDS = 16 + 2 * 512;  // host copy of the sample (smj_msd_sample)
struct MsdPlan {
    int lo, hi;
    u32 shift1, shift2, lo_shift, lbits;
};
__host__ __device__ inline bool smj_msd_plan(const u64 *bits, MsdPlan *p) {
    const u32 vor = (u32)((bits[0] | bits[2]) >> 32), vand = (u32)((bits[1] & bits[3]) >> 32);
    const u32 v   = vor & ~vand;
    if (v == 0) return false;
    int hi = 31, lo = 0;
    while (!((v >> hi) & 1u)) hi--;
    while (!((v >> lo) & 1u)) lo++;
    const int k = hi - lo + 1;
    if (k < 19 || k > 31) return false;
    p->lo       = lo;
    p->hi       = hi;
    p->lbits    = (u32)(k - 18);
    p->shift1   = 32 + (u32)hi - 8;
    p->shift2   = p->shift1 - 9;
    p->lo_shift = 32 + (u32)lo;
    return true;
}

// Capacity of a bin of an optimistic (histogram-free) partition pass from a sample: k of the sampled
// items fell into the bin, every sampled item stands for `scale` items.  The estimate plus seven
// standard deviations of the sampling error, rounded up to 16 items.  shrink (tests only) halves it.
__host__ __device__ inline u64 smj_cap1(u64 k, double scale, int shrink) {
    u64 cap = (u64)ceil(((double)k + 7.0 * sqrt((double)k + 1.0) + 8.0) * scale);
    if (shrink) cap = (u64)((double)k * scale * 0.5);  // test hook: bins that cannot hold their items
    return (cap + 15) & ~15ull;
}

// The join checksum (DESIGN.md / SURVEY.md 8d): sum over matched pairs of A(r)*B(s) mod 2^64.
__host__ __device__ __forceinline__ u64 smj_cks_a(u64 r) {
    return (r >> 32) * 0x9E3779B97F4A7C15ull + (r & 0xffffffffull) * 0xC2B2AE3D27D4EB4Full + 1ull;
}
__host__ __device__ __forceinline__ u64 smj_cks_b(u64 s) {
    return (s >> 32) * 0x165667B19E3779F9ull + (s & 0xffffffffull) * 0x27D4EB2F165667C5ull + 1ull;
}
