// smj_sort.cu -- C-ABI sort entry points (include/smj_device.h).
//
// smjdev_sort replaces src/avxsort/avxsort.c:128-245 (avxsort_aligned/_unaligned behind
// avxsort_tuples / avxsort_int64): whole-word sorts of large inputs take the MSD path
// (smj_msdsort.cu: two histogram-free range-partition passes + in-bucket counting sort), everything
// else LSD one-sweep radix passes over the varying bits (smj_radix.cu).  The m-way join does not go through here: it partitions MSD-first and never
// materialises sorted buckets (smj_join.cu).
#include "smj_common.cuh"
#include <algorithm>
#include "../../include/smj_device.h"

int smj_lsd_sort(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, u64 **sorted);
int smj_msd_sort(SmjCtx *ctx, const u64 *in, u64 n, u64 *out, int *used);

extern "C" int smjdev_sort(uint64_t *buf, uint64_t *tmp, uint64_t n, int key_only, int algo,
                           uint64_t **sorted) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (!sorted) return smj_fail_msg("smjdev_sort: sorted must not be NULL");
    u64 *res = nullptr;
    if (algo != 0 && algo != 1) return smj_fail_msg("smjdev_sort: unknown algorithm");
    int rc;
    if (key_only == 0) {
        // whole-word order, large input: MSD range partition + in-bucket counting sort, buf -> tmp
        // (smj_msdsort.cu); it declines inputs it does not suit
        int used = 0;
        if ((rc = smj_msd_sort(ctx, (const u64 *)buf, n, (u64 *)tmp, &used))) return rc;
        if (used) {
            *sorted = tmp;
            return 0;
        }
    }
    // LSD passes, classic two-buffer ping-pong: pass 1 buf -> tmp, pass 2 tmp -> buf, ...
    rc = smj_lsd_sort(ctx, (const u64 *)buf, (u64 *)tmp, (u64 *)buf, n, key_only, &res);
    if (rc) return rc;
    *sorted = (uint64_t *)res;
    return 0;
}

// ---------------------------------------------------------------------------------------
// int32 sort (src/scalarsort/scalarsort.c:91-99): widen to items with an order preserving key,
// key-only sort, narrow back
// ---------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_widen_i32(const int *__restrict__ in, u64 n,
                                                   u64 *__restrict__ out) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (u64)((u32)in[i] ^ 0x80000000u) << 32;
}
__global__ void __launch_bounds__(256) k_narrow_i32(const u64 *__restrict__ in, u64 n,
                                                    int *__restrict__ out) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (int)((u32)(in[i] >> 32) ^ 0x80000000u);
}
}  // namespace

extern "C" int smjdev_sort_int32(int32_t *data, uint64_t n) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (n <= 1) return 0;
    const int  kind      = smjdev_pointer_kind(data);
    const bool on_device = (kind == 2 || kind == 3);
    SmjScratch sc(ctx);
    int *d32 = on_device ? data : (int *)sc.get(n * 4);
    u64 *a = (u64 *)sc.get(n * 8), *b = (u64 *)sc.get(n * 8);
    if (!d32 || !a || !b) return smj_fail_msg("smjdev_sort_int32: out of device memory");
    if (!on_device)
        SMJ_CHECK(cudaMemcpyAsync(d32, data, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    const unsigned grid = (unsigned)std::min<u64>(div_up(n, 256), (u64)ctx->sm_count * 16);
    SMJ_LAUNCH(ctx, k_widen_i32, grid, 256, 0, d32, n, a);
    u64 *res = nullptr;
    int  rc  = smj_lsd_sort(ctx, a, b, a, n, 1, &res);
    if (rc) return rc;
    SMJ_LAUNCH(ctx, k_narrow_i32, grid, 256, 0, res, n, d32);
    if (!on_device)
        SMJ_CHECK(cudaMemcpyAsync(data, d32, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------
// Device-side verification helpers (full-size property tests: the arrays never leave HBM).
// ---------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_check_order(const u64 *__restrict__ x, u64 n, int key_only,
                                                     u64 *__restrict__ out /* inv, sum, xor */) {
    u64       inv = 0, sum = 0, xr = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 v = x[i];
        sum += v;
        xr ^= v;
        if (i > 0) {
            const u64 p = x[i - 1];
            inv += key_only ? ((v >> 32) < (p >> 32)) : (v < p);
        }
    }
    inv = warp_reduce_add(inv);
    sum = warp_reduce_add(sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xr ^= __shfl_xor_sync(SMJ_FULL_MASK, xr, o);
    if (lane_id() == 0) {
        atomicAdd(out, inv);
        atomicAdd(out + 1, sum);
        atomicXor(out + 2, xr);
    }
}

// every slot of a partitioned output: does the item belong to the partition that owns the slot,
// and (stable scatter of rows whose payload is the row number) do payloads ascend inside it?
__global__ void __launch_bounds__(256) k_check_partition(const u64 *__restrict__ x,
                                                         const u64 *__restrict__ offs,
                                                         const u64 *__restrict__ hist, u32 fanout,
                                                         u64 used, DigitFn fn,
                                                         u64 *__restrict__ out /* misplaced, unstable, items, sum */) {
    u64       bad = 0, unst = 0, cnt = 0, sum = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < used; i += stride) {
        u32 lo = 0, hi = fanout;  // last partition with offs <= i
        while (hi - lo > 1) {
            u32 mid = (lo + hi) >> 1;
            if (offs[mid] <= i)
                lo = mid;
            else
                hi = mid;
        }
        if (i >= offs[lo] + hist[lo]) continue;  // padding slot
        const u64 v = x[i];
        cnt++;
        sum += v;
        bad += fn(v) != lo;
        if (i > offs[lo]) unst += (u32)v <= (u32)x[i - 1];
    }
    bad  = warp_reduce_add(bad);
    unst = warp_reduce_add(unst);
    cnt  = warp_reduce_add(cnt);
    sum  = warp_reduce_add(sum);
    if (lane_id() == 0) {
        atomicAdd(out, bad);
        atomicAdd(out + 1, unst);
        atomicAdd(out + 2, cnt);
        atomicAdd(out + 3, sum);
    }
}
}  // namespace

// stats (host, 3 x u64): number of descents, sum of the items mod 2^64, xor of the items
extern "C" int smjdev_check_order(const uint64_t *items, uint64_t n, int key_only, uint64_t *stats) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    SmjScratch sc(ctx);
    u64 *d = (u64 *)sc.get(24);
    if (!d) return smj_fail_msg("smjdev_check_order: scratch");
    SMJ_CHECK(cudaMemsetAsync(d, 0, 24, ctx->stream));
    if (n) {
        const unsigned grid = (unsigned)std::min<u64>(div_up(n, 256), (u64)ctx->sm_count * 16);
        SMJ_LAUNCH(ctx, k_check_order, grid, 256, 0, (const u64 *)items, n, key_only, d);
    }
    SMJ_CHECK(cudaMemcpyAsync(stats, d, 24, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// stats (host, 4 x u64): items in a wrong partition, payload descents inside partitions, items
// found, sum of the items mod 2^64.  offs/hist: device arrays as written by smjdev_partition.
extern "C" int smjdev_check_partition(const uint64_t *out, const uint64_t *offs, const uint64_t *hist,
                                      uint32_t nbits, uint32_t shift, int key_minus_one,
                                      uint64_t used, uint64_t *stats) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    SmjScratch sc(ctx);
    u64 *d = (u64 *)sc.get(32);
    if (!d) return smj_fail_msg("smjdev_check_partition: scratch");
    SMJ_CHECK(cudaMemsetAsync(d, 0, 32, ctx->stream));
    DigitFn fn = {key_minus_one ? (1ull << 32) : 0ull, 32 + shift, (1u << nbits) - 1};
    if (used) {
        const unsigned grid = (unsigned)std::min<u64>(div_up(used, 256), (u64)ctx->sm_count * 16);
        SMJ_LAUNCH(ctx, k_check_partition, grid, 256, 0, (const u64 *)out, (const u64 *)offs,
                   (const u64 *)hist, 1u << nbits, used, fn, d);
    }
    SMJ_CHECK(cudaMemcpyAsync(stats, d, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}
