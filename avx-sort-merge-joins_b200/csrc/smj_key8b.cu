// smj_key8b.cu -- the KEY_8B build of the join path: 16-byte tuples {int64 payload; int64 key}.
//
// The reference selects 64-bit keys and payloads at compile time (src/types.h:23-29); such a build
// has no AVX kernels -- src/main.c:871-877 insists on --scalarsort --scalarmerge -- so its join is
// std::sort on the key (src/scalarsort/scalarsort.c:35-51) + scalar merges + merge_join
// (src/joins/joincommon.c:239-312).  On the GPU a 16-byte tuple is sorted through its INDEX with the
// 8-byte machinery of this library: words (key bits << 32 | row) go through the stable key sort
// once for the low and -- only if they vary -- once for the high 32 key bits (LSD order), then one
// gather writes the sorted tuples and a dense array of their keys.  The merge-join scan is one
// thread per S tuple: lower / upper bound of its key in R's sorted keys,
//     count    += |run_R|,     checksum += (prefix_A[ub] - prefix_A[lb]) * B(s)
// with prefix_A the running sum of A(r) over sorted R (the checksum is the 64-bit analogue of the
// 8-byte one, smj_cks16_a / _b).  Functionally complete (count and checksum exact for any key
// multiset, signed keys included), not tuned like the 8-byte path: ~3 sort passes over 8-byte words
// per relation + 2 log2|R| probes per S tuple.
#include "smj_internal.cuh"
#include "../../include/smj_device.h"

#include <algorithm>

namespace {
struct Tup16 {
    long long payload, key;
};

__host__ __device__ __forceinline__ u64 smj_cks16_a(u64 key, u64 payload) {
    return key * 0x9E3779B97F4A7C15ull + payload * 0xC2B2AE3D27D4EB4Full + 1ull;
}
__host__ __device__ __forceinline__ u64 smj_cks16_b(u64 key, u64 payload) {
    return key * 0x165667B19E3779F9ull + payload * 0x27D4EB2F165667C5ull + 1ull;
}
// order-preserving map of a signed key to an unsigned one
__device__ __forceinline__ u64 ukey(long long k) { return (u64)k ^ 0x8000000000000000ull; }

__global__ void __launch_bounds__(256) k16_bits(const Tup16 *__restrict__ t, u64 n, u64 *__restrict__ out) {
    u64       vor = 0, vand = ~0ull;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 k = ukey(t[i].key);
        vor |= k;
        vand &= k;
    }
    vor  = warp_reduce_or(vor);
    vand = warp_reduce_and(vand);
    if (lane_id() == 0) {
        atomicOr(out, vor);
        atomicAnd(out + 1, vand);
    }
}
// word = (32 key bits at `shift` << 32) | row.  order == null: row i itself
__global__ void __launch_bounds__(256) k16_words(const Tup16 *__restrict__ t, const u64 *__restrict__ order, u64 n,
                                                 u32 shift, u64 *__restrict__ words) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 row = order ? (order[i] & 0xffffffffull) : i;
        words[i]      = ((ukey(t[row].key) >> shift) << 32) | row;
    }
}
__global__ void __launch_bounds__(256) k16_gather(const Tup16 *__restrict__ t, const u64 *__restrict__ order, u64 n,
                                                  Tup16 *__restrict__ out, u64 *__restrict__ keys,
                                                  u64 *__restrict__ a_of) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const Tup16 x = t[order ? (order[i] & 0xffffffffull) : i];
        if (out) out[i] = x;
        if (keys) keys[i] = ukey(x.key);
        if (a_of) a_of[i] = smj_cks16_a((u64)x.key, (u64)x.payload);
    }
}
__global__ void __launch_bounds__(256) k16_join(const u64 *__restrict__ rkeys, const u64 *__restrict__ prefa, u64 total_a,
                                                u64 nr, const Tup16 *__restrict__ s, const u64 *__restrict__ sorder,
                                                u64 ns, u64 *__restrict__ acc) {
    u64       cnt = 0, cks = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < ns; j += stride) {
        const Tup16 y = s[sorder ? (sorder[j] & 0xffffffffull) : j];
        const u64   k = ukey(y.key);
        u64 lo = 0, hi = nr;
        while (lo < hi) {  // first R key >= k
            const u64 mid = (lo + hi) >> 1;
            if (rkeys[mid] < k) lo = mid + 1; else hi = mid;
        }
        if (lo == nr || rkeys[lo] != k) continue;
        u64 ub = lo + 1, hi2 = nr;
        if (ub < nr && rkeys[ub] == k) {  // duplicates in R: first R key > k
            while (ub < hi2) {
                const u64 mid = (ub + hi2) >> 1;
                if (rkeys[mid] <= k) ub = mid + 1; else hi2 = mid;
            }
        }
        cnt += ub - lo;
        const u64 hi_sum = ub < nr ? prefa[ub] : total_a;
        cks += (hi_sum - prefa[lo]) * smj_cks16_b((u64)y.key, (u64)y.payload);
    }
    cnt = warp_reduce_add(cnt);
    cks = warp_reduce_add(cks);
    if (lane_id() == 0) {
        if (cnt) atomicAdd(acc, cnt);
        if (cks) atomicAdd(acc + 1, cks);
    }
}

// inclusive-to-exclusive scan of n u64 (any n): per-1024 blocks, totals scanned recursively
__global__ void __launch_bounds__(1024) k16_scan_local(const u64 *__restrict__ in, u64 n, u64 *__restrict__ out,
                                                       u64 *__restrict__ totals) {
    __shared__ u64 s_w[32];
    const u64 i    = (u64)blockIdx.x * 1024 + threadIdx.x;
    const u64 v    = i < n ? in[i] : 0ull;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64       incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u64 base = 0;
    for (u32 w = 0; w < warp; w++) base += s_w[w];
    if (i < n) out[i] = base + incl - v;
    if (threadIdx.x == 1023) totals[blockIdx.x] = base + incl;
}
__global__ void __launch_bounds__(1024) k16_scan_add(u64 *__restrict__ out, u64 n, const u64 *__restrict__ block_pre) {
    const u64 i = (u64)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += block_pre[blockIdx.x];
}
int scan_any(SmjCtx *ctx, const u64 *in, u64 n, u64 *out, u64 *total_dev) {
    const u64 nblk = div_up(n, 1024);
    if (nblk == 0) return 0;
    SmjScratch sc(ctx);
    u64 *tot = (u64 *)sc.get((nblk + 1) * 8), *pre = (u64 *)sc.get((nblk + 1) * 8);
    if (!tot || !pre) return smj_fail_msg("key8b: scratch allocation failed");
    SMJ_LAUNCH(ctx, k16_scan_local, (unsigned)nblk, 1024, 0, in, n, out, tot);
    if (nblk > 1) {
        int rc = scan_any(ctx, tot, nblk, pre, total_dev);
        if (rc) return rc;
        SMJ_LAUNCH(ctx, k16_scan_add, (unsigned)nblk, 1024, 0, out, n, pre);
    } else if (total_dev) {
        SMJ_CHECK(cudaMemcpyAsync(total_dev, tot, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}

unsigned grid_for(SmjCtx *ctx, u64 n) { return (unsigned)std::max<u64>(1, std::min<u64>(div_up(n, 256), (u64)ctx->sm_count * 16)); }

// order (device, n words whose low 32 bits are row numbers) of the tuples by signed key, stable;
// *order_out = nullptr when the input is trivially ordered (n <= 1)
int order16(SmjCtx *ctx, SmjScratch &sc, const Tup16 *t, u64 n, u64 **order_out) {
    *order_out = nullptr;
    if (n <= 1) return 0;
    if (n >= (1ull << 32)) return smj_fail_msg("key8b: at most 2^32 - 1 tuples per relation");
    u64 *d_bits = (u64 *)sc.get(16), *wa = (u64 *)sc.get(n * 8), *wb = (u64 *)sc.get(n * 8), *wc = (u64 *)sc.get(n * 8);
    u64 *wd = (u64 *)sc.get(n * 8);
    if (!d_bits || !wa || !wb || !wc || !wd) return smj_fail_msg("key8b: out of device memory");
    const u64 init[2] = {0ull, ~0ull};
    SMJ_CHECK(cudaMemcpyAsync(d_bits, init, 16, cudaMemcpyHostToDevice, ctx->stream));
    SMJ_LAUNCH(ctx, k16_bits, grid_for(ctx, n), 256, 0, t, n, d_bits);
    u64 h[2];
    SMJ_CHECK(cudaMemcpyAsync(h, d_bits, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    const u64 varying = h[0] & ~h[1];
    u64      *cur     = nullptr;
    int       rc;
    if (varying & 0xffffffffull) {  // low half of the key: words in wa, result in wa, wb or wc
        SMJ_LAUNCH(ctx, k16_words, grid_for(ctx, n), 256, 0, t, (const u64 *)nullptr, n, 0u, wa);
        if ((rc = smj_lsd_sort(ctx, wa, wb, wc, n, 1, &cur))) return rc;
    }
    if (varying >> 32) {  // high half, stable on top of the low-half order: words in wd, result in wd, wa or wb
        SMJ_LAUNCH(ctx, k16_words, grid_for(ctx, n), 256, 0, t, (const u64 *)cur, n, 32u, wd);
        if ((rc = smj_lsd_sort(ctx, wd, wa, wb, n, 1, &cur))) return rc;
    }
    if (!cur) {  // all keys equal: any order is sorted; keep the rows as they are
        SMJ_LAUNCH(ctx, k16_words, grid_for(ctx, n), 256, 0, t, (const u64 *)nullptr, n, 0u, wa);
        cur = wa;
    }
    *order_out = cur;
    return 0;
}
}  // namespace

extern "C" {

// Sort of n 16-byte tuples {int64 payload; int64 key} by signed key, equal keys in input order
// (scalarsort_tuples of a KEY_8B build, src/scalarsort/scalarsort.c:35-51).  in != out, both device.
int smjdev_sort_key8b(const void *in, void *out, uint64_t n) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (n == 0) return 0;
    SmjScratch sc(ctx);
    u64       *order = nullptr;
    int        rc    = order16(ctx, sc, (const Tup16 *)in, n, &order);
    if (rc) return rc;
    SMJ_LAUNCH(ctx, k16_gather, grid_for(ctx, n), 256, 0, (const Tup16 *)in, (const u64 *)order, n, (Tup16 *)out,
               (u64 *)nullptr, (u64 *)nullptr);
    return 0;
}

// Join of two relations of 16-byte tuples (device pointers, left unmodified): count of matching
// pairs and checksum = sum over matched pairs of A(r) * B(s) mod 2^64,
//   A(r) = key * 0x9E3779B97F4A7C15 + payload * 0xC2B2AE3D27D4EB4F + 1,
//   B(s) = key * 0x165667B19E3779F9 + payload * 0x27D4EB2F165667C5 + 1     (64-bit wrapping).
// presorted != 0: both inputs are already sorted by key (merge_join of a KEY_8B build).
int smjdev_join_key8b(const void *r_, uint64_t nr, const void *s_, uint64_t ns, int presorted,
                      smj_join_stats_t *stats) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (!stats) return smj_fail_msg("smjdev_join_key8b: stats must not be NULL");
    memset(stats, 0, sizeof(*stats));
    if (nr == 0 || ns == 0) return 0;
    const Tup16 *r = (const Tup16 *)r_, *s = (const Tup16 *)s_;
    const u64    launches0 = ctx->launches;
    cudaEvent_t  e0, e1;
    SMJ_CHECK(cudaEventCreate(&e0));
    SMJ_CHECK(cudaEventCreate(&e1));
    SMJ_CHECK(cudaEventRecord(e0, ctx->stream));
    SmjScratch sc(ctx);
    u64 *ord_r = nullptr, *ord_s = nullptr;
    int  rc;
    if (!presorted) {
        if ((rc = order16(ctx, sc, r, nr, &ord_r))) return rc;
        if ((rc = order16(ctx, sc, s, ns, &ord_s))) return rc;
    }
    u64 *rkeys = (u64 *)sc.get(nr * 8), *a_of = (u64 *)sc.get(nr * 8), *prefa = (u64 *)sc.get(nr * 8);
    u64 *d_tot = (u64 *)sc.get(8), *acc = (u64 *)sc.get(16);
    if (!rkeys || !a_of || !prefa || !d_tot || !acc) return smj_fail_msg("smjdev_join_key8b: out of device memory");
    SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    SMJ_LAUNCH(ctx, k16_gather, grid_for(ctx, nr), 256, 0, r, (const u64 *)ord_r, nr, (Tup16 *)nullptr, rkeys, a_of);
    if ((rc = scan_any(ctx, a_of, nr, prefa, d_tot))) return rc;
    u64 h_tot = 0;  // (the grand total is only needed for runs that end at the last R tuple)
    SMJ_CHECK(cudaMemcpyAsync(&h_tot, d_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    SMJ_ITEMS(ctx, nr + ns);
    SMJ_LAUNCH(ctx, k16_join, grid_for(ctx, ns), 256, 0, (const u64 *)rkeys, (const u64 *)prefa, h_tot, nr, s,
               (const u64 *)ord_s, ns, acc);
    u64 h[2];
    SMJ_CHECK(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaEventRecord(e1, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&stats->ms_total, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    stats->matches   = h[0];
    stats->checksum  = h[1];
    stats->launches  = ctx->launches - launches0;
    stats->hbm_bytes = sc.bytes;
    return 0;
}

}  // extern "C"
