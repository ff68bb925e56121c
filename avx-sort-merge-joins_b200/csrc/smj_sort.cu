// smj_sort.cu -- C-ABI sort entry points (include/smj_device.h).
//
// smjdev_sort replaces src/avxsort/avxsort.c:128-245 (avxsort_aligned/_unaligned behind
// avxsort_tuples / avxsort_int64): it picks between the two device sorts
//   * LSD one-sweep radix passes over the varying bits      (smj_radix.cu, any distribution)
//   * MSD range partition + in-shared-memory segment sort   (smj_segsort.cu, when segments fit)
#include "smj_common.cuh"
#include "../../include/smj_device.h"

int smj_lsd_sort(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, u64 **sorted);
int smj_msd_sort(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, u64 **sorted,
                 int *used);

extern "C" int smjdev_sort(uint64_t *buf, uint64_t *tmp, uint64_t n, int key_only, int algo,
                           uint64_t **sorted) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (!sorted) return smj_fail_msg("smjdev_sort: sorted must not be NULL");
    u64 *res = nullptr;
    int  rc;
    if (algo == 0 || algo == 2) {
        int used = 0;
        rc       = smj_msd_sort(ctx, (const u64 *)buf, (u64 *)tmp, (u64 *)buf, n, key_only, &res, &used);
        if (rc) return rc;
        if (used) {
            *sorted = (uint64_t *)res;
            return 0;
        }
    }
    // classic two-buffer ping-pong: pass 1 buf -> tmp, pass 2 tmp -> buf, ...
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
