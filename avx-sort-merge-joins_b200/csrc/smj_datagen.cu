// smj_datagen.cu -- on-device deterministic relation generators.
//
// Replaces the reference's host generators for the benchmark workloads:
//   * src/datagen/generator.c:83-93, 234-252 (create_relation_pk: keys = shuffled 1..n) and
//     :255-350 (parallel_create_relation).  The reference seeds from time() and shuffles under a
//     race (generator.c:136-139, 164-175), so "same inputs" can only mean a deterministic
//     restatement: keys are a pseudo-random PERMUTATION of 1..n computed independently per row by
//     a 4-round Feistel network with cycle walking, payload = row index.
//   * src/datagen/generator.c:458-488 (create_relation_fk: keys uniformly drawn from 1..maxkey).
// The same integer formulas are restated on the CPU in oracle/smj_oracle.c (orc_gen_*), which is
// how the tests check the generated relations bit for bit.
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>

namespace {
__host__ __device__ __forceinline__ u64 mix64(u64 z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) k_gen_unique(u64 *__restrict__ out, u64 n, u64 first, u64 count,
                                                    u64 seed, int half) {
    const u64 mask   = (1ull << half) - 1;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const u64 row = first + i;
        u64       x   = row;
        do {
            u64 l = x >> half, r = x & mask;
#pragma unroll
            for (int round = 0; round < 4; round++) {
                u64 t = r;
                r     = l ^ (mix64(r + seed * 4 + (u64)round) & mask);
                l     = t;
            }
            x = (l << half) | r;
        } while (x >= n);
        out[i] = ((x + 1) << 32) | (row & 0xffffffffull);
    }
}

__global__ void __launch_bounds__(256) k_gen_fk(u64 *__restrict__ out, u64 first, u64 count,
                                                u32 maxkey, u64 seed) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const u64 row = first + i;
        out[i]        = ((mix64(row + (seed << 32)) % maxkey + 1) << 32) | (row & 0xffffffffull);
    }
}
}  // namespace

// rows [first, first+count) of the n-row relation are written to out[0..count)
extern "C" int smjdev_gen_unique(uint64_t *out, uint64_t n, uint64_t first, uint64_t count,
                                 uint64_t seed) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (count == 0) return 0;
    if (first + count > n) return smj_fail_msg("smjdev_gen_unique: rows out of range");
    int bits = 2;
    while ((1ull << bits) < n) bits++;
    const unsigned grid = (unsigned)std::min<u64>(div_up(count, 256), (u64)ctx->sm_count * 32);
    SMJ_LAUNCH(ctx, k_gen_unique, grid, 256, 0, (u64 *)out, n, first, count, seed, (bits + 1) / 2);
    return 0;
}

extern "C" int smjdev_gen_fk(uint64_t *out, uint64_t first, uint64_t count, uint32_t maxkey,
                             uint64_t seed) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (count == 0) return 0;
    if (maxkey == 0) return smj_fail_msg("smjdev_gen_fk: maxkey must be > 0");
    const unsigned grid = (unsigned)std::min<u64>(div_up(count, 256), (u64)ctx->sm_count * 32);
    SMJ_LAUNCH(ctx, k_gen_fk, grid, 256, 0, (u64 *)out, first, count, maxkey, seed);
    return 0;
}
