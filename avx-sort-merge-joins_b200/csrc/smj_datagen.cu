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

__device__ __forceinline__ u64 permute_row(u64 x, u64 n, u64 seed, int half);

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
        u64       x   = ((mix64(row + (seed << 32)) % maxkey + 1) << 32) | (row & 0xffffffffull);
        // keys with bits 20..30 all set read as NaN when the word is a double: the reference's
        // generators clear key bit 20 of those (avoid_NaN, generator.c:107-121)
        if ((x & (0x7FFull << 52)) == (0x7FFull << 52)) x &= ~(1ull << 52);
        out[i] = x;
    }
}

// S = the rows of R repeated |S| / |R| times (+ remainder) and shuffled (create_relation_fk_from_pk,
// generator.c:458-488): row i of S is R's row pi(i) mod |R|, pi a pseudo-random permutation of [0, |S|)
__global__ void __launch_bounds__(256) k_gen_fk_from(u64 *__restrict__ out, u64 n, u64 first, u64 count,
                                                     const u64 *__restrict__ src, u64 src_n, u64 seed, int half) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        out[i] = src[permute_row(first + i, n, seed, half) % src_n];
}

// Feistel permutation of [0, n) shared by k_gen_unique and the Zipf alphabet
__device__ __forceinline__ u64 permute_row(u64 x, u64 n, u64 seed, int half) {
    const u64 mask = (1ull << half) - 1;
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
    return x;
}

// ---------------------------------------------------------------------------------------
// Zipf generator: src/datagen/genzipf.c:29-159 (gen_alphabet, gen_zipf_lut, gen_zipf).
//   alphabet  a random permutation of 1..A          -> the Feistel permutation above, seed `aseed`
//   lut[i]    (sum_{j<=i+1} j^-theta) / (sum_{j<=A} j^-theta), doubles
//   draw      r = rand() / RAND_MAX, the reference's binary search (genzipf.c:121-143), key =
//             alphabet[pos]; payload = row index (the reference never writes it)
// The reference is sequential in three places: the glibc rand() stream, the running double sum and
// libm pow().  Per-row independence needs a counter-based variate (31 bits of mix64(row, seed):
// the resolution of rand() / RAND_MAX), and a reproducible LUT needs arithmetic that rounds the
// same way everywhere: i^-theta is computed from IEEE add / mul / div / fma only (zipf_weight; every
// step is correctly rounded on the device and on a CPU, so the CPU checker gets the same bits), and
// the running sum is a fixed three-level blocked sum (256 terms left to right, 256 blocks left to
// right, super-blocks left to right), prefix(i) = (S2 + S1) + S0.  lut values agree with the
// reference's pow()-based ones to ~1e-15 relative (tests/test_oracle_vs_ref.py compares them and the
// drawn key frequencies against the compiled reference).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double zipf_weight(u32 i, double theta) {
    // ln(i): i = m * 2^e, m in [sqrt(1/2), sqrt(2)); ln(m) = 2 atanh(z), z = (m - 1) / (m + 1)
    const double x    = (double)i;  // exact
    long long    bits = __double_as_longlong(x);
    int          e    = (int)((bits >> 52) & 0x7ff) - 1023;
    double       m    = __longlong_as_double((bits & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
    if (m > 1.4142135623730951) {
        m = __dmul_rn(m, 0.5);
        e += 1;
    }
    const double z = __ddiv_rn(__dadd_rn(m, -1.0), __dadd_rn(m, 1.0)), z2 = __dmul_rn(z, z);
    double       p = 1.0 / 23.0;
#pragma unroll
    for (int k = 21; k >= 1; k -= 2) p = __fma_rn(p, z2, __ddiv_rn(1.0, (double)k));
    const double ln_m = __dmul_rn(__dmul_rn(2.0, z), p);
    const double ln_i = __fma_rn((double)e, 0.6931471805599453, ln_m);
    // exp(y), y = -theta ln(i) <= 0: y = k ln2 + r, |r| <= ln2 / 2
    const double y  = __dmul_rn(-theta, ln_i);
    const double kf = rint(__dmul_rn(y, 1.4426950408889634));
    double       r  = __fma_rn(-kf, 0.6931471803691238, y);       // ln2 high part
    r               = __fma_rn(-kf, 1.9082149292705877e-10, r);   // ln2 low part
    double q = 1.0 / 6227020800.0;  // 1/13!
    const double inv_fact[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                 1.0 / 40320.0,     1.0 / 5040.0,     1.0 / 720.0,     1.0 / 120.0,
                                 1.0 / 24.0,        1.0 / 6.0,        0.5,             1.0,
                                 1.0};
#pragma unroll
    for (int k = 0; k < 13; k++) q = __fma_rn(q, r, inv_fact[k]);
    const int ki = (int)kf;  // >= -1022 for every theta <= 30
    return __dmul_rn(q, __longlong_as_double((long long)(ki + 1023) << 52));
}

constexpr int ZB = 256;  // terms per block, blocks per super-block

// level 0: thread b sums block b left to right; lut[i] = running sum inside the block
__global__ void __launch_bounds__(128) k_zipf_local(double *__restrict__ lut, u32 A, double theta,
                                                    double *__restrict__ t1) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    const u64 i0 = (u64)b * ZB;
    if (i0 >= A) return;
    const u32 cnt = (u32)min((u64)ZB, (u64)A - i0);
    double    sum = 0.0;
    for (u32 j = 0; j < cnt; j++) {
        sum = __dadd_rn(sum, zipf_weight((u32)(i0 + j) + 1u, theta));
        lut[i0 + j] = sum;
    }
    t1[b] = sum;
}
// level 1: thread sb turns the totals of its 256 blocks into exclusive prefixes, left to right
__global__ void __launch_bounds__(128) k_zipf_scan1(double *__restrict__ t1, u32 nblocks,
                                                    double *__restrict__ t2) {
    const u32 sb = blockIdx.x * blockDim.x + threadIdx.x;
    const u64 b0 = (u64)sb * ZB;
    if (b0 >= nblocks) return;
    const u32 cnt = (u32)min((u64)ZB, (u64)nblocks - b0);
    double    sum = 0.0;
    for (u32 j = 0; j < cnt; j++) {
        const double v = t1[b0 + j];
        t1[b0 + j]     = sum;
        sum            = __dadd_rn(sum, v);
    }
    t2[sb] = sum;
}
// level 2: one thread, super-block totals -> exclusive prefixes; t2[nsuper] = the grand total
__global__ void k_zipf_scan2(double *__restrict__ t2, u32 nsuper) {
    double sum = 0.0;
    for (u32 j = 0; j < nsuper; j++) {
        const double v = t2[j];
        t2[j]          = sum;
        sum            = __dadd_rn(sum, v);
    }
    t2[nsuper] = sum;
}
__global__ void __launch_bounds__(256) k_zipf_finish(double *__restrict__ lut, u32 A,
                                                     const double *__restrict__ t1,
                                                     const double *__restrict__ t2, u32 nsuper) {
    const double total  = t2[nsuper];
    const u64    stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < A; i += stride) {
        const u64 b = i / ZB;
        lut[i]      = __ddiv_rn(__dadd_rn(__dadd_rn(t2[b / ZB], t1[b]), lut[i]), total);
    }
}

__global__ void __launch_bounds__(256) k_gen_zipf(u64 *__restrict__ out, u64 first, u64 count,
                                                  const double *__restrict__ lut, u32 A, u64 seed,
                                                  u64 aseed, int half) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const u64    row = first + i;
        const double r   = __ddiv_rn((double)(mix64(row + (seed << 32)) >> 33), 2147483647.0);
        u32          pos;
        if (lut[0] >= r) {  // genzipf.c:126-143
            pos = 0;
        } else {
            u32 left = 0, right = A - 1;
            while (right - left > 1) {
                const u32 m = (u32)(((u64)left + right) / 2);
                if (lut[m] < r)
                    left = m;
                else
                    right = m;
            }
            pos = right;
        }
        out[i] = ((permute_row(pos, A, aseed, half) + 1) << 32) | (row & 0xffffffffull);
    }
}
}  // namespace

// The cumulative Zipf table (genzipf.c:59-93) of an alphabet of A symbols into lut (device, A
// doubles).  tmp: device scratch of (A / 256 + A / 65536 + 4) doubles.
extern "C" int smjdev_zipf_lut(double *lut, uint32_t A, double theta, double *tmp) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (A == 0) return smj_fail_msg("smjdev_zipf_lut: empty alphabet");
    if (!(theta >= 0.0) || theta > 30.0) return smj_fail_msg("smjdev_zipf_lut: theta must be in [0, 30]");
    const u32 nblocks = (u32)div_up(A, ZB), nsuper = (u32)div_up(nblocks, ZB);
    double   *t1 = tmp, *t2 = tmp + nblocks;
    SMJ_LAUNCH(ctx, k_zipf_local, (unsigned)div_up(nblocks, 128), 128, 0, lut, A, theta, t1);
    SMJ_LAUNCH(ctx, k_zipf_scan1, (unsigned)div_up(nsuper, 128), 128, 0, t1, nblocks, t2);
    SMJ_LAUNCH(ctx, k_zipf_scan2, 1, 1, 0, t2, nsuper);
    SMJ_LAUNCH(ctx, k_zipf_finish, (unsigned)std::min<u64>(div_up(A, 256), (u64)ctx->sm_count * 32), 256, 0, lut, A,
               (const double *)t1, (const double *)t2, nsuper);
    return 0;
}

// Rows [first, first + count) of a Zipf(theta) relation over the alphabet 1..A (genzipf.c:98-159):
// key = alphabet[pos(r)], alphabet = the pseudo-random permutation smjdev_gen_unique(.., n = A,
// seed = alphabet_seed) generates, payload = row index.  lut: NULL (the table is built in scratch
// memory and dropped) or a table smjdev_zipf_lut made for the same A and theta.
extern "C" int smjdev_gen_zipf(uint64_t *out, uint64_t first, uint64_t count, uint32_t A, double theta,
                               uint64_t seed, uint64_t alphabet_seed, const double *lut) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (count == 0) return 0;
    if (A == 0) return smj_fail_msg("smjdev_gen_zipf: empty alphabet");
    SmjScratch sc(ctx);
    if (!lut) {
        double *l = (double *)sc.get((size_t)A * 8);
        double *t = (double *)sc.get((size_t)(A / ZB + A / (ZB * ZB) + 8) * 8);
        if (!l || !t) return smj_fail_msg("smjdev_gen_zipf: out of device memory");
        int rc = smjdev_zipf_lut(l, A, theta, t);
        if (rc) return rc;
        lut = l;
    }
    int bits = 2;
    while ((1ull << bits) < (u64)A) bits++;
    const unsigned grid = (unsigned)std::min<u64>(div_up(count, 256), (u64)ctx->sm_count * 32);
    SMJ_LAUNCH(ctx, k_gen_zipf, grid, 256, 0, (u64 *)out, first, count, lut, A, seed, alphabet_seed,
               (bits + 1) / 2);
    return 0;
}

// rows [first, first + count) of an n-row relation whose rows are rows of src (device, src_n items)
extern "C" int smjdev_gen_fk_from(uint64_t *out, uint64_t n, uint64_t first, uint64_t count, const uint64_t *src,
                                  uint64_t src_n, uint64_t seed) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (count == 0) return 0;
    if (src_n == 0 || first + count > n) return smj_fail_msg("smjdev_gen_fk_from: bad arguments");
    int bits = 2;
    while ((1ull << bits) < n) bits++;
    const unsigned grid = (unsigned)std::min<u64>(div_up(count, 256), (u64)ctx->sm_count * 32);
    SMJ_LAUNCH(ctx, k_gen_fk_from, grid, 256, 0, (u64 *)out, n, first, count, (const u64 *)src, src_n, seed,
               (bits + 1) / 2);
    return 0;
}

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
