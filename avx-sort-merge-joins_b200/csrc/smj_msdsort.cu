// smj_msdsort.cu -- whole-word sort of 8-byte items, MSD first (the fast path of smjdev_sort).
//
// Replaces src/avxsort/avxsort.c:128-245 (in-register sorting networks + bitonic merges, then
// 2-way merge passes over the whole array) for large inputs.  An LSD radix sort moves every item
// once per 9 key bits through a STABLE scatter (0.86 ms per 128 M items); here the array is range
// partitioned by its top 18 varying key bits with the two UNSTABLE, histogram-free passes of the
// m-way join (smj_radix.cu: sampled bin capacities, fixed slots, exact repair of overflowing
// buckets; 0.50 ms per pass) and every bucket is then finished on its own by k_bucket_sort:
// a counting sort over the <= 2^13 key values that are left inside a bucket, written straight to
// its final position.  Equal keys (ties) are ordered by payload afterwards, run by run; inputs
// with long runs of equal keys, a skewed key distribution or an unsuitable key domain are left to
// the LSD sort (smj_lsd_sort), so the result never depends on the guesses.
// HBM traffic: 16 + 16 + 16 B/item whatever the key width (LSD: 8 + 8 + 16 per 9 key bits).
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

int smj_msd_sample(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, u32 max_sectors, u64 *d_st,
                   u32 *d_sub, u64 *host);
size_t smj_msd_sample_bytes(int which);
u64 smj_msd_caps1(const u64 *hist, u64 nsamp, u64 n, int shrink);
int smj_msd_caps1_tables(SmjCtx *ctx, const u64 *d_st, u64 nr, u64 ns, u64 alloc_r, u64 alloc_s, int shrink,
                         u64 *tab_r, u64 *tab_s, u32 *d_flags);
u64 smj_msd_alloc2(u64 n);
int smj_msd_pass_a_opt(SmjCtx *ctx, const u64 *in, u64 n, u64 *a, u32 shift1, u64 *d_tab, u32 *d_flag,
                       u32 kmask, u32 kref, u32 *d_kflag);
int smj_msd_pass_b_opt(SmjCtx *ctx, const u64 *a, u64 n, const u64 *d_tab, u64 *b, u64 alloc2, u32 shift2,
                       u32 dense_slot, u64 *starts, u64 *ends, u64 *d_counts, u32 *d_flag, u64 *d_nfixed,
                       int shrink);
int smj_msd_counts_opt(SmjCtx *ctx, const u64 *d_tab, u64 *d_counts);
int smj_msd_pass_b(SmjCtx *ctx, const u64 *a, u64 n, const u64 *d_counts, const u64 *d_starts, u64 *b,
                   u32 shift2, u64 *offsets);
int smj_exclusive_scan_u64(SmjCtx *ctx, const u64 *in, u64 n, u64 *out);

namespace {

constexpr u32 BS_MAXRUN  = 64;        // longest run of equal keys ordered in place
constexpr u64 BS_MAXBUCKET = 1u << 20;  // larger buckets (skew): leave the input to the LSD sort
constexpr u32 NB2 = 512u * 512u;

// bucket sizes, and the largest one
__global__ void __launch_bounds__(256) k_bucket_sizes(const u64 *__restrict__ starts,
                                                      const u64 *__restrict__ ends, u32 nbuckets,
                                                      u64 *__restrict__ sizes, u64 *__restrict__ largest) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    u64       c = 0;
    if (b < nbuckets) {
        c        = ends[b] - starts[b];
        sizes[b] = c;
    }
    c = max(c, __shfl_xor_sync(SMJ_FULL_MASK, c, 16));
    c = max(c, __shfl_xor_sync(SMJ_FULL_MASK, c, 8));
    c = max(c, __shfl_xor_sync(SMJ_FULL_MASK, c, 4));
    c = max(c, __shfl_xor_sync(SMJ_FULL_MASK, c, 2));
    c = max(c, __shfl_xor_sync(SMJ_FULL_MASK, c, 1));
    if (lane_id() == 0 && c > ld_relaxed(largest)) atomicMax(largest, c);
}

// One CTA per bucket (static round-robin: the buckets are close to equal in size): counting sort
// of the bucket over the 2^lbits key values left inside it.  A bucket of up to BS_THREADS * IPT
// items stays in registers between the counting and the placing pass and is put in order in a
// shared-memory staging area, so that it is read once and written with coalesced stores; a larger
// bucket (skew) is read twice and placed directly in global memory.  The loop is software
// pipelined: the bounds of the next bucket are requested when the current one starts, its items
// as soon as the registers are free (before the current bucket is written out).  Runs of equal
// keys are ordered by payload (insertion sort by the thread that owns the key slot); a run longer
// than BS_MAXRUN raises *flag.
template <int BS_THREADS, int IPT>
__global__ void __launch_bounds__(BS_THREADS) k_bucket_sort(const u64 *__restrict__ in,
                                                            const u64 *__restrict__ starts,
                                                            const u64 *__restrict__ ends,
                                                            const u64 *__restrict__ outoff, u32 nbuckets,
                                                            u32 lo_shift, u32 lbits, u64 *__restrict__ out,
                                                            u32 *flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr u32 CAP = BS_THREADS * IPT;
    const u32 slots = 1u << lbits, mask = slots - 1;
    u64 *stage = (u64 *)smem_raw;       // [CAP] the bucket in order
    u32 *cnt   = (u32 *)(stage + CAP);  // [slots + slots / per] counts, then cursors (= run ends after placing)
    __shared__ u32 s_w[BS_THREADS / 32];
    const u32 per = (slots + BS_THREADS - 1) / BS_THREADS;  // consecutive slots per thread (a power of two)
    // A thread scans its `per` consecutive slots: neighbouring lanes would be `per` words apart (per = 8:
    // eight lanes on one bank, and the scan is two thirds of the kernel's shared-memory wavefronts).
    // One word of padding per thread's group puts them per + 1 apart: slot sl lives at CNT(sl).
    // (measured: 10^9 items, per = 8: 6.14 -> 4.77 ms; with per = 4 the extra index arithmetic costs
    // more than the 4-way conflicts, 0.69 -> 0.76 ms per 128 M items, so small groups stay unpadded)
    const u32 pshift = per >= 8 ? 31u - __clz(per) : 31u;
#define CNT(sl) cnt[(sl) + ((sl) >> pshift)]
    const u32 cnt_words = slots + (per >= 8 ? slots / per : 0u);
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 b = blockIdx.x;
    if (b >= nbuckets) return;
    u64 s0 = starts[b], c = ends[b] - s0, o0 = outoff[b];
    u64 x[IPT];
    auto load_items = [&](u64 from, u64 count) {
#pragma unroll
        for (int j = 0; j < IPT; j++)
            x[j] = ((u64)j * BS_THREADS + threadIdx.x < count) ? ld_stream(in + from + j * BS_THREADS + threadIdx.x)
                                                                : 0ull;
    };
    if (c <= (u64)CAP) load_items(s0, c);
    for (u32 i = threadIdx.x; i < cnt_words; i += BS_THREADS) cnt[i] = 0;
    while (true) {
        // bounds of the next bucket: requested now, needed when this one is done
        const u32 nb = b + gridDim.x;
        u64 ns0 = 0, nc = 0, no0 = 0;
        if (nb < nbuckets) {
            ns0 = starts[nb];
            nc  = ends[nb] - ns0;
            no0 = outoff[nb];
        }
        const bool small = c <= (u64)CAP;
        __syncthreads();  // cnt is zero, the staging area is free
        if (small) {
#pragma unroll
            for (int j = 0; j < IPT; j++)
                if ((u64)j * BS_THREADS + threadIdx.x < c) atomicAdd(&CNT((u32)(x[j] >> lo_shift) & mask), 1u);
        } else {
            for (u64 i = threadIdx.x; i < c; i += BS_THREADS) atomicAdd(&CNT((u32)(in[s0 + i] >> lo_shift) & mask), 1u);
        }
        __syncthreads();
        // exclusive scan over the slots: thread t owns slots [t * per, t * per + per)
        u32 sum = 0, tie = 0;
        for (u32 q = 0; q < per; q++) {
            const u32 sl = threadIdx.x * per + q;
            if (sl < slots) {
                const u32 v = CNT(sl);
                sum += v;
                tie |= v > 1;
            }
        }
        u32 incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
            if (lane >= (u32)o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        const int any_tie = __syncthreads_or((int)tie);
        u32 excl = incl - sum;
        for (u32 w = 0; w < warp; w++) excl += s_w[w];
        for (u32 q = 0; q < per; q++) {
            const u32 sl = threadIdx.x * per + q;
            if (sl < slots) {
                const u32 v = CNT(sl);
                CNT(sl)     = excl;
                excl += v;
            }
        }
        __syncthreads();
        u64 *dst = small ? stage : out + o0;  // where the bucket is put in order
        if (small) {
#pragma unroll
            for (int j = 0; j < IPT; j++)
                if ((u64)j * BS_THREADS + threadIdx.x < c)
                    stage[atomicAdd(&CNT((u32)(x[j] >> lo_shift) & mask), 1u)] = x[j];
        } else {
            for (u64 i = threadIdx.x; i < c; i += BS_THREADS) {
                const u64 y = in[s0 + i];
                out[o0 + atomicAdd(&CNT((u32)(y >> lo_shift) & mask), 1u)] = y;
            }
        }
        // the registers are free: the next bucket's items travel while this one is finished
        if (nb < nbuckets && nc <= (u64)CAP) load_items(ns0, nc);
        __syncthreads();  // the bucket is in place (this CTA's stores are visible to it)
        if (any_tie) {
            // cnt[sl] is now the END of slot sl's run; it starts where the previous slot's ends
            for (u32 q = 0; q < per; q++) {
                const u32 sl = threadIdx.x * per + q;
                if (sl >= slots) break;
                const u32 end = CNT(sl), beg = sl ? CNT(sl - 1) : 0u, l = end - beg;
                if (l < 2) continue;
                if (l > BS_MAXRUN) {
                    *flag = 1u;
                    continue;
                }
                u64 *run = dst + beg;
                for (u32 i = 1; i < l; i++) {  // insertion sort on the whole word (the keys are equal)
                    const u64 v = run[i];
                    u32       j = i;
                    while (j > 0 && run[j - 1] > v) {
                        run[j] = run[j - 1];
                        j--;
                    }
                    run[j] = v;
                }
            }
            __syncthreads();
        }
        for (u32 i = threadIdx.x; i < cnt_words; i += BS_THREADS) cnt[i] = 0;  // for the next bucket
        if (small)
            for (u32 i = threadIdx.x; i < (u32)c; i += BS_THREADS) st_stream(out + o0 + i, stage[i]);
        if (nb >= nbuckets) break;
        b = nb, s0 = ns0, c = nc, o0 = no0;
    }
}
#undef CNT

template <int BS_THREADS, int IPT>
int launch_bucket_sort(SmjCtx *ctx, const u64 *b, const u64 *starts, const u64 *bend, const u64 *outoff,
                       u32 lo_shift, u32 lbits, u64 *out, u32 *flag, u64 n) {
    const size_t smem = (size_t)BS_THREADS * IPT * 8 + ((size_t)(1u << lbits) + BS_THREADS) * 4;
    if (smj_func_smem(ctx, (const void *)k_bucket_sort<BS_THREADS, IPT>, smem)) return -1;
    int per_sm = 1;  // resident CTAs per SM: the grid is exactly one wave (static bucket assignment)
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bucket_sort<BS_THREADS, IPT>, BS_THREADS, smem) !=
            cudaSuccess ||
        per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    SMJ_ITEMS(ctx, n);
    SMJ_LAUNCH(ctx, (k_bucket_sort<BS_THREADS, IPT>), ctx->sm_count * per_sm, BS_THREADS, smem, b, starts, bend, outoff, NB2,
               lo_shift, lbits, out, flag);
    return 0;
}

}  // namespace

// Whole-word MSD sort of in[0..n) into out[0..n) (out must not overlap in).  *used = 0 when the
// input does not suit the path (fewer than 40 M items, fewer than 19 or more than 31 varying key
// bits, skew, long runs of equal keys, a sampling miss): nothing useful is in out then and the
// caller sorts with smj_lsd_sort.  Synchronises.
int smj_msd_sort(SmjCtx *ctx, const u64 *in, u64 n, u64 *out, int *used) {
    *used = 0;
    // 2^18 buckets pay off from ~150 items per bucket on (measured: 16 M items 0.77 ms against
    // 0.41 ms for the LSD sort, 128 M items 1.89 against 3.20 ms, 10^9 items 14.2 against 25.1 ms)
    const char *mode_env = getenv("SMJ_SORT_PATH");  // "lsd" disables this path, "msd" lowers the threshold
    if (mode_env && mode_env[0] == 'l') return 0;
    if (n < ((mode_env && mode_env[0] == 'm') ? (1ull << 21) : 40000000ull)) return 0;
    const char *hook    = getenv("SMJ_MSD_TEST");  // test hook: under-sized slots -> repair kernels
    const int   shrink2 = hook && !strcmp(hook, "shrink2");
    SmjScratch sc(ctx);
    u64 *d_st = (u64 *)sc.get(smj_msd_sample_bytes(0));
    u32 *d_sub = (u32 *)sc.get(smj_msd_sample_bytes(1)), *d_flags = (u32 *)sc.get(64);
    u64 *d_misc = (u64 *)sc.get(64);  // [0] largest bucket, [1] repaired first-level buckets
    if (!d_st || !d_sub || !d_flags || !d_misc) return smj_fail_msg("smj_msd_sort: scratch allocation failed");
    SMJ_CHECK(cudaMemsetAsync(d_flags, 0, 64, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_misc, 0, 64, ctx->stream));
    std::vector<u64> samp(SMJ_MSD_SAMPLE_WORDS);
    int rc = smj_msd_sample(ctx, in, n, nullptr, 0, 1u << 18, d_st, d_sub, samp.data());
    if (rc == 1) return 0;  // no cooperative launch on this device
    if (rc) return rc;
    MsdPlan pl;
    if (!smj_msd_plan(samp.data(), &pl)) return 0;
    const u32 krange = (pl.hi == 31 ? 0xffffffffu : ((1u << (pl.hi + 1)) - 1)) & ~((1u << pl.lo) - 1);
    const u32 kmask = ~krange, kref = (u32)(samp[1] >> 32) & kmask;
    const u64 nsamp = samp[8], dups = samp[6];
    const bool slots = dups <= 16;  // equal second-level slots only without duplicate keys
    const u64  alloc1 = smj_msd_caps1(samp.data() + 16, nsamp, n, 0), alloc2 = slots ? smj_msd_alloc2(n) : n;
    u64 *a = (u64 *)sc.get(std::max(alloc1, n) * 8), *b = (u64 *)sc.get(std::max(alloc2, n) * 8);
    u64 *starts = (u64 *)sc.get((size_t)(NB2 + 1) * 8), *ends = (u64 *)sc.get((size_t)NB2 * 8);
    u64 *sizes = (u64 *)sc.get((size_t)NB2 * 8), *outoff = (u64 *)sc.get((size_t)NB2 * 8);
    u64 *d_tab = (u64 *)sc.get(3 * 512 * 8), *d_counts = (u64 *)sc.get(512 * 8);
    if (!a || !b || !starts || !ends || !sizes || !outoff || !d_tab || !d_counts)
        return smj_fail_msg("smj_msd_sort: out of device memory");
    // flags: [0] level-1 overflow / regions do not fit, [2] key outside the sampled range,
    //        [4] second-level slots do not fit, [6] run of equal keys too long
    if ((rc = smj_msd_caps1_tables(ctx, d_st, n, 0, alloc1, 0, 0, d_tab, nullptr, d_flags))) return rc;
    if ((rc = smj_msd_pass_a_opt(ctx, in, n, a, pl.shift1, d_tab, d_flags, kmask, kref, d_flags + 2))) return rc;
    const u64 *bend = ends;
    if (slots) {
        if ((rc = smj_msd_pass_b_opt(ctx, a, n, d_tab, b, alloc2, pl.shift2, 1u << pl.lbits, starts, ends, d_counts, d_flags + 4,
                                     d_misc + 1, shrink2)))
            return rc;
    } else {
        if ((rc = smj_msd_counts_opt(ctx, d_tab, d_counts))) return rc;
        if ((rc = smj_msd_pass_b(ctx, a, n, d_counts, d_tab + 512, b, pl.shift2, starts))) return rc;
        bend = starts + 1;
    }
    SMJ_LAUNCH(ctx, k_bucket_sizes, NB2 / 256, 256, 0, starts, bend, NB2, sizes, d_misc);
    if ((rc = smj_exclusive_scan_u64(ctx, sizes, NB2, outoff))) return rc;
    // register-resident buckets: 128 threads x 8 items (buckets up to 1024 items) when the average
    // bucket is small, 512 x 10 (up to 5120) otherwise; larger buckets take the kernel's slow path
    if (n / NB2 <= 700)
        rc = launch_bucket_sort<128, 8>(ctx, b, starts, bend, outoff, pl.lo_shift, pl.lbits, out, d_flags + 6, n);
    else
        rc = launch_bucket_sort<512, 10>(ctx, b, starts, bend, outoff, pl.lo_shift, pl.lbits, out, d_flags + 6, n);
    if (rc) return rc;
    u32 hf[16];
    u64 hm[8];
    SMJ_CHECK(cudaMemcpyAsync(hf, d_flags, 64, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaMemcpyAsync(hm, d_misc, 64, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    if (hf[0] || hf[2] || hf[4] || hf[6] || hm[0] > BS_MAXBUCKET) return 0;
    *used = 1;
    return 0;
}
