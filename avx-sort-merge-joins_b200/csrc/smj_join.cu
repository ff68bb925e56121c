// smj_join.cu -- merge-join scan and the three device-resident join pipelines.
//
// B200 restatement of
//   * src/joins/joincommon.c:239-312 merge_join: a sequential two-cursor scan that counts the
//     cross product of every pair of equal-key runs.
//   * src/joins/sortmergejoin_multiway.c:129-328, sortmergejoin_multipass.c:137-292 and the
//     (stubbed, sortmergejoin_mpsm.c:38-45) MPSM driver.
//
// Scan design: the merged order of R and S (by key, R first on ties) is cut into tiles of 4096
// items (k_mj_partition, merge path).  A CTA loads its two pieces into shared memory, builds one
// prefix-sum array of the checksum terms, and every thread that owns the FIRST S item of a key
// run finds the matching R run with a shared-memory binary search:
//     count    += |run_R| * |run_S|                       (what the reference counts one by one)
//     checksum += (sum A(r)) * (sum B(s))                 (bilinear, so no pair enumeration)
// Key runs that may continue in a neighbouring tile (the tile's first and last key) are not
// multiplied in the tile; their partial sums go to a record array that k_mj_fixup stitches.
// Work per key run is O(log tile), so Zipf-skewed inputs cost the same as uniform ones.
//
// HBM traffic: 8 B per input tuple (read only) + 80 B of records per 4096 tuples.
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <string.h>
#include <vector>

int smj_lsd_sort(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, u64 **sorted);
int smj_multiway_merge(SmjCtx *ctx, u64 *out, u64 *tmp, const u64 *const *runs, const u64 *lens,
                       u32 nruns, int mode);

namespace {

constexpr int MJ_THREADS = 512;
constexpr int MJ_IPT     = 8;
constexpr int MJ_TILE    = MJ_THREADS * MJ_IPT;                 // 4096 merged items per CTA
constexpr size_t MJ_SMEM = (size_t)MJ_TILE * 8 + (size_t)(MJ_TILE + 2) * 8;

struct MjRec {
    u64 key;
    u64 cr, cs;  // items of this key seen in R / S
    u64 sa, sb;  // sum of A(r) / B(s) over them
};

__device__ __forceinline__ u32 key_of(u64 x) { return (u32)(x >> 32); }

// first index in [lo,hi) whose key is >= k
__device__ __forceinline__ u32 lower_key(const u64 *a, u32 lo, u32 hi, u32 k) {
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (key_of(a[mid]) < k)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}
// first index in [lo,hi) whose key is > k
__device__ __forceinline__ u32 upper_key(const u64 *a, u32 lo, u32 hi, u32 k) {
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (key_of(a[mid]) <= k)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// merge path on keys, R before S on ties: number of R items among the first diag merged items
__global__ void __launch_bounds__(256) k_mj_partition(const u64 *__restrict__ r, u64 nr,
                                                      const u64 *__restrict__ s, u64 ns, u64 tiles,
                                                      u64 *__restrict__ diag_r) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > tiles) return;
    u64 diag = min(t * (u64)MJ_TILE, nr + ns);
    u64 lo   = diag > ns ? diag - ns : 0;
    u64 hi   = diag < nr ? diag : nr;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if (key_of(r[mid]) <= key_of(s[diag - 1 - mid]))
            lo = mid + 1;
        else
            hi = mid;
    }
    diag_r[t] = lo;
}

__global__ void __launch_bounds__(MJ_THREADS) k_mj_tile(const u64 *__restrict__ r, u64 nr,
                                                        const u64 *__restrict__ s, u64 ns,
                                                        const u64 *__restrict__ diag_r,
                                                        MjRec *__restrict__ recs,
                                                        u64 *__restrict__ acc, int semi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_it = (u64 *)smem_raw;   // [MJ_TILE]  R piece followed by S piece
    u64 *s_pf = s_it + MJ_TILE;    // [MJ_TILE+1] exclusive prefix sums of A(r) ... B(s)
    __shared__ u64 s_wsum[MJ_THREADS / 32];
    __shared__ u64 s_red[2][MJ_THREADS / 32];

    const u64 t  = blockIdx.x;
    const u64 d0 = t * (u64)MJ_TILE;
    const u64 d1 = min(d0 + (u64)MJ_TILE, nr + ns);
    const u64 r0 = diag_r[t], r1 = diag_r[t + 1];
    const u64 s0 = d0 - r0, s1 = d1 - r1;
    const u32 la = (u32)(r1 - r0), lb = (u32)(s1 - s0), cnt = la + lb;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (u32 i = threadIdx.x; i < cnt; i += MJ_THREADS)
        s_it[i] = (i < la) ? ld_stream(r + r0 + i) : ld_stream(s + s0 + (i - la));
    __syncthreads();
    const u64 *sR = s_it, *sS = s_it + la;

    // ---- one exclusive scan over [A(r_0..la-1), B(s_0..lb-1)]
    {
        u64       v[MJ_IPT];
        u64       ts   = 0;
        const u32 base = threadIdx.x * MJ_IPT;
#pragma unroll
        for (int k = 0; k < MJ_IPT; k++) {
            u32 idx = base + k;
            v[k]    = (idx < la) ? smj_cks_a(s_it[idx]) : (idx < cnt ? smj_cks_b(s_it[idx]) : 0ull);
            ts += v[k];
        }
        u64 incl = ts;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u64 x = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
            if (lane >= (u32)o) incl += x;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        u64 run = incl - ts;
        for (u32 w = 0; w < warp; w++) run += s_wsum[w];
#pragma unroll
        for (int k = 0; k < MJ_IPT; k++) {
            s_pf[base + k] = run;
            run += v[k];
        }
        if (threadIdx.x == MJ_THREADS - 1) s_pf[MJ_TILE] = run;
    }
    __syncthreads();

    // ---- keys that may continue in a neighbouring tile
    const u32 kf = min(la ? key_of(sR[0]) : 0xffffffffu, lb ? key_of(sS[0]) : 0xffffffffu);
    const u32 kl = max(la ? key_of(sR[la - 1]) : 0u, lb ? key_of(sS[lb - 1]) : 0u);

    u64 my_cnt = 0, my_cks = 0;
    for (u32 j = threadIdx.x; j < lb; j += MJ_THREADS) {
        const u32 k = key_of(sS[j]);
        if (j > 0 && key_of(sS[j - 1]) == k) continue;  // not the head of its S run
        if (k == kf || k == kl) continue;               // stitched by k_mj_fixup
        u32 lo = lower_key(sR, 0, la, k);
        if (lo >= la || key_of(sR[lo]) != k) continue;  // no partner in R
        u32 hi = lo + 1;
        if (hi < la && key_of(sR[hi]) == k) hi = upper_key(sR, hi, la, k);
        u32 j2 = j + 1;
        if (j2 < lb && key_of(sS[j2]) == k) j2 = upper_key(sS, j2, lb, k);
        my_cnt += semi ? (u64)(j2 - j) : (u64)(hi - lo) * (u64)(j2 - j);
        my_cks += (s_pf[hi] - s_pf[lo]) * (s_pf[la + j2] - s_pf[la + j]);
    }

    if (threadIdx.x == 0) {
        u32   hr = upper_key(sR, 0, la, kf), hs = upper_key(sS, 0, lb, kf);
        MjRec rec;
        rec.key = kf;
        rec.cr  = hr;
        rec.cs  = hs;
        rec.sa  = s_pf[hr] - s_pf[0];
        rec.sb  = s_pf[la + hs] - s_pf[la];
        recs[2 * t] = rec;
    }
    if (threadIdx.x == 32) {
        MjRec rec;
        rec.key = kl;
        if (kl != kf) {
            u32 lr = lower_key(sR, 0, la, kl), ls = lower_key(sS, 0, lb, kl);
            rec.cr = la - lr;
            rec.cs = lb - ls;
            rec.sa = s_pf[la] - s_pf[lr];
            rec.sb = s_pf[la + lb] - s_pf[la + ls];
        } else {
            rec.cr = rec.cs = rec.sa = rec.sb = 0;
        }
        recs[2 * t + 1] = rec;
    }

    my_cnt = warp_reduce_add(my_cnt);
    my_cks = warp_reduce_add(my_cks);
    if (lane == 0) {
        s_red[0][warp] = my_cnt;
        s_red[1][warp] = my_cks;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 c = 0, x = 0;
        for (int w = 0; w < MJ_THREADS / 32; w++) {
            c += s_red[0][w];
            x += s_red[1][w];
        }
        if (c) atomicAdd(acc, c);
        if (x) atomicAdd(acc + 1, x);
    }
}

// Stitch the per-tile boundary records: consecutive records with the same key form one run.
constexpr int FX_THREADS = 512;
__global__ void __launch_bounds__(FX_THREADS) k_mj_fixup(const MjRec *__restrict__ recs, u64 nrec,
                                                         u64 *__restrict__ acc, int semi) {
    __shared__ MjRec         s_open[2 * FX_THREADS];
    __shared__ unsigned char s_valid[2 * FX_THREADS];
    __shared__ u64           s_red[2][FX_THREADS / 32];
    const u64 chunk = (nrec + FX_THREADS - 1) / FX_THREADS;
    const u64 lo    = (u64)threadIdx.x * chunk;
    const u64 hi    = min(nrec, lo + chunk);
    u64       my_cnt = 0, my_cks = 0;
    s_valid[2 * threadIdx.x]     = 0;
    s_valid[2 * threadIdx.x + 1] = 0;
    if (lo < hi) {
        MjRec cur         = recs[lo];
        bool  first_saved = false;
        for (u64 i = lo + 1; i < hi; i++) {
            MjRec x = recs[i];
            if (x.key == cur.key) {
                cur.cr += x.cr;
                cur.cs += x.cs;
                cur.sa += x.sa;
                cur.sb += x.sb;
            } else {
                if (!first_saved) {  // may continue a run of the previous chunk
                    s_open[2 * threadIdx.x]  = cur;
                    s_valid[2 * threadIdx.x] = 1;
                    first_saved              = true;
                } else {
                    my_cnt += semi ? (cur.cr ? cur.cs : 0ull) : cur.cr * cur.cs;
                    my_cks += cur.sa * cur.sb;
                }
                cur = x;
            }
        }
        const int slot = first_saved ? 1 : 0;  // may continue in the next chunk
        s_open[2 * threadIdx.x + slot]  = cur;
        s_valid[2 * threadIdx.x + slot] = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool  have = false;
        MjRec cur;
        cur.key = cur.cr = cur.cs = cur.sa = cur.sb = 0;
        for (int i = 0; i < 2 * FX_THREADS; i++) {
            if (!s_valid[i]) continue;
            MjRec x = s_open[i];
            if (have && x.key == cur.key) {
                cur.cr += x.cr;
                cur.cs += x.cs;
                cur.sa += x.sa;
                cur.sb += x.sb;
            } else {
                if (have) {
                    my_cnt += semi ? (cur.cr ? cur.cs : 0ull) : cur.cr * cur.cs;
                    my_cks += cur.sa * cur.sb;
                }
                cur  = x;
                have = true;
            }
        }
        if (have) {
            my_cnt += semi ? (cur.cr ? cur.cs : 0ull) : cur.cr * cur.cs;
            my_cks += cur.sa * cur.sb;
        }
    }
    my_cnt = warp_reduce_add(my_cnt);
    my_cks = warp_reduce_add(my_cks);
    if (lane_id() == 0) {
        s_red[0][threadIdx.x >> 5] = my_cnt;
        s_red[1][threadIdx.x >> 5] = my_cks;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 c = 0, x = 0;
        for (int w = 0; w < FX_THREADS / 32; w++) {
            c += s_red[0][w];
            x += s_red[1][w];
        }
        if (c) atomicAdd(acc, c);
        if (x) atomicAdd(acc + 1, x);
    }
}

}  // namespace

// accumulate count / checksum of R join S into the device pair acc[0..1] (asynchronous)
// semi != 0: count the S tuples that have a partner in R (merge_join_interpolation semantics);
// the checksum is only defined for semi == 0
int smj_merge_join_acc(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, u64 *acc,
                       int semi = 0) {
    if (nr == 0 || ns == 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        SMJ_CHECK(cudaFuncSetAttribute(k_mj_tile, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)MJ_SMEM));
        attr_set = true;
    }
    const u64  tiles = div_up(nr + ns, MJ_TILE);
    SmjScratch sc(ctx);
    u64   *d_diag = (u64 *)sc.get((tiles + 1) * 8);
    MjRec *d_recs = (MjRec *)sc.get(tiles * 2 * sizeof(MjRec));
    if (!d_diag || !d_recs) return smj_fail_msg("merge_join: scratch allocation failed");
    SMJ_LAUNCH(ctx, k_mj_partition, (unsigned)div_up(tiles + 1, 256), 256, 0, r, nr, s, ns, tiles,
               d_diag);
    SMJ_ITEMS(ctx, nr + ns);
    SMJ_LAUNCH(ctx, k_mj_tile, (unsigned)tiles, MJ_THREADS, MJ_SMEM, r, nr, s, ns, d_diag, d_recs,
               acc, semi);
    SMJ_LAUNCH(ctx, k_mj_fixup, 1, FX_THREADS, 0, d_recs, tiles * 2, acc, semi);
    return 0;
}

extern "C" int smjdev_merge_join_acc(const uint64_t *r, uint64_t nr, const uint64_t *s,
                                     uint64_t ns, uint64_t *acc) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    return smj_merge_join_acc(ctx, (const u64 *)r, nr, (const u64 *)s, ns, (u64 *)acc);
}

extern "C" int smjdev_merge_join_ex(const uint64_t *r, uint64_t nr, const uint64_t *s,
                                    uint64_t ns, int semi, uint64_t *count, uint64_t *checksum);
extern "C" int smjdev_merge_join(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                                 uint64_t *count, uint64_t *checksum) {
    return smjdev_merge_join_ex(r, nr, s, ns, 0, count, checksum);
}
extern "C" int smjdev_merge_join_ex(const uint64_t *r, uint64_t nr, const uint64_t *s,
                                    uint64_t ns, int semi, uint64_t *count, uint64_t *checksum) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    SmjScratch sc(ctx);
    u64 *acc = (u64 *)sc.get(16);
    if (!acc) return smj_fail_msg("merge_join: scratch allocation failed");
    SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    int rc = smj_merge_join_acc(ctx, (const u64 *)r, nr, (const u64 *)s, ns, acc, semi);
    if (rc) return rc;
    u64 h[2];
    SMJ_CHECK(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    if (count) *count = h[0];
    if (checksum) *checksum = h[1];
    return 0;
}

// ---------------------------------------------------------------------------------------
// join pipelines
// ---------------------------------------------------------------------------------------
namespace {
struct PhaseTimer {
    SmjCtx     *ctx;
    cudaEvent_t ev[8];
    int         n;
    explicit PhaseTimer(SmjCtx *c) : ctx(c), n(0) {
        for (auto &e : ev) e = nullptr;
    }
    ~PhaseTimer() {
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
    }
    int mark() {
        if (n >= 8) return -1;
        if (cudaEventCreate(&ev[n]) != cudaSuccess) return -1;
        cudaEventRecord(ev[n], ctx->stream);
        return n++;
    }
    float ms(int a, int b) {
        float v = 0.f;
        if (a < 0 || b < 0) return 0.f;
        cudaEventElapsedTime(&v, ev[a], ev[b]);
        return v;
    }
};

// split [0,n) into `runs` contiguous chunks like the reference splits relations over threads
// (src/joins/joincommon.c:105-139: n/T each, the last one takes the remainder)
void chunk_bounds(u64 n, int runs, std::vector<u64> &off) {
    off.resize(runs + 1);
    u64 per = n / runs;
    for (int i = 0; i < runs; i++) off[i] = per * i;
    off[runs] = n;
}

int sort_runs(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int runs,
              std::vector<const u64 *> &ptrs, std::vector<u64> &lens) {
    std::vector<u64> off;
    chunk_bounds(n, runs, off);
    ptrs.clear();
    lens.clear();
    for (int i = 0; i < runs; i++) {
        u64  len    = off[i + 1] - off[i];
        u64 *sorted = nullptr;
        int  rc = smj_lsd_sort(ctx, in + off[i], a + off[i], b + off[i], len, 1, &sorted);
        if (rc) return rc;
        ptrs.push_back(sorted);
        lens.push_back(len);
    }
    return 0;
}
}  // namespace

extern "C" int smjdev_join(int algo, const uint64_t *r_, uint64_t nr, const uint64_t *s_,
                           uint64_t ns, int nruns, smj_join_stats_t *stats) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (!stats) return smj_fail_msg("smjdev_join: stats must not be NULL");
    if (algo < SMJ_ALGO_MWAY || algo > SMJ_ALGO_MPSM) return smj_fail_msg("smjdev_join: bad algo");
    const u64 *r = (const u64 *)r_, *s = (const u64 *)s_;
    memset(stats, 0, sizeof(*stats));
    if (nruns < 1) nruns = 1;
    if ((u64)nruns > nr && nr) nruns = (int)nr;
    if ((u64)nruns > ns && ns) nruns = (int)ns;
    const u64 launches0 = ctx->launches;

    SmjScratch sc(ctx);
    u64 *acc = (u64 *)sc.get(16);
    if (!acc) return smj_fail_msg("smjdev_join: scratch allocation failed");
    PhaseTimer tm(ctx);
    const int  t0 = tm.mark();
    SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    int t_part = t0, t_sort = t0, t_merge = t0, t_join = t0, rc = 0;

    if (nr && ns) {
        u64 *aR = (u64 *)sc.get(nr * 8), *bR = (u64 *)sc.get(nr * 8);
        u64 *aS = (u64 *)sc.get(ns * 8), *bS = (u64 *)sc.get(ns * 8);
        if (!aR || !bR || !aS || !bS) return smj_fail_msg("smjdev_join: out of device memory");
        if (algo == SMJ_ALGO_MWAY) {
            // range partition + in-partition sort == a global key sort on one GPU: the sorted
            // partitions are disjoint key ranges, so no run merging is needed
            // (the reference's nthreads == 1 path, sortmergejoin_multiway.c:476-489)
            u64 *sr = nullptr, *ss = nullptr;
            t_part = tm.mark();
            if ((rc = smj_lsd_sort(ctx, r, aR, bR, nr, 1, &sr))) return rc;
            if ((rc = smj_lsd_sort(ctx, s, aS, bS, ns, 1, &ss))) return rc;
            t_sort = t_merge = tm.mark();
            if ((rc = smj_merge_join_acc(ctx, sr, nr, ss, ns, acc))) return rc;
            t_join = tm.mark();
        } else if (algo == SMJ_ALGO_MPASS) {
            // sorted runs, then pairwise merge passes (sortmergejoin_multipass.c:410-708)
            std::vector<const u64 *> pr, ps;
            std::vector<u64>         lr, ls;
            t_part = tm.mark();
            if ((rc = sort_runs(ctx, r, aR, bR, nr, nruns, pr, lr))) return rc;
            if ((rc = sort_runs(ctx, s, aS, bS, ns, nruns, ps, ls))) return rc;
            t_sort = tm.mark();
            u64 *cR = (u64 *)sc.get(nr * 8), *dR = (u64 *)sc.get(nr * 8);
            u64 *cS = (u64 *)sc.get(ns * 8), *dS = (u64 *)sc.get(ns * 8);
            if (!cR || !dR || !cS || !dS) return smj_fail_msg("smjdev_join: out of device memory");
            if ((rc = smj_multiway_merge(ctx, cR, dR, pr.data(), lr.data(), (u32)pr.size(), 1)))
                return rc;
            if ((rc = smj_multiway_merge(ctx, cS, dS, ps.data(), ls.data(), (u32)ps.size(), 1)))
                return rc;
            t_merge = tm.mark();
            if ((rc = smj_merge_join_acc(ctx, cR, nr, cS, ns, acc))) return rc;
            t_join = tm.mark();
        } else {
            // P-MPSM: the private input R is range partitioned and sorted (== globally sorted),
            // the public input S is only sorted run by run; every S run is merge-joined against R
            u64 *sr = nullptr;
            std::vector<const u64 *> ps;
            std::vector<u64>         ls;
            t_part = tm.mark();
            if ((rc = smj_lsd_sort(ctx, r, aR, bR, nr, 1, &sr))) return rc;
            if ((rc = sort_runs(ctx, s, aS, bS, ns, nruns, ps, ls))) return rc;
            t_sort = t_merge = tm.mark();
            for (size_t i = 0; i < ps.size(); i++)
                if ((rc = smj_merge_join_acc(ctx, sr, nr, ps[i], ls[i], acc))) return rc;
            t_join = tm.mark();
        }
    } else {
        t_part = t_sort = t_merge = t_join = tm.mark();
    }
    u64 h[2];
    SMJ_CHECK(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    const int t_end = tm.mark();
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    stats->matches      = h[0];
    stats->checksum     = h[1];
    stats->ms_partition = tm.ms(t0, t_part);
    stats->ms_sort      = tm.ms(t_part, t_sort);
    stats->ms_merge     = tm.ms(t_sort, t_merge);
    stats->ms_join      = tm.ms(t_merge, t_join);
    stats->ms_total     = tm.ms(t0, t_end);
    stats->launches     = ctx->launches - launches0;
    stats->hbm_bytes    = sc.bytes;
    return 0;
}
