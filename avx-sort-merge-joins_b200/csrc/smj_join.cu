// smj_join.cu -- merge-join scan and the three device-resident join pipelines.
//
// B200 restatement of
//   * src/joins/joincommon.c:239-312 merge_join: a sequential two-cursor scan that counts the
//     cross product of every pair of equal-key runs.
//   * src/joins/sortmergejoin_multiway.c:129-328, sortmergejoin_multipass.c:137-292 and the
//     (stubbed, sortmergejoin_mpsm.c:38-45) MPSM driver.
//
// Scan design: the merged order of R and S (by key, R first on ties) is cut into tiles of 4096
// items (k_mj_partition, merge path).  A CTA loads its two pieces into shared memory; every thread
// walks its own 8-item slice of the tile's merge path, and the thread that meets the FIRST S item
// of a key run combines it with the R run that ends at its R cursor:
//     count    += |run_R| * |run_S|                       (what the reference counts one by one)
//     checksum += (sum A(r)) * (sum B(s))                 (bilinear, so no pair enumeration)
// Key runs that may continue in a neighbouring tile (the tile's first and last key) are not
// multiplied in the tile; their partial sums go to a record array that k_mj_fixup stitches.
// Work per key run is O(log tile), so Zipf-skewed inputs cost the same as uniform ones.
//
// HBM traffic: 8 B per input tuple (read only); boundary records only where a run really crosses.
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "smj_internal.cuh"

namespace {

constexpr int MJ_THREADS = 256;
constexpr int MJ_IPT     = 18;                                   // see k_mj_tile
constexpr int MJ_TILE    = MJ_THREADS * MJ_IPT;                 // 4608 merged items per CTA

struct MjRec {
    u64 key;
    u64 cr, cs;  // items of this key seen in R / S
    u64 sa, sb;  // sum of A(r) / B(s) over them
};

__device__ __forceinline__ u32 key_of(u64 x) { return (u32)(x >> 32); }

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

// One tile of the merged order (R piece and S piece of the merge path, 4096 items together).
//
// Keys whose runs may continue in a neighbouring tile (the tile's first / last key) are taken out
// first: they are a prefix / suffix of the sorted pieces, four warps add them up into the boundary
// records for k_mj_fixup, and the S piece the walk sees is SHORTENED by them, so that the walk needs
// no per-item test for them.
//
// The walk: every thread finds its 18-item slice of the merge path (R before S on ties) and merges
// with two register-cached heads, one shared-memory load per step.  It carries the R RUN that ends
// at its R cursor -- {key, items, sum of A} -- which is all an S item needs: an S item of that key
// adds run_items to the count and run_sumA * B(s) to the checksum (sum over S items of a key =
// sumA * sumB).  No head-of-run detection, no searches, duplicates on either side cost nothing
// extra.  A thread that starts inside an R run first walks back to the run's start (one step for
// unique keys).
// The two pieces are contiguous ranges of R and of S: ONE thread stages them with two bulk copies
// (TMA, completion on an mbarrier; an LDG + STS per item and thread was a sixth of the kernel's
// instructions).  The layout in shared memory is therefore plain, and bank conflicts are avoided by
// the number of items per thread: with 18, the cursors of neighbouring threads are ~9 items apart in
// either piece and fall on different banks (with 16 they met in two banks: round 1, 146 M of 241 M
// wavefronts were conflicts).  An EVEN number also keeps a warp convergent on primary-key joins,
// where R and S items alternate in the merged order and every slice then starts with an R item
// (17 items per thread: the lanes disagree at every step and the walk costs 76 instead of 52
// instructions per step).  Measured, 128 M x 128 M / 32 M x 320 M: 16 items, padded layout, plain
// loads 0.92 / 1.55 ms; 17 items + TMA 1.00 / 1.44; 18 items + TMA 0.82 / 1.41.

__global__ void __launch_bounds__(MJ_THREADS, 6) k_mj_tile(const u64 *__restrict__ r, u64 nr,
                                                           const u64 *__restrict__ s, u64 ns,
                                                           const u64 *__restrict__ diag_r,
                                                           MjRec *__restrict__ recs,
                                                           unsigned char *__restrict__ rec_valid,
                                                           u64 *__restrict__ acc, int semi) {
    __shared__ __align__(16) u64 s_it[MJ_TILE + 4];  // R piece at [offR ..), S piece at [offS ..)
    __shared__ __align__(8) u64 s_bar;
    __shared__ u64 s_red[2][MJ_THREADS / 32];
    __shared__ u64 s_edge[4][2];  // {items, sum} of: R prefix, S prefix (key kf), R suffix, S suffix (key kl)

    const u64 t  = blockIdx.x;
    const u64 d0 = t * (u64)MJ_TILE;
    const u64 d1 = min(d0 + (u64)MJ_TILE, nr + ns);
    const u64 r0 = diag_r[t], r1 = diag_r[t + 1];
    const u64 s0 = d0 - r0, s1 = d1 - r1;
    const u32 la = (u32)(r1 - r0), lb = (u32)(s1 - s0);
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const u64 *gr = r + r0, *gs = s + s0;
    const u32 misR = la ? (u32)(((uintptr_t)gr >> 3) & 1) : 0u, misS = lb ? (u32)(((uintptr_t)gs >> 3) & 1) : 0u;
    const u32 evenS = (misR + la + 1) & ~1u;  // 16-byte aligned start of S's staging area
    const u32 offR = misR, offS = evenS + misS;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const BulkStage st_r(s_it, gr, la), st_s(s_it + evenS, gs, lb);
        const u32       bytes = st_r.bytes() + st_s.bytes();
        st_r.plain();
        st_s.plain();
        mbar_arrive_expect_tx(&s_bar, bytes);
        st_r.issue(&s_bar);
        st_s.issue(&s_bar);
    }
    mbar_wait(&s_bar, 0);
#define SR(i) s_it[offR + (i)]
#define SS(j) s_it[offS + (j)]

    // ---- the tile's first / last key: their runs may continue in the neighbouring tiles
    const u32 kf = min(la ? key_of(SR(0)) : 0xffffffffu, lb ? key_of(SS(0)) : 0xffffffffu);
    const u32 kl = max(la ? key_of(SR(la - 1)) : 0u, lb ? key_of(SS(lb - 1)) : 0u);
    const bool first_open = (r0 > 0 && key_of(r[r0 - 1]) == kf) || (s0 > 0 && key_of(s[s0 - 1]) == kf);
    const bool last_open  = (r1 < nr && key_of(r[r1]) == kl) || (s1 < ns && key_of(s[s1]) == kl);
    // a key whose run is not complete inside this tile is left to k_mj_fixup
    const bool ex_f = first_open || (kf == kl && last_open);
    const bool ex_l = (kl != kf) && last_open;
    if (ex_f | ex_l) {  // block-uniform; about every second tile when both sides have the key
        if (warp < 4 && ((warp < 2) ? ex_f : ex_l)) {
            const bool is_r = !(warp & 1), pre = warp < 2;
            const u32  len = is_r ? la : lb, base = is_r ? offR : offS, k = pre ? kf : kl;
            u64        c = 0, sum = 0;
            for (u32 i = lane; i < len; i += 32) {  // the run is a prefix / suffix of the sorted piece
                const u64 x = s_it[base + (pre ? i : len - 1 - i)];
                if (key_of(x) != k) break;
                c++;
                sum += is_r ? smj_cks_a(x) : smj_cks_b(x);
            }
            c   = warp_reduce_add(c);
            sum = warp_reduce_add(sum);
            if (lane == 0) {
                s_edge[warp][0] = c;
                s_edge[warp][1] = sum;
            }
        } else if (warp < 4 && lane == 0) {
            s_edge[warp][0] = 0;
            s_edge[warp][1] = 0;
        }
        __syncthreads();
    }
    // the S piece without its open runs
    const u32 sx  = ex_f ? (u32)s_edge[1][0] : 0u;
    const u32 lbw = lb - sx - (ex_l ? (u32)s_edge[3][0] : 0u);
    const u32 cw  = la + lbw;
#define SW(j) s_it[offS + sx + (j)]

    u64 my_cnt = 0, my_cks = 0;
    const u32 d = min((u32)(threadIdx.x * MJ_IPT), cw);
    u32       ai, bi;
    {  // merge path on keys inside the tile, R before S on ties
        u32 lo = d > lbw ? d - lbw : 0, hi = d < la ? d : la;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (key_of(SR(mid)) <= key_of(SW(d - 1 - mid)))
                lo = mid + 1;
            else
                hi = mid;
        }
        ai = lo;
        bi = d - lo;
    }
    // the R run that ends at the cursor
    u32 run_key = 0, run_cnt = 0;
    u64 run_sa = 0;
    if (ai > 0 && d < cw) {
        const u64 x = SR(ai - 1);
        run_key = key_of(x), run_cnt = 1, run_sa = smj_cks_a(x);
        for (u32 i = ai - 1; i > 0; i--) {
            const u64 y = SR(i - 1);
            if (key_of(y) != run_key) break;
            run_cnt++;
            run_sa += smj_cks_a(y);
        }
    }
    u64 a_it = (ai < la) ? SR(ai) : 0ull, b_it = (bi < lbw) ? SW(bi) : 0ull;
    const u32 steps = min((u32)MJ_IPT, cw - d);
#pragma unroll 1
    for (u32 step = 0; step < steps; step++) {
        // selects instead of an R branch and an S branch: the lanes of a warp disagree on take_r at
        // almost every step and would otherwise run both sides
        const bool take_r = (bi >= lbw) || (ai < la && key_of(a_it) <= key_of(b_it));
        const u64  x      = take_r ? a_it : b_it;
        const u32  k      = key_of(x);
        const bool same   = k == run_key;
        // advance the cursor that was taken and load its next head (one shared-memory load)
        ai += take_r ? 1u : 0u;
        bi += take_r ? 0u : 1u;
        const u32 nidx = take_r ? offR + ai : offS + sx + bi;
        const u32 nend = take_r ? offR + la : offS + sx + lbw;
        const u64 nxt  = (nidx < nend) ? s_it[nidx] : 0ull;
        if (take_r) {
            run_sa  = (same ? run_sa : 0ull) + smj_cks_a(x);
            run_cnt = same ? run_cnt + 1u : 1u;
            run_key = k;
            a_it    = nxt;
        } else {
            if (same) {
                my_cnt += semi ? (run_cnt ? 1ull : 0ull) : (u64)run_cnt;
                my_cks += run_sa * smj_cks_b(x);
            }
            b_it = nxt;
        }
    }
#undef SR
#undef SS
#undef SW

    // ---- block reduction of count / checksum; the boundary records come from s_edge
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
        rec_valid[2 * t]     = ex_f ? 1 : 0;
        rec_valid[2 * t + 1] = ex_l ? 1 : 0;
        if (ex_f) {
            MjRec rec = {kf, s_edge[0][0], s_edge[1][0], s_edge[0][1], s_edge[1][1]};
            recs[2 * t] = rec;
        }
        if (ex_l) {
            MjRec rec = {kl, s_edge[2][0], s_edge[3][0], s_edge[2][1], s_edge[3][1]};
            recs[2 * t + 1] = rec;
        }
    }
}

// Stitch the per-tile boundary records: consecutive valid records with the same key form one run.
// Every thread walks a contiguous chunk of the records: complete runs inside the chunk are
// multiplied, the chunk's first and last run may continue next door and are kept as OPEN records.
// open_out == null (one CTA): the open records go to shared memory and thread 0 stitches them.
// open_out != null (any grid): they go to open_out / open_valid (two per thread of the grid), which
// a second, one-CTA launch takes as its record array -- a single CTA walking 130 records per thread
// took 76 us per 128 M x 128 M join, a tenth of the scan itself.
constexpr int FX_THREADS = 512;
__global__ void __launch_bounds__(FX_THREADS) k_mj_fixup(const MjRec *__restrict__ recs,
                                                         const unsigned char *__restrict__ rec_valid,
                                                         u64 nrec, u64 *__restrict__ acc, int semi,
                                                         MjRec *__restrict__ open_out,
                                                         unsigned char *__restrict__ open_valid) {
    __shared__ MjRec         s_open_sh[2 * FX_THREADS];
    __shared__ unsigned char s_valid_sh[2 * FX_THREADS];
    __shared__ u64           s_red[2][FX_THREADS / 32];
    const u64 nthreads = (u64)gridDim.x * FX_THREADS, gtid = (u64)blockIdx.x * FX_THREADS + threadIdx.x;
    const u64 chunk = (nrec + nthreads - 1) / nthreads;
    const u64 lo    = min(nrec, gtid * chunk);
    const u64 hi    = min(nrec, lo + chunk);
    u64       my_cnt = 0, my_cks = 0;
    // this thread's two open slots
    MjRec         *s_open  = open_out ? open_out + 2 * gtid - 2 * threadIdx.x : s_open_sh;
    unsigned char *s_valid = open_out ? open_valid + 2 * gtid - 2 * threadIdx.x : s_valid_sh;
    s_valid[2 * threadIdx.x]     = 0;
    s_valid[2 * threadIdx.x + 1] = 0;
    {
        MjRec cur;
        bool  have = false, first_saved = false;
        cur.key = cur.cr = cur.cs = cur.sa = cur.sb = 0;
        for (u64 i = lo; i < hi; i++) {
            if (!rec_valid[i]) continue;
            const MjRec x = recs[i];
            if (have && x.key == cur.key) {
                cur.cr += x.cr;
                cur.cs += x.cs;
                cur.sa += x.sa;
                cur.sb += x.sb;
                continue;
            }
            if (have) {
                if (!first_saved) {  // may continue a run of the previous chunk
                    s_open[2 * threadIdx.x]  = cur;
                    s_valid[2 * threadIdx.x] = 1;
                    first_saved              = true;
                } else {
                    my_cnt += semi ? (cur.cr ? cur.cs : 0ull) : cur.cr * cur.cs;
                    my_cks += cur.sa * cur.sb;
                }
            }
            cur  = x;
            have = true;
        }
        if (have) {  // may continue in the next chunk
            const int slot                  = first_saved ? 1 : 0;
            s_open[2 * threadIdx.x + slot]  = cur;
            s_valid[2 * threadIdx.x + slot] = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && !open_out) {
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
int smj_merge_join_acc(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, u64 *acc, int semi) {
    if (nr == 0 || ns == 0) return 0;
    const u64  tiles = div_up(nr + ns, MJ_TILE);
    SmjScratch sc(ctx);
    u64           *d_diag  = (u64 *)sc.get((tiles + 1) * 8);
    MjRec         *d_recs  = (MjRec *)sc.get(tiles * 2 * sizeof(MjRec));
    unsigned char *d_valid = (unsigned char *)sc.get(tiles * 2);
    if (!d_diag || !d_recs || !d_valid) return smj_fail_msg("merge_join: scratch allocation failed");
    SMJ_LAUNCH(ctx, k_mj_partition, (unsigned)div_up(tiles + 1, 256), 256, 0, r, nr, s, ns, tiles,
               d_diag);
    SMJ_ITEMS(ctx, nr + ns);
    SMJ_LAUNCH(ctx, k_mj_tile, (unsigned)tiles, MJ_THREADS, 0, r, nr, s, ns, d_diag, d_recs,
               d_valid, acc, semi);
    const u64 nrec = tiles * 2;
    if (nrec <= 8 * FX_THREADS) {
        SMJ_LAUNCH(ctx, k_mj_fixup, 1, FX_THREADS, 0, d_recs, d_valid, nrec, acc, semi, (MjRec *)nullptr,
                   (unsigned char *)nullptr);
    } else {  // two levels: 8 records per thread, then the open records of those threads in one CTA
        const unsigned g     = (unsigned)std::min<u64>(div_up(nrec, 8 * FX_THREADS), 8);
        const u64      nopen = 2ull * g * FX_THREADS;
        MjRec         *d_open  = (MjRec *)sc.get(nopen * sizeof(MjRec));
        unsigned char *d_ovalid = (unsigned char *)sc.get(nopen);
        if (!d_open || !d_ovalid) return smj_fail_msg("merge_join: scratch allocation failed");
        SMJ_LAUNCH(ctx, k_mj_fixup, g, FX_THREADS, 0, d_recs, d_valid, nrec, acc, semi, d_open, d_ovalid);
        SMJ_LAUNCH(ctx, k_mj_fixup, 1, FX_THREADS, 0, (const MjRec *)d_open, (const unsigned char *)d_ovalid, nopen, acc, semi,
                   (MjRec *)nullptr, (unsigned char *)nullptr);
    }
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
// In-partition sort + merge-join, fused, for the 2^18 key-range buckets of smj_msd_partition.
//
// Inside a bucket only L <= 10 key bits still differ, so the in-partition sort degenerates to a
// counting sort over 2^L key slots, and the merge-join scan over two sorted buckets only needs,
// per key, |run_R| and sum A(r) -- exactly the counting sort's histogram.  One warp owns a bucket:
// it builds {count, sum A} per key slot from the R bucket in shared memory and streams the S
// bucket past it:  count += |run_R(key(s))|,  checksum += sum A(run_R(key(s))) * B(s).
// Sorted buckets are never written back: 8 B per input tuple of HBM traffic instead of 16 + 8.
// Buckets are handed out through an atomic counter, so uneven bucket sizes balance themselves.
// ---------------------------------------------------------------------------------------
namespace {
constexpr int BJ_WARPS = 8;

__global__ void __launch_bounds__(BJ_WARPS * 32, 4) k_bucket_join(const u64 *__restrict__ r,
                                                               const u64 *__restrict__ rs_,
                                                               const u64 *__restrict__ re_,
                                                               const u64 *__restrict__ s,
                                                               const u64 *__restrict__ ss_,
                                                               const u64 *__restrict__ se_,
                                                               u32 nbuckets, u32 lo_shift, u32 lbits,
                                                               u64 big, u32 *next_bucket,
                                                               u64 *__restrict__ acc,
                                                               const u32 *__restrict__ range) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // range != null (join group): only the buckets [range[0], range[1]) can hold anything on this GPU
    u32 first_bucket = 0;
    if (range) {  // (clamped: a layout that was abandoned half way must not send the walk out of bounds)
        nbuckets     = min(range[1], nbuckets);
        first_bucket = min(range[0], nbuckets);
    }
    const u32 slots = 1u << lbits, mask = slots - 1;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per warp: u64 sumA[slots] followed by u32 cnt[slots]
    u64 *t_sum = (u64 *)(smem_raw + (size_t)warp * slots * 12);
    u32 *t_cnt = (u32 *)(t_sum + slots);
    u64  my_cnt = 0, my_cks = 0;
    constexpr u32 CHUNK = 4;  // buckets claimed per atomic: one round trip for 4 buckets' bounds
    constexpr int UN    = 8;  // loads in flight per lane and relation
    while (true) {
        u32 b0 = 0;
        if (lane == 0) b0 = first_bucket + atomicAdd(next_bucket, CHUNK);
        b0 = __shfl_sync(SMJ_FULL_MASK, b0, 0);
        if (b0 >= nbuckets) break;
        // lanes 0..15 fetch the 4 x {r0, r1, s0, s1} of the chunk in one go
        u64 bound = 0;
        {
            const u32  k   = lane >> 2, which = lane & 3, bb = b0 + k;
            const u64 *arr = which == 0 ? rs_ : which == 1 ? re_ : which == 2 ? ss_ : se_;
            if (lane < 4 * CHUNK && bb < nbuckets) bound = arr[bb];
        }
#pragma unroll 1
        for (u32 k = 0; k < CHUNK; k++) {
        const u64 r0 = __shfl_sync(SMJ_FULL_MASK, bound, 4 * k), r1 = __shfl_sync(SMJ_FULL_MASK, bound, 4 * k + 1);
        const u64 s0 = __shfl_sync(SMJ_FULL_MASK, bound, 4 * k + 2), s1 = __shfl_sync(SMJ_FULL_MASK, bound, 4 * k + 3);
        if (r0 == r1 || s0 == s1) continue;
        if (r1 - r0 > big || s1 - s0 > big) continue;  // heavy hitters: k_big_bucket_join
        // Fast path, R keys unique inside the bucket (the usual primary-key side): the slot holds
        // the R item itself (stored complemented, 0 = empty), one exchange per R item and one load
        // per S item.  A second R item on the same slot sends the bucket down the general path.
        // The first batches of R and of S are requested together, so a bucket of average size
        // costs three DRAM round trips instead of one per 128 items.
        u64 x[UN], y[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) x[u] = (r0 + lane + 32 * u < r1) ? ld_stream(r + r0 + lane + 32 * u) : 0ull;
#pragma unroll
        for (int u = 0; u < UN; u++) y[u] = (s0 + lane + 32 * u < s1) ? ld_stream(s + s0 + lane + 32 * u) : 0ull;
        for (u32 i = lane; i < slots; i += 32) t_sum[i] = 0;
        __syncwarp();
        bool dup = false;
        for (u64 i = r0 + lane;;) {
#pragma unroll
            for (int u = 0; u < UN; u++) {
                if (i + 32 * u < r1) {
                    const u32 slot = (u32)(x[u] >> lo_shift) & mask;
                    dup |= atomicExch(&t_sum[slot], ~x[u]) != 0ull;
                }
            }
            i += 32 * UN;
            if (i - lane >= r1) break;
#pragma unroll
            for (int u = 0; u < UN; u++) x[u] = (i + 32 * u < r1) ? ld_stream(r + i + 32 * u) : 0ull;
        }
        __syncwarp();
        if (!__any_sync(SMJ_FULL_MASK, dup)) {
            for (u64 j = s0 + lane;;) {
#pragma unroll
                for (int u = 0; u < UN; u++) {
                    if (j + 32 * u < s1) {
                        const u64 v = t_sum[(u32)(y[u] >> lo_shift) & mask];
                        if (v) {
                            my_cnt += 1;
                            my_cks += smj_cks_a(~v) * smj_cks_b(y[u]);
                        }
                    }
                }
                j += 32 * UN;
                if (j - lane >= s1) break;
#pragma unroll
                for (int u = 0; u < UN; u++) y[u] = (j + 32 * u < s1) ? ld_stream(s + j + 32 * u) : 0ull;
            }
            __syncwarp();
            continue;
        }
        // General path: {count, sum A} per key slot
        for (u32 i = lane; i < slots * 3 / 2; i += 32) ((u64 *)t_sum)[i] = 0;  // sums and counts
        __syncwarp();
        for (u64 i = r0 + lane; i < r1; i += 128) {  // 4 independent loads in flight per lane
            u64 x[4];
#pragma unroll
            for (int u = 0; u < 4; u++) x[u] = (i + 32 * u < r1) ? ld_stream(r + i + 32 * u) : 0ull;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (i + 32 * u < r1) {
                    const u32 slot = (u32)(x[u] >> lo_shift) & mask;
                    atomicAdd(&t_cnt[slot], 1u);
                    atomicAdd(&t_sum[slot], smj_cks_a(x[u]));
                }
            }
        }
        __syncwarp();
        for (u64 j = s0 + lane; j < s1; j += 128) {
            u64 y[4];
#pragma unroll
            for (int u = 0; u < 4; u++) y[u] = (j + 32 * u < s1) ? ld_stream(s + j + 32 * u) : 0ull;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (j + 32 * u < s1) {
                    const u32 slot = (u32)(y[u] >> lo_shift) & mask;
                    const u32 c    = t_cnt[slot];
                    if (c) {
                        my_cnt += c;
                        my_cks += t_sum[slot] * smj_cks_b(y[u]);
                    }
                }
            }
        }
        __syncwarp();
        }
    }
    my_cnt = warp_reduce_add(my_cnt);
    my_cks = warp_reduce_add(my_cks);
    if (lane == 0) {
        if (my_cnt) atomicAdd(acc, my_cnt);
        if (my_cks) atomicAdd(acc + 1, my_cks);
    }
}
}  // namespace

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

// Sorts the `runs` chunks of a relation by key.  The varying key bits are probed ONCE for the whole
// relation (one host synchronisation instead of one per run; a bit that varies in the relation but
// not in a run only costs that run a pass), small runs take the one-CTA sort.
int sort_runs(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int runs,
              std::vector<const u64 *> &ptrs, std::vector<u64> &lens) {
    std::vector<u64> off;
    chunk_bounds(n, runs, off);
    ptrs.clear();
    lens.clear();
    u64  h_bits[3] = {0, 0, 0};
    bool have_bits = false;
    for (int i = 0; i < runs; i++) {
        u64  len    = off[i + 1] - off[i];
        u64 *sorted = nullptr;
        int  rc;
        if (len <= SMJ_SMALL_SORT_MAX) {
            rc = smj_lsd_sort(ctx, in + off[i], a + off[i], b + off[i], len, 2, &sorted);
        } else {
            if (!have_bits) {
                if ((rc = smj_item_bits(ctx, in, n, h_bits))) return rc;
                h_bits[2] |= 1;  // "already sorted" of the whole relation is not probed per run: sort
                have_bits = true;
            }
            rc = smj_lsd_sort_bits(ctx, in + off[i], a + off[i], b + off[i], len, 2, h_bits, &sorted);
        }
        if (rc) return rc;
        ptrs.push_back(sorted);
        lens.push_back(len);
    }
    return 0;
}
}  // namespace

// Wide key domains (11..13 key bits left inside a bucket) are usually SPARSE: 128 M random 31-bit
// keys put ~500 tuples into a bucket with 8192 possible key values, and clearing a 64 KB
// direct-address table per bucket would cost more than the bucket.  k_bucket_join_bitmap keeps the
// direct addressing but tracks occupancy in a BITMAP (1 KB to clear per bucket): a slot holds only
// the R payload (the R key of a match is the S key), atomicOr on the bitmap word claims it and
// detects a duplicate R key, which marks the bucket in dupflag for k_bucket_join_cta (the general
// {count, sum A} table).  33 KB of shared memory per CTA instead of 96: six CTAs per SM.
// (An open-addressing hash table per warp was tried first: its divergent probe loops made it as
// slow as clearing the big table, 2.9 ms for 128 M x 128 M.)
namespace {
__global__ void __launch_bounds__(256)
    k_bucket_join_bitmap(const u64 *__restrict__ r, const u64 *__restrict__ rs_, const u64 *__restrict__ re_,
                         const u64 *__restrict__ s, const u64 *__restrict__ ss_, const u64 *__restrict__ se_,
                         u32 nbuckets, u32 lo_shift, u32 lbits, u64 big, u32 *next_bucket,
                         u32 *__restrict__ dupflag, u64 *__restrict__ acc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 slots = 1u << lbits, mask = slots - 1, words = slots >> 5;
    u32 *tpay = (u32 *)smem_raw;  // [slots] R payload
    u32 *bm   = tpay + slots;     // [slots / 32] occupancy
    constexpr u32 CLAIM = 64;
    constexpr int UN    = 4;
    __shared__ u32 s_b, s_n, s_todo[CLAIM];
    __shared__ u64 s_red[2][8];
    u64 my_cnt = 0, my_cks = 0;
    u32 q = 0, nq = 0;
    while (true) {
        if (q == nq) {  // next block of buckets: threads 0..63 test one bucket each
            __syncthreads();
            if (threadIdx.x == 0) {
                s_b = atomicAdd(next_bucket, CLAIM);
                s_n = 0;
            }
            __syncthreads();
            const u32 base = s_b;
            if (base >= nbuckets) break;
            if (threadIdx.x < CLAIM && base + threadIdx.x < nbuckets) {
                const u32 bb  = base + threadIdx.x;
                const u64 nr_ = re_[bb] - rs_[bb], ns_ = se_[bb] - ss_[bb];
                if (nr_ && ns_ && nr_ <= big && ns_ <= big) s_todo[atomicAdd(&s_n, 1u)] = bb;
            }
            __syncthreads();
            q  = 0;
            nq = s_n;
            continue;
        }
        __syncthreads();  // the previous bucket is done with the table
        const u32 b = s_todo[q++];
        const u64 r0 = rs_[b], r1 = re_[b], s0 = ss_[b], s1 = se_[b];
        for (u32 i = threadIdx.x; i < words; i += 256) bm[i] = 0;
        __syncthreads();
        int dup = 0;
        for (u64 i = r0 + threadIdx.x; i < r1; i += UN * 256) {
            u64 x[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) x[u] = (i + u * 256 < r1) ? ld_stream(r + i + u * 256) : 0ull;
#pragma unroll
            for (int u = 0; u < UN; u++) {
                if (i + u * 256 < r1) {
                    const u32 slot = (u32)(x[u] >> lo_shift) & mask, bit = 1u << (slot & 31);
                    dup |= (atomicOr(&bm[slot >> 5], bit) & bit) != 0;
                    tpay[slot] = (u32)x[u];
                }
            }
        }
        if (__syncthreads_or(dup)) {  // a duplicate R key: the general table of k_bucket_join_cta
            if (threadIdx.x == 0) dupflag[b] = 1u;
            continue;
        }
        for (u64 j = s0 + threadIdx.x; j < s1; j += UN * 256) {
            u64 y[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) y[u] = (j + u * 256 < s1) ? ld_stream(s + j + u * 256) : 0ull;
#pragma unroll
            for (int u = 0; u < UN; u++) {
                if (j + u * 256 < s1) {
                    const u32 slot = (u32)(y[u] >> lo_shift) & mask;
                    if ((bm[slot >> 5] >> (slot & 31)) & 1u) {  // equal keys: R item = S key + stored payload
                        my_cnt += 1;
                        my_cks += smj_cks_a((y[u] & 0xffffffff00000000ull) | tpay[slot]) * smj_cks_b(y[u]);
                    }
                }
            }
        }
    }
    my_cnt = warp_reduce_add(my_cnt);
    my_cks = warp_reduce_add(my_cks);
    if ((threadIdx.x & 31) == 0) {
        s_red[0][threadIdx.x >> 5] = my_cnt;
        s_red[1][threadIdx.x >> 5] = my_cks;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 c = 0, x = 0;
        for (int w = 0; w < 8; w++) {
            c += s_red[0][w];
            x += s_red[1][w];
        }
        if (c) atomicAdd(acc, c);
        if (x) atomicAdd(acc + 1, x);
    }
}
}  // namespace

// Same join for 11..13 remaining key bits (key domains up to 2^31): the 2^L-slot table no longer
// fits a warp's share of shared memory, so a whole CTA owns a bucket.
__global__ void __launch_bounds__(256) k_bucket_join_cta(const u64 *__restrict__ r,
                                                         const u64 *__restrict__ rs_,
                                                         const u64 *__restrict__ re_,
                                                         const u64 *__restrict__ s,
                                                         const u64 *__restrict__ ss_,
                                                         const u64 *__restrict__ se_, u32 nbuckets,
                                                         u32 lo_shift, u32 lbits, u64 big, u32 min_r,
                                                         const u32 *__restrict__ dupflag,
                                                         u32 *next_bucket, u64 *__restrict__ acc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 slots = 1u << lbits, mask = slots - 1;
    u64 *t_sum = (u64 *)smem_raw;
    u32 *t_cnt = (u32 *)(t_sum + slots);
    constexpr u32 CLAIM = 64;  // buckets claimed per atomic; the ones this kernel has to do are compacted
    __shared__ u32 s_b, s_n, s_todo[CLAIM];
    __shared__ u64 s_red[2][8];
    u64 my_cnt = 0, my_cks = 0;
    constexpr int UN = 4;  // loads in flight per thread
    u32 q = 0, nq = 0;
    while (true) {
        if (q == nq) {  // next block of buckets: threads 0..63 test one bucket each
            __syncthreads();
            if (threadIdx.x == 0) {
                s_b = atomicAdd(next_bucket, CLAIM);
                s_n = 0;
            }
            __syncthreads();
            const u32 base = s_b;
            if (base >= nbuckets) break;
            if (threadIdx.x < CLAIM && base + threadIdx.x < nbuckets) {
                const u32 bb = base + threadIdx.x;
                const u64 nr_ = re_[bb] - rs_[bb], ns_ = se_[bb] - ss_[bb];
                // empty: nothing; heavy: the host takes the general path; without duplicate R keys:
                // done by k_bucket_join_bitmap
                if (nr_ && ns_ && nr_ <= big && ns_ <= big && (nr_ > (u64)min_r || dupflag[bb]))
                    s_todo[atomicAdd(&s_n, 1u)] = bb;
            }
            __syncthreads();
            q  = 0;
            nq = s_n;
            continue;
        }
        __syncthreads();  // the previous bucket is done with the table
        const u32 b = s_todo[q++];
        const u64 r0 = rs_[b], r1 = re_[b], s0 = ss_[b], s1 = se_[b];
        // fast path as in k_bucket_join: the slot holds the (complemented) R item; a second R item
        // on a slot sends the bucket down the {count, sum A} path.  (Wiping only the used slots of a
        // sparse bucket instead of zeroing the table was measured: slower, 3.6 against 3.05 ms.)
        for (u32 i = threadIdx.x; i < slots; i += 256) t_sum[i] = 0;
        __syncthreads();
        int dup = 0;
        for (u64 i = r0 + threadIdx.x; i < r1; i += UN * 256) {
            u64 x[UN];
#pragma unroll
            for (int u = 0; u < UN; u++) x[u] = (i + u * 256 < r1) ? ld_stream(r + i + u * 256) : 0ull;
#pragma unroll
            for (int u = 0; u < UN; u++)
                if (i + u * 256 < r1) dup |= atomicExch(&t_sum[(u32)(x[u] >> lo_shift) & mask], ~x[u]) != 0ull;
        }
        if (!__syncthreads_or(dup)) {
            for (u64 j = s0 + threadIdx.x; j < s1; j += UN * 256) {
                u64 y[UN];
#pragma unroll
                for (int u = 0; u < UN; u++) y[u] = (j + u * 256 < s1) ? ld_stream(s + j + u * 256) : 0ull;
#pragma unroll
                for (int u = 0; u < UN; u++) {
                    if (j + u * 256 < s1) {
                        const u64 v = t_sum[(u32)(y[u] >> lo_shift) & mask];
                        if (v) {
                            my_cnt += 1;
                            my_cks += smj_cks_a(~v) * smj_cks_b(y[u]);
                        }
                    }
                }
            }
            continue;
        }
        for (u32 i = threadIdx.x; i < slots * 3 / 2; i += 256) ((u64 *)t_sum)[i] = 0;
        __syncthreads();
        for (u64 i = r0 + threadIdx.x; i < r1; i += 256) {
            const u64 x    = ld_stream(r + i);
            const u32 slot = (u32)(x >> lo_shift) & mask;
            atomicAdd(&t_cnt[slot], 1u);
            atomicAdd(&t_sum[slot], smj_cks_a(x));
        }
        __syncthreads();
        for (u64 j = s0 + threadIdx.x; j < s1; j += 256) {
            const u64 y    = ld_stream(s + j);
            const u32 slot = (u32)(y >> lo_shift) & mask;
            const u32 c    = t_cnt[slot];
            if (c) {
                my_cnt += c;
                my_cks += t_sum[slot] * smj_cks_b(y);
            }
        }
    }
    my_cnt = warp_reduce_add(my_cnt);
    my_cks = warp_reduce_add(my_cks);
    if ((threadIdx.x & 31) == 0) {
        s_red[0][threadIdx.x >> 5] = my_cnt;
        s_red[1][threadIdx.x >> 5] = my_cks;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 c = 0, x = 0;
        for (int w = 0; w < 8; w++) {
            c += s_red[0][w];
            x += s_red[1][w];
        }
        if (c) atomicAdd(acc, c);
        if (x) atomicAdd(acc + 1, x);
    }
}

// ---- heavy hitters (Zipf-skewed S): buckets with more than `big` tuples on either side are
// listed by k_collect_big and joined by ALL warps of the grid together: every warp builds the
// (small) R table of the bucket and takes an equal slice of its S side.
namespace {
__global__ void __launch_bounds__(256) k_collect_big(const u64 *__restrict__ rs_,
                                                     const u64 *__restrict__ re_,
                                                     const u64 *__restrict__ ss_,
                                                     const u64 *__restrict__ se_, u32 nbuckets,
                                                     u64 big, u64 rmax, u32 *__restrict__ list,
                                                     u32 *__restrict__ info /* [0] count, [1] unsupported */) {
    const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    const u64 nr = re_[b] - rs_[b], ns = se_[b] - ss_[b];
    if (nr == 0 || ns == 0) return;
    if (nr > big || ns > big) {
        if (nr > rmax) atomicExch(&info[1], 1u);  // both sides heavy: leave it to the general path
        list[atomicAdd(&info[0], 1u)] = b;
    }
}

// The S side of every listed (heavy) bucket is cut into chunks of BIG_CHUNK tuples; pref[q] = first
// chunk of the q-th listed bucket, pref[nbig] = number of chunks.  One CTA.  (The prefix array lives
// in the upper half of the list array: more than 2^17 - 1 heavy buckets go to the general path.)
constexpr u32 BIG_CHUNK = 1u << 14;
__global__ void __launch_bounds__(1024) k_big_chunks(const u64 *__restrict__ ss_, const u64 *__restrict__ se_,
                                                     const u32 *__restrict__ list, u32 *__restrict__ info,
                                                     u32 *__restrict__ pref, u32 pref_cap) {
    __shared__ u32 s_w[32];
    const u32 nbig = info[0], lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (nbig + 1 > pref_cap) {
        if (threadIdx.x == 0) info[1] = 1u;
        return;
    }
    u32 carry = 0;
    for (u32 base = 0; base < nbig; base += 1024) {
        const u32 q = base + threadIdx.x;
        u32       c = 0;
        if (q < nbig) {
            const u32 b = list[q];
            c           = (u32)((se_[b] - ss_[b] + BIG_CHUNK - 1) / BIG_CHUNK);
        }
        u32 incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 v = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
            if (lane >= (u32)o) incl += v;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        u32 excl = incl - c + carry, tot = 0;
        for (u32 w = 0; w < 32; w++) {
            if (w < warp) excl += s_w[w];
            tot += s_w[w];
        }
        if (q < nbig) pref[q] = excl;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) pref[nbig] = carry;
}

// Every warp of the grid takes an equal, contiguous share of the chunks: it builds the {count, sum A}
// table of the chunk's bucket from the (small) R side -- once per bucket, consecutive chunks of one
// bucket reuse it -- and streams the chunk's S tuples past it.  Work is proportional to the heavy S
// tuples (the first version gave every bucket to ALL warps: every warp rebuilt every table, 11 ms of
// the 26 ms of a Zipf(1.0) join).
__global__ void __launch_bounds__(BJ_WARPS * 32) k_big_bucket_join(const u64 *__restrict__ r,
                                                                   const u64 *__restrict__ rs_,
                                                                   const u64 *__restrict__ re_,
                                                                   const u64 *__restrict__ s,
                                                                   const u64 *__restrict__ ss_,
                                                                   const u64 *__restrict__ se_,
                                                                   const u32 *__restrict__ list,
                                                                   const u32 *__restrict__ info,
                                                                   const u32 *__restrict__ pref,
                                                                   u32 lo_shift, u32 lbits,
                                                                   u64 *__restrict__ acc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 slots = 1u << lbits, mask = slots - 1;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u64 *t_sum = (u64 *)(smem_raw + (size_t)warp * slots * 12);
    u32 *t_cnt = (u32 *)(t_sum + slots);
    const u64 gw = (u64)blockIdx.x * BJ_WARPS + warp, nw = (u64)gridDim.x * BJ_WARPS;
    if (info[1]) return;  // a bucket heavy on both sides: the host takes the general path
    const u32 nbig = info[0];
    if (nbig == 0) return;
    const u64 total = pref[nbig], per = (total + nw - 1) / nw;
    const u64 c0 = gw * per, c1 = min(total, c0 + per);
    u64       my_cnt = 0, my_cks = 0;
    u32       q = 0, have_b = 0xffffffffu;
    if (c0 < c1) {  // listed bucket of the first chunk
        u32 lo = 0, hi = nbig;
        while (hi - lo > 1) {
            const u32 mid = (lo + hi) >> 1;
            if ((u64)pref[mid] <= c0) lo = mid; else hi = mid;
        }
        q = lo;
    }
    for (u64 c = c0; c < c1; c++) {
        while ((u64)pref[q + 1] <= c) q++;
        const u32 b  = list[q];
        const u64 s0 = ss_[b], s1 = se_[b];
        const u64 j0 = s0 + (c - pref[q]) * BIG_CHUNK, j1 = min(s1, j0 + BIG_CHUNK);
        if (b != have_b) {
            const u64 r0 = rs_[b], r1 = re_[b];
            for (u32 i = lane; i < slots * 3 / 2; i += 32) ((u64 *)t_sum)[i] = 0;
            __syncwarp();
            for (u64 i = r0 + lane; i < r1; i += 32) {
                const u64 x    = r[i];
                const u32 slot = (u32)(x >> lo_shift) & mask;
                atomicAdd(&t_cnt[slot], 1u);
                atomicAdd(&t_sum[slot], smj_cks_a(x));
            }
            __syncwarp();
            have_b = b;
        }
        for (u64 j = j0 + lane; j < j1; j += 256) {
            u64 y[8];
#pragma unroll
            for (int u = 0; u < 8; u++) y[u] = (j + 32 * u < j1) ? ld_stream(s + j + 32 * u) : 0ull;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (j + 32 * u < j1) {
                    const u32 slot = (u32)(y[u] >> lo_shift) & mask;
                    const u32 cc   = t_cnt[slot];
                    if (cc) {
                        my_cnt += cc;
                        my_cks += t_sum[slot] * smj_cks_b(y[u]);
                    }
                }
            }
        }
    }
    my_cnt = warp_reduce_add(my_cnt);
    my_cks = warp_reduce_add(my_cks);
    if (lane == 0) {
        if (my_cnt) atomicAdd(acc, my_cnt);
        if (my_cks) atomicAdd(acc + 1, my_cks);
    }
}
}  // namespace

// Buckets of one relation after the second partitioning level: bucket q is items[starts[q] ..
// ends[q]).  An exact pass B leaves them back to back (ends = starts + 1 of one offsets array), an
// optimistic one in fixed slots.

// classification of the 2^18 bucket pairs: heavy ones (more than BIG tuples on a side) are listed
// for the cooperative kernel; d_state[6] = {count, flag "heavy on both sides"}
constexpr u64 BJ_BIG = 1u << 16, BJ_RMAX = 1u << 16;
int smj_bucket_collect(SmjCtx *ctx, const BucketSet &R, const BucketSet &S, u64 *d_state, u32 *d_list) {
    const u32 nb2 = 512u * 512u;
    SMJ_CHECK(cudaMemsetAsync(d_state + 6, 0, 8, ctx->stream));
    SMJ_LAUNCH(ctx, k_collect_big, nb2 / 256, 256, 0, R.starts, R.ends, S.starts, S.ends, nb2, BJ_BIG, BJ_RMAX,
               d_list, (u32 *)(d_state + 6));
    return 0;
}

// Join of the 2^18 bucket pairs, enqueued behind bucket_collect without asking the host: the
// bucket kernel skips the heavy buckets, the cooperative kernel takes the listed ones (it returns
// at once when there are none).  Whether the result may be used is decided afterwards from
// d_state[6] (bucket_join_valid): a key that is heavy on BOTH sides, or any heavy bucket in a wide
// key domain (lbits > 10), is left to the general sort + scan path.  d_state[1] is the bucket
// counter (zero on entry).
int smj_bucket_join(SmjCtx *ctx, const BucketSet &R, const BucketSet &S, u32 lo_shift, u32 lbits,
                    u64 *d_state, u32 *d_list, u64 *acc, const u32 *d_range) {
    const u32 nb2 = 512u * 512u;
    u32 *d_next = (u32 *)(d_state + 1);
    if (lbits > 10) {
        // wide key domains (up to 2^31): one CTA per bucket with a bitmap-tracked payload table;
        // buckets with duplicate R keys are flagged (d_list doubles as the flags) and done by the
        // general-table kernel afterwards; no cooperative heavy-hitter kernel.
        const size_t bsmem = (size_t)(1u << lbits) * 4 + (size_t)(1u << lbits) / 8;
        if (smj_func_smem(ctx, (const void *)k_bucket_join_bitmap, 48 * 1024)) return -1;
        const int bm_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200u * 1024u) / bsmem));
        SMJ_CHECK(cudaMemsetAsync(d_list, 0, (size_t)nb2 * 4, ctx->stream));
        SMJ_ITEMS(ctx, R.n + S.n);
        SMJ_LAUNCH(ctx, k_bucket_join_bitmap, ctx->sm_count * bm_per_sm, 256, bsmem, R.items, R.starts, R.ends,
                   S.items, S.starts, S.ends, nb2, lo_shift, lbits, BJ_BIG, d_next, d_list, acc);
        const size_t smem = (size_t)(1u << lbits) * 12;
        if (smj_func_smem(ctx, (const void *)k_bucket_join_cta, smem)) return -1;
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200u * 1024u) / smem));
        SMJ_ITEMS(ctx, R.n + S.n);
        SMJ_LAUNCH(ctx, k_bucket_join_cta, ctx->sm_count * per_sm, 256, smem, R.items, R.starts, R.ends,
                   S.items, S.starts, S.ends, nb2, lo_shift, lbits, BJ_BIG, 0xffffffffu, d_list, d_next + 1, acc);
        return 0;
    }
    const size_t smem = (size_t)BJ_WARPS * (1u << lbits) * 12;
    if (smj_func_smem(ctx, (const void *)k_bucket_join, smem)) return -1;
    if (smj_func_smem(ctx, (const void *)k_big_bucket_join, smem)) return -1;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200u * 1024u) / std::max<size_t>(smem, 1)));
    SMJ_ITEMS(ctx, R.n + S.n);
    SMJ_LAUNCH(ctx, k_bucket_join, ctx->sm_count * per_sm, BJ_WARPS * 32, smem, R.items, R.starts, R.ends,
               S.items, S.starts, S.ends, nb2, lo_shift, lbits, BJ_BIG, d_next, acc, d_range);
    u32 *d_pref = d_list + nb2 / 2;  // (the list of heavy buckets is short: its upper half holds the chunk prefix)
    SMJ_LAUNCH(ctx, k_big_chunks, 1, 1024, 0, S.starts, S.ends, d_list, (u32 *)(d_state + 6), d_pref, nb2 / 2);
    SMJ_ITEMS(ctx, S.n);
    SMJ_LAUNCH(ctx, k_big_bucket_join, ctx->sm_count * per_sm, BJ_WARPS * 32, smem, R.items, R.starts,
               R.ends, S.items, S.starts, S.ends, d_list, (u32 *)(d_state + 6), (const u32 *)d_pref, lo_shift, lbits, acc);
    return 0;
}
bool smj_bucket_join_valid(const u64 *h_state, u32 lbits) {
    const u32 nbig = (u32)(h_state[6] & 0xffffffffu), unsupported = (u32)(h_state[6] >> 32);
    return lbits > 10 ? nbig == 0 : unsupported == 0;
}

// no key bit outside [lo, hi] may vary in R or S (exact = {OR_R, AND_R, OR_S, AND_S})
static bool msd_plan_holds(const u64 *exact, int lo, int hi) {
    const u32 xor_ = (u32)((exact[0] | exact[2]) >> 32), xand = (u32)((exact[1] & exact[3]) >> 32);
    const u32 range = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1)) & ~((1u << lo) - 1);
    return ((xor_ & ~xand) & ~range) == 0;
}

// d_state layout (16 x u64): [1] bucket counter, [2..5] exact OR / AND of R and S, [6] heavy-bucket
// info, [8..11] overflow flags of the optimistic passes {R level 1, S level 1, R level 2, S level 2},
// [12..13] first-level buckets of R / S whose second level was repeated exactly on the device,
// [14..15] R / S hold a key outside the sampled key range (optimistic first pass)
extern const u64 MSD_STATE_INIT[16] = {0ull, 0ull, 0ull, ~0ull, 0ull, ~0ull, 0ull, 0ull,
                                       0ull, 0ull, 0ull, 0ull,  0ull, 0ull,  0ull, 0ull};

struct Upload;
// One relation of the m-way fast path.  Optimistic: no histogram passes (sampled capacities, fixed
// slots); exact: histogram + scatter per level.
struct MsdRel {
    const u64 *in;             // device-resident input (the staging buffer of an upload)
    u64        n;
    Upload    *up;             // non-null: `in` arrives chunk by chunk
    const u64 *sample_from;    // what the sample kernel reads (the pinned host copy of an upload)
    u64       *a, *b;          // level-1 / level-2 outputs
    u64        alloc1, alloc2; // capacities of a and b in items
    u64       *starts, *ends;  // 2^18 (+1) entries each
    u64       *d_tab, *d_counts;
    bool       opt1, opt2;     // level 1 / level 2 ran optimistically
    BucketSet set() const { return BucketSet{b, starts, opt2 ? ends : starts + 1, n}; }
};

// A relation in pinned host memory on its way to HBM: copied chunk by chunk (128 MiB) on the copy stream,
// one event per chunk, so that the first partitioning pass can consume chunk k while chunk k + 1
// is still crossing PCIe (the pass needs no histogram, hence no complete relation).
struct Upload {
    SmjCtx                  *ctx  = nullptr;
    const u64               *host = nullptr;
    u64                     *dev  = nullptr;
    u64                      n = 0, chunk = 0;  // items, items per chunk
    std::vector<cudaEvent_t> ev;                // ev[k]: chunk k has arrived (cached, not owned)
    bool                     waited = false;
    bool                     host_visible = false;  // kernels can read `host` directly
    bool active() const { return host != nullptr; }
    u64  nchunks() const { return ev.size(); }
    // every later operation on ctx->stream sees the whole relation
    int wait_all() {
        if (active() && !waited && !ev.empty()) SMJ_CHECK(cudaStreamWaitEvent(ctx->stream, ev.back(), 0));
        waited = true;
        return 0;
    }
    ~Upload() { wait_all(); }  // the staging buffer is freed in stream order behind the copies
};
// (entries 0 and 1 of the context's event cache belong to smjdev_exchange_begin / _end)
static cudaEvent_t upload_event(SmjCtx *ctx, size_t idx) { return smj_pool_event(ctx, idx + 2); }
// starts the copies; ev_base: first index of this relation's events in the cache
static int upload_begin(SmjCtx *ctx, SmjScratch &sc, const u64 *host, u64 n, size_t ev_base, Upload &up) {
    u64 CHUNK = 0;  // items per chunk, a multiple of the scatter tile (8192 items)
    {
        // measured end to end (128 M x 128 M): 32 MiB 38.3 ms, 64 MiB 38.2, 128 MiB 38.1, 256 MiB 38.1
        const char *e = getenv("SMJ_UPLOAD_CHUNK_MB");
        CHUNK         = ((e && atoi(e) > 0) ? (u64)atoi(e) : 128ull) << 17;
    }
    up.ctx   = ctx;
    up.host  = host;
    up.n     = n;
    // kernels may read the pinned copy directly (the sample) only if it is mapped into the device's
    // address space at the same address; registered memory on some systems is not
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, const_cast<u64 *>(host), 0) != cudaSuccess) {
        cudaGetLastError();
        dp = nullptr;
    }
    up.host_visible = dp == (const void *)host;
    up.chunk = CHUNK;
    up.dev   = (u64 *)sc.get(n * 8);
    if (!up.dev) return smj_fail_msg("smjdev_join: out of device memory");
    // the staging buffer comes from the stream-ordered pool of ctx->stream
    cudaEvent_t e_alloc = upload_event(ctx, ev_base);
    if (!e_alloc) return smj_fail_msg("smjdev_join: cudaEventCreate failed");
    SMJ_CHECK(cudaEventRecord(e_alloc, ctx->stream));
    SMJ_CHECK(cudaStreamWaitEvent(ctx->copy_stream, e_alloc, 0));
    const u64 nch = div_up(n, CHUNK);
    for (u64 k = 0; k < nch; k++) {
        const u64   off = k * CHUNK, len = std::min(CHUNK, n - off);
        cudaEvent_t e   = upload_event(ctx, ev_base + 1 + k);
        if (!e) return smj_fail_msg("smjdev_join: cudaEventCreate failed");
        SMJ_CHECK(cudaMemcpyAsync(up.dev + off, host + off, len * 8, cudaMemcpyHostToDevice, ctx->copy_stream));
        SMJ_CHECK(cudaEventRecord(e, ctx->copy_stream));
        up.ev.push_back(e);
    }
    return 0;
}

// m-way fast path: two-level MSD range partition of R and S by the same 18 key bits, then the
// fused in-bucket counting sort + merge join.  *used = 0 when the key domain or the bucket sizes
// do not suit it; the caller then takes the general sort + scan path.  On success h_acc = {count,
// checksum} (the call has synchronised).
//
// The partition passes run OPTIMISTICALLY: one sample of R and S gives the key range, the
// first-level bin capacities and a skew test; both levels then scatter without a histogram pass
// into regions with slack, every key is checked against the sampled key range on the way, and the
// join kernels are enqueued right behind.  Only then does the host look: overflow flags, key
// range, heavy buckets.  What did not hold is repeated exactly (histogram first), so the result
// never depends on the guess.  In the good case the host synchronises twice per join (sample,
// result).
static int mway_msd(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, Upload *up_r, Upload *up_s,
                    u64 *acc, int *used, PhaseTimer &tm, int &t_part, int &t_sort, int &t_join,
                    u64 *scratch_bytes, u32 *path, u32 *repeats, u64 *h_acc) {
    *used = 0;
    if (nr + ns < (1u << 20)) return 0;  // 2^18 buckets only pay off on large inputs
    const char *mode_env = getenv("SMJ_MSD_MODE");  // "exact" disables the optimistic passes
    const int   mode     = (mode_env && mode_env[0] == 'e') ? 0 : 1;
    // test hooks: SMJ_MSD_TEST=shrink1 / shrink2 under-sizes the first / second level regions so
    // that the overflow -> exact repeat paths run on ordinary inputs
    const char *hook    = getenv("SMJ_MSD_TEST");
    const int   shrink1 = hook && !strcmp(hook, "shrink1"), shrink2 = hook && !strcmp(hook, "shrink2");
    const u32   nb2     = 512u * 512u;
    SmjScratch  sc(ctx);
    u64 *d_state = (u64 *)sc.get(128), *d_st = (u64 *)sc.get(smj_msd_sample_bytes(0));
    u32 *d_list = (u32 *)sc.get((size_t)nb2 * 4), *d_sub = (u32 *)sc.get(smj_msd_sample_bytes(1));
    if (!d_state || !d_list || !d_st || !d_sub) return smj_fail_msg("mway_msd: scratch allocation failed");
    SMJ_CHECK(cudaMemcpyAsync(d_state, MSD_STATE_INIT, 128, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<u64> samp(SMJ_MSD_SAMPLE_WORDS);
    // an uploading relation is sampled where it is complete: in pinned host memory, over PCIe
    // (a smaller sample: 2 MiB per relation next to the 128 MiB chunks of the upload)
    for (Upload *u : {up_r, up_s})  // a pinned copy the device cannot address: sample the uploaded one
        if (u && u->active() && !u->host_visible && u->wait_all()) return -1;
    const bool live_r = up_r && up_r->active() && !up_r->waited, live_s = up_s && up_s->active() && !up_s->waited;
    const bool uploading = live_r || live_s;
    const u64 *samp_r = live_r ? up_r->host : r;
    const u64 *samp_s = live_s ? up_s->host : s;
    int rc = smj_msd_sample(ctx, samp_r, nr, samp_s, ns, uploading ? 1u << 16 : 1u << 18, d_st, d_sub,
                            samp.data());
    bool have_sample = true;
    if (rc == 1) {
        // no cooperative launch on this device: plain sampled OR/AND, exact (histogram) passes
        std::fill(samp.begin(), samp.end(), 0ull);
        if ((rc = smj_key_bits_sampled(ctx, samp_r, nr, samp_s, ns, samp.data()))) return rc;
        have_sample = false;
    } else if (rc)
        return rc;
    MsdPlan pl;
    if (!smj_msd_plan(samp.data(), &pl)) return 0;
    // key bits outside [lo, hi] are constant in the sample: every key must carry the same value there
    const u32 krange = (pl.hi == 31 ? 0xffffffffu : ((1u << (pl.hi + 1)) - 1)) & ~((1u << pl.lo) - 1);
    const u32 kmask = ~krange, kref = (u32)((samp[1] & samp[3]) >> 32) & kmask;
    MsdRel rel[2] = {{r, nr, up_r, samp_r}, {s, ns, up_s, samp_s}};
    for (int q = 0; q < 2; q++) {
        MsdRel &x = rel[q];
        // equal slots at the second level are trusted only for relations whose sample shows no
        // duplicate keys (see k_msd_sample); the first level adapts its capacities to any skew
        const u64 nsamp = samp[8 + q], dups = samp[6 + q];
        x.opt1 = mode && have_sample && x.n >= (1u << 20);
        x.opt2 = x.opt1 && dups <= 16;
        x.alloc1 = x.opt1 ? smj_msd_caps1(samp.data() + 16 + 512 * q, nsamp, x.n, shrink1) : x.n;
        x.alloc2 = x.opt2 ? smj_msd_alloc2(x.n) : x.n;
        x.a      = (u64 *)sc.get(std::max(x.alloc1, x.n) * 8);
        x.b      = (u64 *)sc.get(std::max(x.alloc2, x.n) * 8);
        x.starts = (u64 *)sc.get((size_t)(nb2 + 1) * 8);
        x.ends   = (u64 *)sc.get((size_t)nb2 * 8);
        x.d_tab  = (u64 *)sc.get(3 * 512 * 8);
        x.d_counts = (u64 *)sc.get(512 * 8);
        if (!x.a || !x.b || !x.starts || !x.ends || !x.d_tab || !x.d_counts)
            return smj_fail_msg("smjdev_join: out of device memory");
    }
    u32 *flags = (u32 *)(d_state + 8);  // one u64 slot per flag, low word used
    if (rel[0].opt1 || rel[1].opt1)
        if ((rc = smj_msd_caps1_tables(ctx, d_st, nr, ns, rel[0].alloc1, rel[1].alloc1, shrink1,
                                       rel[0].opt1 ? rel[0].d_tab : nullptr,
                                       rel[1].opt1 ? rel[1].d_tab : nullptr, flags)))
            return rc;
    for (int q = 0; q < 2; q++) {
        MsdRel &x = rel[q];
        const bool chunked = x.opt1 && x.up && x.up->active() && !x.up->waited;
        if (chunked) {
            // the pass consumes the upload chunk by chunk (the bin cursors carry over)
            for (u64 k = 0; k < x.up->nchunks(); k++) {
                const u64 off = k * x.up->chunk, len = std::min(x.up->chunk, x.n - off);
                SMJ_CHECK(cudaStreamWaitEvent(ctx->stream, x.up->ev[k], 0));
                if ((rc = smj_msd_pass_a_opt(ctx, x.in + off, len, x.a, pl.shift1, x.d_tab,
                                             flags + 2 * q, kmask, kref, flags + 12 + 2 * q)))
                    return rc;
            }
            x.up->waited = true;
        } else if (x.up && (rc = x.up->wait_all()))
            return rc;
        if (x.opt1) {
            if (!chunked &&
                (rc = smj_msd_pass_a_opt(ctx, x.in, x.n, x.a, pl.shift1, x.d_tab, flags + 2 * q, kmask,
                                         kref, flags + 12 + 2 * q)))
                return rc;
            if (x.opt2) {
                if ((rc = smj_msd_pass_b_opt(ctx, x.a, x.n, x.d_tab, x.b, x.alloc2, pl.shift2, 1u << pl.lbits, x.starts,
                                             x.ends, x.d_counts, flags + 4 + 2 * q, d_state + 12 + q, shrink2)))
                    return rc;
            } else {
                if ((rc = smj_msd_counts_opt(ctx, x.d_tab, x.d_counts))) return rc;
                if ((rc = smj_msd_pass_b(ctx, x.a, x.n, x.d_counts, x.d_tab + 512, x.b, pl.shift2, x.starts)))
                    return rc;
            }
        } else if ((rc = smj_msd_partition(ctx, x.in, x.n, x.a, x.b, pl.shift1, pl.shift2, x.starts,
                                           d_state + 2 + 2 * q)))
            return rc;
    }
    *path = (rel[0].opt1 || rel[1].opt1) ? SMJ_PATH_MSD_OPT : SMJ_PATH_MSD;
    u64 h[18];
    for (int attempt = 0;; attempt++) {
        if ((rc = smj_bucket_collect(ctx, rel[0].set(), rel[1].set(), d_state, d_list))) return rc;
        const int t_mid = tm.mark();
        if ((rc = smj_bucket_join(ctx, rel[0].set(), rel[1].set(), pl.lo_shift, pl.lbits, d_state, d_list, acc)))
            return rc;
        SMJ_CHECK(cudaMemcpyAsync(h, d_state, 128, cudaMemcpyDeviceToHost, ctx->stream));
        SMJ_CHECK(cudaMemcpyAsync(h + 16, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
        const int t_done = tm.mark();
        SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
        t_part = t_sort = t_mid;  // the in-partition sort is fused into the join kernel
        t_join = t_done;
        // the sample missed key bits that vary (seen by an exact histogram pass or by the key check)
        bool ok = msd_plan_holds(h + 2, pl.lo, pl.hi) && !h[14] && !h[15];
        bool redo = false;
        for (int q = 0; q < 2 && ok; q++) {
            MsdRel &x = rel[q];
            const bool over1 = x.opt1 && h[8 + q], over2 = x.opt2 && h[10 + q];
            if (over1) {  // the whole relation again, exactly (a and b are large enough)
                if ((rc = smj_msd_partition(ctx, x.in, x.n, x.a, x.b, pl.shift1, pl.shift2, x.starts, nullptr)))
                    return rc;
                x.opt1 = x.opt2 = false;
                redo = true;
                *repeats += 1;
            } else if (over2) {  // the slots did not fit b: level 2 again from the intact level-1 output
                if ((rc = smj_msd_pass_b(ctx, x.a, x.n, x.d_counts, x.d_tab + 512, x.b, pl.shift2, x.starts)))
                    return rc;
                x.opt2 = false;
                redo   = true;
                *repeats += 1;
            }
        }
        if (ok && redo && attempt == 0) {  // join again on the repaired buckets
            SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
            SMJ_CHECK(cudaMemsetAsync(d_state + 1, 0, 8, ctx->stream));
            continue;
        }
        ok = ok && !redo && smj_bucket_join_valid(h, pl.lbits);
        if (!ok) {
            SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
            return 0;
        }
        break;
    }
    for (int q = 0; q < 2; q++)
        if (rel[q].opt2) *repeats += (u32)h[12 + q];
    h_acc[0] = h[16];
    h_acc[1] = h[17];
    *used    = 1;
    if (scratch_bytes) *scratch_bytes = sc.bytes;
    return 0;
}

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

    // pinned host relations are uploaded chunk by chunk on the copy stream (the m-way fast path
    // overlaps its first partitioning pass with the transfer); everything else sees device arrays
    Upload up_r, up_s;
    struct EventPair {  // timing events around the copies, destroyed on every return path
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } h2d;
    cudaEvent_t &h2d0 = h2d.a, &h2d1 = h2d.b;
    if (nr && ns && (smj_is_host_pointer(r) || smj_is_host_pointer(s))) {
        SMJ_CHECK(cudaEventCreate(&h2d0));
        SMJ_CHECK(cudaEventCreate(&h2d1));
        SMJ_CHECK(cudaEventRecord(h2d0, ctx->copy_stream));
        if (smj_is_host_pointer(r)) {
            if ((rc = upload_begin(ctx, sc, r, nr, 0, up_r))) return rc;
            r = up_r.dev;
        }
        if (smj_is_host_pointer(s)) {
            if ((rc = upload_begin(ctx, sc, s, ns, 1 + up_r.ev.size() + 1, up_s))) return rc;
            s = up_s.dev;
        }
        SMJ_CHECK(cudaEventRecord(h2d1, ctx->copy_stream));
    }
    u64 msd_bytes = 0, h_fast[2] = {0, 0};
    u32 path = SMJ_PATH_GENERAL, repeats = 0;
    int fast_done = 0;
    if (nr && ns) {
        u64 *aR = nullptr, *bR = nullptr, *aS = nullptr, *bS = nullptr;
        auto ping_pong = [&]() {
            aR = (u64 *)sc.get(nr * 8), bR = (u64 *)sc.get(nr * 8);
            aS = (u64 *)sc.get(ns * 8), bS = (u64 *)sc.get(ns * 8);
            return aR && bR && aS && bS;
        };
        if (algo == SMJ_ALGO_MWAY) {
            // range partition + in-partition sort == a global key sort on one GPU: the sorted
            // partitions are disjoint key ranges, so no run merging is needed
            // (the reference's nthreads == 1 path, sortmergejoin_multiway.c:476-489)
            int used = 0;
            const char *force = getenv("SMJ_MWAY_PATH");  // "lsd" forces the general path
            if (!(force && force[0] == 'l')) {
                if ((rc = mway_msd(ctx, r, nr, s, ns, &up_r, &up_s, acc, &used, tm, t_part, t_sort, t_join,
                                   &msd_bytes, &path, &repeats, h_fast)))
                    return rc;
                t_merge = t_sort;
            }
            fast_done = used;
            if (!used) path = SMJ_PATH_GENERAL;
            if (!used) {
                u64 *sr = nullptr, *ss = nullptr;
                if (up_r.wait_all() || up_s.wait_all()) return -1;
                if (!ping_pong()) return smj_fail_msg("smjdev_join: out of device memory");
                SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
                t_part = tm.mark();
                if ((rc = smj_lsd_sort(ctx, r, aR, bR, nr, 2, &sr))) return rc;
                if ((rc = smj_lsd_sort(ctx, s, aS, bS, ns, 2, &ss))) return rc;
                t_sort = t_merge = tm.mark();
                if ((rc = smj_merge_join_acc(ctx, sr, nr, ss, ns, acc))) return rc;
                t_join = tm.mark();
            }
        } else if (up_r.wait_all() || up_s.wait_all()) {
            return -1;
        } else if (algo == SMJ_ALGO_MPASS) {
            // sorted runs, then pairwise merge passes (sortmergejoin_multipass.c:410-708)
            std::vector<const u64 *> pr, ps;
            std::vector<u64>         lr, ls;
            if (!ping_pong()) return smj_fail_msg("smjdev_join: out of device memory");
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
            if (!ping_pong()) return smj_fail_msg("smjdev_join: out of device memory");
            t_part = tm.mark();
            if ((rc = smj_lsd_sort(ctx, r, aR, bR, nr, 2, &sr))) return rc;
            if ((rc = sort_runs(ctx, s, aS, bS, ns, nruns, ps, ls))) return rc;
            t_sort = t_merge = tm.mark();
            // all runs against R in one pass over R (smj_mpsm.cu); one merge-path scan per run when that
            // does not apply (long runs of equal keys in R, heavy hitters in S)
            int probed = 0;
            if ((rc = smj_mpsm_probe(ctx, sr, nr, ps.data(), ls.data(), (u32)ps.size(), acc, &probed))) return rc;
            for (size_t i = 0; i < ps.size() && !probed; i++)
                if ((rc = smj_merge_join_acc(ctx, sr, nr, ps[i], ls[i], acc))) return rc;
            t_join = tm.mark();
        }
    } else {
        t_part = t_sort = t_merge = t_join = tm.mark();
    }
    u64 h[2] = {h_fast[0], h_fast[1]};
    int t_end = t_join;
    if (!fast_done) {  // (the fast path has already synchronised and brought the result along)
        SMJ_CHECK(cudaMemcpyAsync(h, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
        t_end = tm.mark();
        SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    if (h2d0) {
        cudaEventSynchronize(h2d1);
        cudaEventElapsedTime(&stats->ms_h2d, h2d0, h2d1);
    }
    stats->matches      = h[0];
    stats->checksum     = h[1];
    stats->ms_partition = tm.ms(t0, t_part);
    stats->ms_sort      = tm.ms(t_part, t_sort);
    stats->ms_merge     = tm.ms(t_sort, t_merge);
    stats->ms_join      = tm.ms(t_merge, t_join);
    stats->ms_total     = tm.ms(t0, t_end);
    stats->launches     = ctx->launches - launches0;
    stats->hbm_bytes    = sc.bytes + msd_bytes;
    stats->path         = path;
    stats->repeats      = repeats;
    return 0;
}

// ---------------------------------------------------------------------------------------
// Multi-GPU building blocks: on G GPUs the FIRST pass of the MSD partition is the exchange.
// Every rank histograms its slice by the top 9 key bits (smjdev_msd_hist), the ranks agree on
// bucket owners and write offsets, and each rank's pass-A scatter stores its items directly into
// the owners' buffers over NVLink (smjdev_msd_scatter).  What a rank then holds is exactly the
// input of pass B: smjdev_join_level1 finishes the join from there.
// ---------------------------------------------------------------------------------------
extern "C" int smjdev_sample_key_bits(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                                      uint64_t *host4) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    return smj_key_bits_sampled(ctx, (const u64 *)r, nr, (const u64 *)s, ns, (u64 *)host4);
}

extern "C" int smjdev_msd_hist(const uint64_t *in, uint64_t n, uint32_t shift1, uint64_t *hist512,
                               uint64_t *bits2) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (shift1 < 32 || shift1 > 55) return smj_fail_msg("smjdev_msd_hist: digit outside the key");
    return smj_msd_hist1(ctx, (const u64 *)in, n, shift1, (u64 *)hist512, (u64 *)bits2);
}

extern "C" int smjdev_msd_scatter(const uint64_t *in, uint64_t n, uint32_t shift1,
                                  uint64_t *const *dest_ptrs512, const uint64_t *dest_offsets512) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (shift1 < 32 || shift1 > 55) return smj_fail_msg("smjdev_msd_scatter: digit outside the key");
    return smj_msd_scatter_to(ctx, (const u64 *)in, n, shift1, (u64 *const *)dest_ptrs512,
                              (const u64 *)dest_offsets512);
}

extern "C" int smjdev_msd_scatter_local(const uint64_t *in, uint64_t n, uint32_t shift1,
                                        const uint64_t *hist512_dev, uint64_t *out) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (shift1 < 32 || shift1 > 55) return smj_fail_msg("smjdev_msd_scatter_local: digit outside the key");
    return smj_msd_scatter_local(ctx, (const u64 *)in, n, shift1, (const u64 *)hist512_dev, (u64 *)out);
}

// common body of smjdev_join_level1 / smjdev_join_level1_seg: nseg == 0 -> counts512 describe 512
// back-to-back buckets, else seg = {start[nseg], count[nseg], bucket[nseg]} (host arrays)
static int join_level1(const uint64_t *r_, const uint64_t *rcounts512, uint32_t nseg_r, const uint64_t *rseg,
                       const uint64_t *s_, const uint64_t *scounts512, uint32_t nseg_s, const uint64_t *sseg,
                       uint32_t shift1, uint32_t lo_bit, smj_join_stats_t *stats);

extern "C" int smjdev_join_level1(const uint64_t *r_, const uint64_t *rcounts512, const uint64_t *s_,
                                  const uint64_t *scounts512, uint32_t shift1, uint32_t lo_bit,
                                  smj_join_stats_t *stats) {
    return join_level1(r_, rcounts512, 0, nullptr, s_, scounts512, 0, nullptr, shift1, lo_bit, stats);
}
extern "C" int smjdev_join_level1_seg(const uint64_t *r_, uint32_t nseg_r, const uint64_t *rseg,
                                      const uint64_t *s_, uint32_t nseg_s, const uint64_t *sseg,
                                      uint32_t shift1, uint32_t lo_bit, smj_join_stats_t *stats) {
    if (!nseg_r || !nseg_s || !rseg || !sseg) return smj_fail_msg("smjdev_join_level1_seg: no segments");
    return join_level1(r_, nullptr, nseg_r, rseg, s_, nullptr, nseg_s, sseg, shift1, lo_bit, stats);
}

static int join_level1(const uint64_t *r_, const uint64_t *rcounts512, uint32_t nseg_r, const uint64_t *rseg,
                       const uint64_t *s_, const uint64_t *scounts512, uint32_t nseg_s, const uint64_t *sseg,
                       uint32_t shift1, uint32_t lo_bit, smj_join_stats_t *stats) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (!stats) return smj_fail_msg("smjdev_join_level1: stats must not be NULL");
    if (shift1 < 41 || shift1 > 55 || 32 + lo_bit + 1 > shift1 - 9 || shift1 - 9 - (32 + lo_bit) > 13)
        return smj_fail_msg("smjdev_join_level1: key digits out of range");
    const u64 *r = (const u64 *)r_, *s = (const u64 *)s_;
    memset(stats, 0, sizeof(*stats));
    u64 nr = 0, ns = 0;
    if (nseg_r) {
        for (u32 i = 0; i < nseg_r; i++) nr += rseg[nseg_r + i];
        for (u32 i = 0; i < nseg_s; i++) ns += sseg[nseg_s + i];
    } else {
        for (int b = 0; b < 512; b++) {
            nr += rcounts512[b];
            ns += scounts512[b];
        }
    }
    const u64  launches0 = ctx->launches;
    const u32  shift2 = shift1 - 9, lo_shift = 32 + lo_bit, lbits = shift2 - lo_shift, nb2 = 512u * 512u;
    SmjScratch sc(ctx);
    u64 *acc = (u64 *)sc.get(16);
    if (!acc) return smj_fail_msg("smjdev_join_level1: scratch allocation failed");
    PhaseTimer tm(ctx);
    const int  t0 = tm.mark();
    int        t_part = t0, t_join = t0;
    SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    if (nr && ns) {
        u64 *bR = (u64 *)sc.get(nr * 8), *bS = (u64 *)sc.get(ns * 8);
        u64 *offR = (u64 *)sc.get((size_t)(nb2 + 1) * 8), *offS = (u64 *)sc.get((size_t)(nb2 + 1) * 8);
        u64 *d_h1 = (u64 *)sc.get(2 * 512 * 8), *d_state = (u64 *)sc.get(128);
        u32 *d_list = (u32 *)sc.get((size_t)nb2 * 4);
        if (!bR || !bS || !offR || !offS || !d_h1 || !d_state || !d_list)
            return smj_fail_msg("smjdev_join_level1: out of device memory");
        SMJ_CHECK(cudaMemcpyAsync(d_state, MSD_STATE_INIT, 128, cudaMemcpyHostToDevice, ctx->stream));
        int rc;
        if (nseg_r) {
            u64 *d_segr = (u64 *)sc.get((size_t)nseg_r * 24), *d_segs = (u64 *)sc.get((size_t)nseg_s * 24);
            if (!d_segr || !d_segs) return smj_fail_msg("smjdev_join_level1: out of device memory");
            SMJ_CHECK(cudaMemcpyAsync(d_segr, rseg, (size_t)nseg_r * 24, cudaMemcpyHostToDevice, ctx->stream));
            SMJ_CHECK(cudaMemcpyAsync(d_segs, sseg, (size_t)nseg_s * 24, cudaMemcpyHostToDevice, ctx->stream));
            if ((rc = smj_msd_pass_b_seg(ctx, r, nr, nseg_r, d_segr, bR, shift2, offR))) return rc;
            if ((rc = smj_msd_pass_b_seg(ctx, s, ns, nseg_s, d_segs, bS, shift2, offS))) return rc;
        } else {
            SMJ_CHECK(cudaMemcpyAsync(d_h1, rcounts512, 512 * 8, cudaMemcpyHostToDevice, ctx->stream));
            SMJ_CHECK(cudaMemcpyAsync(d_h1 + 512, scounts512, 512 * 8, cudaMemcpyHostToDevice, ctx->stream));
            if ((rc = smj_msd_pass_b(ctx, r, nr, d_h1, nullptr, bR, shift2, offR))) return rc;
            if ((rc = smj_msd_pass_b(ctx, s, ns, d_h1 + 512, nullptr, bS, shift2, offS))) return rc;
        }
        const BucketSet R = {bR, offR, offR + 1, nr}, S = {bS, offS, offS + 1, ns};
        u64 h[16];
        if ((rc = smj_bucket_collect(ctx, R, S, d_state, d_list))) return rc;
        t_part = tm.mark();
        if ((rc = smj_bucket_join(ctx, R, S, lo_shift, lbits, d_state, d_list, acc))) return rc;
        SMJ_CHECK(cudaMemcpyAsync(h, d_state, 128, cudaMemcpyDeviceToHost, ctx->stream));
        SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
        const int done = smj_bucket_join_valid(h, lbits);
        if (!done) {
            // heavy on both sides: general path on the same items (partitioning does not matter to it)
            u64 *aR = (u64 *)sc.get(nr * 8), *aS = (u64 *)sc.get(ns * 8), *sr = nullptr, *ss = nullptr;
            if (!aR || !aS) return smj_fail_msg("smjdev_join_level1: out of device memory");
            SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
            if ((rc = smj_lsd_sort(ctx, r, aR, bR, nr, 2, &sr))) return rc;
            if ((rc = smj_lsd_sort(ctx, s, aS, bS, ns, 2, &ss))) return rc;
            if ((rc = smj_merge_join_acc(ctx, sr, nr, ss, ns, acc))) return rc;
        }
        t_join = tm.mark();
    }
    u64 hres[2];
    SMJ_CHECK(cudaMemcpyAsync(hres, acc, 16, cudaMemcpyDeviceToHost, ctx->stream));
    const int t_end = tm.mark();
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    stats->matches      = hres[0];
    stats->checksum     = hres[1];
    stats->ms_partition = tm.ms(t0, t_part);
    stats->ms_join      = tm.ms(t_part, t_join);
    stats->ms_total     = tm.ms(t0, t_end);
    stats->launches     = ctx->launches - launches0;
    stats->hbm_bytes    = sc.bytes;
    return 0;
}
