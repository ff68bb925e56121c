// smj_merge.cu -- merge-path 2-way merge and batched k-way merge of sorted runs.
//
// B200 restatement of
//   * src/merge/merge.c:28-235 (scalar_merge_*, avx_merge_*) and the bitonic merge kernels
//     merge{4,8,16}_{eqlen,varlen} of src/avxsort/avxsort_core.h:111-1211
//   * src/merge/avx_multiwaymerge.c:199-338 avx_multiway_merge (binary tree of cache resident
//     FIFOs)  ->  ceil(log2 k) rounds of pairwise merges, every pair of a round in ONE launch.
//
// Instead of a sequential two-cursor loop (CPU) the output is cut into fixed tiles along the
// merge path: k_mp_partition binary-searches where each tile starts in A and B, k_mp_merge loads
// the two pieces into shared memory, every thread finds its own 8-item slice by a second
// (shared-memory) search and merges it serially; the tile is written back coalesced.
//
// HBM traffic: 16 B per item per round (+ 2 binary searches per 4096-item tile).
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <stdlib.h>
#include <vector>

int smj_msd_sort(SmjCtx *ctx, const u64 *in, u64 n, u64 *out, int *used);
int smj_kway_merge(SmjCtx *ctx, u64 *out, const u64 *const *runs, const u64 *lens, u32 nruns, int *used);

namespace {

constexpr int MP_THREADS = 512;
constexpr int MP_IPT     = 9;  // odd: see k_mp_merge
constexpr int MP_TILE    = MP_THREADS * MP_IPT;  // 4608 items = 36 KB of shared memory

struct MergeJob {
    const u64 *a;
    const u64 *b;
    u64       *out;
    u64        na, nb;
    u64        tile_begin;  // index of this job's first tile in the launch
};

// mode 0: whole-word compare (AVX kernels, src/avxsort/avxcommon.h:186), A first on ties
// mode 1: key-only compare, B first on ties (scalar_merge_tuples, src/merge/merge.c:78-87)
template <int MODE>
__device__ __forceinline__ bool a_goes_first(u64 a, u64 b) {
    if (MODE == 0) return a <= b;
    return (a >> 32) < (b >> 32);
}

// number of A items among the first `diag` items of the merged sequence
template <int MODE, typename PA, typename PB>
__device__ __forceinline__ u64 merge_path(PA a, u64 na, PB b, u64 nb, u64 diag) {
    u64 lo = diag > nb ? diag - nb : 0;
    u64 hi = diag < na ? diag : na;
    while (lo < hi) {
        u64 mid = (lo + hi) >> 1;
        if (a_goes_first<MODE>(a[mid], b[diag - 1 - mid]))
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

__device__ __forceinline__ u32 find_job(const MergeJob *jobs, u32 njobs, u64 tile) {
    u32 lo = 0, hi = njobs;  // last job with tile_begin <= tile
    while (hi - lo > 1) {
        u32 mid = (lo + hi) >> 1;
        if (jobs[mid].tile_begin <= tile)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_mp_partition(const MergeJob *__restrict__ jobs, u32 njobs,
                                                      u64 total_tiles, u64 *__restrict__ diag_a) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_tiles) return;
    const MergeJob j    = jobs[find_job(jobs, njobs, t)];
    u64            diag = (t - j.tile_begin) * MP_TILE;
    diag_a[t]           = merge_path<MODE>(j.a, j.na, j.b, j.nb, diag);
}

// The two pieces of a tile are contiguous ranges of A and of B: ONE thread stages them with two bulk
// copies (TMA, completion on an mbarrier) instead of an LDG + STS per item and thread.  The layout in
// shared memory is therefore plain (no padding); bank conflicts are avoided by the ODD number of items
// per thread: the write-back of a thread's 9 consecutive outputs and the cursors of neighbouring
// threads (4.5 items apart in either piece) spread over the banks by themselves.
template <int MODE>
__global__ void __launch_bounds__(MP_THREADS, 4) k_mp_merge(const MergeJob *__restrict__ jobs,
                                                         u32 njobs, u64 total_tiles,
                                                         const u64 *__restrict__ diag_a) {
    __shared__ __align__(16) u64 s_in[MP_TILE + 4];  // A piece at [misA ..), B piece at [offB ..)
    __shared__ __align__(8) u64 s_bar;
    const u64      t  = blockIdx.x;
    const u32      ji = find_job(jobs, njobs, t);
    const MergeJob j  = jobs[ji];
    const u64 local   = t - j.tile_begin;
    const u64 total   = j.na + j.nb;
    const u64 d0      = local * MP_TILE;
    const u64 d1      = min(d0 + (u64)MP_TILE, total);
    const bool last   = (ji + 1 < njobs) ? (t + 1 == jobs[ji + 1].tile_begin) : (t + 1 == total_tiles);
    const u64 a0 = diag_a[t], a1 = last ? j.na : diag_a[t + 1];
    const u64 b0 = d0 - a0, b1 = d1 - a1;
    const u32 la = (u32)(a1 - a0), lb = (u32)(b1 - b0), cnt = la + lb;
    const u64 *ga = j.a + a0, *gb = j.b + b0;
    const u32 misA = la ? (u32)(((uintptr_t)ga >> 3) & 1) : 0u, misB = lb ? (u32)(((uintptr_t)gb >> 3) & 1) : 0u;
    const u32 evenB = (misA + la + 1) & ~1u;  // 16-byte aligned start of B's staging area
    const u32 offA = misA, offB = evenB + misB;

    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const BulkStage sa_(s_in, ga, la), sb_(s_in + evenB, gb, lb);
        const u32       bytes = sa_.bytes() + sb_.bytes();
        sa_.plain();
        sb_.plain();
        mbar_arrive_expect_tx(&s_bar, bytes);
        sa_.issue(&s_bar);
        sb_.issue(&s_bar);
    }
    mbar_wait(&s_bar, 0);

    const u32 d = min((u32)(threadIdx.x * MP_IPT), cnt);
    u32       ai, bi;
    {  // merge path inside the tile
        u32 lo = d > lb ? d - lb : 0, hi = d < la ? d : la;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (a_goes_first<MODE>(s_in[offA + mid], s_in[offB + d - 1 - mid]))
                lo = mid + 1;
            else
                hi = mid;
        }
        ai = lo;
        bi = d - lo;
    }
    u64 r[MP_IPT];
    // heads in registers: one shared-memory load per output item
    u64 ha = ai < la ? s_in[offA + ai] : 0ull, hb = bi < lb ? s_in[offB + bi] : 0ull;
#pragma unroll
    for (int k = 0; k < MP_IPT; k++) {
        r[k] = 0;
        if (d + k < cnt) {
            const bool take_a = (bi >= lb) || (ai < la && a_goes_first<MODE>(ha, hb));
            r[k]              = take_a ? ha : hb;
            ai += take_a ? 1u : 0u;
            bi += take_a ? 0u : 1u;
            const bool more = take_a ? ai < la : bi < lb;
            const u64  nxt  = more ? s_in[take_a ? offA + ai : offB + bi] : 0ull;
            if (take_a) ha = nxt; else hb = nxt;
        }
    }
    __syncthreads();  // everyone is done reading the inputs: reuse the buffer for the output
    // the merged tile is one contiguous range of the output: it is put together at the 16-byte phase of
    // its destination and leaves as ONE bulk copy shared -> global (instead of an LDS + STG per item)
    u64      *gout = j.out + d0;
    const u32 ph   = (u32)(((uintptr_t)gout >> 3) & 1);
#pragma unroll
    for (int k = 0; k < MP_IPT; k++)
        if (d + k < cnt) s_in[ph + d + k] = r[k];
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_store_items(gout, s_in + ph, cnt);
        bulk_wait_read();
    }
}

int run_jobs(SmjCtx *ctx, const std::vector<MergeJob> &jobs_h, int mode) {
    u64                   total_tiles = 0, total_items = 0;
    std::vector<MergeJob> jobs;
    for (MergeJob j : jobs_h) {
        if (j.na + j.nb == 0) continue;
        total_items += j.na + j.nb;
        j.tile_begin = total_tiles;
        total_tiles += div_up(j.na + j.nb, MP_TILE);
        jobs.push_back(j);
    }
    if (jobs.empty()) return 0;
    SmjScratch sc(ctx);
    MergeJob  *d_jobs = (MergeJob *)sc.get(jobs.size() * sizeof(MergeJob));
    u64       *d_diag = (u64 *)sc.get((total_tiles + 1) * 8);
    if (!d_jobs || !d_diag) return smj_fail_msg("merge: scratch allocation failed");
    SMJ_CHECK(cudaMemcpyAsync(d_jobs, jobs.data(), jobs.size() * sizeof(MergeJob),
                              cudaMemcpyHostToDevice, ctx->stream));
    const u32 nj = (u32)jobs.size();
    const unsigned pgrid = (unsigned)div_up(total_tiles, 256);
    if (mode == 0) {
        SMJ_LAUNCH(ctx, k_mp_partition<0>, pgrid, 256, 0, d_jobs, nj, total_tiles, d_diag);
        SMJ_ITEMS(ctx, total_items);
        SMJ_LAUNCH(ctx, k_mp_merge<0>, (unsigned)total_tiles, MP_THREADS, 0, d_jobs, nj,
                   total_tiles, d_diag);
    } else {
        SMJ_LAUNCH(ctx, k_mp_partition<1>, pgrid, 256, 0, d_jobs, nj, total_tiles, d_diag);
        SMJ_ITEMS(ctx, total_items);
        SMJ_LAUNCH(ctx, k_mp_merge<1>, (unsigned)total_tiles, MP_THREADS, 0, d_jobs, nj,
                   total_tiles, d_diag);
    }
    // jobs_h lives in pageable memory: the async copy above has already staged it
    return 0;
}

}  // namespace

int smj_merge2(SmjCtx *ctx, const u64 *a, u64 na, const u64 *b, u64 nb, u64 *out, int mode) {
    std::vector<MergeJob> jobs(1);
    jobs[0] = {a, b, out, na, nb, 0};
    return run_jobs(ctx, jobs, mode);
}

// runs[i] / lens[i]: host arrays of device pointers.  Result in out.
int smj_multiway_merge(SmjCtx *ctx, u64 *out, u64 *tmp, const u64 *const *runs, const u64 *lens,
                       u32 nruns, int mode) {
    struct Run {
        const u64 *p;
        u64        n;
    };
    std::vector<Run> cur;
    for (u32 i = 0; i < nruns; i++)
        if (lens[i]) cur.push_back({runs[i], lens[i]});
    if (cur.empty()) return 0;
    // Whole-word order, 3 .. 256 runs: ONE pass over DRAM (smj_kway.cu: value-range tiles from regular
    // samples, every tile gathered, put in order in shared memory by an interpolating counting sort
    // and written once; 16 B/item + samples).  It declines tiny inputs and inputs with piles of equal
    // words; then, or under SMJ_MERGE_PATH=sort / rounds, the paths below run (SMJ_MERGE_PATH=kway
    // forces it for 3 and 4 runs too).
    {
        const char *env = getenv("SMJ_MERGE_PATH");
        bool inplace = false;  // a run inside out would be overwritten while other tiles still read it
        u64  total   = 0;
        for (const Run &r : cur) total += r.n;
        for (const Run &r : cur) inplace |= (r.p < out + total && r.p + r.n > out);
        // (up to 16 runs the merge-path rounds of 0.45 ms per 128 M items are faster than the one pass:
        // 8 runs 1.34 against 1.54 ms, 16 runs 1.80 against 1.76; 32 runs 2.25 against 1.75)
        const bool want = env ? env[0] == 'k' : cur.size() > 16;
        if (mode == 0 && want && !inplace) {
            std::vector<const u64 *> ps;
            std::vector<u64>         ls;
            for (const Run &r : cur) {
                ps.push_back(r.p);
                ls.push_back(r.n);
            }
            int used = 0, rc = smj_kway_merge(ctx, out, ps.data(), ls.data(), (u32)cur.size(), &used);
            if (rc) return rc;
            if (used) return 0;
        }
    }
    // More than 16 runs in whole-word order that the one-pass merge declined (or SMJ_MERGE_PATH=sort):
    // the merged sequence is the sorted union of the runs (ties are bit-identical items), and
    // range-partitioning the union MSD-first (smj_msdsort.cu, 48 B/item) beats log2(k) merge-path
    // rounds (16 B/item each): 128 runs of 1 M tuples 3.2 -> 2.2 ms.  The runs are gathered into tmp (they need not be
    // adjacent) and sorted into out; the sort declines what it does not suit and the rounds run.
    {
        u64 total = 0;
        for (const Run &r : cur) total += r.n;
        const char *env = getenv("SMJ_MERGE_PATH");  // "rounds" disables this shortcut
        if (mode == 0 && (env ? env[0] == 's' && cur.size() >= 8 : cur.size() > 16)) {
            u64 off = 0;
            bool aliased = false;  // a run inside tmp or out cannot be gathered in place
            for (const Run &r : cur)
                aliased |= (r.p < tmp + total && r.p + r.n > tmp) || (r.p < out + total && r.p + r.n > out);
            bool adjacent = true;  // runs that lie back to back need no gathering
            for (size_t i = 0; i + 1 < cur.size(); i++) adjacent &= cur[i].p + cur[i].n == cur[i + 1].p;
            const u64 *src = cur[0].p;
            if (adjacent) aliased = src < out + total && src + total > out;
            if (!aliased) {
                if (!adjacent) {
                    for (const Run &r : cur) {
                        SMJ_CHECK(cudaMemcpyAsync(tmp + off, r.p, r.n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                        off += r.n;
                    }
                    src = tmp;
                }
                int used = 0, rc = smj_msd_sort(ctx, src, total, out, &used);
                if (rc) return rc;
                if (used) return 0;
            }
        }
    }
    int rounds = 0;
    while ((1ull << rounds) < cur.size()) rounds++;
    if (rounds == 0) rounds = 1;  // a single run is copied by a merge with an empty partner
    for (int r = 1; r <= rounds; r++) {
        u64 *dst = ((rounds - r) % 2 == 0) ? out : tmp;
        std::vector<MergeJob> jobs;
        std::vector<Run>      next;
        u64                   off = 0;
        for (size_t i = 0; i < cur.size(); i += 2) {
            MergeJob j;
            j.a   = cur[i].p;
            j.na  = cur[i].n;
            j.b   = (i + 1 < cur.size()) ? cur[i + 1].p : nullptr;
            j.nb  = (i + 1 < cur.size()) ? cur[i + 1].n : 0;
            j.out = dst + off;
            j.tile_begin = 0;
            jobs.push_back(j);
            next.push_back({dst + off, j.na + j.nb});
            off += j.na + j.nb;
        }
        int rc = run_jobs(ctx, jobs, mode);
        if (rc) return rc;
        cur.swap(next);
    }
    return 0;
}

extern "C" int smjdev_merge(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb,
                            uint64_t *out, int mode) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    return smj_merge2(ctx, (const u64 *)a, na, (const u64 *)b, nb, (u64 *)out, mode);
}

extern "C" int smjdev_multiway_merge(uint64_t *out, uint64_t *tmp, const uint64_t *const *runs,
                                     const uint64_t *lens, uint32_t nruns, int mode) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    std::vector<const u64 *> r(nruns);
    std::vector<u64>         l(nruns);
    for (u32 i = 0; i < nruns; i++) {
        r[i] = (const u64 *)runs[i];
        l[i] = lens[i];
    }
    return smj_multiway_merge(ctx, (u64 *)out, (u64 *)tmp, r.data(), l.data(), nruns, mode);
}
