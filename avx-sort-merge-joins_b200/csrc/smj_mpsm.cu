// smj_mpsm.cu -- the join phase of P-MPSM: every sorted run of the public input S against the
// sorted private input R, in ONE pass over R.
//
// src/joins/sortmergejoin_mpsm.h:32-33 (a stub in the reference; the algorithm is Albutiu et al.'s
// P-MPSM as the reference's README describes it): R is sorted completely, S only run by run, and
// every S run is merge-joined against R.  Doing that literally costs one scan of R per run
// (16 runs of 8 M tuples against 128 M tuples of R: 16 x 0.49 ms, all of it reading R again).
// Here the scan is cut along R instead: a tile is a range of ~3000 consecutive R tuples that starts
// and ends on a key boundary, staged in shared memory by one TMA bulk copy, and the pieces of ALL
// runs that fall into the tile's key range -- found by binary search, one search per (tile, run) --
// stream past it: a warp takes a piece, every lane looks its S tuple up in the R tile by binary
// search (lower bound; the equal-key R run is walked, it is one tuple for a primary-key R) and adds
//     count += |run_R(key)|,  checksum += sum A(run_R(key)) * B(s)
// exactly like the merge-join scan (smj_join.cu).  R is read once, S once.
// An R run longer than the tile capacity (a key with thousands of duplicates in R) cannot be kept in
// one tile: the partition kernel raises a flag and the caller falls back to one scan per run.
#include "smj_internal.cuh"

#include <algorithm>
#include <vector>

namespace {
constexpr int MP_MAXRUNS = 256;                           // S runs handled in one pass
constexpr int PB_THREADS = 256, PB_TR = 3072, PB_CAP = 4096;  // R tuples per tile (nominal / capacity)

struct MpsmRuns {
    const u64 *p[MP_MAXRUNS];
    u64        n[MP_MAXRUNS];
};

__device__ __forceinline__ u32 kkey(u64 x) { return (u32)(x >> 32); }

// first index in a[0..n) whose key is >= k
__device__ __forceinline__ u64 lower_key64(const u64 *__restrict__ a, u64 n, u32 k) {
    u64 lo = 0, hi = n;
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (kkey(a[mid]) < k) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Tile t starts at the first R tuple whose key equals key(R[t * PB_TR]) (so that no R run straddles a
// tile).  bounds[t * (k + 1)] = that R index, bounds[t * (k + 1) + 1 + i] = first tuple of run i whose
// key is >= that key; row ntiles = the ends.  A tile with more than PB_CAP R tuples raises *flag.
__global__ void __launch_bounds__(256) k_mpsm_partition(const u64 *__restrict__ r, u64 nr,
                                                        const MpsmRuns *__restrict__ runs, u32 k, u64 ntiles,
                                                        u64 *__restrict__ bounds, u32 *__restrict__ flag) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (ntiles + 1) * (k + 1)) return;
    const u64 t = idx / (k + 1);
    const u32 c = (u32)(idx % (k + 1));
    u64       v;
    if (t == ntiles) {
        v = c == 0 ? nr : runs->n[c - 1];
    } else if (t == 0) {
        v = 0;  // everything below R's first key has no partner anyway, but the first tile takes it
    } else {
        const u32 key = kkey(r[t * (u64)PB_TR]);
        if (c == 0) {
            v = lower_key64(r, t * (u64)PB_TR + 1, key);
            // the previous tile then ends here: it holds [start of t - 1, v); too long -> not this path
        } else {
            v = lower_key64(runs->p[c - 1], runs->n[c - 1], key);
        }
    }
    bounds[idx] = v;
    if (c == 0 && t > 0) {
        // tile t - 1 ends here; its length needs its own start: recompute it (cheap, and keeps the kernel one pass)
        const u64 prev = (t - 1 == 0) ? 0ull : lower_key64(r, (t - 1) * (u64)PB_TR + 1, kkey(r[(t - 1) * (u64)PB_TR]));
        if (v - prev > (u64)PB_CAP) *flag = 1u;
    }
}

// A tile whose pieces of S add up to more than PB_SMAX tuples (a heavy hitter: every S tuple of a hot
// key meets in ONE tile, i.e. one CTA) would serialise the pass: raise the flag, the merge-path scans
// (which split the work evenly whatever the skew) take over.
constexpr u64 PB_SMAX = 1ull << 20;
__global__ void __launch_bounds__(256) k_mpsm_check(const u64 *__restrict__ bounds, u32 k, u64 ntiles, u32 *__restrict__ flag) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const u64 *row0 = bounds + t * (k + 1), *row1 = row0 + (k + 1);
    u64        tot = 0;
    for (u32 i = 1; i <= k; i++) tot += row1[i] - row0[i];
    if (tot > PB_SMAX) *flag = 1u;
}

__global__ void __launch_bounds__(PB_THREADS, 4)
    k_mpsm_probe(const u64 *__restrict__ r, const MpsmRuns *__restrict__ runs, u32 k, const u64 *__restrict__ bounds,
                 u64 *__restrict__ acc) {
    __shared__ __align__(16) u64 s_r[PB_CAP + 4];
    __shared__ __align__(8) u64  s_bar;
    __shared__ u64               s_red[2][PB_THREADS / 32];
    const u64  t    = blockIdx.x;
    const u64 *row0 = bounds + t * (k + 1), *row1 = row0 + (k + 1);
    const u64  r0 = row0[0], r1 = row1[0];
    const u32  la = (u32)(r1 - r0);
    const u32  warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u64 *gr   = r + r0;
    const u32  mis  = la ? (u32)(((uintptr_t)gr >> 3) & 1) : 0u;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        const BulkStage st(s_r, gr, la);
        st.plain();
        mbar_arrive_expect_tx(&s_bar, st.bytes());
        st.issue(&s_bar);
    }
    mbar_wait(&s_bar, 0);
    const u64 *sr = s_r + mis;  // the tile: la R tuples in key order, no run of equal keys is cut
    u64        my_cnt = 0, my_cks = 0;
    if (la) {
        // pieces: run i contributes [row0[1 + i], row1[1 + i]); a warp takes a piece, 32 tuples at a time
        for (u32 i = warp; i < k; i += PB_THREADS / 32) {
            const u64  s0 = row0[1 + i], s1 = row1[1 + i];
            const u64 *gs = runs->p[i];
            for (u64 j = s0 + lane; j < s1; j += 32) {
                const u64 y   = ld_stream(gs + j);
                const u32 key = kkey(y);
                u32       lo = 0, hi = la;  // lower bound of key in the tile
                while (lo < hi) {
                    const u32 mid = (lo + hi) >> 1;
                    if (kkey(sr[mid]) < key) lo = mid + 1; else hi = mid;
                }
                u64 sa = 0;
                u32 cr = 0;
                while (lo < la && kkey(sr[lo]) == key) {  // the R run of this key (one tuple for a primary key)
                    sa += smj_cks_a(sr[lo]);
                    cr++, lo++;
                }
                if (cr) {
                    my_cnt += cr;
                    my_cks += sa * smj_cks_b(y);
                }
            }
        }
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
        for (int w = 0; w < PB_THREADS / 32; w++) {
            c += s_red[0][w];
            x += s_red[1][w];
        }
        if (c) atomicAdd(acc, c);
        if (x) atomicAdd(acc + 1, x);
    }
}
}  // namespace

// acc[0..1] += {count, checksum} of R join (S_0 u ... u S_{k-1}); r and every run are key-sorted device
// arrays.  *used = 0 when the path does not apply (too many runs, an R run longer than a tile): nothing
// was added then and the caller scans run by run.  Synchronises once (the flag).
int smj_mpsm_probe(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *const *runs, const u64 *lens, u32 nruns, u64 *acc,
                   int *used) {
    *used = 0;
    if (nruns == 0 || nruns > (u32)MP_MAXRUNS || nr < (u64)PB_TR || getenv("SMJ_MPSM_PER_RUN")) return 0;
    MpsmRuns h;
    memset(&h, 0, sizeof(h));
    for (u32 i = 0; i < nruns; i++) h.p[i] = runs[i], h.n[i] = lens[i];
    const u64  ntiles = div_up(nr, PB_TR);
    SmjScratch sc(ctx);
    MpsmRuns *d_runs = (MpsmRuns *)sc.get(sizeof(MpsmRuns));
    u64      *d_bounds = (u64 *)sc.get((ntiles + 1) * (nruns + 1) * 8);
    u32      *d_flag = (u32 *)sc.get(4);
    if (!d_runs || !d_bounds || !d_flag) return smj_fail_msg("smj_mpsm_probe: out of device memory");
    // (h is pageable memory: the driver has staged it when the call returns)
    SMJ_CHECK(cudaMemcpyAsync(d_runs, &h, sizeof(MpsmRuns), cudaMemcpyHostToDevice, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_flag, 0, 4, ctx->stream));
    SMJ_LAUNCH(ctx, k_mpsm_partition, (unsigned)div_up((ntiles + 1) * (nruns + 1), 256), 256, 0, r, nr, d_runs, nruns, ntiles,
               d_bounds, d_flag);
    SMJ_LAUNCH(ctx, k_mpsm_check, (unsigned)div_up(ntiles, 256), 256, 0, d_bounds, nruns, ntiles, d_flag);
    u32 h_flag = 0;
    SMJ_CHECK(cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    if (h_flag) return 0;  // an R run does not fit a tile, or a hot key of S would land on one CTA
    u64 total = nr;
    for (u32 i = 0; i < nruns; i++) total += lens[i];
    SMJ_ITEMS(ctx, total);
    SMJ_LAUNCH(ctx, k_mpsm_probe, (unsigned)ntiles, PB_THREADS, 0, r, d_runs, nruns, d_bounds, acc);
    *used = 1;
    return 0;
}
