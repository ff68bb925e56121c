// smj_materialize.cu -- the join with its result written out.
//
// The reference only counts unless it is built with JOIN_MATERIALIZE, and then its merge_join
// appends one tuple <S.key, S.payload> per matching (r, s) pair to a per-thread chained buffer
// (src/joins/joincommon.c:256-278; that build lacks tuple_buffer.h in the reference's tree -- the
// parity tests compile it with a stand-in header, oracle/shim/, and compare the emitted tuples).  Here:
//
//   1. R and S are sorted by key (LSD radix sort, smj_lsd_sort).
//   2. S is cut into chunks of 1024 tuples; k_mat_bounds finds, for every chunk, the range of R
//      that can match it (two binary searches per chunk, all chunks in parallel).
//   3. k_mat_emit, one CTA per chunk: every S tuple finds its run of equal keys in the chunk's R
//      range (binary search over the R keys staged in shared memory), a block scan turns the run
//      lengths into output offsets, the chunk reserves its slice of the output with one atomicAdd
//      and the pairs are written with one thread per OUTPUT tuple, so a heavy run is spread over
//      the whole CTA.  Output order is unspecified (the reference's is per thread, too).
//
// Layouts: SMJ_OUT_S_TUPLE = the reference's <S.key, S.payload> per match (one 64-bit item),
//          SMJ_OUT_RID_PAIR = <R.payload, S.payload> (row-id pair: R.payload << 32 | S.payload).
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>

int smj_lsd_sort(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, u64 **sorted);
int smj_msd_sort(SmjCtx *ctx, const u64 *in, u64 n, u64 *out, int *used);

namespace {

constexpr int MT_THREADS = 256, MT_S = 1024, MT_SPT = MT_S / MT_THREADS, MT_RCAP = 2048;

__device__ __forceinline__ u32 key_of(u64 x) { return (u32)(x >> 32); }

// first index in r[lo, hi) whose key is >= k (UPPER: > k)
template <bool UPPER>
__device__ __forceinline__ u64 bound_global(const u64 *__restrict__ r, u64 lo, u64 hi, u32 k) {
    while (lo < hi) {
        const u64 mid = lo + ((hi - lo) >> 1);
        const u32 km  = key_of(r[mid]);
        if (UPPER ? km <= k : km < k)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

// bounds[c] = first R index that can match chunk c, bounds2[c] = one past the last
__global__ void __launch_bounds__(256) k_mat_bounds(const u64 *__restrict__ r, u64 nr,
                                                    const u64 *__restrict__ s, u64 ns, u64 nchunks,
                                                    u64 *__restrict__ rlo, u64 *__restrict__ rhi) {
    const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const u64 j0 = c * MT_S, j1 = min(ns, j0 + MT_S);
    rlo[c] = bound_global<false>(r, 0, nr, key_of(s[j0]));
    rhi[c] = bound_global<true>(r, 0, nr, key_of(s[j1 - 1]));
}

__global__ void __launch_bounds__(MT_THREADS) k_mat_emit(const u64 *__restrict__ r,
                                                         const u64 *__restrict__ s, u64 ns,
                                                         const u64 *__restrict__ rlo_,
                                                         const u64 *__restrict__ rhi_, int layout,
                                                         u64 *__restrict__ out, u64 capacity,
                                                         u64 *total) {
    __shared__ u64 s_item[MT_S];   // the chunk's S tuples
    __shared__ u64 s_off[MT_S];    // exclusive scan of the run lengths
    __shared__ u32 s_lb[MT_S];     // start of each S tuple's run, relative to rlo
    __shared__ u32 s_rkey[MT_RCAP];
    __shared__ u64 s_wsum[MT_THREADS / 32];
    __shared__ u64 s_base;
    const u64 c = blockIdx.x, j0 = c * MT_S;
    const u32 m = (u32)min((u64)MT_S, ns - j0);
    const u64 rlo = rlo_[c], rhi = rhi_[c];
    const u64 rn    = rhi - rlo;
    const bool staged = rn <= (u64)MT_RCAP;
    if (staged)
        for (u32 i = threadIdx.x; i < (u32)rn; i += MT_THREADS) s_rkey[i] = key_of(r[rlo + i]);
    for (u32 j = threadIdx.x; j < m; j += MT_THREADS) s_item[j] = s[j0 + j];
    __syncthreads();
    // run of equal keys per S tuple: thread t owns the CONSECUTIVE tuples t*SPT .. t*SPT+SPT-1
    u64 cnt[MT_SPT], tsum = 0;
#pragma unroll
    for (int q = 0; q < MT_SPT; q++) {
        const u32 j = threadIdx.x * MT_SPT + q;
        cnt[q]      = 0;
        if (j < m && rn) {
            const u32 k = key_of(s_item[j]);
            u64       lb, ub;
            if (staged) {
                u32 lo = 0, hi = (u32)rn;
                while (lo < hi) {
                    const u32 mid = (lo + hi) >> 1;
                    if (s_rkey[mid] < k) lo = mid + 1; else hi = mid;
                }
                lb = lo;
                hi = (u32)rn;
                while (lo < hi) {
                    const u32 mid = (lo + hi) >> 1;
                    if (s_rkey[mid] <= k) lo = mid + 1; else hi = mid;
                }
                ub = lo;
            } else {  // a long R range (duplicates in R): search it where it is
                lb = bound_global<false>(r, rlo, rhi, k) - rlo;
                ub = bound_global<true>(r, rlo + lb, rhi, k) - rlo;
            }
            s_lb[j] = (u32)lb;
            cnt[q]  = ub - lb;
        }
        tsum += cnt[q];
    }
    // block exclusive scan of the run lengths
    u64       incl = tsum;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    u64 excl = incl - tsum, T = 0;
    for (u32 w = 0; w < MT_THREADS / 32; w++) {
        if (w < warp) excl += s_wsum[w];
        T += s_wsum[w];
    }
#pragma unroll
    for (int q = 0; q < MT_SPT; q++) {
        const u32 j = threadIdx.x * MT_SPT + q;
        if (j < m) s_off[j] = excl;
        excl += cnt[q];
    }
    if (threadIdx.x == 0) s_base = T ? atomicAdd(total, T) : 0ull;
    __syncthreads();
    if (!out || T == 0) return;
    const u64 base = s_base;
    // one thread per output tuple: its S tuple is the last one whose offset is <= o
    for (u64 o = threadIdx.x; o < T; o += MT_THREADS) {
        if (base + o >= capacity) break;
        u32 lo = 0, hi = m;  // invariant: s_off[lo] <= o, answer in [lo, hi)
        while (hi - lo > 1) {
            const u32 mid = (lo + hi) >> 1;
            if (s_off[mid] <= o) lo = mid; else hi = mid;
        }
        // tuples without a partner share their offset with the next one: take the last such index
        const u64 sitem = s_item[lo];
        u64       v     = sitem;
        if (layout == SMJ_OUT_RID_PAIR) {
            const u64 ritem = r[rlo + s_lb[lo] + (o - s_off[lo])];
            v               = ((ritem & 0xffffffffull) << 32) | (sitem & 0xffffffffull);
        }
        out[base + o] = v;
    }
}

}  // namespace

extern "C" int smjdev_join_materialize(const uint64_t *r_, uint64_t nr, const uint64_t *s_,
                                       uint64_t ns, int layout, uint64_t *out, uint64_t capacity,
                                       uint64_t *count) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (!count) return smj_fail_msg("smjdev_join_materialize: count must not be NULL");
    if (layout != SMJ_OUT_S_TUPLE && layout != SMJ_OUT_RID_PAIR)
        return smj_fail_msg("smjdev_join_materialize: bad layout");
    *count = 0;
    if (nr == 0 || ns == 0) return 0;
    const u64 *r = (const u64 *)r_, *s = (const u64 *)s_;
    SmjScratch sc(ctx);
    u64 *aR = (u64 *)sc.get(nr * 8), *bR = (u64 *)sc.get(nr * 8);
    u64 *aS = (u64 *)sc.get(ns * 8), *bS = (u64 *)sc.get(ns * 8);
    const u64 nchunks = div_up(ns, MT_S);
    u64 *d_lo = (u64 *)sc.get(nchunks * 8), *d_hi = (u64 *)sc.get(nchunks * 8), *d_total = (u64 *)sc.get(8);
    if (!aR || !bR || !aS || !bS || !d_lo || !d_hi || !d_total)
        return smj_fail_msg("smjdev_join_materialize: out of device memory");
    u64 *sr = nullptr, *ss = nullptr;
    int  rc;
    // key order is all that is needed; the MSD sort (whole-word order) is the faster one when it applies
    int used = 0;
    if ((rc = smj_msd_sort(ctx, r, nr, aR, &used))) return rc;
    sr = aR;
    if (!used && (rc = smj_lsd_sort(ctx, r, aR, bR, nr, 2, &sr))) return rc;
    if ((rc = smj_msd_sort(ctx, s, ns, aS, &used))) return rc;
    ss = aS;
    if (!used && (rc = smj_lsd_sort(ctx, s, aS, bS, ns, 2, &ss))) return rc;
    SMJ_CHECK(cudaMemsetAsync(d_total, 0, 8, ctx->stream));
    SMJ_LAUNCH(ctx, k_mat_bounds, (unsigned)div_up(nchunks, 256), 256, 0, sr, nr, ss, ns, nchunks, d_lo, d_hi);
    SMJ_ITEMS(ctx, nr + ns);
    SMJ_LAUNCH(ctx, k_mat_emit, (unsigned)nchunks, MT_THREADS, 0, sr, ss, ns, d_lo, d_hi, layout, (u64 *)out,
               out ? capacity : 0ull, d_total);
    u64 h = 0;
    SMJ_CHECK(cudaMemcpyAsync(&h, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    *count = h;
    return 0;
}
