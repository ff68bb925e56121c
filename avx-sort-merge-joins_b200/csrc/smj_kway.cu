// smj_kway.cu -- one-pass k-way merge of sorted runs (whole-word order).
//
// B200 restatement of src/merge/avx_multiwaymerge.c:199-338 avx_multiway_merge: there a binary
// tree of cache-resident FIFOs merges k runs in ONE pass over DRAM (read 8 + write 8 bytes per
// tuple, the tree traffic stays in L3).  Here the one pass is cut along the VALUE axis:
//   1. k_kw_sample   every s-th item of every run (the last item of each block of s) -> samples
//   2. the samples are sorted (they are n / s items: the library's own sort)
//   3. every q-th sorted sample is a splitter; tile t holds the values in (v_t, v_t+1].  Regular
//      sampling bounds a tile: a run contributes at most (c + 1) s items when c of its samples fall
//      into the interval, so a tile holds at most (q + k) s items, and q s items give or take
//      s sqrt(k / 6) on average -- q is chosen so that this fits shared memory with a wide margin
//   4. k_kw_partition   upper_bound of every splitter in every run (one thread per pair): tile t
//      of the output is the concatenation of k contiguous pieces, run j's piece is
//      [pos[t][j], pos[t+1][j])
//   5. k_kw_tile_count   one CTA per tile: the k pieces are gathered into shared memory (a warp per
//      batch of 8 pieces, the 8 loads in flight together) and put in order by an INTERPOLATING
//      COUNTING SORT: slot = (x - min) >> shift with 2^13 slots over the tile's own value range,
//      count / scan / place with shared-memory atomics, the few items that share a slot ordered by
//      insertion -- O(1) work per item instead of the O(log^2) of a comparison sort -- and written
//      to their final place with coalesced stores.  A tile whose values are not spread evenly (a
//      slot with more than 64 items) is put on a list and redone by
//   6. k_kw_tile   the same tile sorted in shared memory by compare-exchange (register network +
//      merge-path rounds, smj_smemsort.cuh): any distribution, ~6x the time.
// HBM traffic: 16 B per item + the samples (8 / s B per item read, sorted) + the splitter searches
// (k log n loads per tile, mostly L2).  Ties are bit-identical words, so the merged sequence is
// the sorted union whatever the tie order.  Many equal words can overfill a tile (the bound
// assumes distinct sample values at the splitters): the tile kernel then raises a flag and the
// caller falls back to the pairwise rounds.
#include "smj_internal.cuh"
#include "smj_smemsort.cuh"

#include <algorithm>
#include <math.h>
#include <vector>

namespace {
constexpr int KW_MAXK = 256, KW_THREADS = 512, KW_CAP = KW_THREADS * SS_IPT;  // 8192 items = 64 KB

struct KwRuns {  // device resident
    const u64 *p[KW_MAXK];
    u64        n[KW_MAXK];
    u64        samp_off[KW_MAXK + 1];  // first sample of every run in the sample array
};

__global__ void __launch_bounds__(256) k_kw_sample(const KwRuns *__restrict__ runs, u32 k, u32 s, u64 m,
                                                   u64 *__restrict__ samples) {
    __shared__ u64 s_off[KW_MAXK + 1];
    for (u32 i = threadIdx.x; i <= k; i += blockDim.x) s_off[i] = runs->samp_off[i];
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < m; g += stride) {
        u32 lo = 0, hi = k;  // last run with samp_off <= g
        while (hi - lo > 1) {
            const u32 mid = (lo + hi) >> 1;
            if (s_off[mid] <= g) lo = mid; else hi = mid;
        }
        samples[g] = runs->p[lo][(g - s_off[lo] + 1) * s - 1];
    }
}

// pos[(t * k) + j] = number of items of run j that are <= splitter t (t = 1 .. ntiles-1),
// row 0 = 0, row ntiles = the run lengths.  A thread owns KW_GROUP consecutive splitters of one run:
// the first one costs a binary search over the run, the following ones gallop on from the previous
// answer (a piece holds ~ tile / k items: two or three probes and a short search, in cache lines
// that are already there) -- a third of the loads of one full search per (splitter, run) pair.
constexpr u32 KW_GROUP = 16;
__global__ void __launch_bounds__(256) k_kw_partition(const KwRuns *__restrict__ runs, u32 k, u32 q, u32 off, u64 m,
                                                      u64 ntiles, const u64 *__restrict__ sorted_samples,
                                                      u64 *__restrict__ pos) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 ngroups = (ntiles + KW_GROUP) / KW_GROUP;  // rows 0 .. ntiles
    if (idx >= ngroups * k) return;
    // the threads of a warp search ONE run at neighbouring places: same pages, shared upper search levels
    const u32  j = (u32)(idx / ngroups);
    const u64  g = idx % ngroups;
    const u64  n = runs->n[j];
    const u64 *a = runs->p[j];
    u64        prev = 0;
    for (u32 u = 0; u < KW_GROUP; u++) {
        const u64 t = g * KW_GROUP + u;
        if (t > ntiles) break;
        u64 r = 0;
        if (t == ntiles) {
            r = n;
        } else if (t > 0) {
            const u64 v = sorted_samples[min(t * q - off, m) - 1];  // tile t starts after that many samples
            u64       lo = prev, hi = n;
            if (u > 0) {  // gallop: the answer is >= prev, usually by a few dozen items
                u64 step = 32;
                while (lo + step <= n && a[lo + step - 1] <= v) {
                    lo += step;
                    step <<= 1;
                }
                hi = min(n, lo + step - 1);  // a[hi] > v (or hi == n)
            }
            while (lo < hi) {
                const u64 mid = (lo + hi) >> 1;
                if (a[mid] <= v) lo = mid + 1; else hi = mid;
            }
            r = lo;
        }
        pos[t * k + j] = r;
        prev           = r;
    }
}

// list != null: CTA b takes tile list[b] (the tiles the counting kernel gave up on)
__global__ void __launch_bounds__(KW_THREADS) k_kw_tile(const KwRuns *__restrict__ runs, u32 k,
                                                        const u64 *__restrict__ pos, u64 *__restrict__ out,
                                                        u32 *__restrict__ overflow, const u32 *__restrict__ list) {
    extern __shared__ __align__(16) u64 s_tile[];  // KW_CAP items
    __shared__ u64 s_lo[KW_MAXK];
    __shared__ u32 s_len[KW_MAXK], s_off[KW_MAXK + 1];
    __shared__ u64 s_base;
    const u64 t = list ? list[blockIdx.x] : blockIdx.x;
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // piece bounds; base = items in the tiles before this one
    u64 before = 0;
    for (u32 j = tid; j < k; j += KW_THREADS) {
        const u64 a = pos[t * k + j], b = pos[(t + 1) * k + j];
        s_lo[j]  = a;
        s_len[j] = (u32)min(b - a, (u64)(KW_CAP + 1));
        before += a;
    }
    before = warp_reduce_add(before);
    __shared__ u64 s_w[KW_THREADS / 32];
    if (lane == 0) s_w[warp] = before;
    __syncthreads();
    if (tid == 0) {
        u64 tot = 0;
        for (int w = 0; w < KW_THREADS / 32; w++) tot += s_w[w];
        s_base  = tot;
        u32 acc = 0;
        for (u32 j = 0; j < k; j++) {  // k <= 256: a serial scan is a few hundred cycles
            s_off[j] = acc;
            acc      = min(acc + s_len[j], (u32)KW_CAP + 1);
        }
        s_off[k] = acc;
    }
    __syncthreads();
    const u32 size = s_off[k];
    if (size > (u32)KW_CAP) {  // equal words piled up between two splitters: the caller takes the rounds
        if (tid == 0) *overflow = 1u;
        return;
    }
    // gather: a warp per piece
    for (u32 j = warp; j < k; j += KW_THREADS / 32) {
        const u64 *src = runs->p[j] + s_lo[j];
        const u32  len = s_len[j], off = s_off[j];
        for (u32 i = lane; i < len; i += 32) s_tile[SSP(off + i)] = ld_stream(src + i);
    }
    for (u32 i = size + tid; i < (u32)KW_CAP; i += KW_THREADS) s_tile[SSP(i)] = ~0ull;
    __syncthreads();
    ss_sort_in_smem(s_tile, size);
    u64 *dst = out + s_base;
    for (u32 i = tid; i < size; i += KW_THREADS) st_stream(dst + i, s_tile[SSP(i)]);
}
// ---- the tile kernel of the default path: interpolating counting sort -------------------------
constexpr int KC_SLOT_BITS = 13, KC_SLOTS = 1 << KC_SLOT_BITS, KC_PER = KC_SLOTS / KW_THREADS, KC_MAXRUN = 64;
// a thread's KC_PER = 16 consecutive slots lie 17 words from its neighbour's (no bank conflicts in the scan)
#define KCP(sl) ((sl) + ((sl) >> 4))
constexpr size_t KC_SMEM = (size_t)(KW_CAP + 2) * 8 + (size_t)(KCP(KC_SLOTS) + 1) * 4;

__global__ void __launch_bounds__(KW_THREADS, 2)
    k_kw_tile_count(const KwRuns *__restrict__ runs, u32 k, const u64 *__restrict__ pos, u64 *__restrict__ out,
                    u32 *__restrict__ state /* [0] a tile overflowed, [1] tiles on the redo list */,
                    u32 *__restrict__ redo) {
    extern __shared__ __align__(16) u64 s_dyn[];
    u64 *stage = s_dyn;                      // [KW_CAP + 2] the tile: gathered, then in order (at the phase of its destination)
    u32 *cnt   = (u32 *)(s_dyn + KW_CAP + 2);  // [KCP(KC_SLOTS) + 1] counts, then cursors
    __shared__ u64 s_lo[KW_MAXK], s_w64[KW_THREADS / 32], s_mm[2], s_base;
    __shared__ u32 s_len[KW_MAXK], s_off[KW_MAXK + 1], s_coff[KW_MAXK + 1], s_w[KW_THREADS / 32], s_wc[KW_THREADS / 32], s_bad;
    const u64 t = blockIdx.x;
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // piece bounds (k <= 256 < threads): exclusive scan of the lengths, base = items in earlier tiles
    u64 a = 0;
    u32 len = 0;
    if (tid < k) {
        a = pos[t * k + tid];
        const u64 b = pos[(t + 1) * k + tid];
        len        = (u32)min(b - a, (u64)(KW_CAP + 1));
        s_lo[tid]  = a;
        s_len[tid] = len;
    }
    const u32 nchunk = (len + 31) >> 5;
    u32       incl = len, cincl = nchunk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(SMJ_FULL_MASK, incl, o), c = __shfl_up_sync(SMJ_FULL_MASK, cincl, o);
        if (lane >= (u32)o) incl += v, cincl += c;
    }
    const u64 asum = warp_reduce_add(a);
    if (lane == 31) s_w[warp] = incl, s_wc[warp] = cincl;
    if (lane == 0) s_w64[warp] = asum;
    for (u32 i = tid; i < (u32)KCP(KC_SLOTS) + 1; i += KW_THREADS) cnt[i] = 0;
    if (tid == 0) {
        s_mm[0] = ~0ull;
        s_mm[1] = 0;
        s_bad   = 0;
    }
    __syncthreads();
    {
        u32 excl = incl - len, cexcl = cincl - nchunk;
        for (u32 w = 0; w < warp; w++) excl += s_w[w], cexcl += s_wc[w];
        if (tid < k) s_off[tid] = excl, s_coff[tid] = cexcl;
        if (tid == k - 1) s_off[k] = excl + len, s_coff[k] = cexcl + nchunk;
        if (tid == 0) {
            u64 tot = 0;
            for (int w = 0; w < KW_THREADS / 32; w++) tot += s_w64[w];
            s_base = tot;
        }
    }
    __syncthreads();
    const u32 size = s_off[k];
    if (size > (u32)KW_CAP) {  // equal words piled up between two splitters: the caller takes another path
        if (tid == 0) state[0] = 1u;
        return;
    }
    // the tile is one contiguous range of the output: it is put in order at the 16-byte phase of its
    // destination and leaves as ONE bulk copy shared -> global (TMA) instead of an LDS + STG per item
    u64 *dst = out + s_base;
    stage += ((uintptr_t)dst >> 3) & 1;
    // gather -- three shapes, each the fastest in its range of k (measured, 128 M items: k = 4 1.45 ms,
    // 16 1.73, 32 1.79, 64 1.77, 128 1.85, 256 2.13):
    if (k <= 8) {
        // a few long pieces: the whole CTA copies one piece after the other
        for (u32 j = 0; j < k; j++) {
            const u64 *src = runs->p[j] + s_lo[j];
            const u32  ln = s_len[j], off = s_off[j];
            for (u32 i = tid; i < ln; i += KW_THREADS) stage[off + i] = ld_stream(src + i);
        }
    } else if (k < 64) {
        // chunks of 32 consecutive items of one piece, numbered through all pieces (s_coff): warp w takes
        // chunks w, w + 16, ... four at a time and finds a chunk's piece by binary search
        const u32 nch = s_coff[k];
        for (u32 c0 = warp; c0 < nch; c0 += (KW_THREADS / 32) * 4) {
            u64 x[4];
            u32 dst[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const u32 c = c0 + u * (KW_THREADS / 32);
                dst[u]      = 0xffffffffu;
                x[u]        = 0;
                if (c < nch) {
                    u32 lo = 0, hi = k;  // last piece whose first chunk is <= c
                    while (hi - lo > 1) {
                        const u32 mid = (lo + hi) >> 1;
                        if (s_coff[mid] <= c) lo = mid; else hi = mid;
                    }
                    const u32 i = (c - s_coff[lo]) * 32 + lane;
                    if (i < s_len[lo]) {
                        x[u]   = ld_stream(runs->p[lo] + s_lo[lo] + i);
                        dst[u] = s_off[lo] + i;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (dst[u] != 0xffffffffu) stage[dst[u]] = x[u];
        }
    } else {
        // many short pieces: a warp takes 8 pieces at a time, their first 32 items in flight together
        for (u32 j0 = warp * 8; j0 < k; j0 += (KW_THREADS / 32) * 8) {
            u64 x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const u32 j = j0 + u;
                x[u]        = 0;
                if (j < k && lane < s_len[j]) x[u] = ld_stream(runs->p[j] + s_lo[j] + lane);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const u32 j = j0 + u;
                if (j < k && lane < s_len[j]) stage[s_off[j] + lane] = x[u];
            }
            for (int u = 0; u < 8; u++) {
                const u32 j = j0 + u;
                if (j < k && s_len[j] > 32) {
                    const u64 *src = runs->p[j] + s_lo[j];
                    const u32  ln = s_len[j], off = s_off[j];
                    for (u32 i = lane + 32; i < ln; i += 32) stage[off + i] = ld_stream(src + i);
                }
            }
        }
    }
    __syncthreads();
    // into registers (thread-strided, conflict-free) + the tile's value range
    u64 x[SS_IPT], mn = ~0ull, mx = 0;
#pragma unroll
    for (int j = 0; j < SS_IPT; j++) {
        const u32 i = tid + j * KW_THREADS;
        x[j]        = 0;
        if (i < size) {
            x[j] = stage[i];
            mn   = min(mn, x[j]);
            mx   = max(mx, x[j]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(SMJ_FULL_MASK, mn, o));
        mx = max(mx, __shfl_xor_sync(SMJ_FULL_MASK, mx, o));
    }
    if (lane == 0) {
        atomicMin((unsigned long long *)&s_mm[0], (unsigned long long)mn);
        atomicMax((unsigned long long *)&s_mm[1], (unsigned long long)mx);
    }
    __syncthreads();  // everybody holds its items: the staging area is free again
    const u64 lo = s_mm[0], range = s_mm[1] - lo;
    const u32 nbits = range ? 64u - (u32)__clzll((long long)range) : 0u;
    const u32 shift = nbits > (u32)KC_SLOT_BITS ? nbits - KC_SLOT_BITS : 0u;  // slot = (x - lo) >> shift < 2^13, monotone in x
#pragma unroll
    for (int j = 0; j < SS_IPT; j++)
        if (tid + j * KW_THREADS < size) atomicAdd(&cnt[KCP((u32)((x[j] - lo) >> shift))], 1u);
    __syncthreads();
    // exclusive scan over the slots: thread t owns slots [t * KC_PER, (t + 1) * KC_PER)
    u32 sum = 0, longest = 0;
#pragma unroll
    for (int q = 0; q < KC_PER; q++) {
        const u32 v = cnt[KCP(tid * KC_PER + q)];
        sum += v;
        longest = max(longest, v);
    }
    incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 v = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += v;
    }
    if (lane == 31) s_w[warp] = incl;
    if (longest > (u32)KC_MAXRUN) s_bad = 1u;
    const int any_tie = __syncthreads_or((int)(longest > 1));
    if (s_bad) {  // values crowd into a few slots: leave the tile to the comparison sort
        if (tid == 0) redo[atomicAdd(&state[1], 1u)] = (u32)t;
        return;
    }
    {
        u32 excl = incl - sum;
        for (u32 w = 0; w < warp; w++) excl += s_w[w];
#pragma unroll
        for (int q = 0; q < KC_PER; q++) {
            const u32 v = cnt[KCP(tid * KC_PER + q)];
            cnt[KCP(tid * KC_PER + q)] = excl;
            excl += v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SS_IPT; j++)
        if (tid + j * KW_THREADS < size) stage[atomicAdd(&cnt[KCP((u32)((x[j] - lo) >> shift))], 1u)] = x[j];
    fence_proxy_async();
    __syncthreads();
    if (any_tie) {
        // cnt[slot] is now the END of the slot's run; it starts where the previous slot's ends
        u32 beg = tid ? cnt[KCP(tid * KC_PER - 1)] : 0u;
        for (int q = 0; q < KC_PER; q++) {
            const u32 end = cnt[KCP(tid * KC_PER + q)], l = end - beg;
            u64      *run = stage + beg;
            for (u32 i = 1; i < l; i++) {  // insertion sort (runs are a handful of items)
                const u64 v = run[i];
                u32       p = i;
                while (p > 0 && run[p - 1] > v) {
                    run[p] = run[p - 1];
                    p--;
                }
                run[p] = v;
            }
            beg = end;
        }
        fence_proxy_async();
        __syncthreads();
    }
    if (tid == 0) {
        bulk_store_items(dst, stage, size);
        bulk_wait_read();
    }
}
#undef KCP
}  // namespace

// One-pass merge of nruns <= 256 sorted runs (host arrays of device pointers / lengths) into out.
// *used = 0 when it does not apply (too many runs, too few items, a tile overflowed): the caller
// merges in rounds.  Synchronises once (the overflow flag) when it ran.
int smj_kway_merge(SmjCtx *ctx, u64 *out, const u64 *const *runs, const u64 *lens, u32 nruns, int *used) {
    *used = 0;
    if (nruns < 3 || nruns > (u32)KW_MAXK) return 0;
    u64 total = 0;
    for (u32 j = 0; j < nruns; j++) total += lens[j];
    if (total < (1u << 16)) return 0;
    // Regular sampling bounds a tile by (q + k) s items, but that bound needs every run to overshoot
    // by a whole block at both splitters; a run's piece is (its samples in the interval) * s give or
    // take a uniform (-s, s), so a tile holds q s items with a standard deviation of s sqrt(k / 3).
    // q leaves six of those (never more than the strict k s) below the capacity: 81 % average fill
    // for 128 runs instead of 50 %.  A tile that overflows all the same raises the flag (the caller
    // takes another path).
    const u32 k = nruns, s = k <= 128 ? 32u : 16u;
    const u32 margin = std::min<u32>(k * s, (u32)(6.0 * s * sqrt((double)k / 3.0)) + s);
    const u32 q = ((u32)KW_CAP - margin) / s;
    // The first tile also holds what lies before every run's first sample and the last one what lies
    // behind the last samples (k s / 2 items each on average): they get that many samples fewer.
    const u32 off = std::min<u32>(q - 1, (k + 1) / 2);
    KwRuns h;
    memset(&h, 0, sizeof(h));
    u64 m = 0;
    for (u32 j = 0; j < k; j++) {
        h.p[j]        = runs[j];
        h.n[j]        = lens[j];
        h.samp_off[j] = m;
        m += lens[j] / s;
    }
    h.samp_off[k] = m;
    const u64  ntiles = std::max<u64>(1, div_up(m + 2 * off, q));  // tile t starts after min(t q - off, m) samples
    SmjScratch sc(ctx);
    KwRuns *d_runs = (KwRuns *)sc.get(sizeof(KwRuns));
    u64    *d_samp = (u64 *)sc.get((m + 1) * 8), *d_sa = (u64 *)sc.get((m + 1) * 8), *d_sb = (u64 *)sc.get((m + 1) * 8);
    u64    *d_pos  = (u64 *)sc.get((ntiles + 1) * k * 8);
    u32    *d_state = (u32 *)sc.get(8), *d_redo = (u32 *)sc.get(ntiles * 4);
    if (!d_runs || !d_samp || !d_sa || !d_sb || !d_pos || !d_state || !d_redo)
        return smj_fail_msg("smj_kway_merge: out of device memory");
    // (h is pageable memory: the driver has staged it when the call returns)
    SMJ_CHECK(cudaMemcpyAsync(d_runs, &h, sizeof(KwRuns), cudaMemcpyHostToDevice, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_state, 0, 8, ctx->stream));
    u64 *sorted = d_samp;
    if (m) {
        SMJ_LAUNCH(ctx, k_kw_sample, (unsigned)std::min<u64>(div_up(m, 256), (u64)ctx->sm_count * 16), 256, 0, d_runs, k, s, m,
                   d_samp);
        // by key only (half the radix passes): a splitter is a sample's whole word whatever its place
        // among samples of the same key; with many equal keys the tiles just come out less even
        int rc = smj_lsd_sort(ctx, d_samp, d_sa, d_sb, m, 2, &sorted);
        if (rc) return rc;
    }
    SMJ_LAUNCH(ctx, k_kw_partition, (unsigned)div_up((ntiles + KW_GROUP) / KW_GROUP * k, 256), 256, 0, d_runs, k, q, off, m,
               ntiles, sorted, d_pos);
    const size_t smem_sort = (size_t)SS_WORDS(KW_CAP) * 8;
    if (smj_func_smem(ctx, (const void *)k_kw_tile, smem_sort)) return -1;
    if (smj_func_smem(ctx, (const void *)k_kw_tile_count, KC_SMEM)) return -1;
    const char *tile_env = getenv("SMJ_KWAY_TILE");  // "sort": every tile by the comparison sort (round-2 first version)
    SMJ_ITEMS(ctx, total);
    if (tile_env && tile_env[0] == 's')
        SMJ_LAUNCH(ctx, k_kw_tile, (unsigned)ntiles, KW_THREADS, smem_sort, d_runs, k, d_pos, out, d_state, (const u32 *)nullptr);
    else
        SMJ_LAUNCH(ctx, k_kw_tile_count, (unsigned)ntiles, KW_THREADS, KC_SMEM, d_runs, k, d_pos, out, d_state, d_redo);
    u32 h_state[2] = {0, 0};
    SMJ_CHECK(cudaMemcpyAsync(h_state, d_state, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    if (h_state[0]) return 0;  // a tile overflowed: not used
    if (h_state[1]) {           // tiles whose values are not evenly spread: comparison sort (sizes are known to fit)
        SMJ_LAUNCH(ctx, k_kw_tile, h_state[1], KW_THREADS, smem_sort, d_runs, k, d_pos, out, d_state, (const u32 *)d_redo);
    }
    *used = 1;
    return 0;
}
