// smj_kway.cu -- one-pass k-way merge of sorted runs (whole-word order).
//
// B200 restatement of src/merge/avx_multiwaymerge.c:199-338 avx_multiway_merge: there a binary
// tree of cache-resident FIFOs merges k runs in ONE pass over DRAM (read 8 + write 8 bytes per
// tuple, the tree traffic stays in L3).  Here the one pass is cut along the VALUE axis:
//   1. k_kw_sample   every s-th item of every run (the last item of each block of s) -> samples
//   2. the samples are sorted (they are n / s items: the library's own sort)
//   3. every q-th sorted sample is a splitter; tile t holds the values in (v_t, v_t+1].  Regular
//      sampling bounds a tile: a run contributes at most (c + 1) s items when c of its samples fall
//      into the interval, so a tile holds at most (q + k) s items -- chosen to fit shared memory
//   4. k_kw_partition   upper_bound of every splitter in every run (one thread per pair): tile t
//      of the output is the concatenation of k contiguous pieces, run j's piece is
//      [pos[t][j], pos[t+1][j])
//   5. k_kw_tile   one CTA per tile: the k pieces are gathered into shared memory (a warp per
//      piece, coalesced), sorted there (register network + merge-path rounds, smj_smemsort.cuh)
//      and written to their final place with coalesced stores.
// HBM traffic: 16 B per item + the samples (8 / s B per item read, sorted) + the splitter searches
// (k log n loads per tile, mostly L2).  Ties are bit-identical words, so the merged sequence is
// the sorted union whatever the tie order.  Many equal words can overfill a tile (the bound
// assumes distinct sample values at the splitters): the tile kernel then raises a flag and the
// caller falls back to the pairwise rounds.
#include "smj_internal.cuh"
#include "smj_smemsort.cuh"

#include <algorithm>
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
// row 0 = 0, row ntiles = the run lengths
__global__ void __launch_bounds__(256) k_kw_partition(const KwRuns *__restrict__ runs, u32 k, u32 q, u64 ntiles,
                                                      const u64 *__restrict__ sorted_samples,
                                                      u64 *__restrict__ pos) {
    const u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (ntiles + 1) * k) return;
    const u64 t = idx / k;
    const u32 j = (u32)(idx % k);
    const u64 n = runs->n[j];
    u64       r = 0;
    if (t == ntiles) {
        r = n;
    } else if (t > 0) {
        const u64  v = sorted_samples[t * q - 1];
        const u64 *a = runs->p[j];
        u64        lo = 0, hi = n;
        while (lo < hi) {
            const u64 mid = (lo + hi) >> 1;
            if (a[mid] <= v) lo = mid + 1; else hi = mid;
        }
        r = lo;
    }
    pos[idx] = r;
}

__global__ void __launch_bounds__(KW_THREADS) k_kw_tile(const KwRuns *__restrict__ runs, u32 k,
                                                        const u64 *__restrict__ pos, u64 *__restrict__ out,
                                                        u32 *__restrict__ overflow) {
    extern __shared__ __align__(16) u64 s_tile[];  // KW_CAP items
    __shared__ u64 s_lo[KW_MAXK];
    __shared__ u32 s_len[KW_MAXK], s_off[KW_MAXK + 1];
    __shared__ u64 s_base;
    const u64 t = blockIdx.x;
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
    // (q + k) s <= capacity; the average tile holds q s items
    const u32 k = nruns, s = k <= 128 ? 32u : 16u;
    const u32 q = (u32)KW_CAP / s - k;
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
    const u64  ntiles = m / q + 1;
    SmjScratch sc(ctx);
    KwRuns *d_runs = (KwRuns *)sc.get(sizeof(KwRuns));
    u64    *d_samp = (u64 *)sc.get((m + 1) * 8), *d_sa = (u64 *)sc.get((m + 1) * 8), *d_sb = (u64 *)sc.get((m + 1) * 8);
    u64    *d_pos  = (u64 *)sc.get((ntiles + 1) * k * 8);
    u32    *d_flag = (u32 *)sc.get(4);
    if (!d_runs || !d_samp || !d_sa || !d_sb || !d_pos || !d_flag) return smj_fail_msg("smj_kway_merge: out of device memory");
    SMJ_CHECK(cudaMemcpyAsync(d_runs, &h, sizeof(KwRuns), cudaMemcpyHostToDevice, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));  // h is a stack object
    SMJ_CHECK(cudaMemsetAsync(d_flag, 0, 4, ctx->stream));
    u64 *sorted = d_samp;
    if (m) {
        SMJ_LAUNCH(ctx, k_kw_sample, (unsigned)std::min<u64>(div_up(m, 256), (u64)ctx->sm_count * 16), 256, 0, d_runs, k, s, m,
                   d_samp);
        int rc = smj_lsd_sort(ctx, d_samp, d_sa, d_sb, m, 0, &sorted);
        if (rc) return rc;
    }
    SMJ_LAUNCH(ctx, k_kw_partition, (unsigned)div_up((ntiles + 1) * k, 256), 256, 0, d_runs, k, q, ntiles, sorted, d_pos);
    if (smj_func_smem(ctx, (const void *)k_kw_tile, (size_t)SS_WORDS(KW_CAP) * 8)) return -1;
    SMJ_ITEMS(ctx, total);
    SMJ_LAUNCH(ctx, k_kw_tile, (unsigned)ntiles, KW_THREADS, (size_t)SS_WORDS(KW_CAP) * 8, d_runs, k, d_pos, out, d_flag);
    u32 h_flag = 0;
    SMJ_CHECK(cudaMemcpyAsync(&h_flag, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    *used = h_flag == 0;
    return 0;
}
