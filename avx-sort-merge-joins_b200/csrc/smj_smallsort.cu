// smj_smallsort.cu -- one-launch sort of small inputs (n <= 16384 items) inside one CTA.
//
// B200 restatement of the reference's in-cache sort (src/avxsort/avxsort_core.h:1213-1274
// inregister_sort_keyval32: sorting network on registers; :2056-2214 avxsort_block: bitonic merges
// that double the run length up to the 16384-item L2 block, src/avxsort/avxsort.c:17):
//   1. every thread sorts its 16 items in registers with a bitonic sorting network (static
//      register indices, no memory traffic),
//   2. log2(threads) merge rounds in shared memory double the run length: a thread owns 16
//      consecutive output positions of a pair of runs, finds its start by a merge-path binary
//      search and merges serially into registers; after a barrier the registers replace the
//      shared-memory contents (one buffer: 128 KB hold the whole 16384-item block),
//   3. coalesced write of the first n items.
// No histogram, no global scratch, no host synchronisation: what smj_lsd_sort's grid-wide radix
// passes + bit probe cost for a few thousand items is launch latency, not bandwidth.
// Stable key-only order (scalarsort twins): the network sorts (key << 32 | position) words, which are
// unique, and the payloads are gathered at the end.
#include "smj_smemsort.cuh"

namespace {
template <bool STABLE_KEY>
__global__ void __launch_bounds__(1024) k_small_sort(const u64 *__restrict__ in, u64 *__restrict__ out, u32 n) {
    extern __shared__ __align__(16) u64 s_ss[];
    const u32 t = threadIdx.x, T = blockDim.x, total = T * SS_IPT;
    // coalesced load into shared memory; positions past n hold the largest word
    for (u32 i = t; i < total; i += T) {
        u64 v = ~0ull;
        if (i < n) v = STABLE_KEY ? ((in[i] & 0xffffffff00000000ull) | i) : in[i];
        s_ss[SSP(i)] = v;
    }
    __syncthreads();
    ss_sort_in_smem(s_ss, n);
    for (u32 i = t; i < n; i += T) {
        const u64 v = s_ss[SSP(i)];
        out[i]      = STABLE_KEY ? in[(u32)v] : v;
    }
}
}  // namespace

// Sorts in[0..n) into out (in != out), n <= SMJ_SMALL_SORT_MAX.  stable_key: order by the key only,
// equal keys in input order; else by the whole word.  One launch, asynchronous.
int smj_small_sort(SmjCtx *ctx, const u64 *in, u64 *out, u64 n, bool stable_key) {
    if (n == 0) return 0;
    if (n > 16384) return smj_fail_msg("smj_small_sort: input too large");
    unsigned threads = 32;
    while ((u64)threads * SS_IPT < n) threads <<= 1;
    const size_t smem = (size_t)SS_WORDS(threads * SS_IPT) * 8;
    SMJ_ITEMS(ctx, n);
    if (stable_key) {
        if (smj_func_smem(ctx, (const void *)k_small_sort<true>, (size_t)SS_WORDS(1024 * SS_IPT) * 8)) return -1;
        SMJ_LAUNCH(ctx, k_small_sort<true>, 1, threads, smem, in, out, (u32)n);
    } else {
        if (smj_func_smem(ctx, (const void *)k_small_sort<false>, (size_t)SS_WORDS(1024 * SS_IPT) * 8)) return -1;
        SMJ_LAUNCH(ctx, k_small_sort<false>, 1, threads, smem, in, out, (u32)n);
    }
    return 0;
}
