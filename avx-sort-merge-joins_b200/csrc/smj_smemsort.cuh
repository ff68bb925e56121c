// smj_smemsort.cuh -- sort of one shared-memory tile by one CTA (internal).
//
// The B200 counterpart of the reference's in-cache sort (src/avxsort/avxsort_core.h:1213-1274
// inregister_sort_keyval32: sorting network on registers; :2056-2214 avxsort_block: bitonic merges
// that double the run length): every thread sorts 16 items in registers with a bitonic network,
// then log2(threads) merge rounds in shared memory double the run length -- a thread owns 16
// consecutive output positions of a pair of runs, finds its start with a merge-path binary search
// and merges serially INTO REGISTERS; after a barrier the registers replace the shared-memory
// contents, so one buffer holds the tile.  Used by k_small_sort (smj_smallsort.cu) and by the tiles
// of the one-pass k-way merge (smj_kway.cu).
#pragma once
#include "smj_common.cuh"

constexpr int SS_IPT = 16;
// A thread's 16 consecutive items would start 128 bytes apart from its neighbour's: every lane of a
// warp on the same bank.  One word of padding per 16 items moves lane l's items by l words:
// logical position i lives at SSP(i); a tile of n logical positions needs SS_WORDS(n) words.
#define SSP(i) ((i) + ((i) >> 4))
#define SS_WORDS(n) ((n) + ((n) >> 4) + 1)

__device__ __forceinline__ void ss_cswap(u64 &a, u64 &b, bool asc) {
    const bool sw = asc ? (a > b) : (a < b);
    const u64  t  = sw ? b : a;
    b             = sw ? a : b;
    a             = t;
}

// Logical positions 0 .. blockDim.x * 16 of s (see SSP) hold the tile, positions >= n hold ~0 (the
// largest word).  On return positions 0 .. n are ascending.  All threads of the CTA call it; ends with a barrier.
__device__ __forceinline__ void ss_sort_in_smem(u64 *s, u32 n) {
    const u32 t = threadIdx.x, total = blockDim.x * SS_IPT;
    u64       x[SS_IPT];
#pragma unroll
    for (int j = 0; j < SS_IPT; j++) x[j] = s[SSP(t * SS_IPT + j)];
    // bitonic network on registers (80 compare-exchanges, all indices static)
#pragma unroll
    for (int k = 2; k <= SS_IPT; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < SS_IPT; i++) {
                const int l = i ^ j;
                if (l > i) ss_cswap(x[i], x[l], (i & k) == 0);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SS_IPT; j++) s[SSP(t * SS_IPT + j)] = x[j];
    __syncthreads();
    // merge rounds: runs of L items -> runs of 2L items.  Runs that lie entirely past n hold only
    // padding, so the rounds stop as soon as one run covers the input.
    for (u32 L = SS_IPT; L < total && L < n; L <<= 1) {
        const u32 base = (t * SS_IPT) / (2 * L) * (2 * L), d = t * SS_IPT - base, bB = base + L;
        const bool live = base < n;  // pairs of runs past n hold only padding: nothing to do
        u32        lo = d > L ? d - L : 0, hi = d < L ? d : L;
        if (!live) lo = hi = 0;
        while (lo < hi) {  // merge path, A first on ties
            const u32 mid = (lo + hi) >> 1;
            if (s[SSP(base + mid)] <= s[SSP(bB + d - 1 - mid)])
                lo = mid + 1;
            else
                hi = mid;
        }
        u32 ai = lo, bi = d - lo;
        u64 av = ai < L ? s[SSP(base + ai)] : ~0ull, bv = bi < L ? s[SSP(bB + bi)] : ~0ull;
#pragma unroll
        for (int j = 0; j < SS_IPT; j++) {
            if (!live) break;
            const bool take_a = bi >= L || (ai < L && av <= bv);
            x[j]              = take_a ? av : bv;
            if (take_a) {
                ai++;
                av = ai < L ? s[SSP(base + ai)] : ~0ull;
            } else {
                bi++;
                bv = bi < L ? s[SSP(bB + bi)] : ~0ull;
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int j = 0; j < SS_IPT; j++) s[SSP(t * SS_IPT + j)] = x[j];
        }
        __syncthreads();
    }
}
