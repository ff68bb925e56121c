// smj_radix.cu -- radix partitioning and LSD radix sort for 8-byte key|payload items.
//
// B200 restatement of two reference subsystems:
//   * src/partition/partition.c:94-287   radix_cluster{,_optimized,_optimized_V2}: histogram +
//     software-write-combining scatter  ->  k_hist_* + scatter kernels that stage a tile in shared
//     memory in bin order and write it as runs of consecutive addresses:
//       k_onesweep           STABLE (partition API, LSD passes): per-warp histograms, peers found by
//                            ballot / shared-memory bitmask, decoupled look-back between tiles;
//       k_onesweep_unstable  order inside a bin irrelevant (MSD passes, first LSD pass, exchange):
//                            rank = return value of a shared-memory atomicAdd, one global atomicAdd
//                            per (tile, bin) on the bin cursor, no look-back chain; optional bin
//                            capacity check for the histogram-free (optimistic) passes.
//   * src/avxsort/avxsort.c:128-209      avxsort_aligned: in-register networks + bitonic merges
//     ->  LSD radix sort over the VARYING bits only (k_bits_reduce finds them); large whole-word
//     sorts go MSD first (smj_msdsort.cu).
//   * the m-way join's two-level MSD range partition (exact: k_msd_hist1/2 + scatter; optimistic:
//     k_msd_sample, k_msd_caps1, k_msd_level2_plan, k_msd_scatter2<true>, k_msd_fix_*).
//
// HBM traffic per item: k_bits_reduce 8 B, histograms 8 B, every scatter pass 16 B.
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <type_traits>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int RADIX_NB = 512;  // stride of every per-digit table; a pass uses 2^8 or 2^9 bins

inline int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// ---------------------------------------------------------------------------------------
// OR / AND of all items: tells which bit positions vary at all (constant bits need no pass)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bits_reduce(const u64 *__restrict__ in, u64 n,
                                                     u64 *__restrict__ out /* [0]=OR [1]=AND [2]=order flags */) {
    // order flags: bit 0 = some item is smaller than its predecessor (whole word), bit 1 = some
    // payload (low 32 bits) is smaller than its predecessor's.  An input whose payloads already
    // ascend needs no payload passes: a stable sort on the key alone gives the whole-word order.
    u64 vor = 0, vand = ~0ull, flags = 0;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    constexpr int UN = 8;  // independent loads in flight per thread
    // (the loop condition is the same for all lanes of a warp: the shuffles below need every lane)
    for (u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x; i0 - lane_id() < n; i0 += UN * stride) {
        u64 a[UN], p0[UN];
#pragma unroll
        for (int u = 0; u < UN; u++) {
            const u64 i = i0 + u * stride;
            a[u]        = i < n ? ld_stream(in + i) : 0ull;
            // the predecessor is the neighbouring lane's item; only lane 0 loads it (same or previous sector)
            p0[u] = (i < n && i && lane_id() == 0) ? in[i - 1] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < UN; u++) {
            const u64 i  = i0 + u * stride;
            const u64 up = __shfl_up_sync(SMJ_FULL_MASK, a[u], 1);
            const u64 p  = lane_id() ? up : p0[u];
            if (i < n) {
                vor |= a[u];
                vand &= a[u];
                flags |= (a[u] < p ? 1ull : 0ull) | ((u32)a[u] < (u32)p ? 2ull : 0ull);
            }
        }
    }
    vor   = warp_reduce_or(vor);
    vand  = warp_reduce_and(vand);
    flags = warp_reduce_or(flags);
    __shared__ u64 s_or[8], s_and[8], s_fl[8];
    if (lane_id() == 0) {
        s_or[threadIdx.x >> 5]  = vor;
        s_and[threadIdx.x >> 5] = vand;
        s_fl[threadIdx.x >> 5]  = flags;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
            vor |= s_or[w];
            vand &= s_and[w];
            flags |= s_fl[w];
        }
        atomicOr(out, vor);
        atomicAnd(out + 1, vand);
        if (flags) atomicOr(out + 2, flags);
    }
}

// ---------------------------------------------------------------------------------------
// histograms of up to 8 digits in one read of the data
// ---------------------------------------------------------------------------------------
struct DigitPlan {
    int     nd;
    DigitFn d[8];
};

__global__ void __launch_bounds__(512) k_hist_digits(const u64 *__restrict__ in, u64 n,
                                                     DigitPlan plan,
                                                     u64 *__restrict__ hist /* [nd][RADIX_NB] */,
                                                     u64 *__restrict__ bits /* OR, AND or null */) {
    __shared__ u32 sh[8 * RADIX_NB];
    u64 vor = 0, vand = ~0ull;
    for (int i = threadIdx.x; i < plan.nd * RADIX_NB; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
        u64  x[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            u64 idx = i + (u64)u * stride;
            ok[u]   = idx < n;
            x[u]    = ok[u] ? ld_stream(in + idx) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (ok[u]) {
                vor |= x[u];
                vand &= x[u];
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k < plan.nd) atomicAdd(&sh[k * RADIX_NB + plan.d[k](x[u])], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.nd * RADIX_NB; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (u64)sh[i]);
    if (bits) {
        vor  = warp_reduce_or(vor);
        vand = warp_reduce_and(vand);
        if (lane_id() == 0) {
            atomicOr(bits, vor);
            atomicAnd(bits + 1, vand);
        }
    }
}

// OR / AND of every stride-th item: a cheap guess of the varying bits (verified later)
__global__ void __launch_bounds__(256) k_bits_sample(const u64 *__restrict__ in, u64 n, u64 stride,
                                                     u64 *__restrict__ out) {
    u64 vor = 0, vand = ~0ull;
    for (u64 i = ((u64)blockIdx.x * blockDim.x + threadIdx.x) * stride; i < n;
         i += (u64)gridDim.x * blockDim.x * stride) {
        u64 a = in[i];
        vor |= a;
        vand &= a;
    }
    vor  = warp_reduce_or(vor);
    vand = warp_reduce_and(vand);
    if (lane_id() == 0) {
        atomicOr(out, vor);
        atomicAnd(out + 1, vand);
    }
}

// one digit with up to 2^14 bins in dynamic shared memory (partition API, fan-out sweep)
__global__ void __launch_bounds__(512) k_hist_wide_smem(const u64 *__restrict__ in, u64 n,
                                                        DigitFn fn, u32 nbins,
                                                        u64 *__restrict__ hist) {
    extern __shared__ u32 shw[];
    for (u32 i = threadIdx.x; i < nbins; i += blockDim.x) shw[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
        u64  x[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            u64 idx = i + (u64)u * stride;
            ok[u]   = idx < n;
            x[u]    = ok[u] ? ld_stream(in + idx) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (ok[u]) atomicAdd(&shw[fn(x[u])], 1u);
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < nbins; i += blockDim.x)
        if (shw[i]) atomicAdd(&hist[i], (u64)shw[i]);
}

// more than 2^14 bins: warp-aggregated atomics straight into the (L2 resident) table
__global__ void __launch_bounds__(256) k_hist_wide_global(const u64 *__restrict__ in, u64 n,
                                                          DigitFn fn, u64 *__restrict__ hist) {
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 base = (u64)blockIdx.x * blockDim.x; base < n; base += stride) {
        u64  idx = base + threadIdx.x;
        bool ok  = idx < n;
        u32  d   = ok ? fn(ld_stream(in + idx)) : 0xffffffffu;
        u32  m   = __match_any_sync(SMJ_FULL_MASK, d);
        if (ok && lane_id() == (u32)(__ffs(m) - 1)) atomicAdd(&hist[d], (u64)__popc(m));
    }
}

// exclusive scan of each 256-bin histogram (one block per digit)
__global__ void __launch_bounds__(RADIX_NB) k_bases_multi(const u64 *__restrict__ hist,
                                                          u64 *__restrict__ bases) {
    __shared__ u64 s[RADIX_NB];
    const int      b = threadIdx.x;
    const u64      v = hist[blockIdx.x * RADIX_NB + b];
    s[b]             = v;
    __syncthreads();
    for (int o = 1; o < RADIX_NB; o <<= 1) {
        u64 t = (b >= o) ? s[b - o] : 0;
        __syncthreads();
        s[b] += t;
        __syncthreads();
    }
    bases[blockIdx.x * RADIX_NB + b] = s[b] - v;
}

// Partition starts for the partition API.  offs[p] = start of partition p in the output
// (dense, or rounded up to 8 tuples per partition: reference partition.c:183-187);
// padshift[p] = offs[p] - dense start of p.
__global__ void __launch_bounds__(1024) k_bases_wide(const u64 *__restrict__ hist, u32 nbins,
                                                     int padded, u64 *__restrict__ offs,
                                                     u64 *__restrict__ padshift) {
    __shared__ u64 s_pad[1024], s_dense[1024];
    const u32      chunk = (nbins + 1023) / 1024;
    const u32      lo    = threadIdx.x * chunk;
    const u32      hi    = min(nbins, lo + chunk);
    u64            sp = 0, sd = 0;
    for (u32 p = lo; p < hi; p++) {
        u64 h = hist[p];
        sd += h;
        sp += padded ? ((h + 7) & ~7ull) : h;
    }
    s_pad[threadIdx.x]   = sp;
    s_dense[threadIdx.x] = sd;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        u64 tp = 0, td = 0;
        if ((int)threadIdx.x >= o) {
            tp = s_pad[threadIdx.x - o];
            td = s_dense[threadIdx.x - o];
        }
        __syncthreads();
        s_pad[threadIdx.x] += tp;
        s_dense[threadIdx.x] += td;
        __syncthreads();
    }
    u64 rp = s_pad[threadIdx.x] - sp, rd = s_dense[threadIdx.x] - sd;  // exclusive
    for (u32 p = lo; p < hi; p++) {
        u64 h = hist[p];
        offs[p] = rp;
        if (padshift) padshift[p] = rp - rd;
        rd += h;
        rp += padded ? ((h + 7) & ~7ull) : h;
    }
}

// fold a 2^nbits histogram into the histograms of its low / high sub-digits
__global__ void __launch_bounds__(256) k_fold_hist(const u64 *__restrict__ hist, u32 nbins,
                                                   u32 lo_bits, u64 *__restrict__ hist2 /*[2][256]*/) {
    const u32 lo_mask = (1u << lo_bits) - 1;
    for (u32 p = blockIdx.x * blockDim.x + threadIdx.x; p < nbins; p += gridDim.x * blockDim.x) {
        u64 h = hist[p];
        if (h) {
            atomicAdd(&hist2[p & lo_mask], h);
            atomicAdd(&hist2[RADIX_NB + (p >> lo_bits)], h);
        }
    }
}

// ---------------------------------------------------------------------------------------
// one stable partitioning pass: read once, write once (16 B / item)
// ---------------------------------------------------------------------------------------
constexpr u64 ST_AGG    = 1ull << 62;  // tile-local count published
constexpr u64 ST_PREFIX = 2ull << 62;  // inclusive prefix over tiles 0..t published
constexpr u64 ST_VAL    = (1ull << 62) - 1;

// Decoupled look-back for one bin: sum of the bin's counts over all earlier tiles.  The states of
// LB predecessors are fetched with independent loads, so one L2 round trip covers LB tiles; with
// one predecessor per round trip the whole pass is limited to about 2*depth tiles per L2 latency
// (45 % of the warp time of the first versions was spent here, profiles/r01_onesweep.md).
template <int NB>
__device__ __forceinline__ u64 lookback_sum(const u64 *status, u32 tile, u32 b) {
    constexpr int LB   = 8;
    u64           prev = 0;
    long long     t    = (long long)tile - 1;
    while (true) {
        u64 sv[LB];
#pragma unroll
        for (int i = 0; i < LB; i++)
            sv[i] = (t - i >= 0) ? ld_relaxed(status + (u64)(t - i) * NB + b) : ST_PREFIX;
#pragma unroll
        for (int i = 0; i < LB; i++) {
            u64 x = sv[i];
            while ((x >> 62) == 0) x = ld_relaxed(status + (u64)(t - i) * NB + b);
            prev += x & ST_VAL;
            if ((x >> 62) == 2) return prev;
        }
        t -= LB;
    }
}

template <int THREADS, int IPT, int BITS>
struct OneSweepCfg {
    static constexpr int    NB    = 1 << BITS;
    static constexpr int    WARPS = THREADS / 32;
    static constexpr int    TILE  = THREADS * IPT;
    static constexpr int    BPT   = (NB + THREADS - 1) / THREADS;  // bins owned by a thread
    static constexpr size_t SMEM  = (size_t)TILE * 8 + (size_t)WARPS * NB * 4 + NB * 8;
};

// Ranks are computed with warp votes: one ballot per digit bit gives every lane the mask of the
// lanes that hold the same digit ("peers").  match.any would do the same in one instruction but
// executes on the XU pipe once per DISTINCT value in the warp (~28 for random 8-bit digits): the
// first version of this kernel was XU-bound at 24 % of the HBM roofline because of it
// (profiles/r01_onesweep_matchany.txt).
// One tile of a stable partitioning pass:
//   1. load the tile (warp-striped: position order == memory order)
//   2. count: per-warp histogram with shared-memory atomics                -> publish the tile's
//      per-bin counts for the look-back of later tiles as early as possible
//   3. scan: counts -> position of every (warp, bin) group inside the tile
//   4. rank: position of every item = group position + number of equal-digit items before it.
//      RANK 0 finds the equal-digit lanes ("peers") with one ballot per digit bit, RANK 1 through a
//      per-warp bitmask in shared memory (atomicOr of the lane bit into mask[digit]); the mask
//      array aliases the staging buffer, which is idle until step 5.  match.any would do it in
//      one instruction but runs on the XU pipe once per DISTINCT value in the warp: the first
//      version of this kernel was bound by it (profiles/r01_onesweep.md).
//   5. stage the tile in bin order in shared memory, finish the look-back (global position of the
//      tile's first item of every bin), write runs of consecutive addresses.
template <int THREADS, int IPT, int BITS, class F, bool FULL, int RANK>
__device__ __forceinline__ void onesweep_tile(const u64 *__restrict__ in, u64 *__restrict__ out,
                                              const F &fn, const u64 *__restrict__ bases,
                                              u64 *status, const DigitFn &fn_full,
                                              const u64 *__restrict__ padshift, u32 tile,
                                              u32 valid, u64 *s_items, u32 *s_whist,
                                              u64 *s_goff, u32 *s_wsum) {
    using Cfg = OneSweepCfg<THREADS, IPT, BITS>;
    constexpr int NB  = Cfg::NB;
    constexpr int BPT = Cfg::BPT;
    const u64 tile_start = (u64)tile * Cfg::TILE;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 last_bin   = fn.last_bin();  // bins in use: 0..last_bin (< NB)
    const u32 lt         = lanemask_lt();

    // ---- 1. load
    u64       item[IPT];
    const u32 wbase = warp * 32 * IPT;
    const u64 *src  = in + tile_start + wbase + lane;
#pragma unroll
    for (int j = 0; j < IPT; j++)
        item[j] = (FULL || wbase + j * 32 + lane < valid) ? ld_stream(src + j * 32) : ~0ull;

    // ---- 2. count
    u32 *wh = s_whist + warp * NB;
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 d = (FULL || wbase + j * 32 + lane < valid) ? fn(item[j]) : last_bin;
        atomicAdd(&wh[d], 1u);
    }
    __syncthreads();

    // ---- 3. scan: every thread owns BPT consecutive bins
    u32 cntb[BPT], tsum = 0;
#pragma unroll
    for (int q = 0; q < BPT; q++) {
        const u32 b = threadIdx.x * BPT + q;
        u32       c = 0;
        if (b < NB) {
#pragma unroll
            for (int w = 0; w < Cfg::WARPS; w++) c += s_whist[w * NB + b];
        }
        cntb[q] = c;
        tsum += c;
    }
    // publish the counts now: later tiles can add them up while this one is still ranking
#pragma unroll
    for (int q = 0; q < BPT; q++) {
        const u32 b   = threadIdx.x * BPT + q;
        if (b <= last_bin) {
            // padding items of a partial tile were counted in the last bin; they are not data
            const u64 cnt = (u64)cntb[q] - ((b == last_bin) ? (u64)(Cfg::TILE - valid) : 0ull);
            if (tile == 0) {
                st_relaxed(status + b, ST_PREFIX | cnt);
            } else {
                st_relaxed(status + (u64)tile * NB + b, ST_AGG | cnt);
            }
        }
    }
    u32 incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    u32 excl = incl - tsum;
    for (u32 w = 0; w < warp; w++) excl += s_wsum[w];
    u32 excl_b[BPT];
#pragma unroll
    for (int q = 0; q < BPT; q++) {
        const u32 b = threadIdx.x * BPT + q;
        excl_b[q]   = excl;
        if (b < NB) {
            u32 run = excl;
#pragma unroll
            for (int w = 0; w < Cfg::WARPS; w++) {
                const u32 t         = s_whist[w * NB + b];
                s_whist[w * NB + b] = run;  // tile position of warp w's first item of bin b
                run += t;
            }
        }
        excl += cntb[q];
    }
    __syncthreads();

    // ---- 4. rank: tile position of every item (two 16-bit positions per register)
    u32  pos2[IPT / 2];
    u32 *wm = (u32 *)s_items + warp * NB;  // RANK 1 only; zeroed by the kernel prologue
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        const u32 d = (FULL || wbase + j * 32 + lane < valid) ? fn(item[j]) : last_bin;
        u32       peers, base;
        if (RANK == 0) {
            peers = SMJ_FULL_MASK;
#pragma unroll
            for (int b = 0; b < BITS; b++) {
                const bool bit = (d >> b) & 1u;
                const u32  v   = __ballot_sync(SMJ_FULL_MASK, bit);
                peers &= bit ? v : ~v;
            }
            base = wh[d];  // the same word for all peers: a broadcast
            __syncwarp();
            if ((peers >> lane) == 1u) wh[d] = base + __popc(peers);  // highest peer bumps it
            __syncwarp();
        } else {
            atomicOr(&wm[d], 1u << lane);
            __syncwarp();
            peers = wm[d];
            base  = wh[d];
            __syncwarp();
            if ((peers >> lane) == 1u) {
                wh[d] = base + __popc(peers);
                wm[d] = 0;
            }
            __syncwarp();
        }
        const u32 p = base + __popc(peers & lt);  // peers in lower lanes come earlier in memory
        if (j & 1)
            pos2[j / 2] |= p << 16;
        else
            pos2[j / 2] = p;
    }
    if (RANK == 1) __syncthreads();  // the mask arrays alias the staging buffer

    // ---- 5. stage in bin order
#pragma unroll
    for (int j = 0; j < IPT; j++)
        s_items[(j & 1) ? (pos2[j / 2] >> 16) : (pos2[j / 2] & 0xffffu)] = item[j];

    // finish the decoupled look-back: global rank of this tile's first item in every bin
#pragma unroll
    for (int q = 0; q < BPT; q++) {
        const u32 b = threadIdx.x * BPT + q;
        if (b <= last_bin) {
            u64 prev = 0;
            if (tile != 0) {
                const u64 cnt = (u64)cntb[q] - ((b == last_bin) ? (u64)(Cfg::TILE - valid) : 0ull);
                prev = lookback_sum<NB>(status, tile, b);
                st_relaxed(status + (u64)tile * NB + b, ST_PREFIX | (prev + cnt));
            }
            s_goff[b] = bases[b] + prev - (u64)excl_b[q];
        }
    }
    __syncthreads();  // staged items and s_goff visible

    if (padshift == nullptr) {
        if (FULL) {
#pragma unroll 4
            for (u32 i = threadIdx.x; i < (u32)Cfg::TILE; i += THREADS) {
                u64 x = s_items[i];
                st_stream(out + s_goff[fn(x)] + i, x);
            }
        } else {
            for (u32 i = threadIdx.x; i < valid; i += THREADS) {
                u64 x = s_items[i];
                st_stream(out + s_goff[fn(x)] + i, x);
            }
        }
    } else {
        for (u32 i = threadIdx.x; i < valid; i += THREADS) {
            u64 x = s_items[i];
            st_stream(out + s_goff[fn(x)] + i + padshift[fn_full(x)], x);
        }
    }
}

template <int THREADS, int IPT, int BITS, class F, int RANK, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    k_onesweep(const u64 *__restrict__ in, u64 *__restrict__ out, u64 n, F fn,
               const u64 *__restrict__ bases,  // [NB] global start of every bin
               u64 *status,                    // [tiles][NB] look-back state, zeroed
               u32 *tile_counter,              // zeroed
               DigitFn fn_full, const u64 *__restrict__ padshift) {
    using Cfg = OneSweepCfg<THREADS, IPT, BITS>;
    constexpr int NB = Cfg::NB;
    static_assert(IPT % 2 == 0, "positions are packed in pairs");
    static_assert(Cfg::TILE <= 65536, "positions are 16 bit");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_items = (u64 *)smem_raw;                       // TILE staged items
    u32 *s_whist = (u32 *)(s_items + Cfg::TILE);          // [WARPS][NB]
    u64 *s_goff  = (u64 *)(s_whist + Cfg::WARPS * NB);    // [NB] global offset - tile position
    __shared__ u32 s_tile;
    __shared__ u32 s_wsum[32];

    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    // (16-byte stores: the two arrays are 2 x 32 KB, a word per thread and store was an eighth of the
    // kernel's shared-memory instructions)
    static_assert((Cfg::WARPS * NB) % 4 == 0 && (Cfg::TILE * 8) % 16 == 0, "16-byte zeroing");
    for (int i = threadIdx.x; i < Cfg::WARPS * NB / 4; i += THREADS) ((uint4 *)s_whist)[i] = make_uint4(0, 0, 0, 0);
    if (RANK == 1) {
        static_assert(Cfg::WARPS * NB * 4 <= Cfg::TILE * 8, "mask array must fit the staging area");
        for (int i = threadIdx.x; i < Cfg::WARPS * NB / 4; i += THREADS) ((uint4 *)s_items)[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const u32 tile   = s_tile;
    const u64 remain = n - (u64)tile * Cfg::TILE;
    if (remain >= (u64)Cfg::TILE)
        onesweep_tile<THREADS, IPT, BITS, F, true, RANK>(in, out, fn, bases, status, fn_full,
                                                         padshift, tile, (u32)Cfg::TILE, s_items,
                                                         s_whist, s_goff, s_wsum);
    else
        onesweep_tile<THREADS, IPT, BITS, F, false, RANK>(in, out, fn, bases, status, fn_full,
                                                          padshift, tile, (u32)remain, s_items,
                                                          s_whist, s_goff, s_wsum);
}

// ---------------------------------------------------------------------------------------
// UNSTABLE partitioning pass: same contract as k_onesweep except that items of one bin keep no
// particular order inside a tile.  Good enough for the first pass of a key-only sort (nothing to
// preserve yet) and for MSD passes; much cheaper, because the rank of an item is simply the value
// a shared-memory atomicAdd on the bin counter returns: no per-warp histograms, no peer search.
// ---------------------------------------------------------------------------------------
// AGG: rank through warp aggregation (match.any) instead of one atomic per lane; for bucket
// functions with few bins (range partitioning over the GPUs), where per-lane atomics would all
// hit the same few counters.
// Bin capacity policies of the unstable pass.  With exact bin starts (a histogram pass ran first)
// there is nothing to check (NoLimit).  The OPTIMISTIC passes of the m-way join run without a
// histogram: every bin owns a region whose capacity was estimated from a sample; a tile whose
// reservation does not fit does not store that bin's items at all and raises a flag, and the pass
// (or the bucket) is repeated exactly, histogram first.  limit(b) = one past the last item index
// bin b may use.
// Per-bin offsets of a tile are (global start of the tile's slice) - (tile position of the bin): any
// value in [-8192, 2^62).  A bin whose reservation did not fit is marked with SKIP64, which is not
// such a value (an all-ones marker would collide with the legitimate offset -1).
constexpr u64 SKIP64 = 1ull << 62;
// MEASUREMENT HOOK ONLY (tools/p2p_scatter_bw.py): reservations rounded up to this many items, which
// leaves holes in the output; 1 (no rounding) everywhere else.
__constant__ u32 c_dbg_reserve_align = 1;

struct NoLimit {
    static constexpr bool on = false, keys = false;
    __device__ __forceinline__ u64 operator()(u32) const { return ~0ull; }
    __device__ __forceinline__ void raise() const {}
    __device__ __forceinline__ u32 key_bad(u64) const { return 0; }
    __device__ __forceinline__ void raise_key() const {}
};
// first level: limit[b] from an array.  It also verifies the sampled key range on the way: every
// key must agree with kref on the bit positions kmask (the positions the digit plan treats as
// constant); one that does not raises *kflag and the host drops the plan.
struct LimitArray {
    static constexpr bool on = true, keys = true;
    const u64 *lim;
    u32       *flag;
    u32        kmask, kref;
    u32       *kflag;
    __device__ __forceinline__ u64 operator()(u32 b) const { return lim[b]; }
    __device__ __forceinline__ void raise() const { *flag = 1u; }
    __device__ __forceinline__ u32 key_bad(u64 x) const { return ((u32)(x >> 32) ^ kref) & kmask; }
    __device__ __forceinline__ void raise_key() const { *kflag = 1u; }
};
// second level: the 512 sub-buckets of first-level bucket `bucket` are slots of cap2[bucket] items
// from base2[bucket] on (read when needed: once per thread and tile); a slot may be used in full
// except under the test hook that provokes overflows (shrink)
struct LimitSlots {
    static constexpr bool on = true, keys = false;
    const u64 *base2, *cap2;
    u32        bucket, shrink;
    u32       *flag;
    u32       *any;  // optional: one word that says "some bucket overflowed"
    __device__ __forceinline__ u64 operator()(u32 b) const {
        const u64 cap = cap2[bucket];
        return base2[bucket] + (u64)b * cap + (shrink ? cap / 4 : cap);
    }
    __device__ __forceinline__ void raise() const {
        flag[bucket] = 1u;
        if (any) *any = 1u;
    }
    __device__ __forceinline__ u32 key_bad(u64) const { return 0; }
    __device__ __forceinline__ void raise_key() const {}
};

// Asynchronous copies global -> shared (LDGSTS): the next tile of a persistent CTA lands in the
// staging buffer slot by slot, behind the output loop that has just emptied the slot, so that its
// DRAM latency is hidden behind this tile's stores and costs no registers.
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// the tile's items, thread-strided: item[j] = src[j * THREADS + thread]
template <int THREADS, int IPT, bool FULL>
__device__ __forceinline__ void unstable_load(u64 (&item)[IPT], const u64 *__restrict__ src, u32 valid) {
#pragma unroll
    for (int j = 0; j < IPT; j++)
        item[j] = (FULL || j * THREADS + threadIdx.x < valid) ? ld_stream(src + threadIdx.x + j * THREADS)
                                                              : ~0ull;
}

// One tile whose items are already in registers (unstable_load).  s_cnt must be zero on entry and
// is zero again on return (after a __syncthreads() of the caller).  next_src != null: the CTA's
// next tile (a FULL one) is requested into item[] as soon as the registers are free, i.e. before
// this tile is written out, so its DRAM latency hides behind the output phase.
// NEXT = 0: next_src (if any) is loaded into item[] before the output loop (register prefetch);
// NEXT = 1: next_src (a FULL tile; this tile must be FULL too) is copied asynchronously into s_items,
// every slot right after the output loop has read it; the caller waits (cp_async_wait_all +
// __syncthreads) and reads item[] from s_items.
// BULK (FULL tiles only; s_items has TILE + NB slots): every bin's run is staged at a slot of the
// same 16-byte phase as its destination and leaves as one bulk copy issued by the bin's thread
// (plus at most one plain store at either end); the caller must bulk_wait_read() + __syncthreads()
// before s_items is written again and before the CTA exits.
template <int THREADS, int IPT, int BITS, class F, bool FULL, bool AGG = false, class LIM = NoLimit, int NEXT = 0,
          bool BULK = false>
__device__ __forceinline__ void unstable_process(u64 (&item)[IPT], u64 *__restrict__ out, const F &fn,
                                                 u64 *cursors, u32 valid, u64 *s_items, u32 *s_cnt,
                                                 u64 *s_goff, u32 *s_wsum, const LIM &lim,
                                                 const u64 *__restrict__ next_src) {
    constexpr int NB   = 1 << BITS;
    constexpr int TILE = THREADS * IPT;
    constexpr int BPT  = (NB + THREADS - 1) / THREADS;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 last_bin = fn.last_bin();
    u32       pos2[IPT / 2];
    if (LIM::keys) {
        u32 bad = 0;
#pragma unroll
        for (int j = 0; j < IPT; j++)
            if (FULL || j * THREADS + threadIdx.x < valid) bad |= lim.key_bad(item[j]);
        if (bad) lim.raise_key();
    }
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        u32 r = 0;  // rank inside the tile's bin; items past the end of the input take no part
        if (AGG) {
            const bool ok = FULL || j * THREADS + threadIdx.x < valid;
            const u32  d  = ok ? fn(item[j]) : 0xffffffffu;
            const u32  m  = __match_any_sync(SMJ_FULL_MASK, d);
            const u32  ld = __ffs(m) - 1;
            u32        base = 0;
            if (ok && lane == ld) base = atomicAdd(&s_cnt[d], (u32)__popc(m));
            base = __shfl_sync(SMJ_FULL_MASK, base, ld);
            r    = base + __popc(m & lanemask_lt());
        } else if (FULL || j * THREADS + threadIdx.x < valid) {
            r = atomicAdd(&s_cnt[fn(item[j])], 1u);
        }
        if (j & 1)
            pos2[j / 2] |= r << 16;
        else
            pos2[j / 2] = r;
    }
    __syncthreads();

    // Order between tiles does not matter either, so there is no look-back chain: the tile
    // reserves its slice of every bin with one atomicAdd on the bin's global cursor.
    u32 cntb[BPT], tsum = 0;
    u64 gbase[BPT];
#pragma unroll
    for (int q = 0; q < BPT; q++) {
        const u32 b = threadIdx.x * BPT + q;
        cntb[q]     = (b < NB) ? s_cnt[b] : 0u;
        gbase[q]    = 0;
        tsum += cntb[q];
        if (b <= last_bin && cntb[q]) {
            // the limit is requested before the atomic, so that the two round trips to L2 overlap
            const u64 limit = LIM::on ? lim(b) : ~0ull;
            const u32 al    = c_dbg_reserve_align;
            gbase[q]        = atomicAdd(&cursors[b], (u64)((cntb[q] + al - 1) & ~(al - 1)));
            if (LIM::on && gbase[q] + cntb[q] > limit) {  // does not fit: store nothing, raise the flag
                gbase[q] = SKIP64;
                lim.raise();
            }
        }
    }
    u32 incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    u32 excl = incl - tsum;
    for (u32 w = 0; w < warp; w++) excl += s_wsum[w];
    u32 spos[BPT];
#pragma unroll
    for (int q = 0; q < BPT; q++) {
        const u32 b = threadIdx.x * BPT + q;
        if (BULK && FULL) {
            // one spare slot per bin: the run starts on the slot whose 16-byte phase is the destination's
            const u32 first = excl + b;
            spos[q]         = first + ((first ^ (u32)((uintptr_t)(out + gbase[q]) >> 3)) & 1u);
            if (b < NB) s_cnt[b] = spos[q];
        } else if (b < NB) {
            s_cnt[b]  = excl;  // tile position of the bin's first item
            s_goff[b] = (LIM::on && gbase[q] == SKIP64) ? SKIP64 : gbase[q] - (u64)excl;
        }
        excl += cntb[q];
    }
    __syncthreads();

#pragma unroll
    for (int j = 0; j < IPT; j++) {
        if (FULL || j * THREADS + threadIdx.x < valid) {
            const u32 r = (j & 1) ? (pos2[j / 2] >> 16) : (pos2[j / 2] & 0xffffu);
            s_items[s_cnt[fn(item[j])] + r] = item[j];
        }
    }
    if (BULK && FULL) fence_proxy_async();
    __syncthreads();
    for (int i = threadIdx.x; i < NB; i += THREADS) s_cnt[i] = 0;  // ready for the next tile
    if (NEXT == 0 && next_src) unstable_load<THREADS, IPT, true>(item, next_src, (u32)TILE);
    if (BULK && FULL) {
#pragma unroll
        for (int q = 0; q < BPT; q++) {
            u32 c = cntb[q];
            if (!c || threadIdx.x * BPT + q > last_bin || (LIM::on && gbase[q] == SKIP64)) continue;
            u64       *dst = out + gbase[q];
            const u64 *src = s_items + spos[q];
            if ((uintptr_t)dst & 8) {  // odd first item
                st_stream(dst, *src);
                dst++, src++, c--;
            }
            if (c & 1) {  // odd last item
                c--;
                st_stream(dst + c, src[c]);
            }
            if (c) bulk_s2g(dst, src, c * 8);
        }
        bulk_commit();
    } else if (FULL) {
#pragma unroll 4
        for (u32 i = threadIdx.x; i < (u32)TILE; i += THREADS) {
            const u64 x = s_items[i], g = s_goff[fn(x)];
            if (!LIM::on || g != SKIP64) st_stream(out + g + i, x);
            if (NEXT == 1 && next_src) cp_async8(s_items + i, next_src + i);
        }
        if (NEXT == 1 && next_src) cp_async_commit();
    } else {
        for (u32 i = threadIdx.x; i < valid; i += THREADS) {
            const u64 x = s_items[i], g = s_goff[fn(x)];
            if (!LIM::on || g != SKIP64) st_stream(out + g + i, x);
        }
    }
}

// one tile, not pipelined: in[0 .. valid) -> out
template <int THREADS, int IPT, int BITS, class F, bool FULL, bool AGG = false, class LIM = NoLimit>
__device__ __forceinline__ void unstable_tile(const u64 *__restrict__ in, u64 *__restrict__ out,
                                              const F &fn, u64 *cursors, u32 valid, u64 *s_items,
                                              u32 *s_cnt, u64 *s_goff, u32 *s_wsum,
                                              const LIM &lim = LIM()) {
    u64 item[IPT];
    unstable_load<THREADS, IPT, FULL>(item, in, valid);
    unstable_process<THREADS, IPT, BITS, F, FULL, AGG, LIM>(item, out, fn, cursors, valid, s_items, s_cnt,
                                                            s_goff, s_wsum, lim, nullptr);
}

template <int THREADS, int IPT, int MINB, int BITS, class F, bool AGG, class LIM, bool ASYNC = false, bool BULK = false>
__global__ void __launch_bounds__(THREADS, MINB)
    k_onesweep_unstable(const u64 *__restrict__ in, u64 *__restrict__ out, u64 n,
                        const __grid_constant__ F fn,
                        u64 *cursors /* [NB], initialised to the start of every bin */,
                        const __grid_constant__ LIM lim, u32 pf_dist) {
    constexpr int NB   = 1 << BITS;
    constexpr int TILE = THREADS * IPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_items = (u64 *)smem_raw;                            // TILE staged items (BULK: + NB spare slots)
    u32 *s_cnt   = (u32 *)(s_items + TILE + (BULK ? NB : 0));  // [NB]
    u64 *s_goff  = (u64 *)(s_cnt + NB);                        // [NB]
    __shared__ u32 s_wsum[32];
    for (int i = threadIdx.x; i < NB; i += THREADS) s_cnt[i] = 0;
    __syncthreads();
    // persistent CTAs: tile t, t + grid, ...; the next tile is in flight while this one is written
    const u64 nfull = n / TILE;
    u64       tile  = blockIdx.x;
    if (tile < nfull) {
        u64 item[IPT];
        unstable_load<THREADS, IPT, true>(item, in + tile * TILE, (u32)TILE);
        // one CTA per tile (the usual launch): the CTA that will take this CTA's slot next starts
        // with tile + pf_dist; one line per thread covers its whole tile
        if (pf_dist && tile + pf_dist < nfull && threadIdx.x * 16 < TILE)
            prefetch_l2(in + (tile + pf_dist) * TILE + threadIdx.x * 16);
        while (true) {
            const u64  next = tile + gridDim.x;
            const u64 *nsrc = next < nfull ? in + next * TILE : nullptr;
            unstable_process<THREADS, IPT, BITS, F, true, AGG, LIM, ASYNC ? 1 : 0, BULK>(
                item, out, fn, cursors, (u32)TILE, s_items, s_cnt, s_goff, s_wsum, lim, nsrc);
            if (ASYNC && nsrc) cp_async_wait_all();
            if (BULK) bulk_wait_read();
            __syncthreads();
            if (!nsrc) break;
            if (ASYNC) {
#pragma unroll
                for (int j = 0; j < IPT; j++) item[j] = s_items[threadIdx.x + j * THREADS];
            }
            tile = next;
        }
    }
    if (n % TILE && blockIdx.x == nfull % gridDim.x)  // the ragged last tile
        unstable_tile<THREADS, IPT, BITS, F, false, AGG, LIM>(in + nfull * TILE, out, fn, cursors,
                                                              (u32)(n - nfull * TILE), s_items, s_cnt,
                                                              s_goff, s_wsum, lim);
}

// max_grid != 0: at most that many (persistent) CTAs, so that the pass leaves SMs to kernels of
// another stream (join group: the owner side of R runs under the NVLink-bound exchange of S)
template <int THREADS, int IPT, int MINB, int BITS, class F, bool AGG, class LIM>
int launch_unstable_shape(SmjCtx *ctx, const u64 *in, u64 *out, u64 n, F fn, u64 *bases, LIM lim, unsigned max_grid = 0,
                          int bulk = -1 /* -1: $SMJ_SCATTER_BULK decides */) {
    constexpr int    NB = 1 << BITS, TILE = THREADS * IPT;
    constexpr size_t SMEM = (size_t)TILE * 8 + NB * 4 + NB * 8;
    if (smj_func_smem(ctx, (const void *)k_onesweep_unstable<THREADS, IPT, MINB, BITS, F, AGG, LIM>, SMEM)) return -1;
    // One CTA per tile: measured faster on HBM-resident input than persistent CTAs with the next
    // tile in flight (0.49 vs 0.52 ms per 128 M items; the kernel is bound by the shared-memory
    // pipe, not by load latency).  Persistent CTAs (the kernel supports any grid) pay off when the
    // input is host memory read over PCIe, where a tile's load latency is several microseconds.
    const bool     host_in = smj_is_host_pointer(in);
    const unsigned grid = host_in ? (unsigned)std::min<u64>(div_up(n, TILE), (u64)ctx->sm_count * MINB)
                                  : (unsigned)div_up(n, TILE);
    SMJ_ITEMS(ctx, n);
    if constexpr (!AGG) {
        if (bulk >= 0 ? bulk != 0 : env_int("SMJ_SCATTER_BULK", 0) != 0) {  // runs leave shared memory as bulk copies (TMA)
            constexpr size_t SMEM_B = SMEM + NB * 8;
            if (smj_func_smem(ctx, (const void *)k_onesweep_unstable<THREADS, IPT, MINB, BITS, F, AGG, LIM, false, true>, SMEM_B))
                return -1;
            const bool     persist = host_in || (max_grid && max_grid < grid);
            const unsigned g = persist ? (max_grid && max_grid < grid ? max_grid : grid) : grid;
            SMJ_LAUNCH(ctx, (k_onesweep_unstable<THREADS, IPT, MINB, BITS, F, AGG, LIM, false, true>), g, THREADS, SMEM_B, in,
                       out, n, fn, bases, lim, persist ? 0u : (u32)(ctx->sm_count * MINB));
            return 0;
        }
    }
    if (LIM::on && !host_in && env_int("SMJ_SCATTER_ASYNC", 0)) {  // persistent CTAs, next tile by cp.async
        if (smj_func_smem(ctx, (const void *)k_onesweep_unstable<THREADS, IPT, MINB, BITS, F, AGG, LIM, true>, SMEM)) return -1;
        SMJ_LAUNCH(ctx, (k_onesweep_unstable<THREADS, IPT, MINB, BITS, F, AGG, LIM, true>),
                   (unsigned)std::min<u64>(div_up(n, TILE), (u64)ctx->sm_count * MINB), THREADS, SMEM, in, out, n, fn, bases,
                   lim, 0u);
        return 0;
    }
    if (max_grid && max_grid < grid) {
        SMJ_LAUNCH(ctx, (k_onesweep_unstable<THREADS, IPT, MINB, BITS, F, AGG, LIM>), max_grid, THREADS, SMEM, in, out,
                   n, fn, bases, lim, 0u);
        return 0;
    }
    const u32 pf_dist = host_in ? 0u : (u32)(ctx->sm_count * MINB);  // resident CTAs = tiles in flight
    SMJ_LAUNCH(ctx, (k_onesweep_unstable<THREADS, IPT, MINB, BITS, F, AGG, LIM>), grid, THREADS, SMEM, in, out,
               n, fn, bases, lim, pf_dist);
    return 0;
}

// bases: [NB] start of every bin; it is used as the cursor array and is CONSUMED (on return
// bases[b] is the end of bin b)
template <int BITS, class F, bool AGG = false, class LIM = NoLimit>
int launch_unstable(SmjCtx *ctx, const u64 *in, u64 *out, u64 n, F fn, u64 *bases, LIM lim = LIM()) {
    return launch_unstable_shape<512, 16, 2, BITS, F, AGG, LIM>(ctx, in, out, n, fn, bases, lim);
}

// Wide digits want big tiles: with 512 bins a 4096-item tile has only 8 items per bin to amortise
// the per-bin scan and look-back over.
constexpr int OS_MIN_TILE = 4096;

inline size_t status_bytes(u64 n) {
    // sized for the smallest tile and the largest bin count: fits every configuration
    return (size_t)div_up(n, OS_MIN_TILE) * RADIX_NB * 8;
}

template <int THREADS, int IPT, int BITS, class F, int RANK, int MINB>
int launch_onesweep_cfg(SmjCtx *ctx, const u64 *in, u64 *out, u64 n, F fn, const u64 *bases,
                        u64 *status, u32 *counter, DigitFn fn_full, const u64 *padshift) {
    using Cfg = OneSweepCfg<THREADS, IPT, BITS>;
    static_assert(Cfg::TILE >= OS_MIN_TILE, "status_bytes assumes tiles of at least OS_MIN_TILE");
    if (smj_func_smem(ctx, (const void *)k_onesweep<THREADS, IPT, BITS, F, RANK, MINB>, Cfg::SMEM)) return -1;
    const u64 tiles = div_up(n, Cfg::TILE);
    SMJ_ITEMS(ctx, n);
    SMJ_LAUNCH(ctx, (k_onesweep<THREADS, IPT, BITS, F, RANK, MINB>), (unsigned)tiles, THREADS,
               Cfg::SMEM, in, out, n, fn, bases, status, counter, fn_full, padshift);
    return 0;
}

// SMJ_RANK_MODE=0 selects the ballot ranking, 1 (default) the shared-memory bitmask ranking
inline int rank_mode() {
    static int mode = env_int("SMJ_RANK_MODE", 1);
    return mode;
}

template <int BITS, int RANK, class F>
int launch_onesweep_shape(SmjCtx *ctx, const u64 *in, u64 *out, u64 n, F fn, const u64 *bases,
                          u64 *status, u32 *counter, DigitFn fn_full, const u64 *padshift) {
    // 512 threads x 16 items, 2 CTAs/SM: the best of the four shapes measured in round 1
    // (256x16x4: 1.25 ms, 256x24x3: 1.06 ms, 384x16x2: 1.19 ms, 512x16x2: 0.89 ms per 128M-item pass)
    return launch_onesweep_cfg<512, 16, BITS, F, RANK, 2>(ctx, in, out, n, fn, bases, status, counter,
                                                          fn_full, padshift);
}

template <class F>
int launch_onesweep(SmjCtx *ctx, const u64 *in, u64 *out, u64 n, F fn, const u64 *bases,
                    u64 *status, u32 *counter, DigitFn fn_full, const u64 *padshift) {
    const bool wide = fn.last_bin() >= 256;
    if (rank_mode() == 0) {
        if (wide)
            return launch_onesweep_shape<9, 0, F>(ctx, in, out, n, fn, bases, status, counter,
                                                  fn_full, padshift);
        return launch_onesweep_shape<8, 0, F>(ctx, in, out, n, fn, bases, status, counter, fn_full,
                                              padshift);
    }
    if (wide)
        return launch_onesweep_shape<9, 1, F>(ctx, in, out, n, fn, bases, status, counter, fn_full,
                                              padshift);
    return launch_onesweep_shape<8, 1, F>(ctx, in, out, n, fn, bases, status, counter, fn_full,
                                          padshift);
}

// histogram of one bucket function with few (<= 32) bins, e.g. destinations of a range
// partition: lanes with the same bin are found with match.any (cheap when there are only a few
// distinct values in a warp) and their leader adds the whole group with one shared-memory atomic;
// plain per-lane atomics would serialise 32 ways on the same counter
template <class F>
__global__ void __launch_bounds__(512) k_hist_fn(const u64 *__restrict__ in, u64 n, F fn,
                                                 u64 *__restrict__ hist) {
    __shared__ u32 sh[RADIX_NB];
    for (int i = threadIdx.x; i < RADIX_NB; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    const u64 rounds = (n + stride - 1) / stride;
    u64       i      = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (u64 k = 0; k < rounds; k++, i += stride) {
        const u32 d = (i < n) ? fn(ld_stream(in + i)) : 0xffffffffu;
        const u32 m = __match_any_sync(SMJ_FULL_MASK, d);
        if (i < n && lane_id() == (u32)(__ffs(m) - 1)) atomicAdd(&sh[d], (u32)__popc(m));
    }
    __syncthreads();
    for (int i2 = threadIdx.x; i2 < RADIX_NB; i2 += blockDim.x)
        if (sh[i2]) atomicAdd(&hist[i2], (u64)sh[i2]);
}

int hist_grid(SmjCtx *ctx, u64 n, int threads, int per_thread) {
    u64 want = div_up(n, (u64)threads * per_thread);
    u64 cap  = (u64)ctx->sm_count * 8;
    return (int)std::max<u64>(1, std::min(want, cap));
}

}  // namespace

// ---------------------------------------------------------------------------------------
// LSD radix sort over the varying bits (internal entry used by smjdev_sort and the joins)
// ---------------------------------------------------------------------------------------
// key_only: 0 = order by the whole 64-bit word, 1 = by the key, stable, 2 = by the key with an
// unspecified order among equal keys (first pass may then be the cheaper unstable kernel).
// Reads `in`, ping-pongs between the scratch buffers a and b (pass 1: in -> a, pass 2: a -> b,
// pass 3: b -> a, ...).  `in` may alias b (classic two-buffer sort) but not a.  *sorted is set
// to the buffer that holds the result, which is `in` itself when no pass was necessary.
int smj_small_sort(SmjCtx *ctx, const u64 *in, u64 *out, u64 n, bool stable_key);
constexpr u64 SMJ_SMALL_SORT_MAX = 16384;  // the reference's in-cache block (src/avxsort/avxsort.c:17)

// OR / AND / order flags of n items (host array of 3 words); synchronises
int smj_item_bits(SmjCtx *ctx, const u64 *in, u64 n, u64 *h_bits) {
    SmjScratch sc(ctx);
    u64 *d_bits = (u64 *)sc.get(24);
    if (!d_bits) return smj_fail_msg("smj_lsd_sort: scratch allocation failed");
    const u64 init[3] = {0ull, ~0ull, 0ull};
    SMJ_CHECK(cudaMemcpyAsync(d_bits, init, 24, cudaMemcpyHostToDevice, ctx->stream));
    SMJ_ITEMS(ctx, n);
    SMJ_LAUNCH(ctx, k_bits_reduce, hist_grid(ctx, n, 256, 16), 256, 0, in, n, d_bits);
    SMJ_CHECK(cudaMemcpyAsync(h_bits, d_bits, 24, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int smj_lsd_sort_bits(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, const u64 *h_bits,
                      u64 **sorted);
static int lsd_sort_bits_impl(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, const u64 *h_bits,
                              u64 **sorted, u64 *exact_bits);

int smj_lsd_sort(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, u64 **sorted) {
    *sorted = const_cast<u64 *>(in);
    if (n <= 1) return 0;
    if (n <= SMJ_SMALL_SORT_MAX && !getenv("SMJ_NO_SMALL_SORT")) {
        // one CTA, one launch, no bit probe and no host synchronisation (smj_smallsort.cu)
        *sorted = a;
        return smj_small_sort(ctx, in, a, n, key_only == 1);
    }
    u64 h_bits[3];
    // Key-only sorts of large inputs do not read the data just to learn which key bits vary (0.25 ms per
    // 128 M items): the digit plan comes from a SAMPLE of 2^17 items, the histogram pass -- which reads
    // everything anyway -- accumulates the exact OR / AND on the way, and the host checks before the
    // scatter passes that no key bit outside the plan varies; if one does (a sampling miss) the plan
    // is redone from the exact bits the histogram pass has just delivered.  (The order flags of the
    // full probe only matter to whole-word sorts.)
    if (key_only && n >= (1ull << 22) && !getenv("SMJ_NO_SAMPLED_BITS")) {
        SmjScratch sc(ctx);
        u64 *d_bits = (u64 *)sc.get(16);
        if (!d_bits) return smj_fail_msg("smj_lsd_sort: scratch allocation failed");
        const u64 init[2] = {0ull, ~0ull};
        SMJ_CHECK(cudaMemcpyAsync(d_bits, init, 16, cudaMemcpyHostToDevice, ctx->stream));
        SMJ_LAUNCH(ctx, k_bits_sample, 128, 256, 0, in, n, std::max<u64>(1, n >> 17), d_bits);
        SMJ_CHECK(cudaMemcpyAsync(h_bits, d_bits, 16, cudaMemcpyDeviceToHost, ctx->stream));
        SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
        h_bits[2] = 3;  // nothing is known about the order
        u64 exact[2] = {0, 0};
        int rc = lsd_sort_bits_impl(ctx, in, a, b, n, key_only, h_bits, sorted, exact);
        if (rc <= 0) return rc;  // done (or failed)
        h_bits[0] = exact[0], h_bits[1] = exact[1];  // rc == 1: the sample missed a varying key bit
        return lsd_sort_bits_impl(ctx, in, a, b, n, key_only, h_bits, sorted, nullptr);
    }
    int rc = smj_item_bits(ctx, in, n, h_bits);
    if (rc) return rc;
    return smj_lsd_sort_bits(ctx, in, a, b, n, key_only, h_bits, sorted);
}

// The passes of the LSD sort for items whose OR / AND / order flags are known (h_bits: exact for
// these items, or of a superset of them -- extra varying bits only cost passes; order flags of a
// superset that says "ascending" hold for every contiguous part).  No host synchronisation.
int smj_lsd_sort_bits(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, const u64 *h_bits,
                      u64 **sorted) {
    return lsd_sort_bits_impl(ctx, in, a, b, n, key_only, h_bits, sorted, nullptr);
}
// exact_bits != null (key-only sorts): h_bits is a guess; the histogram pass delivers the exact OR / AND
// into exact_bits[0..1], the host synchronises and returns 1 -- before any scatter pass -- when a key bit
// outside the guess varies.
static int lsd_sort_bits_impl(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, const u64 *h_bits,
                              u64 **sorted, u64 *exact_bits) {
    *sorted = const_cast<u64 *>(in);
    const u64 *buf = in;
    if (n <= 1) return 0;
    SmjScratch sc(ctx);
    u64 varying = h_bits[0] & ~h_bits[1];
    if (!key_only && !(h_bits[2] & 1)) return 0;  // already in whole-word order
    // payloads already ascend in input order (e.g. row numbers): the stable key passes suffice
    if (key_only || !(h_bits[2] & 2)) varying &= 0xffffffff00000000ull;

    // plan: cover the varying bits, least significant first.  Bits that are closer than 9
    // positions form a group; a group of span s takes ceil(s/9) passes of equal width (27 key
    // bits -> 9+9+9, 30 bits -> 8+8+7+7), each window trimmed to the bits that really vary.
    DigitPlan plan;
    plan.nd = 0;
    u64 v   = varying;
    while (v && plan.nd < 8) {
        const int lo  = __builtin_ctzll(v);
        int       end = lo + 1;  // one past the highest bit of the group
        for (int b = lo + 1; b < 64 && b < end + 9; b++)
            if ((v >> b) & 1) end = b + 1;
        const int span   = end - lo;
        const int passes = (span + 8) / 9;
        const int width  = (span + passes - 1) / passes;
        for (int p = 0; p < passes && plan.nd < 8; p++) {
            const int wlo = lo + p * width;
            int       w   = std::min(width, end - wlo);
            u64       win = (varying >> wlo) & ((1ull << w) - 1);
            if (win == 0) continue;  // nothing varies inside this window
            const int skip = __builtin_ctzll(win);
            win >>= skip;
            w                     = 64 - __builtin_clzll(win);
            plan.d[plan.nd].sub   = 0;
            plan.d[plan.nd].shift = (u32)(wlo + skip);
            plan.d[plan.nd].mask  = (u32)((1ull << w) - 1);
            plan.nd++;
        }
        v = (end >= 64) ? 0 : (v >> end) << end;
    }
    if (plan.nd == 0 && !exact_bits) return 0;  // all items equal (on the compared bits): already sorted
    if (plan.nd == 0) {  // a guess that says "all keys equal" is checked with a plan of one digit
        plan.d[0].sub = 0, plan.d[0].shift = 32, plan.d[0].mask = 1;
        plan.nd = 1;
    }

    const size_t sbytes = status_bytes(n);
    u64 *d_hist   = (u64 *)sc.get((size_t)plan.nd * RADIX_NB * 8);
    u64 *d_bases  = (u64 *)sc.get((size_t)plan.nd * RADIX_NB * 8);
    u64 *d_status = (u64 *)sc.get(sbytes * plan.nd);
    u32 *d_count  = (u32 *)sc.get(sizeof(u32) * 8);
    if (!d_hist || !d_bases || !d_status || !d_count)
        return smj_fail_msg("smj_lsd_sort: scratch allocation failed");
    SMJ_CHECK(cudaMemsetAsync(d_hist, 0, (size_t)plan.nd * RADIX_NB * 8, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_status, 0, sbytes * plan.nd, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_count, 0, sizeof(u32) * 8, ctx->stream));
    SMJ_ITEMS(ctx, n);
    u64 *d_exact = nullptr;
    if (exact_bits) {
        d_exact = (u64 *)sc.get(16);
        if (!d_exact) return smj_fail_msg("smj_lsd_sort: scratch allocation failed");
        const u64 init[2] = {0ull, ~0ull};
        SMJ_CHECK(cudaMemcpyAsync(d_exact, init, 16, cudaMemcpyHostToDevice, ctx->stream));
    }
    SMJ_LAUNCH(ctx, k_hist_digits, hist_grid(ctx, n, 512, 16), 512, 0, buf, n, plan, d_hist, d_exact);
    SMJ_LAUNCH(ctx, k_bases_multi, plan.nd, RADIX_NB, 0, d_hist, d_bases);
    if (exact_bits) {
        SMJ_CHECK(cudaMemcpyAsync(exact_bits, d_exact, 16, cudaMemcpyDeviceToHost, ctx->stream));
        SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
        const u64 key_bits = 0xffffffff00000000ull;
        if ((exact_bits[0] & ~exact_bits[1] & key_bits) & ~(h_bits[0] & ~h_bits[1])) return 1;  // the plan misses a varying bit
    }

    const u64 *cur = in;
    for (int k = 0; k < plan.nd; k++) {
        u64 *dst = (k & 1) ? b : a;
        FieldFn f  = {plan.d[k].shift, plan.d[k].mask};
        int     rc;
        // The first pass may scramble items that agree on its digit when nothing depends on their
        // input order: key-only sorts that promise no order among equal keys, and whole-word sorts
        // (ties are then bit-identical items) unless the payload passes were dropped above.
        if (k == 0 && (key_only == 2 || (key_only == 0 && (h_bits[2] & 2)))) {
            rc = (f.mask >= 256)
                     ? launch_unstable<9>(ctx, cur, dst, n, f, d_bases)
                     : launch_unstable<8>(ctx, cur, dst, n, f, d_bases);
            if (rc) return rc;
            cur = dst;
            continue;
        }
        rc = launch_onesweep(ctx, cur, dst, n, f, d_bases + (size_t)k * RADIX_NB,
                                     (u64 *)((char *)d_status + sbytes * k), d_count + k, plan.d[k],
                                     nullptr);
        if (rc) return rc;
        cur = dst;
    }
    *sorted = const_cast<u64 *>(cur);
    return 0;
}

// ---------------------------------------------------------------------------------------
// C-ABI: histogram / partition
// ---------------------------------------------------------------------------------------
static int run_hist_wide(SmjCtx *ctx, const u64 *in, u64 n, DigitFn fn, u32 nbins, u64 *hist) {
    SMJ_CHECK(cudaMemsetAsync(hist, 0, (size_t)nbins * 8, ctx->stream));
    if (n == 0) return 0;
    if (nbins <= (1u << 14)) {
        if (smj_func_smem(ctx, (const void *)k_hist_wide_smem, 65536)) return -1;
        SMJ_LAUNCH(ctx, k_hist_wide_smem, hist_grid(ctx, n, 512, 16), 512, (size_t)nbins * 4, in, n,
                   fn, nbins, hist);
    } else {
        SMJ_LAUNCH(ctx, k_hist_wide_global, hist_grid(ctx, n, 256, 16), 256, 0, in, n, fn, hist);
    }
    return 0;
}

extern "C" int smjdev_histogram(const uint64_t *in, uint64_t n, uint32_t nbits, uint32_t shift,
                                int key_minus_one, uint64_t *hist) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (nbits > 16 || shift + nbits > 32) return smj_fail_msg("smjdev_histogram: bad radix bits");
    DigitFn fn = {key_minus_one ? (1ull << 32) : 0ull, 32 + shift, (1u << nbits) - 1};
    return run_hist_wide(ctx, (const u64 *)in, n, fn, 1u << nbits, (u64 *)hist);
}

extern "C" int smjdev_partition(const uint64_t *in_, uint64_t *out_, uint64_t *tmp_, uint64_t n,
                                uint32_t nbits, uint32_t shift, int key_minus_one, int padded,
                                uint64_t *hist_, uint64_t *offs_) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (nbits > 16 || shift > 31) return smj_fail_msg("smjdev_partition: bad radix bits");
    const u64 *in = (const u64 *)in_;
    u64 *out = (u64 *)out_, *tmp = (u64 *)tmp_, *hist = (u64 *)hist_, *offs = (u64 *)offs_;
    const u32 nbins = 1u << nbits;
    const u64 sub   = key_minus_one ? (1ull << 32) : 0ull;
    DigitFn   full  = {sub, 32 + shift, nbins - 1};

    int rc = run_hist_wide(ctx, in, n, full, nbins, hist);
    if (rc) return rc;
    SmjScratch sc(ctx);
    const bool two_pass = nbits > 9;
    u64 *d_padshift = (padded && two_pass) ? (u64 *)sc.get((size_t)nbins * 8) : nullptr;
    if (padded && two_pass && !d_padshift) return smj_fail_msg("smjdev_partition: scratch");
    SMJ_LAUNCH(ctx, k_bases_wide, 1, 1024, 0, hist, nbins, padded, offs, d_padshift);
    if (n == 0) return 0;

    const size_t sbytes = status_bytes(n);
    u64 *d_status = (u64 *)sc.get(sbytes * 2);
    u32 *d_count  = (u32 *)sc.get(8 * sizeof(u32));
    if (!d_status || !d_count) return smj_fail_msg("smjdev_partition: scratch");
    SMJ_CHECK(cudaMemsetAsync(d_status, 0, sbytes * (two_pass ? 2 : 1), ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_count, 0, 8 * sizeof(u32), ctx->stream));

    if (!two_pass) {
        // bins above nbins-1 are never touched by the kernel, so offs can serve as bases
        return launch_onesweep(ctx, in, out, n, full, offs, d_status, d_count, full, nullptr);
    }
    if (!tmp) return smj_fail_msg("smjdev_partition: nbits > 8 needs a scratch buffer");
    // two stable passes, low sub-digit first: the result is ordered by the full digit and
    // keeps the input order inside each partition, exactly like one stable pass would
    const u32 lo_bits = nbits / 2, hi_bits = nbits - lo_bits;
    u64 *d_hist2  = (u64 *)sc.get(2 * RADIX_NB * 8);
    u64 *d_bases2 = (u64 *)sc.get(2 * RADIX_NB * 8);
    if (!d_hist2 || !d_bases2) return smj_fail_msg("smjdev_partition: scratch");
    SMJ_CHECK(cudaMemsetAsync(d_hist2, 0, 2 * RADIX_NB * 8, ctx->stream));
    SMJ_LAUNCH(ctx, k_fold_hist, std::max(1u, std::min(nbins / 256, 64u)), 256, 0, hist, nbins,
               lo_bits, d_hist2);
    SMJ_LAUNCH(ctx, k_bases_multi, 2, RADIX_NB, 0, d_hist2, d_bases2);
    DigitFn lo = {sub, 32 + shift, (1u << lo_bits) - 1};
    DigitFn hi = {sub, 32 + shift + lo_bits, (1u << hi_bits) - 1};
    if (hi.shift > 63) {  // the reference shifts a 32-bit key: everything up there is zero
        hi.shift = 63;
        hi.mask  = 0;
    }
    rc = launch_onesweep(ctx, in, tmp, n, lo, d_bases2, d_status, d_count, full, nullptr);
    if (rc) return rc;
    return launch_onesweep(ctx, tmp, out, n, hi, d_bases2 + RADIX_NB,
                           (u64 *)((char *)d_status + sbytes), d_count + 1, full, d_padshift);
}

// ---------------------------------------------------------------------------------------
// C-ABI: range partitioning for the multi-GPU exchange (replaces the "pull my partitions from
// every thread" step of src/joins/sortmergejoin_multiway.c:492-517)
// ---------------------------------------------------------------------------------------
extern "C" int smjdev_range_partition(const uint64_t *in, uint64_t *out, uint64_t n,
                                      const uint32_t *splitters, uint32_t nparts,
                                      uint64_t *counts) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (nparts < 1 || nparts > 32) return smj_fail_msg("smjdev_range_partition: 1..32 parts");
    SplitFn fn;
    fn.nparts = nparts;
    for (u32 i = 0; i < 31; i++) fn.sp[i] = (i + 1 < nparts) ? splitters[i] : 0xffffffffu;
    SmjScratch sc(ctx);
    u64 *d_hist  = (u64 *)sc.get(RADIX_NB * 8);
    u64 *d_bases = (u64 *)sc.get(RADIX_NB * 8);
    if (!d_hist || !d_bases) return smj_fail_msg("smjdev_range_partition: scratch");
    SMJ_CHECK(cudaMemsetAsync(d_hist, 0, RADIX_NB * 8, ctx->stream));
    if (n) {
        SMJ_LAUNCH(ctx, k_hist_fn<SplitFn>, hist_grid(ctx, n, 512, 8), 512, 0, (const u64 *)in, n, fn,
                   d_hist);
    }
    SMJ_LAUNCH(ctx, k_bases_multi, 1, RADIX_NB, 0, d_hist, d_bases);
    SMJ_CHECK(cudaMemcpyAsync(counts, d_hist, (size_t)nparts * 8, cudaMemcpyDeviceToDevice,
                              ctx->stream));
    if (n == 0) return 0;
    const size_t sbytes   = status_bytes(n);
    u64         *d_status = (u64 *)sc.get(sbytes);
    u32         *d_count  = (u32 *)sc.get(sizeof(u32));
    if (!d_status || !d_count) return smj_fail_msg("smjdev_range_partition: scratch");
    SMJ_CHECK(cudaMemsetAsync(d_status, 0, sbytes, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_count, 0, sizeof(u32), ctx->stream));
    DigitFn none = {0, 0, 0};
    return launch_onesweep(ctx, (const u64 *)in, (u64 *)out, n, fn, d_bases, d_status, d_count, none,
                           nullptr);
}

// ---------------------------------------------------------------------------------------
// Two-level MSD range partition (the m-way join's partitioning phase on one GPU).
//
// The reference partitions each thread's chunk once by PARTFANOUT key bits and then needs a
// multiway merge to bring the threads' runs together (sortmergejoin_multiway.c:331-386, 463-556).
// Here the whole relation is range-partitioned by its top 18 varying key bits in two unstable
// passes (9 bits each, atomic bin cursors, no look-back chain): pass A scatters by the top 9
// bits, pass B partitions every pass-A bucket by the next 9 bits.  Tiles of pass B never straddle
// a pass-A bucket, so a tile only needs the 512 cursors of its own bucket.  The result is 2^18
// key-range buckets in key order; what remains of the key inside a bucket is <= 10 bits.
// ---------------------------------------------------------------------------------------
namespace {

constexpr int MSD_THREADS = 512, MSD_IPT = 16, MSD_TILE = MSD_THREADS * MSD_IPT, MSD_NB = 512;

// histogram of ONE 9-bit digit + exact OR / AND of the items (first pass of the MSD partition)
__global__ void __launch_bounds__(512) k_msd_hist1(const u64 *__restrict__ in, u64 n, FieldFn fn,
                                                   u64 *__restrict__ hist, u64 *__restrict__ bits) {
    __shared__ u32 sh[MSD_NB];
    sh[threadIdx.x] = 0;
    __syncthreads();
    u64       vor = 0, vand = ~0ull;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 8 * stride) {
        u64 x[8];  // eight independent loads in flight per thread
#pragma unroll
        for (int u = 0; u < 8; u++) x[u] = (i + u * stride < n) ? ld_stream(in + i + u * stride) : 0ull;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (i + u * stride < n) {
                vor |= x[u];
                vand &= x[u];
                atomicAdd(&sh[fn(x[u])], 1u);
            }
        }
    }
    __syncthreads();
    const u32 c = sh[threadIdx.x];
    if (c) atomicAdd(&hist[threadIdx.x], (u64)c);
    if (bits) {
        vor  = warp_reduce_or(vor);
        vand = warp_reduce_and(vand);
        if (lane_id() == 0) {
            atomicOr(bits, vor);
            atomicAnd(bits + 1, vand);
        }
    }
}

// per pass-A bucket: first tile index (exclusive scan of the tile counts), first item index and
// item count.  starts == null: the buckets lie back to back (exclusive scan of the counts).
__global__ void __launch_bounds__(MSD_NB) k_msd_tile_prefix(const u64 *__restrict__ counts,
                                                            const u64 *__restrict__ starts,
                                                            u64 *__restrict__ tile_prefix,
                                                            u64 *__restrict__ bucket_start,
                                                            u64 *__restrict__ bucket_count) {
    __shared__ u64 s_t[MSD_NB], s_c[MSD_NB];
    const int      b = threadIdx.x;
    const u64      c = counts[b], t = (c + MSD_TILE - 1) / MSD_TILE;
    s_t[b]           = t;
    s_c[b]           = c;
    __syncthreads();
    for (int o = 1; o < MSD_NB; o <<= 1) {
        u64 at = (b >= o) ? s_t[b - o] : 0, ac = (b >= o) ? s_c[b - o] : 0;
        __syncthreads();
        s_t[b] += at;
        s_c[b] += ac;
        __syncthreads();
    }
    tile_prefix[b]  = s_t[b] - t;
    bucket_start[b] = starts ? starts[b] : s_c[b] - c;
    bucket_count[b] = c;
    if (b == MSD_NB - 1) tile_prefix[MSD_NB] = s_t[b];
}

// The tiles of pass B: tile t covers items [start, start + valid) of first-level bucket `bucket`
// (a tile never straddles two buckets).  CTA b writes the entries of bucket b's tiles.
struct __align__(16) MsdTile {
    u64 start;
    u32 bucket, valid;
};
__device__ __forceinline__ void msd_write_tiles(MsdTile *__restrict__ tiles, u32 b, u64 first_tile,
                                                u64 start, u64 count) {
    const u64 nt = (count + MSD_TILE - 1) / MSD_TILE;
    for (u64 t = threadIdx.x; t < nt; t += blockDim.x) {
        MsdTile e;
        e.start  = start + t * MSD_TILE;
        e.bucket = b;
        e.valid  = (u32)min((u64)MSD_TILE, count - t * MSD_TILE);
        tiles[first_tile + t] = e;
    }
}
__global__ void __launch_bounds__(256) k_msd_tile_table(const u64 *__restrict__ tile_prefix,
                                                        const u64 *__restrict__ bucket_start,
                                                        const u64 *__restrict__ bucket_count,
                                                        MsdTile *__restrict__ tiles) {
    const u32 b = blockIdx.x;
    msd_write_tiles(tiles, b, tile_prefix[b], bucket_start[b], bucket_count[b]);
}

// Tile table from SEGMENTS: segment i = items [start[i], start[i] + count[i]) of first-level
// bucket bucket[i].  A bucket may come in several segments (multi-GPU: one per source rank).
// One CTA; seg = {start[nseg], count[nseg], bucket[nseg]}; ntiles_at[MSD_NB] receives the total.
constexpr int SEG_THREADS = 1024, SEG_PER = 8, SEG_MAX = SEG_THREADS * SEG_PER;
// Segment q (q = thread * SEG_PER + i would serialise every thread over consecutive segments and their
// tiles; instead) is handled by thread q % 1024 in round q / 1024, so that loads are coalesced, and
// the tile entries are written by a flat loop over the TILES (tile -> segment by binary search in the
// shared-memory prefix array), so that the stores are coalesced too: 49 -> ~10 us for 4096 segments.
__global__ void __launch_bounds__(SEG_THREADS) k_msd_seg_tiles(const u64 *__restrict__ seg, u32 nseg,
                                                               MsdTile *__restrict__ tiles,
                                                               u64 *__restrict__ ntiles_at) {
    __shared__ u32 s_pre[SEG_MAX + 1];  // exclusive prefix of the tile counts
    __shared__ u32 s_w[SEG_THREADS / 32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32       carry = 0;
    for (u32 base = 0; base < nseg; base += SEG_THREADS) {
        const u32 i = base + threadIdx.x;
        const u32 t = i < nseg ? (u32)((seg[nseg + i] + MSD_TILE - 1) / MSD_TILE) : 0u;
        u32       incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 v = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
            if (lane >= (u32)o) incl += v;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        u32 excl = incl - t + carry, tot = 0;
        for (u32 w = 0; w < SEG_THREADS / 32; w++) {
            if (w < warp) excl += s_w[w];
            tot += s_w[w];
        }
        if (i < nseg) s_pre[i] = excl;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        s_pre[nseg]       = carry;
        ntiles_at[MSD_NB] = carry;
    }
    __syncthreads();
    const u32 total = s_pre[nseg];
    for (u32 f = threadIdx.x; f < total; f += SEG_THREADS) {
        u32 lo = 0, hi = nseg;  // s_pre[lo] <= f < s_pre[hi]
        while (hi - lo > 1) {
            const u32 mid = (lo + hi) >> 1;
            if (s_pre[mid] <= f) lo = mid; else hi = mid;
        }
        const u64 k = f - s_pre[lo], start = seg[lo], count = seg[nseg + lo];
        MsdTile   e;
        e.start  = start + k * MSD_TILE;
        e.bucket = (u32)seg[2 * nseg + lo];
        e.valid  = (u32)min((u64)MSD_TILE, count - k * MSD_TILE);
        tiles[f] = e;
    }
}

__global__ void __launch_bounds__(MSD_THREADS) k_msd_hist2(const u64 *__restrict__ in,
                                                           const MsdTile *__restrict__ tiles,
                                                           const u64 *__restrict__ tile_prefix,
                                                           FieldFn fn2, u64 *__restrict__ hist2) {
    __shared__ u32 sh[MSD_NB];
    if (blockIdx.x >= tile_prefix[MSD_NB]) return;
    const MsdTile ti = tiles[blockIdx.x];
    sh[threadIdx.x]  = 0;
    __syncthreads();
    const u64 *src = in + ti.start;
#pragma unroll 4
    for (u32 i = threadIdx.x; i < ti.valid; i += MSD_THREADS) atomicAdd(&sh[fn2(ld_stream(src + i))], 1u);
    __syncthreads();
    const u32 c = sh[threadIdx.x];
    if (c) atomicAdd(&hist2[(u64)ti.bucket * MSD_NB + threadIdx.x], (u64)c);
}

// OPT = false: exact pass (cursors2 start at the bucket offsets a histogram pass produced).
// OPT = true: optimistic pass, no histogram: bucket b's 512 sub-buckets are equal slots of cap2[b]
// items from base2[b] on (cursors2 start at the slot starts); a reservation that does not fit its
// slot goes to the dump area and raises flag[bucket] (see LimitSlots); flagged buckets are
// repartitioned exactly by the k_msd_fix_* kernels, on the device, without asking the host.
// Persistent CTAs over the tile table; a full next tile is requested while this one is written.
template <bool OPT, bool ASYNC = false, bool BULK = false>
__global__ void __launch_bounds__(MSD_THREADS, 2)
    k_msd_scatter2(const u64 *__restrict__ in, u64 *__restrict__ out, const MsdTile *__restrict__ tiles,
                   const u64 *__restrict__ tile_prefix, const __grid_constant__ FieldFn fn2,
                   u64 *cursors2 /* [512][512] */, const u64 *__restrict__ base2,
                   const u64 *__restrict__ cap2, u32 *flag, int shrink, u32 pf_dist, u32 *any_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_items = (u64 *)smem_raw;
    u32 *s_cnt   = (u32 *)(s_items + MSD_TILE + (BULK ? MSD_NB : 0));
    u64 *s_goff  = (u64 *)(s_cnt + MSD_NB);
    __shared__ u32 s_wsum[32];
    // the tile entry is requested together with the tile count (the table has an entry, possibly
    // unused, for every CTA of the grid): one dependent round trip to L2 less before the items
    u64       t      = blockIdx.x;
    MsdTile   ti     = tiles[t];
    const u64 ntiles = tile_prefix[MSD_NB];
    if (t >= ntiles) return;
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    using Lim = typename std::conditional<OPT, LimitSlots, NoLimit>::type;
    u64     item[MSD_IPT];
    bool    have = false;  // item[] already holds tile t
    if (pf_dist && t + pf_dist < ntiles) {  // the tile of the CTA that takes this slot next (see level 1)
        const MsdTile pf = tiles[t + pf_dist];
        if (threadIdx.x * 16 < pf.valid) prefetch_l2(in + pf.start + threadIdx.x * 16);
    }
    while (true) {
        const u64 tn  = t + gridDim.x;
        MsdTile   nx  = ti;
        bool      pre = false;
        if (tn < ntiles) {
            nx  = tiles[tn];
            pre = nx.valid == (u32)MSD_TILE;
        }
        // (asynchronous staging needs this tile and the next one to be full tiles)
        if (ASYNC && ti.valid != (u32)MSD_TILE) pre = false;
        const u64 *nsrc = pre ? in + nx.start : nullptr;
        u64       *cur  = cursors2 + (u64)ti.bucket * MSD_NB;
        Lim        lim;
        if constexpr (OPT) lim = LimitSlots{base2, cap2, ti.bucket, (u32)shrink, flag, any_flag};
        if (ti.valid == (u32)MSD_TILE) {
            if (!have) {
                unstable_load<MSD_THREADS, MSD_IPT, true>(item, in + ti.start, ti.valid);
            } else if (ASYNC) {  // the previous iteration copied this tile into the staging buffer
                cp_async_wait_all();
                __syncthreads();
#pragma unroll
                for (int j = 0; j < MSD_IPT; j++) item[j] = s_items[threadIdx.x + j * MSD_THREADS];
            }
            unstable_process<MSD_THREADS, MSD_IPT, 9, FieldFn, true, false, Lim, ASYNC ? 1 : 0, BULK>(
                item, out, fn2, cur, ti.valid, s_items, s_cnt, s_goff, s_wsum, lim, nsrc);
            if (BULK) bulk_wait_read();
        } else {
            unstable_load<MSD_THREADS, MSD_IPT, false>(item, in + ti.start, ti.valid);
            unstable_process<MSD_THREADS, MSD_IPT, 9, FieldFn, false, false, Lim>(
                item, out, fn2, cur, ti.valid, s_items, s_cnt, s_goff, s_wsum, lim, nsrc);
        }
        if (tn >= ntiles) break;
        __syncthreads();
        t    = tn;
        ti   = nx;
        have = pre;
    }
}

// ---- optimistic (histogram-free) partitioning: sample, capacities, slots ----------------------
// One read of a jittered strided sample of R and S (128-byte lines, at most max_sectors / 4 per
// relation) gives (1) the OR / AND of the sampled keys -> the digit plan, computed by every CTA
// after a grid barrier exactly as the host computes it, (2) the sampled first-level histogram ->
// bin capacities of the first pass, (3) the number of sampled keys seen twice (a lossy hash table:
// atomicExch of the key, a hit when the previous occupant was the same key) -> whether the
// relation has duplicate keys at all.  Equal slots at the second level are only trusted for
// relations without duplicates; with duplicates the key frequencies may be skewed in ways a
// sample cannot see (Zipf 0.4 over 128 M keys overflows most first-level buckets although its
// hottest key shows up nine times in 2^20 samples), and the exact second level costs 3 % more.
// st (u64 words): [0..3] OR_R AND_R OR_S AND_S, [4] barrier, [6..7] duplicate keys seen R S,
// [8..9] items sampled R S, [16..528) hist R, [528..1040) hist S.
constexpr int SAMP_TAB_BITS = 20;  // hash table slots per relation
constexpr int SAMP_CTAS = 192, SAMP_THREADS = 512, SAMP_SPT = 3, SAMP_WORDS = 16 + 2 * MSD_NB;

__global__ void __launch_bounds__(SAMP_THREADS)
    k_msd_sample(const u64 *__restrict__ r, u64 nr, const u64 *__restrict__ s, u64 ns,
                 u32 max_sectors, u64 *st, u32 *tab /* [2][2^SAMP_TAB_BITS], zeroed */) {
    __shared__ u32 s_h[MSD_NB];
    const u32  rel  = blockIdx.y;
    const u64 *in   = rel ? s : r;
    const u64  n    = rel ? ns : nr;
    // whole 128-byte lines (16 items) are sampled, four consecutive threads per line: a quarter
    // of the distinct DRAM rows / TLB entries of isolated sectors for the same number of items
    const u64  nline = n >> 4;
    const u64  L     = min(nline, (u64)(max_sectors >> 2));
    const u64  S     = L * 4;
    const u64  step  = L ? nline / L : 0;
    const u32  per_cta = ((u32)((S + gridDim.x - 1) / gridDim.x) + 3u) & ~3u;  // <= SAMP_THREADS * SAMP_SPT
    u32        kor = 0, kand = ~0u;
    u64        x[SAMP_SPT][4];  // this thread's sampled items: all loads in flight together
    bool       ok[SAMP_SPT];
    for (int i = threadIdx.x; i < MSD_NB; i += SAMP_THREADS) s_h[i] = 0;
#pragma unroll
    for (int k = 0; k < SAMP_SPT; k++) {
        // CTA c samples the c-th part of the relation: an SM touches a few pages, not all of them
        const u32 w = (u32)k * SAMP_THREADS + threadIdx.x;
        const u64 i = (u64)blockIdx.x * per_cta + w;
        ok[k]       = w < per_cta && i < S;
        if (ok[k]) {
            const u64 l   = i >> 2;
            const u32 jit = ((u32)l * 2654435761u) >> 16;
            const u64 sec = ((l * nline) / L + (((u64)jit * step) >> 16)) * 4 + (i & 3);
#pragma unroll
            for (int u = 0; u < 4; u++) x[k][u] = in[sec * 4 + u];
        }
    }
#pragma unroll
    for (int k = 0; k < SAMP_SPT; k++) {
        if (ok[k]) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                kor |= (u32)(x[k][u] >> 32);
                kand &= (u32)(x[k][u] >> 32);
            }
        }
    }
    kor  = __reduce_or_sync(SMJ_FULL_MASK, kor);
    kand = __reduce_and_sync(SMJ_FULL_MASK, kand);
    if (lane_id() == 0) {
        const u64 vor = (u64)kor << 32, vand = ((u64)kand << 32) | 0xffffffffull;
        if (vor & ~ld_relaxed(st + 2 * rel)) atomicOr(st + 2 * rel, vor);
        if (~vand & ld_relaxed(st + 2 * rel + 1)) atomicAnd(st + 2 * rel + 1, vand);
    }
    // grid barrier (cooperative launch: all CTAs are resident)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(st + 4, 1ull);
        const u64 all = (u64)gridDim.x * gridDim.y;
        while (ld_relaxed(st + 4) < all) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();
    u64 bits[4];
#pragma unroll
    for (int q = 0; q < 4; q++) bits[q] = ld_relaxed(st + q);
    MsdPlan pl;
    if (!smj_msd_plan(bits, &pl)) return;
    u32 dups = 0;
#pragma unroll
    for (int k = 0; k < SAMP_SPT; k++) {
        if (ok[k]) {
            u32 old[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                atomicAdd(&s_h[(u32)(x[k][u] >> pl.shift1) & (MSD_NB - 1)], 1u);
                const u32 key = (u32)(x[k][u] >> 32);
                const u32 h   = (key * 2654435761u) >> (32 - SAMP_TAB_BITS);
                old[u]        = atomicExch(&tab[(rel << SAMP_TAB_BITS) | h], key + 1u);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) dups += old[u] == (u32)(x[k][u] >> 32) + 1u;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MSD_NB; i += SAMP_THREADS)
        if (s_h[i]) atomicAdd(st + 16 + rel * MSD_NB + i, (u64)s_h[i]);
    dups = __reduce_add_sync(SMJ_FULL_MASK, dups);
    if (lane_id() == 0 && dups) atomicAdd(st + 6 + rel, (u64)dups);
    if (blockIdx.x == 0 && threadIdx.x == 0) st[8 + rel] = S * 4;
}

// First-level bin regions from the sampled histogram, on the device (the host computes the same
// capacities, smj_msd_caps1, to size the buffer): tab[rel] = {cursors, starts, limits} x 512.
// One CTA per relation.  Raises flag[2 * rel] when the regions do not fit alloc[rel] items.
__global__ void __launch_bounds__(MSD_NB) k_msd_caps1(const u64 *__restrict__ st, u64 nr, u64 ns,
                                                      u64 alloc_r, u64 alloc_s, int shrink,
                                                      u64 *__restrict__ tab_r, u64 *__restrict__ tab_s,
                                                      u32 *flag) {
    __shared__ u64 s_c[MSD_NB];
    const u32 rel = blockIdx.x, b = threadIdx.x;
    const u64 n = rel ? ns : nr, alloc = rel ? alloc_s : alloc_r, nsamp = st[8 + rel];
    u64      *tab = rel ? tab_s : tab_r;
    if (!tab) return;
    const double scale = nsamp ? fmax(1.0, (double)n / (double)nsamp) : 1.0;
    const u64    cap   = smj_cap1(st[16 + rel * MSD_NB + b], scale, shrink);
    s_c[b]             = cap;
    __syncthreads();
    for (int o = 1; o < MSD_NB; o <<= 1) {
        u64 ac = (b >= (u32)o) ? s_c[b - o] : 0;
        __syncthreads();
        s_c[b] += ac;
        __syncthreads();
    }
    const u64 start     = s_c[b] - cap;
    tab[b]              = start;
    tab[MSD_NB + b]     = start;
    tab[2 * MSD_NB + b] = s_c[b];
    if (b == MSD_NB - 1 && s_c[b] > alloc) flag[2 * rel] = 1u;
}

// Second-level layout of the optimistic pass, from the first pass's cursors: per first-level bucket
// its item count (clipped to the capacity: an overflowing pass is repeated anyway), its tiles, and
// the slot capacity cap2 = avg + 7 sqrt(avg) + slack of its 512 sub-buckets (a multiple of 4 items,
// so every slot starts on a 32-byte sector).  CTA b redoes the (cheap) scans and writes the slot
// starts and initial cursors of bucket b's sub-buckets; CTA 0 also writes the per-bucket arrays.
// Raises *flag and plans no tiles when the slots do not fit alloc2 items.
// SEG = false: bucket b lies contiguously at tab1's start (one optimistic pass A on this GPU).
// SEG = true: bucket b is the `world` segments b * world .. of seg = {start[nseg], count[nseg],
// bucket[nseg]} (the regions the ranks of a join group filled); the tile table then comes from
// k_msd_seg_tiles and tab1 / tiles / tile_prefix are not used.
template <bool SEG>
__global__ void __launch_bounds__(MSD_NB) k_msd_level2_plan(const u64 *__restrict__ tab1, u64 alloc2,
                                                            u32 dense_slot,
                                                            u64 *__restrict__ tile_prefix,
                                                            u64 *__restrict__ bucket_count,
                                                            u64 *__restrict__ cap2,
                                                            u64 *__restrict__ base2,
                                                            u64 *__restrict__ starts,
                                                            u64 *__restrict__ cursors,
                                                            MsdTile *__restrict__ tiles, u32 *flag,
                                                            const u64 *__restrict__ seg, u32 world) {
    __shared__ u64 s_t[MSD_NB], s_c[MSD_NB], s_d[MSD_NB], s_cap[MSD_NB], s_cnt1[MSD_NB];
    const int b   = threadIdx.x;
    u64 c = 0;
    if (SEG) {
        const u32 nseg = MSD_NB * world;
        for (u32 g = 0; g < world; g++) c += seg[nseg + (u32)b * world + g];
    } else {
        const u64 st1 = tab1[MSD_NB + b];
        c             = min(tab1[b] - st1, tab1[2 * MSD_NB + b] - st1);
    }
    const u64 t   = (c + MSD_TILE - 1) / MSD_TILE;
    const u64 avg = (c + MSD_NB - 1) / MSD_NB;
    u64       cap = c ? ((avg + (u64)ceil(7.0 * sqrt((double)avg + 1.0)) + 8 + 3) & ~3ull) : 0ull;
    // Relations that come here have (practically) unique keys, so a sub-bucket cannot hold more
    // than dense_slot = 2^lbits items, and a first-level bucket that is only partly filled (the top
    // of a dense key domain) fills the sub-buckets it does use completely: give such a bucket slots
    // of that size -- if the whole layout still fits (sparse key domains: 2^lbits is large).
    const u64 dcap = c ? max(cap, (min((u64)dense_slot, c) + 3) & ~3ull) : 0ull;
    s_t[b]        = t;
    s_c[b]        = cap * MSD_NB;
    s_d[b]        = dcap * MSD_NB;
    s_cnt1[b]     = c;
    __syncthreads();
    for (int o = 1; o < MSD_NB; o <<= 1) {
        u64 at = (b >= o) ? s_t[b - o] : 0, ac = (b >= o) ? s_c[b - o] : 0, ad = (b >= o) ? s_d[b - o] : 0;
        __syncthreads();
        s_t[b] += at;
        s_c[b] += ac;
        s_d[b] += ad;
        __syncthreads();
    }
    if (s_d[MSD_NB - 1] <= alloc2) {  // the roomier layout fits: use it (uniform for the whole CTA)
        cap = dcap;
        __syncthreads();
        s_c[b] = s_d[b];
    }
    s_cap[b] = cap;
    __syncthreads();
    const bool fits = s_c[MSD_NB - 1] <= alloc2;
    if (blockIdx.x == 0) {
        if (!SEG) tile_prefix[b] = fits ? s_t[b] - t : 0;
        bucket_count[b] = c;
        cap2[b]         = cap;
        base2[b]        = s_c[b] - cap * MSD_NB;
        if (b == MSD_NB - 1) {
            if (!SEG) tile_prefix[MSD_NB] = fits ? s_t[b] : 0;
            if (SEG && !fits) tile_prefix[MSD_NB] = 0;  // (k_msd_seg_tiles ran before: plan no tiles)
            if (!fits) *flag = 1u;
        }
    }
    const u64 mycap = s_cap[blockIdx.x];
    const u64 v     = s_c[blockIdx.x] - mycap * MSD_NB + (u64)b * mycap;
    starts[blockIdx.x * MSD_NB + b]  = v;
    cursors[blockIdx.x * MSD_NB + b] = v;
    if (!SEG && fits) {
        const u64 mycnt = s_cnt1[blockIdx.x], myt = (mycnt + MSD_TILE - 1) / MSD_TILE;
        msd_write_tiles(tiles, blockIdx.x, s_t[blockIdx.x] - myt, tab1[MSD_NB + blockIdx.x], mycnt);
    }
}

// first-level bucket sizes from the table an optimistic pass A left (cursors, starts, limits)
__global__ void __launch_bounds__(MSD_NB) k_msd_counts_opt(const u64 *__restrict__ tab,
                                                           u64 *__restrict__ counts) {
    const int b = threadIdx.x;
    counts[b]   = min(tab[b] - tab[MSD_NB + b], tab[2 * MSD_NB + b] - tab[MSD_NB + b]);
}

// ---- exact repeat of the second level for the flagged first-level buckets, all on the device ----
// A first-level bucket whose items are not spread evenly over its 512 sub-buckets (the partly
// filled bucket at the edge of a dense key domain, a locally clustered key range) overflows some
// of its equal slots.  Its region still holds the whole bucket (512 slots >= its item count), so
// the bucket is simply partitioned again, exactly, inside its own region: histogram of the
// bucket, exclusive scan -> sub-bucket starts, scatter.  The cost is proportional to the flagged
// buckets; with none flagged the three launches return at once.
__device__ __forceinline__ u32 msd_flagged_list(const u32 *__restrict__ flag, u32 *s_list, u32 *s_wcnt) {
    const u32 b = threadIdx.x, f = flag[b] != 0, warp = b >> 5, lane = b & 31;
    const u32 m = __ballot_sync(SMJ_FULL_MASK, f);
    if (lane == 0) s_wcnt[warp] = __popc(m);
    __syncthreads();
    u32 base = 0, total = 0;
    for (u32 w = 0; w < MSD_NB / 32; w++) {
        const u32 c = s_wcnt[w];
        if (w < warp) base += c;
        total += c;
    }
    if (f) s_list[base + __popc(m & lanemask_lt())] = b;
    __syncthreads();
    return total;
}

// the tiles of all flagged buckets as one index space: s_tp[q] = first tile of the q-th flagged
// bucket, returns the total; (bucket, tile) of a flat index by binary search in s_tp
__device__ __forceinline__ u32 msd_flagged_tiles(const u64 *__restrict__ bucket_count, const u32 *s_list,
                                                 u32 nflag, u32 *s_tp) {
    if (threadIdx.x == 0) {
        u32 acc = 0;
        for (u32 q = 0; q < nflag; q++) {
            s_tp[q] = acc;
            acc += (u32)((bucket_count[s_list[q]] + MSD_TILE - 1) / MSD_TILE);
        }
        s_tp[nflag] = acc;
    }
    __syncthreads();
    return s_tp[nflag];
}
__device__ __forceinline__ u32 msd_flat_find(const u32 *s_tp, u32 nflag, u32 f) {
    u32 lo = 0, hi = nflag;  // s_tp[lo] <= f < s_tp[hi]
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (s_tp[mid] <= f) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(MSD_THREADS) k_msd_fix_hist(const u64 *__restrict__ in,
                                                              const u64 *__restrict__ bucket_start,
                                                              const u64 *__restrict__ bucket_count,
                                                              const u32 *__restrict__ flag, FieldFn fn2,
                                                              u32 *__restrict__ hist2 /* [512][512], zeroed */) {
    __shared__ u32 s_list[MSD_NB], s_wcnt[MSD_NB / 32], sh[MSD_NB], s_tp[MSD_NB + 1];
    const u32 nflag = msd_flagged_list(flag, s_list, s_wcnt);
    if (nflag == 0) return;
    const u32 total = msd_flagged_tiles(bucket_count, s_list, nflag, s_tp);
    for (u32 f = blockIdx.x; f < total; f += gridDim.x) {
        const u32 q = msd_flat_find(s_tp, nflag, f), b = s_list[q];
        const u64 first = bucket_start[b], cnt = bucket_count[b];
        const u64 lo = (u64)(f - s_tp[q]) * MSD_TILE, hi = min(cnt, lo + MSD_TILE);
        sh[threadIdx.x] = 0;
        __syncthreads();
        for (u64 i = lo + threadIdx.x; i < hi; i += MSD_THREADS)
            atomicAdd(&sh[fn2(ld_stream(in + first + i))], 1u);
        __syncthreads();
        const u32 c = sh[threadIdx.x];
        if (c) atomicAdd(&hist2[b * MSD_NB + threadIdx.x], c);
        __syncthreads();
    }
}

// one CTA per first-level bucket: for a flagged bucket the exact sub-bucket starts inside its
// region (and the cursors); *nfixed counts the flagged buckets for the statistics
__global__ void __launch_bounds__(MSD_NB) k_msd_fix_offsets(const u32 *__restrict__ flag,
                                                            const u32 *__restrict__ hist2,
                                                            const u64 *__restrict__ base2,
                                                            u64 *__restrict__ starts,
                                                            u64 *__restrict__ cursors, u64 *nfixed) {
    const u32 b = blockIdx.x, d = threadIdx.x;
    if (!flag[b]) return;
    __shared__ u32 s_w[MSD_NB / 32];
    const u32 c = hist2[b * MSD_NB + d], lane = d & 31, warp = d >> 5;
    u32       incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u32 excl = incl - c;
    for (u32 w = 0; w < warp; w++) excl += s_w[w];
    const u64 v = base2[b] + excl;
    starts[b * MSD_NB + d]  = v;
    cursors[b * MSD_NB + d] = v;
    if (d == 0) atomicAdd(nfixed, 1ull);
}

__global__ void __launch_bounds__(MSD_THREADS, 2)
    k_msd_fix_scatter(const u64 *__restrict__ in, u64 *__restrict__ out,
                      const u64 *__restrict__ bucket_start, const u64 *__restrict__ bucket_count,
                      const u32 *__restrict__ flag, const __grid_constant__ FieldFn fn2, u64 *cursors2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_items = (u64 *)smem_raw;
    u32 *s_cnt   = (u32 *)(s_items + MSD_TILE);
    u64 *s_goff  = (u64 *)(s_cnt + MSD_NB);
    __shared__ u32 s_wsum[32];
    __shared__ u32 s_list[MSD_NB], s_wcnt[MSD_NB / 32], s_tp[MSD_NB + 1];
    const u32 nflag = msd_flagged_list(flag, s_list, s_wcnt);
    if (nflag == 0) return;
    const u32 total = msd_flagged_tiles(bucket_count, s_list, nflag, s_tp);
    s_cnt[threadIdx.x] = 0;
    for (u32 f = blockIdx.x; f < total; f += gridDim.x) {
        const u32 q = msd_flat_find(s_tp, nflag, f), b = s_list[q];
        const u64 first = bucket_start[b], cnt = bucket_count[b];
        const u64 lo    = (u64)(f - s_tp[q]) * MSD_TILE;
        const u32 valid = (u32)min((u64)MSD_TILE, cnt - lo);
        u64      *cur   = cursors2 + (u64)b * MSD_NB;
        __syncthreads();  // the previous tile is done with the staging area, s_cnt is zero
        if (valid == (u32)MSD_TILE)
            unstable_tile<MSD_THREADS, MSD_IPT, 9, FieldFn, true>(in + first + lo, out, fn2, cur, valid,
                                                                  s_items, s_cnt, s_goff, s_wsum);
        else
            unstable_tile<MSD_THREADS, MSD_IPT, 9, FieldFn, false>(in + first + lo, out, fn2, cur, valid,
                                                                   s_items, s_cnt, s_goff, s_wsum);
    }
}

// exclusive scan of nb entries in two launches: per-block scan + block totals, then add the
// totals of the preceding blocks (nb / 1024 <= 1024 blocks)
__global__ void __launch_bounds__(1024) k_scan_local(const u64 *__restrict__ in, u64 nb,
                                                     u64 *__restrict__ out, u64 *__restrict__ totals) {
    __shared__ u64 s_w[32];
    const u64 i    = (u64)blockIdx.x * 1024 + threadIdx.x;
    const u64 v    = i < nb ? in[i] : 0ull;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64       incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 t = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    u64 base = 0;
    for (u32 w = 0; w < warp; w++) base += s_w[w];
    if (i < nb) out[i] = base + incl - v;
    if (threadIdx.x == 1023) totals[blockIdx.x] = base + incl;
}
__global__ void __launch_bounds__(1024) k_scan_add(u64 *__restrict__ out, u64 nb,
                                                   const u64 *__restrict__ totals) {
    __shared__ u64 s_w[32];
    u64 p = (threadIdx.x < blockIdx.x) ? totals[threadIdx.x] : 0ull;  // blockIdx.x <= 1024
    p     = warp_reduce_add(p);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = p;
    __syncthreads();
    u64 base = 0;
    for (int w = 0; w < 32; w++) base += s_w[w];
    const u64 i = (u64)blockIdx.x * 1024 + threadIdx.x;
    if (i < nb) out[i] += base;
}

}  // namespace

// exclusive scan of n <= 2^20 u64 entries (device arrays; out may not alias in)
int smj_exclusive_scan_u64(SmjCtx *ctx, const u64 *in, u64 n, u64 *out) {
    const unsigned nblk = (unsigned)div_up(n, 1024);
    if (nblk == 0) return 0;
    if (nblk > 1024) return smj_fail_msg("smj_exclusive_scan_u64: too many entries");
    SmjScratch sc(ctx);
    u64 *d_tot = (u64 *)sc.get((size_t)nblk * 8);
    if (!d_tot) return smj_fail_msg("smj_exclusive_scan_u64: scratch");
    SMJ_LAUNCH(ctx, k_scan_local, nblk, 1024, 0, in, n, out, d_tot);
    SMJ_LAUNCH(ctx, k_scan_add, nblk, 1024, 0, out, n, d_tot);
    return 0;
}

static int msd_scatter2_attr(SmjCtx *ctx) {
    constexpr size_t SMEM = (size_t)MSD_TILE * 8 + MSD_NB * 4 + MSD_NB * 8;
    if (smj_func_smem(ctx, (const void *)k_msd_scatter2<false>, SMEM)) return -1;
    if (smj_func_smem(ctx, (const void *)k_msd_scatter2<true>, SMEM)) return -1;
    if (smj_func_smem(ctx, (const void *)k_msd_scatter2<true, true>, SMEM)) return -1;
    if (smj_func_smem(ctx, (const void *)k_msd_scatter2<true, false, true>, SMEM + MSD_NB * 8)) return -1;
    return smj_func_smem(ctx, (const void *)k_msd_fix_scatter, SMEM);
}
constexpr size_t MSD_SCATTER_SMEM = (size_t)MSD_TILE * 8 + MSD_NB * 4 + MSD_NB * 8;

// Pass B of the MSD partition: `a` holds n items already partitioned by the first-level digit,
// d_counts (device, 512 entries) the sizes of the first-level buckets, d_starts (device, 512
// entries, or null = back to back) where each begins in a.  Every bucket is partitioned by the
// 9-bit digit at shift2 into b.  offsets (device, 2^18 + 1 entries) receives the start of every
// second-level bucket in b (and n): bucket q is b[offsets[q] .. offsets[q + 1]).
// d_n (device, may be null): the exact number of items when n is only an upper bound (multi-GPU
// join: what a rank received is known on the device only)
static int msd_pass_b_common(SmjCtx *ctx, const u64 *a, u64 n, const u64 *d_counts, const u64 *d_starts,
                             u32 nseg, const u64 *d_seg, u64 *b, u32 shift2, u64 *offsets,
                             const u64 *d_n = nullptr) {
    const u32  NB2 = MSD_NB * MSD_NB;
    SmjScratch sc(ctx);
    u64 *d_tpref = (u64 *)sc.get((MSD_NB + 1) * 8), *d_bstart = (u64 *)sc.get(MSD_NB * 8);
    u64 *d_bcount = (u64 *)sc.get(MSD_NB * 8);
    u64 *d_hist2 = (u64 *)sc.get((size_t)NB2 * 8), *d_cur2 = (u64 *)sc.get((size_t)NB2 * 8);
    u64 *d_tot   = (u64 *)sc.get((NB2 / 1024) * 8);
    const unsigned max_tiles = (unsigned)(n / MSD_TILE + std::max<u32>(MSD_NB, nseg));
    MsdTile *d_tiles = (MsdTile *)sc.get((size_t)max_tiles * sizeof(MsdTile));
    if (!d_tpref || !d_bstart || !d_bcount || !d_hist2 || !d_cur2 || !d_tot || !d_tiles)
        return smj_fail_msg("smj_msd_pass_b: scratch allocation failed");
    SMJ_CHECK(cudaMemsetAsync(d_hist2, 0, (size_t)NB2 * 8, ctx->stream));
    FieldFn f2 = {shift2, (u32)MSD_NB - 1};
    if (d_seg) {
        SMJ_LAUNCH(ctx, k_msd_seg_tiles, 1, SEG_THREADS, 0, d_seg, nseg, d_tiles, d_tpref);
    } else {
        SMJ_LAUNCH(ctx, k_msd_tile_prefix, 1, MSD_NB, 0, d_counts, d_starts, d_tpref, d_bstart, d_bcount);
        SMJ_LAUNCH(ctx, k_msd_tile_table, MSD_NB, 256, 0, d_tpref, d_bstart, d_bcount, d_tiles);
    }
    const unsigned grid2 = max_tiles;  // one CTA per tile (see launch_unstable_shape)
    if (msd_scatter2_attr(ctx)) return -1;
    SMJ_ITEMS(ctx, n);
    SMJ_LAUNCH(ctx, k_msd_hist2, max_tiles, MSD_THREADS, 0, a, d_tiles, d_tpref, f2, d_hist2);
    SMJ_LAUNCH(ctx, k_scan_local, NB2 / 1024, 1024, 0, d_hist2, (u64)NB2, offsets, d_tot);
    SMJ_LAUNCH(ctx, k_scan_add, NB2 / 1024, 1024, 0, offsets, (u64)NB2, d_tot);
    SMJ_CHECK(cudaMemcpyAsync(d_cur2, offsets, (size_t)NB2 * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_n)
        SMJ_CHECK(cudaMemcpyAsync(offsets + NB2, d_n, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    else
        SMJ_CHECK(cudaMemcpyAsync(offsets + NB2, &n, 8, cudaMemcpyHostToDevice, ctx->stream));
    SMJ_ITEMS(ctx, n);
    SMJ_LAUNCH(ctx, k_msd_scatter2<false>, grid2, MSD_THREADS, MSD_SCATTER_SMEM, a, b, d_tiles, d_tpref, f2,
               d_cur2, (const u64 *)nullptr, (const u64 *)nullptr, (u32 *)nullptr, 0, (u32)(ctx->sm_count * 2), (u32 *)nullptr);
    return 0;
}

int smj_msd_pass_a_opt(SmjCtx *ctx, const u64 *in, u64 n, u64 *a, u32 shift1, u64 *d_tab, u32 *d_flag,
                       u32 kmask, u32 kref, u32 *d_kflag);
int smj_msd_pass_b(SmjCtx *ctx, const u64 *a, u64 n, const u64 *d_counts, const u64 *d_starts, u64 *b,
                   u32 shift2, u64 *offsets) {
    return msd_pass_b_common(ctx, a, n, d_counts, d_starts, 0, nullptr, b, shift2, offsets);
}

// Pass B over SEGMENTS: the n items of `a` lie in nseg segments, d_seg (device) = {start[nseg],
// count[nseg], bucket[nseg]}; the segments of one first-level bucket need not be adjacent (after a
// multi-GPU exchange every source rank contributes one segment per bucket).
int smj_msd_pass_b_seg(SmjCtx *ctx, const u64 *a, u64 n, u32 nseg, const u64 *d_seg, u64 *b, u32 shift2,
                       u64 *offsets) {
    if (nseg == 0 || nseg > (u32)SEG_MAX) return smj_fail_msg("smj_msd_pass_b_seg: bad segment count");
    return msd_pass_b_common(ctx, a, n, nullptr, nullptr, nseg, d_seg, b, shift2, offsets);
}
// the same with the segment table AND the item count on the device (n_upper bounds the count)
int smj_msd_pass_b_seg_dev(SmjCtx *ctx, const u64 *a, u64 n_upper, u32 nseg, const u64 *d_seg, const u64 *d_n,
                           u64 *b, u32 shift2, u64 *offsets) {
    if (nseg == 0 || nseg > (u32)SEG_MAX) return smj_fail_msg("smj_msd_pass_b_seg: bad segment count");
    return msd_pass_b_common(ctx, a, n_upper, nullptr, nullptr, nseg, d_seg, b, shift2, offsets, d_n);
}
// first-level pass with caller-built tables on the DEVICE: d_tab = {cursors, starts, limits} x 512,
// cursors / limits are absolute item addresses (pointer / 8 + offset, possibly peer memory)
// big_tile: tiles of 16384 items (one CTA of 1024 threads per SM) instead of 8192: a (tile, bin) run
// is then 256 bytes on average instead of 128.  Into local HBM the smaller tile is faster (two CTAs
// per SM overlap their phases); into PEER memory the pass is bound by NVLink, whose packets carry up
// to 256 bytes, and the longer runs cross it faster.
int smj_msd_pass_a_abs(SmjCtx *ctx, const u64 *in, u64 n, u32 shift1, u64 *d_tab, u32 *d_flag, u32 kmask,
                       u32 kref, u32 *d_kflag, int big_tile, int max_grid) {
    if (n == 0) return 0;
    FieldFn          f1  = {shift1, (u32)MSD_NB - 1};
    const LimitArray lim = {d_tab + 2 * MSD_NB, d_flag, kmask, kref, d_kflag};
    if (big_tile == 2)  // experiment: 12288-item tiles, two CTAs per SM (register-starved: spills)
        return launch_unstable_shape<512, 24, 2, 9, FieldFn, false, LimitArray>(ctx, in, (u64 *)nullptr, n, f1, d_tab, lim);
    // Peer destinations: the runs of a bin leave as bulk copies (one TMA instruction per run instead of
    // an LDS + LDS + STG per item); measured at N = 2: pass A 1.206 -> 1.157 ms.  (Into local HBM the
    // same variant is SLOWER, 0.46 -> 0.54 ms: 512 small copies per tile bound the copy engine.)
    if (big_tile)
        return launch_unstable_shape<1024, 16, 1, 9, FieldFn, false, LimitArray>(ctx, in, (u64 *)nullptr, n, f1, d_tab, lim,
                                                                                 (unsigned)max_grid, env_int("SMJ_XG_BULK", 1));
    return launch_unstable<9, FieldFn, false, LimitArray>(ctx, in, (u64 *)nullptr, n, f1, d_tab, lim);
}
u64 smj_smj_cap1(u64 k, double scale) { return smj_cap1(k, scale, 0); }

// ---- optimistic passes -------------------------------------------------------------------------
// Sample of R and S (see k_msd_sample).  host: SMJ_MSD_SAMPLE_WORDS u64.  Synchronises.
// Returns 1 (no error recorded) when the cooperative kernel cannot run on this device.
int smj_msd_sample(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, u32 max_sectors, u64 *d_st,
                   u32 *d_sub, u64 *host) {
    // the kernel contains a grid barrier: size the grid to what the device can hold at once (a
    // smaller GPU, a MIG slice) and return 1 when even that fails, so that the caller can fall back
    // to the exact passes instead of failing the join
    auto occ = ctx->func_occ.find((const void *)k_msd_sample);
    if (occ == ctx->func_occ.end()) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_msd_sample, SAMP_THREADS, 0) != cudaSuccess) {
            cudaGetLastError();
            per_sm = 0;
        }
        occ = ctx->func_occ.emplace((const void *)k_msd_sample, per_sm * ctx->sm_count).first;
    }
    const int max_ctas = occ->second;
    const int   ctas = std::min(SAMP_CTAS, max_ctas / 2);
    const char *hook = getenv("SMJ_MSD_TEST");  // "nocoop": behave like a device without cooperative launch
    if (ctas < 1 || (hook && !strcmp(hook, "nocoop"))) return 1;
    max_sectors = std::min<u32>(max_sectors, (u32)ctas * SAMP_THREADS * SAMP_SPT);
    SMJ_CHECK(cudaMemsetAsync(d_st, 0, SAMP_WORDS * 8, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_sub, 0, ((size_t)2 << SAMP_TAB_BITS) * 4, ctx->stream));
    const u64 init[4] = {0ull, ~0ull, 0ull, ~0ull};
    SMJ_CHECK(cudaMemcpyAsync(d_st, init, 32, cudaMemcpyHostToDevice, ctx->stream));
    // cooperative launch: the kernel contains a grid barrier, so the runtime must guarantee (or
    // refuse) that all 2 x SAMP_CTAS CTAs are resident at once
    void *args[] = {(void *)&r, (void *)&nr, (void *)&s, (void *)&ns, (void *)&max_sectors, (void *)&d_st,
                    (void *)&d_sub};
    if (ctx->profiling) smj_prof_before(ctx, "k_msd_sample");
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_msd_sample, dim3(ctas, 2),
                                                dim3(SAMP_THREADS), args, 0, ctx->stream);
    if (ctx->profiling) smj_prof_after(ctx);
    ctx->next_items = 0;
    smj_count_launch(ctx);
    if (e != cudaSuccess) {  // e.g. cooperative launches not available here
        cudaGetLastError();
        return 1;
    }
    if (host) {
        SMJ_CHECK(cudaMemcpyAsync(host, d_st, SAMP_WORDS * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}
size_t smj_msd_sample_bytes(int which) {
    return which == 0 ? (size_t)SAMP_WORDS * 8 : ((size_t)2 << SAMP_TAB_BITS) * 4;
}

// First-level bin regions from a sampled histogram (k of nsamp items fell into the bin):
// capacity = (k + 7 sqrt(k) + slack) n / nsamp, i.e. the estimate plus seven standard deviations
// of the sampling error, rounded up to 16 items.  Returns the number of items the regions span
// (what k_msd_caps1 lays out on the device).  shrink (tests only) halves the capacities.
u64 smj_msd_caps1(const u64 *hist, u64 nsamp, u64 n, int shrink) {
    const double scale = nsamp ? std::max(1.0, (double)n / (double)nsamp) : 1.0;
    u64          tot   = 0;
    for (int b = 0; b < MSD_NB; b++) tot += smj_cap1(hist[b], scale, shrink);
    return tot;
}
// device tables of both relations (null = relation not partitioned optimistically)
int smj_msd_caps1_tables(SmjCtx *ctx, const u64 *d_st, u64 nr, u64 ns, u64 alloc_r, u64 alloc_s, int shrink,
                         u64 *tab_r, u64 *tab_s, u32 *d_flags) {
    SMJ_LAUNCH(ctx, k_msd_caps1, 2, MSD_NB, 0, d_st, nr, ns, alloc_r, alloc_s, shrink, tab_r, tab_s, d_flags);
    return 0;
}
// upper bound of the items the second-level slots of an n-item relation span (k_msd_level2_plan)
u64 smj_msd_alloc2(u64 n) {
    const double per = (double)n / (double)(MSD_NB * MSD_NB);
    return n + (u64)((double)(MSD_NB * MSD_NB) * (7.0 * sqrt(per + 3.0) + 16.0)) + 4096;
}

// Optimistic pass A: in -> a, bin b into a[start[b] .. limit[b]), no histogram.  d_tab (device,
// 3 x 512 u64 = {cursors, starts, limits}, laid out by k_msd_caps1; the cursors end up as the bin
// ends); *d_flag is raised when a bin overflowed (its items are then missing from a).  Every key is
// checked against the sampled key range on the way: a key
// that differs from kref on a bit position of kmask raises *d_kflag.
int smj_msd_pass_a_opt(SmjCtx *ctx, const u64 *in, u64 n, u64 *a, u32 shift1, u64 *d_tab, u32 *d_flag,
                       u32 kmask, u32 kref, u32 *d_kflag) {
    if (n == 0) return 0;
    FieldFn          f1  = {shift1, (u32)MSD_NB - 1};
    const LimitArray lim = {d_tab + 2 * MSD_NB, d_flag, kmask, kref, d_kflag};
    // (tile shapes 256x16, 512x12, 384x16 and 1024x8 were measured: 512x16 with 2 CTAs per SM wins)
    return launch_unstable<9, FieldFn, false, LimitArray>(ctx, in, a, n, f1, d_tab, lim);
}

int smj_msd_counts_opt(SmjCtx *ctx, const u64 *d_tab, u64 *d_counts) {
    SMJ_LAUNCH(ctx, k_msd_counts_opt, 1, MSD_NB, 0, d_tab, d_counts);
    return 0;
}

// Optimistic pass B: a (as left by smj_msd_pass_a_opt, d_tab its table) -> b, every first-level
// bucket into 512 equal slots, no histogram; buckets with an overflowing slot are repartitioned
// exactly inside their region by the k_msd_fix_* kernels.  starts / ends (device, 2^18 entries
// each): sub-bucket q is b[starts[q] .. ends[q]).  d_counts (device, 512) receives the
// first-level bucket sizes.  alloc2 = items b can hold; dense_slot = 2^(key bits left inside a
// sub-bucket): the most items a sub-bucket of a relation with unique keys can hold.  *d_flag is raised when the slots do not fit alloc2 (nothing is partitioned then);
// *d_nfixed (u64) counts the repartitioned buckets.
int smj_msd_pass_b_opt(SmjCtx *ctx, const u64 *a, u64 n, const u64 *d_tab, u64 *b, u64 alloc2, u32 shift2,
                       u32 dense_slot, u64 *starts, u64 *ends, u64 *d_counts, u32 *d_flag, u64 *d_nfixed,
                       int shrink) {
    SmjScratch sc(ctx);
    u64 *d_tpref = (u64 *)sc.get((MSD_NB + 1) * 8), *d_cap2 = (u64 *)sc.get(MSD_NB * 8);
    u64 *d_base2 = (u64 *)sc.get(MSD_NB * 8);
    u32 *d_bflag = (u32 *)sc.get(MSD_NB * 4), *d_hist2 = (u32 *)sc.get((size_t)MSD_NB * MSD_NB * 4);
    if (!d_tpref || !d_cap2 || !d_base2 || !d_bflag || !d_hist2)
        return smj_fail_msg("smj_msd_pass_b_opt: scratch allocation failed");
    SMJ_CHECK(cudaMemsetAsync(d_bflag, 0, MSD_NB * 4, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(d_hist2, 0, (size_t)MSD_NB * MSD_NB * 4, ctx->stream));
    FieldFn f2 = {shift2, (u32)MSD_NB - 1};
    const unsigned max_tiles = (unsigned)(n / MSD_TILE + MSD_NB);
    MsdTile *d_tiles = (MsdTile *)sc.get((size_t)max_tiles * sizeof(MsdTile));
    if (!d_tiles) return smj_fail_msg("smj_msd_pass_b_opt: scratch allocation failed");
    SMJ_LAUNCH(ctx, k_msd_level2_plan<false>, MSD_NB, MSD_NB, 0, d_tab, alloc2, dense_slot, d_tpref, d_counts, d_cap2,
               d_base2, starts, ends, d_tiles, d_flag, (const u64 *)nullptr, 0u);
    const unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 4, max_tiles);
    if (msd_scatter2_attr(ctx)) return -1;
    SMJ_ITEMS(ctx, n);
    // SMJ_SCATTER_ASYNC=1: persistent CTAs (two per SM), the next tile of a CTA is copied into the
    // staging buffer with cp.async behind the output loop; else one CTA per tile
    if (env_int("SMJ_SCATTER_ASYNC", 0)) {
        SMJ_LAUNCH(ctx, (k_msd_scatter2<true, true>), (unsigned)std::min<u64>(max_tiles, (u64)ctx->sm_count * 2), MSD_THREADS,
                   MSD_SCATTER_SMEM, a, b, d_tiles, d_tpref, f2, ends, d_base2, d_cap2, d_bflag, shrink, 0u, (u32 *)nullptr);
    } else if (env_int("SMJ_SCATTER_BULK", 0)) {
        SMJ_LAUNCH(ctx, (k_msd_scatter2<true, false, true>), max_tiles, MSD_THREADS, MSD_SCATTER_SMEM + MSD_NB * 8, a, b, d_tiles,
                   d_tpref, f2, ends, d_base2, d_cap2, d_bflag, shrink, (u32)(ctx->sm_count * 2), (u32 *)nullptr);
    } else {
        SMJ_LAUNCH(ctx, k_msd_scatter2<true>, max_tiles, MSD_THREADS, MSD_SCATTER_SMEM, a, b, d_tiles, d_tpref, f2, ends,
                   d_base2, d_cap2, d_bflag, shrink, (u32)(ctx->sm_count * 2), (u32 *)nullptr);
    }
    SMJ_LAUNCH(ctx, k_msd_fix_hist, grid, MSD_THREADS, 0, a, d_tab + MSD_NB, d_counts, d_bflag, f2, d_hist2);
    SMJ_LAUNCH(ctx, k_msd_fix_offsets, MSD_NB, MSD_NB, 0, d_bflag, d_hist2, d_base2, starts, ends, d_nfixed);
    SMJ_LAUNCH(ctx, k_msd_fix_scatter, grid, MSD_THREADS, MSD_SCATTER_SMEM, a, b, d_tab + MSD_NB, d_counts,
               d_bflag, f2, ends);
    return 0;
}

// Optimistic pass B over SEGMENTS (join group): first-level bucket b = the `world` segments b * world ..
// of d_seg = {start[nseg], count[nseg], bucket[nseg]}, nseg = 512 * world, all on the device; a holds
// them, n_upper bounds the item count.  Every bucket goes into 512 equal slots of b (alloc2 items),
// no histogram.  A bucket whose slot overflowed cannot be repaired in place here (its items are not
// contiguous): *d_flag is raised instead -- also when the slots do not fit alloc2 -- and the caller
// repeats exactly.  starts / ends: 2^18 entries each, sub-bucket q is b[starts[q] .. ends[q]).
int smj_msd_pass_b_opt_seg(SmjCtx *ctx, const u64 *a, u64 n_upper, u32 world, const u64 *d_seg, u64 *b, u64 alloc2,
                           u32 shift2, u32 dense_slot, u64 *starts, u64 *ends, u32 *d_flag) {
    const u32 nseg = (u32)MSD_NB * world;
    if (world == 0 || nseg > (u32)SEG_MAX) return smj_fail_msg("smj_msd_pass_b_opt_seg: bad segment count");
    SmjScratch sc(ctx);
    u64 *d_tpref = (u64 *)sc.get((MSD_NB + 1) * 8), *d_cap2 = (u64 *)sc.get(MSD_NB * 8);
    u64 *d_base2 = (u64 *)sc.get(MSD_NB * 8), *d_counts = (u64 *)sc.get(MSD_NB * 8);
    u32 *d_bflag = (u32 *)sc.get(MSD_NB * 4);
    const unsigned max_tiles = (unsigned)(n_upper / MSD_TILE + nseg);
    MsdTile *d_tiles = (MsdTile *)sc.get((size_t)max_tiles * sizeof(MsdTile));
    if (!d_tpref || !d_cap2 || !d_base2 || !d_counts || !d_bflag || !d_tiles)
        return smj_fail_msg("smj_msd_pass_b_opt_seg: scratch allocation failed");
    SMJ_CHECK(cudaMemsetAsync(d_bflag, 0, MSD_NB * 4, ctx->stream));
    FieldFn f2 = {shift2, (u32)MSD_NB - 1};
    SMJ_LAUNCH(ctx, k_msd_seg_tiles, 1, SEG_THREADS, 0, d_seg, nseg, d_tiles, d_tpref);
    SMJ_LAUNCH(ctx, k_msd_level2_plan<true>, MSD_NB, MSD_NB, 0, (const u64 *)nullptr, alloc2, dense_slot, d_tpref, d_counts,
               d_cap2, d_base2, starts, ends, d_tiles, d_flag, d_seg, world);
    if (msd_scatter2_attr(ctx)) return -1;
    SMJ_ITEMS(ctx, n_upper);
    SMJ_LAUNCH(ctx, k_msd_scatter2<true>, max_tiles, MSD_THREADS, MSD_SCATTER_SMEM, a, b, d_tiles, d_tpref, f2, ends,
               d_base2, d_cap2, d_bflag, 0, (u32)(ctx->sm_count * 2), d_flag);
    return 0;
}

// first-level histogram (512 bins, digit at shift1) + exact OR / AND of the items.
// hist: device, 512 entries, overwritten.  exact_bits: device pair {OR, AND} that is ACCUMULATED
// into (the caller initialises it to {0, ~0}); may be null.
int smj_msd_hist1(SmjCtx *ctx, const u64 *in, u64 n, u32 shift1, u64 *hist, u64 *exact_bits) {
    SMJ_CHECK(cudaMemsetAsync(hist, 0, MSD_NB * 8, ctx->stream));
    if (n == 0) return 0;
    FieldFn f1 = {shift1, (u32)MSD_NB - 1};
    SMJ_ITEMS(ctx, n);
    SMJ_LAUNCH(ctx, k_msd_hist1, hist_grid(ctx, n, 512, 16), 512, 0, in, n, f1, hist, exact_bits);
    return 0;
}

// in -> a (pass A) -> b (pass B).  shift1 = position of the top 9-bit digit, shift2 = shift1 - 9.
// offsets: device array of 2^18 + 1 entries, receives the start of every bucket in b (and n).
// exact_bits: see smj_msd_hist1.
int smj_msd_partition(SmjCtx *ctx, const u64 *in, u64 n, u64 *a, u64 *b, u32 shift1, u32 shift2,
                      u64 *offsets, u64 *exact_bits) {
    SmjScratch sc(ctx);
    u64 *d_hist1 = (u64 *)sc.get(RADIX_NB * 8), *d_bases1 = (u64 *)sc.get(RADIX_NB * 8);
    if (!d_hist1 || !d_bases1) return smj_fail_msg("smj_msd_partition: scratch allocation failed");
    SMJ_CHECK(cudaMemsetAsync(d_hist1, 0, RADIX_NB * 8, ctx->stream));
    int rc = smj_msd_hist1(ctx, in, n, shift1, d_hist1, exact_bits);
    if (rc) return rc;
    SMJ_LAUNCH(ctx, k_bases_multi, 1, RADIX_NB, 0, d_hist1, d_bases1);
    FieldFn f1 = {shift1, (u32)MSD_NB - 1};
    if (n && (rc = launch_unstable<9>(ctx, in, a, n, f1, d_bases1))) return rc;  // consumes d_bases1
    return smj_msd_pass_b(ctx, a, n, d_hist1, nullptr, b, shift2, offsets);
}

// Pass A of the MSD partition with caller-supplied destinations: bucket d of the 9-bit digit at
// shift1 goes to dest_ptrs[d] (device pointer, possibly a peer GPU's buffer) starting at item
// dest_offsets[d].  Host arrays of 512 entries.  Used by the multi-GPU join, where the first
// partitioning pass IS the exchange.
int smj_msd_scatter_to(SmjCtx *ctx, const u64 *in, u64 n, u32 shift1, u64 *const *dest_ptrs,
                       const u64 *dest_offsets) {
    if (n == 0) return 0;
    u64 h_cur[RADIX_NB];
    for (int d = 0; d < MSD_NB; d++) {
        if ((uintptr_t)dest_ptrs[d] & 7) return smj_fail_msg("smj_msd_scatter_to: unaligned buffer");
        h_cur[d] = (u64)((uintptr_t)dest_ptrs[d] >> 3) + dest_offsets[d];
    }
    SmjScratch sc(ctx);
    u64 *d_cur = (u64 *)sc.get(RADIX_NB * 8);
    if (!d_cur) return smj_fail_msg("smj_msd_scatter_to: scratch");
    SMJ_CHECK(cudaMemcpyAsync(d_cur, h_cur, RADIX_NB * 8, cudaMemcpyHostToDevice, ctx->stream));
    FieldFn f1 = {shift1, (u32)MSD_NB - 1};
    // measurement hooks (tools/p2p_scatter_bw.py): 8-bit digit (256 bins: runs twice as long) and / or
    // 16384-item tiles
    const int bits = env_int("SMJ_SCATTER_TO_BITS", 9), big = env_int("SMJ_SCATTER_TO_BIG", 0);
    const u32 al = (u32)env_int("SMJ_SCATTER_TO_ALIGN", 1);
    SMJ_CHECK(cudaMemcpyToSymbolAsync(c_dbg_reserve_align, &al, 4, 0, cudaMemcpyHostToDevice, ctx->stream));
    struct Reset {  // back to exact reservations for every other caller
        SmjCtx *c;
        ~Reset() {
            const u32 one = 1;
            cudaMemcpyToSymbolAsync(c_dbg_reserve_align, &one, 4, 0, cudaMemcpyHostToDevice, c->stream);
            cudaStreamSynchronize(c->stream);
        }
    } reset{ctx};
    if (bits == 8) {
        f1.mask = 255;
        if (big) return launch_unstable_shape<1024, 16, 1, 8, FieldFn, false, NoLimit>(ctx, in, (u64 *)nullptr, n, f1, d_cur, NoLimit());
        return launch_unstable<8>(ctx, in, (u64 *)nullptr, n, f1, d_cur);
    }
    if (big) return launch_unstable_shape<1024, 16, 1, 9, FieldFn, false, NoLimit>(ctx, in, (u64 *)nullptr, n, f1, d_cur, NoLimit());
    return launch_unstable<9>(ctx, in, (u64 *)nullptr, n, f1, d_cur);
}

// Pass A into a LOCAL buffer laid out bucket-major (bucket d starts at the exclusive prefix of
// hist512, a DEVICE array as smj_msd_hist1 leaves it): nothing goes through the host, so the pass
// can be enqueued while the histograms of the ranks are still being all-gathered.
int smj_msd_scatter_local(SmjCtx *ctx, const u64 *in, u64 n, u32 shift1, const u64 *d_hist512, u64 *out) {
    if (n == 0) return 0;
    SmjScratch sc(ctx);
    u64 *d_cur = (u64 *)sc.get(RADIX_NB * 8);
    if (!d_cur) return smj_fail_msg("smj_msd_scatter_local: scratch");
    SMJ_LAUNCH(ctx, k_bases_multi, 1, RADIX_NB, 0, d_hist512, d_cur);
    FieldFn f1 = {shift1, (u32)MSD_NB - 1};
    return launch_unstable<9>(ctx, in, out, n, f1, d_cur);
}

// Same from a strided sample of about 2^16 items per relation (a guess, to be verified).
int smj_key_bits_sampled(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, u64 *host) {
    SmjScratch sc(ctx);
    u64 *d = (u64 *)sc.get(32);
    if (!d) return smj_fail_msg("smj_key_bits: scratch allocation failed");
    const u64 init[4] = {0ull, ~0ull, 0ull, ~0ull};
    SMJ_CHECK(cudaMemcpyAsync(d, init, 32, cudaMemcpyHostToDevice, ctx->stream));
    if (nr) SMJ_LAUNCH(ctx, k_bits_sample, 64, 256, 0, r, nr, std::max<u64>(1, nr >> 16), d);
    if (ns) SMJ_LAUNCH(ctx, k_bits_sample, 64, 256, 0, s, ns, std::max<u64>(1, ns >> 16), d + 2);
    SMJ_CHECK(cudaMemcpyAsync(host, d, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------
// Fused range partition + exchange: the scatter kernel stores every item straight into the
// receive buffer of the GPU that owns its key range.  dest_ptrs[d] may be a peer GPU's buffer
// (CUDA IPC mapping over NVLink) or a local one; dest_offsets[d] is where this rank's slice
// starts inside it (computed by the caller from the all-gathered count matrix).  The order of
// items inside a destination is irrelevant (the receiver partitions/sorts them anyway), so this
// is the unstable kernel with one cursor per destination holding an ABSOLUTE item index
// (address / 8): no pack buffer, no separate copy kernel, the transfer overlaps the partitioning
// tile by tile.
// ---------------------------------------------------------------------------------------
extern "C" int smjdev_range_count(const uint64_t *in, uint64_t n, const uint32_t *splitters,
                                  uint32_t nparts, uint64_t *counts) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (nparts < 1 || nparts > 32) return smj_fail_msg("smjdev_range_count: 1..32 parts");
    SplitFn fn;
    fn.nparts = nparts;
    for (u32 i = 0; i < 31; i++) fn.sp[i] = (i + 1 < nparts) ? splitters[i] : 0xffffffffu;
    SmjScratch sc(ctx);
    u64 *d_hist = (u64 *)sc.get(RADIX_NB * 8);
    if (!d_hist) return smj_fail_msg("smjdev_range_count: scratch");
    SMJ_CHECK(cudaMemsetAsync(d_hist, 0, RADIX_NB * 8, ctx->stream));
    if (n) {
        SMJ_ITEMS(ctx, n);
        SMJ_LAUNCH(ctx, k_hist_fn<SplitFn>, hist_grid(ctx, n, 512, 8), 512, 0, (const u64 *)in, n, fn,
                   d_hist);
    }
    SMJ_CHECK(cudaMemcpyAsync(counts, d_hist, (size_t)nparts * 8, cudaMemcpyDeviceToDevice,
                              ctx->stream));
    return 0;
}

extern "C" int smjdev_range_scatter(const uint64_t *in, uint64_t n, const uint32_t *splitters,
                                    uint32_t nparts, uint64_t *const *dest_ptrs,
                                    const uint64_t *dest_offsets) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (nparts < 1 || nparts > 32) return smj_fail_msg("smjdev_range_scatter: 1..32 parts");
    if (n == 0) return 0;
    SplitFn fn;
    fn.nparts = nparts;
    for (u32 i = 0; i < 31; i++) fn.sp[i] = (i + 1 < nparts) ? splitters[i] : 0xffffffffu;
    u64 h_cur[RADIX_NB];
    memset(h_cur, 0, sizeof(h_cur));
    for (u32 d = 0; d < nparts; d++) {
        if ((uintptr_t)dest_ptrs[d] & 7) return smj_fail_msg("smjdev_range_scatter: unaligned buffer");
        h_cur[d] = (u64)((uintptr_t)dest_ptrs[d] >> 3) + dest_offsets[d];
    }
    SmjScratch sc(ctx);
    u64 *d_cur = (u64 *)sc.get(RADIX_NB * 8);
    if (!d_cur) return smj_fail_msg("smjdev_range_scatter: scratch");
    SMJ_CHECK(cudaMemcpyAsync(d_cur, h_cur, RADIX_NB * 8, cudaMemcpyHostToDevice, ctx->stream));
    return launch_unstable<8, SplitFn, true>(ctx, (const u64 *)in, (u64 *)nullptr, n, fn, d_cur);
}
