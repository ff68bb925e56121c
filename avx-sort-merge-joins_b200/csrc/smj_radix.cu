// smj_radix.cu -- radix partitioning and LSD radix sort for 8-byte key|payload items.
//
// B200 restatement of two reference subsystems:
//   * src/partition/partition.c:94-287   radix_cluster{,_optimized,_optimized_V2}: histogram +
//     software-write-combining scatter  ->  k_hist_* + k_onesweep (shared-memory staged scatter,
//     warp-aggregated ranks via match.any, decoupled look-back instead of a second data pass).
//   * src/avxsort/avxsort.c:128-209      avxsort_aligned: in-register networks + bitonic merges
//     ->  one-sweep style LSD radix sort over the VARYING bits only (k_bits_reduce finds them).
//
// Every pass is STABLE (the reference's partition tests require stability,
// tests/check_partitioning.c:82-90, and LSD sorting needs it).
//
// HBM traffic per item: k_bits_reduce 8 B, k_hist_digits 8 B, each k_onesweep pass 16 B.
#include "smj_common.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <string.h>

namespace {

constexpr int RADIX_NB = 256;  // bins per pass (8-bit digits)

// ---------------------------------------------------------------------------------------
// OR / AND of all items: tells which bit positions vary at all (constant bits need no pass)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bits_reduce(const u64 *__restrict__ in, u64 n,
                                                     u64 *__restrict__ out /* [0]=OR [1]=AND */) {
    u64 vor = 0, vand = ~0ull;
    const u64 stride = (u64)gridDim.x * blockDim.x;
    u64       i      = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        u64 a = ld_stream(in + i), b = ld_stream(in + i + stride);
        u64 c = ld_stream(in + i + 2 * stride), d = ld_stream(in + i + 3 * stride);
        vor |= (a | b) | (c | d);
        vand &= (a & b) & (c & d);
    }
    for (; i < n; i += stride) {
        u64 a = ld_stream(in + i);
        vor |= a;
        vand &= a;
    }
    vor  = warp_reduce_or(vor);
    vand = warp_reduce_and(vand);
    __shared__ u64 s_or[8], s_and[8];
    if (lane_id() == 0) {
        s_or[threadIdx.x >> 5]  = vor;
        s_and[threadIdx.x >> 5] = vand;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
            vor |= s_or[w];
            vand &= s_and[w];
        }
        atomicOr(out, vor);
        atomicAnd(out + 1, vand);
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
                                                     u64 *__restrict__ hist /* [nd][256] */) {
    __shared__ u32 sh[8 * RADIX_NB];
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
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k < plan.nd) atomicAdd(&sh[k * RADIX_NB + plan.d[k](x[u])], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < plan.nd * RADIX_NB; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (u64)sh[i]);
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

template <int THREADS, int IPT>
struct OneSweepCfg {
    static constexpr int    WARPS = THREADS / 32;
    static constexpr int    TILE  = THREADS * IPT;
    static constexpr size_t SMEM =
        (size_t)TILE * 8 + (size_t)WARPS * RADIX_NB * 4 + RADIX_NB * 4 + RADIX_NB * 8;
};

template <int THREADS, int IPT, class F>
__global__ void __launch_bounds__(THREADS, 2)
    k_onesweep(const u64 *__restrict__ in, u64 *__restrict__ out, u64 n, F fn,
               const u64 *__restrict__ bases,  // [256] global start of every bin
               u64 *status,                    // [tiles][256] look-back state, zeroed
               u32 *tile_counter,              // zeroed
               DigitFn fn_full, const u64 *__restrict__ padshift) {
    using Cfg = OneSweepCfg<THREADS, IPT>;
    static_assert(THREADS >= RADIX_NB, "need one thread per bin");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_items     = (u64 *)smem_raw;                        // TILE staged items
    u32 *s_whist     = (u32 *)(s_items + Cfg::TILE);           // [WARPS][256]
    u32 *s_tile_base = s_whist + Cfg::WARPS * RADIX_NB;        // [256] exclusive scan of the tile
    u64 *s_goff      = (u64 *)(s_tile_base + RADIX_NB);        // [256] global offset - tile_base
    __shared__ u32 s_tile;
    __shared__ u32 s_wsum[32];

    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = threadIdx.x; i < Cfg::WARPS * RADIX_NB; i += THREADS) s_whist[i] = 0;
    __syncthreads();

    const u32 tile       = s_tile;
    const u64 tile_start = (u64)tile * Cfg::TILE;
    const u64 remain     = n - tile_start;
    const u32 valid      = remain < (u64)Cfg::TILE ? (u32)remain : (u32)Cfg::TILE;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 last_bin   = fn.last_bin();  // bins in use: 0..last_bin (<= 255)
    const u32 lt         = lanemask_lt();

    // ---- load (warp-striped so that rank order == memory order) and rank
    u64       item[IPT];
    u32       rank[IPT];
    const u32 wbase = warp * 32 * IPT;
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        u32 idx = wbase + j * 32 + lane;
        item[j] = (idx < valid) ? ld_stream(in + tile_start + idx) : ~0ull;
    }
    u32 *wh = s_whist + warp * RADIX_NB;
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        u32 idx    = wbase + j * 32 + lane;
        u32 d      = (idx < valid) ? fn(item[j]) : last_bin;
        u32 m      = __match_any_sync(SMJ_FULL_MASK, d);
        u32 leader = __ffs(m) - 1;
        u32 prev   = 0;
        if (lane == leader) {
            prev  = wh[d];
            wh[d] = prev + __popc(m);
        }
        __syncwarp();
        prev    = __shfl_sync(SMJ_FULL_MASK, prev, leader);
        rank[j] = prev + __popc(m & lt);
    }
    __syncthreads();

    // ---- per bin: exclusive scan over warps, then exclusive scan over bins
    u32 tile_cnt = 0;
    if (threadIdx.x < RADIX_NB) {
        u32 run = 0;
#pragma unroll
        for (int w = 0; w < Cfg::WARPS; w++) {
            u32 t                                = s_whist[w * RADIX_NB + threadIdx.x];
            s_whist[w * RADIX_NB + threadIdx.x] = run;
            run += t;
        }
        tile_cnt = run;
    }
    u32 incl = tile_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(SMJ_FULL_MASK, incl, o);
        if (lane >= (u32)o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    u32 excl = 0;
    if (threadIdx.x < RADIX_NB) {
        u32 woff = 0;
        for (u32 w = 0; w < warp; w++) woff += s_wsum[w];
        excl                     = woff + incl - tile_cnt;
        s_tile_base[threadIdx.x] = excl;
    }

    // ---- decoupled look-back: global rank of this tile's first item in every bin
    if (threadIdx.x <= last_bin) {
        const u32 b = threadIdx.x;
        // padding items of a partial tile were counted in the last bin; they are not data
        const u64 cnt  = (u64)tile_cnt - ((b == last_bin) ? (u64)(Cfg::TILE - valid) : 0ull);
        u64      *st   = status + (u64)tile * RADIX_NB + b;
        u64       prev = 0;
        if (tile == 0) {
            st_relaxed(st, ST_PREFIX | cnt);
        } else {
            st_relaxed(st, ST_AGG | cnt);
            long long t = (long long)tile - 1;
            while (true) {
                u64 s;
                do {
                    s = ld_relaxed(status + (u64)t * RADIX_NB + b);
                } while ((s >> 62) == 0);
                prev += s & ST_VAL;
                if ((s >> 62) == 2) break;
                t--;
            }
            st_relaxed(st, ST_PREFIX | (prev + cnt));
        }
        s_goff[b] = bases[b] + prev - (u64)excl;
    }
    __syncthreads();  // s_tile_base, s_whist (warp bases) and s_goff visible

    // ---- stage the tile in bin order, then write runs of consecutive addresses
#pragma unroll
    for (int j = 0; j < IPT; j++) {
        u32 idx = wbase + j * 32 + lane;
        u32 d   = (idx < valid) ? fn(item[j]) : last_bin;
        u32 pos = s_tile_base[d] + wh[d] + rank[j];
        s_items[pos] = item[j];
    }
    __syncthreads();
    if (padshift == nullptr) {
        for (u32 i = threadIdx.x; i < valid; i += THREADS) {
            u64 x = s_items[i];
            st_stream(out + s_goff[fn(x)] + i, x);
        }
    } else {
        for (u32 i = threadIdx.x; i < valid; i += THREADS) {
            u64 x = s_items[i];
            st_stream(out + s_goff[fn(x)] + i + padshift[fn_full(x)], x);
        }
    }
}

// tuning point: 384 threads x 16 items = 6144 items (48 KB) per tile, 2 CTAs / SM
constexpr int OS_THREADS = 384;
constexpr int OS_IPT     = 16;
using OSCfg              = OneSweepCfg<OS_THREADS, OS_IPT>;

struct PassState {
    u64 *status;
    u32 *counter;
};

template <class F>
int launch_onesweep(SmjCtx *ctx, const u64 *in, u64 *out, u64 n, F fn, const u64 *bases,
                    u64 *status, u32 *counter, DigitFn fn_full, const u64 *padshift) {
    static bool attr_set = false;
    if (!attr_set) {
        SMJ_CHECK(cudaFuncSetAttribute(k_onesweep<OS_THREADS, OS_IPT, F>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)OSCfg::SMEM));
        attr_set = true;
    }
    const u64 tiles = div_up(n, OSCfg::TILE);
    SMJ_ITEMS(ctx, n);
    SMJ_LAUNCH(ctx, (k_onesweep<OS_THREADS, OS_IPT, F>), (unsigned)tiles, OS_THREADS, OSCfg::SMEM,
               in, out, n, fn, bases, status, counter, fn_full, padshift);
    return 0;
}

// histogram of one bucket function with <= 256 bins
template <class F>
__global__ void __launch_bounds__(512) k_hist_fn(const u64 *__restrict__ in, u64 n, F fn,
                                                 u64 *__restrict__ hist) {
    __shared__ u32 sh[RADIX_NB];
    for (int i = threadIdx.x; i < RADIX_NB; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const u64 stride = (u64)gridDim.x * blockDim.x;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&sh[fn(ld_stream(in + i))], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX_NB; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (u64)sh[i]);
}

inline size_t status_bytes(u64 n) { return (size_t)div_up(n, OSCfg::TILE) * RADIX_NB * 8; }

int hist_grid(SmjCtx *ctx, u64 n, int threads, int per_thread) {
    u64 want = div_up(n, (u64)threads * per_thread);
    u64 cap  = (u64)ctx->sm_count * 8;
    return (int)std::max<u64>(1, std::min(want, cap));
}

}  // namespace

// ---------------------------------------------------------------------------------------
// LSD radix sort over the varying bits (internal entry used by smjdev_sort and the joins)
// ---------------------------------------------------------------------------------------
// Reads `in`, ping-pongs between the scratch buffers a and b (pass 1: in -> a, pass 2: a -> b,
// pass 3: b -> a, ...).  `in` may alias b (classic two-buffer sort) but not a.  *sorted is set
// to the buffer that holds the result, which is `in` itself when no pass was necessary.
int smj_lsd_sort(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, u64 **sorted) {
    *sorted = const_cast<u64 *>(in);
    const u64 *buf = in;
    if (n <= 1) return 0;
    SmjScratch sc(ctx);
    u64 *d_bits = (u64 *)sc.get(16);
    if (!d_bits) return smj_fail_msg("smj_lsd_sort: scratch allocation failed");
    const u64 init[2] = {0ull, ~0ull};
    SMJ_CHECK(cudaMemcpyAsync(d_bits, init, 16, cudaMemcpyHostToDevice, ctx->stream));
    SMJ_ITEMS(ctx, n);
    SMJ_LAUNCH(ctx, k_bits_reduce, hist_grid(ctx, n, 256, 16), 256, 0, buf, n, d_bits);
    u64 h_bits[2];
    SMJ_CHECK(cudaMemcpyAsync(h_bits, d_bits, 16, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    u64 varying = h_bits[0] & ~h_bits[1];
    if (key_only) varying &= 0xffffffff00000000ull;

    // plan: cover the varying bits, least significant first, with windows of <= 8 bits
    DigitPlan plan;
    plan.nd = 0;
    u64 v   = varying;
    while (v) {
        int lo   = __builtin_ctzll(v);
        int bits = std::min(8, 64 - lo);
        u64 win  = (v >> lo) & ((1ull << bits) - 1);
        bits     = 64 - __builtin_clzll(win);  // trim constant bits at the top of the window
        plan.d[plan.nd].sub   = 0;
        plan.d[plan.nd].shift = (u32)lo;
        plan.d[plan.nd].mask  = (u32)((1ull << bits) - 1);
        plan.nd++;
        v &= ~(((1ull << bits) - 1) << lo);
    }
    if (plan.nd == 0) return 0;  // all items equal (on the compared bits): already sorted

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
    SMJ_LAUNCH(ctx, k_hist_digits, hist_grid(ctx, n, 512, 16), 512, 0, buf, n, plan, d_hist);
    SMJ_LAUNCH(ctx, k_bases_multi, plan.nd, RADIX_NB, 0, d_hist, d_bases);

    const u64 *cur = in;
    for (int k = 0; k < plan.nd; k++) {
        u64 *dst = (k & 1) ? b : a;
        int  rc  = launch_onesweep(ctx, cur, dst, n, plan.d[k], d_bases + (size_t)k * RADIX_NB,
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
        static bool attr_set = false;
        if (!attr_set) {
            SMJ_CHECK(cudaFuncSetAttribute(k_hist_wide_smem,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            attr_set = true;
        }
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
    const bool two_pass = nbits > 8;
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
