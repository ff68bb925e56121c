// smj_mgpu.cu -- the m-way join over the G GPUs of one box: a "join group".
//
// Replaces the reference's cross-socket step.  There, the fan-out over workers happens INSIDE
// sortmergejoin_multiway(): sortmergejoin_initrun cuts R and S into NTHREADS chunks
// (src/joins/joincommon.c:105-173), every thread partitions and sorts its chunk, and thread t then
// PULLS partitions [t F/T, (t+1) F/T) out of every other thread's memory in NUMA-shuffle order
// (src/joins/sortmergejoin_multiway.c:492-517, src/util/numa_shuffle.c:56-85) before it merges and
// joins them.  Here a rank (one GPU; one process per GPU, or one host thread per GPU inside one
// process) owns a contiguous run of the 512 first-level key buckets, and the FIRST partitioning
// pass of the m-way join IS the exchange: the scatter kernel of every rank stores each item
// straight into the receive buffer of the GPU that owns its bucket -- peer stores over NVLink /
// NVSwitch, no send buffer, no pack kernel, no data collective -- and the owner finishes with the
// second pass and the fused bucket join on what arrived.
//
// Everything between the sample and the result is device-driven; the host of a rank synchronises
// twice per join, exactly like the single-GPU path (sample -> digit plan, result):
//   1. k_msd_sample on the local slices, host reads the LOCAL digit plan (sync 1);
//   2. k_xg_post_sample: the sampled first-level histograms, key-bit OR/AND, sizes and buffer
//      capacities go into a mailbox in EVERY rank's control block (stores over NVLink, then a
//      release flag per source rank): an all-gather without a collective library;
//   3. k_xg_layout (waits for the W flags): every rank computes the same layout from the same W
//      posts -- global digit plan, bucket owners (contiguous runs of equal estimated load), and
//      inside every owner's buffer one REGION per (bucket, source rank) sized from the source's
//      sample (estimate + 7 sigma, smj_cap1), so that the scatter needs no histogram pass and no
//      cross-GPU cursor: a source only bumps cursors it owns;
//   4. pass A of R, then of S: the optimistic scatter kernel of the single-GPU path
//      (k_onesweep_unstable<LimitArray>) with absolute peer addresses as bin cursors; after each
//      relation k_xg_post_counts tells every owner how many items each of its regions received;
//   5. the owner waits for the W count posts of R, builds its segment table (k_xg_segs) and runs
//      pass B over the segments; same for S; bucket join; k_xg_post_result / k_xg_reduce give every
//      rank the global {count, checksum, status} (sync 2).
// Whatever the sample could not promise -- a region overflowed, a key outside the sampled key
// range, ranks that disagree on the digit plan, a key domain the MSD plan does not cover, a bucket
// that is heavy on both sides, receive buffers too small for the layout -- shows up in the status
// word, which is the OR over all ranks and therefore the same everywhere; all ranks then repeat the
// join through the exact path: range partition by equal-width splitters with exact counts
// (smjdev_range_count / _scatter, counts exchanged through the same mailboxes) + the general
// single-GPU join on what arrived.  A result never depends on a guess.
#include "smj_internal.cuh"
#include "../../include/smj_device.h"

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

namespace {

constexpr int XG_MAXW  = 16;
constexpr int XG_NB    = 512;
constexpr int XG_POSTW = SMJ_MSD_SAMPLE_WORDS;  // a post = the sample block of k_msd_sample + the words below
// words of a post beyond what k_msd_sample fills ([0..3] OR/AND, [6..7] duplicates, [8..9] sampled items,
// [16..1040) the two histograms)
constexpr int PW_NR = 10, PW_NS = 11, PW_SHIFT1 = 12, PW_CAPR = 13, PW_CAPS = 14;
constexpr int XG_CNTW = XG_NB + 8;  // per-bin counts + [512] overflow flag + [513] key-range flag
constexpr u32 ST_PLAN = 1, ST_CAPACITY = 2, ST_OVERFLOW = 4, ST_KEYRANGE = 8, ST_HEAVY = 16, ST_TIMEOUT = 32;

struct XgCtl {  // one per rank, in memory every rank can write
    u64 flag[4][XG_MAXW];  // [phase][source]: epoch of the source's latest post of that phase
    u64 post[XG_MAXW][XG_POSTW];
    u64 cnt[2][XG_MAXW][XG_CNTW];
    u64 res[XG_MAXW][8];
};
struct XgPeers {  // how THIS rank reaches every rank (device resident)
    XgCtl *ctl[XG_MAXW];
    u64   *recv[2][XG_MAXW];
};
constexpr int XG_NSEG_MAX = XG_NB * XG_MAXW;
struct XgLayout {  // what k_xg_layout leaves on the device
    u32 status, abort;
    u64 gbits[4];                 // global sampled OR_R AND_R OR_S AND_S
    u64 n_recv[2];                // items this rank received (k_xg_segs)
    u32 own_range[2];             // second-level buckets this rank owns: [first, one past the last)
    u64 sent[2];                  // items this rank stored into OTHER ranks' buffers
    u32 owner[XG_NB];
    u64 tab[2][3][XG_NB];         // pass A: cursors, starts, limits (absolute item addresses)
    u64 seg[2][3 * XG_NSEG_MAX];  // pass B: {start[nseg], count[nseg], bucket[nseg]}, nseg = 512 W
    u64 out[8];                   // k_xg_reduce: count, checksum, status, sent bytes
};

__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 global_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// threads 0 .. world-1 wait until every source has posted `phase` of this epoch.  A peer that never
// posts (its process died) must not hang the GPU: after 4 s the wait gives up and raises *abort.
__device__ __forceinline__ void xg_wait(const XgCtl *ctl, int phase, u64 epoch, int world, u32 *abort) {
    if ((int)threadIdx.x < world) {
        const u64 *f  = &ctl->flag[phase][threadIdx.x];
        const u64  t0 = global_ns();
        while (ld_acquire_sys(f) < epoch) {
            if (global_ns() - t0 > 4000000000ull) {
                atomicOr(abort, 1u);
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
}

// CTA p copies nwords words to rank p's control block at byte offset dst_off, then raises the flag
__global__ void __launch_bounds__(256) k_xg_post(const XgPeers *__restrict__ peers, int me, const u64 *__restrict__ src,
                                                 u32 nwords, size_t dst_off, int phase, u64 epoch) {
    XgCtl *ctl = peers->ctl[blockIdx.x];
    u64   *dst = (u64 *)((char *)ctl + dst_off);
    for (u32 i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(&ctl->flag[phase][me], epoch);
}

__device__ __forceinline__ u64 block_excl_scan_512(u64 v, u64 *s_tmp, u64 *total) {
    const int b = threadIdx.x;
    s_tmp[b]    = v;
    __syncthreads();
    for (int o = 1; o < XG_NB; o <<= 1) {
        const u64 t = (b >= o) ? s_tmp[b - o] : 0;
        __syncthreads();
        s_tmp[b] += t;
        __syncthreads();
    }
    const u64 incl = s_tmp[b];
    *total         = s_tmp[XG_NB - 1];
    __syncthreads();
    return incl - v;
}

// One CTA of 512 threads, thread b = first-level bucket b.  See the file header, step 3.
// Launched with `world` CTAs: CTA p first publishes this rank's sample in rank p's mailbox (what
// k_xg_post_sample does); CTA 0 then goes on to compute the layout.  One launch instead of two.
__global__ void __launch_bounds__(XG_NB) k_xg_layout(const XgPeers *__restrict__ peers, XgCtl *ctl, int me, int world,
                                                     u64 epoch, u32 my_shift1, XgLayout *__restrict__ lay,
                                                     const u64 *__restrict__ st, u64 nr, u64 ns, u64 cap_r, u64 cap_s) {
    __shared__ u64 s_tmp[XG_NB], s_p[2][XG_NB + 1];
    __shared__ u32 s_first[XG_MAXW + 1], s_owner[XG_NB], s_status;
    const int b = threadIdx.x;
    {
        XgCtl *pc  = peers->ctl[blockIdx.x];
        u64   *dst = pc->post[me];
        for (u32 i = threadIdx.x; i < (u32)XG_POSTW; i += blockDim.x) {
            u64 v = st[i];
            if (i == PW_NR) v = nr;
            if (i == PW_NS) v = ns;
            if (i == PW_SHIFT1) v = my_shift1;
            if (i == PW_CAPR) v = cap_r;
            if (i == PW_CAPS) v = cap_s;
            dst[i] = v;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(&pc->flag[0][me], epoch);
        if (blockIdx.x != 0) return;
    }
    if (b == 0) s_status = 0;
    if (b <= XG_MAXW) s_first[b] = XG_NB;
    xg_wait(ctl, 0, epoch, world, &lay->abort);
    // ---- the global digit plan: every rank derived its own from its own sample; they must agree
    u64 gb[4] = {0ull, ~0ull, 0ull, ~0ull}, tot_n = 0;
    bool same = true;
    for (int g = 0; g < world; g++) {
        const u64 *p = ctl->post[g];
        gb[0] |= p[0];
        gb[1] &= p[1];
        gb[2] |= p[2];
        gb[3] &= p[3];
        tot_n += p[PW_NR] + p[PW_NS];
        same = same && p[PW_SHIFT1] == (u64)my_shift1;
    }
    MsdPlan pl;
    const bool ok = my_shift1 != 0 && same && smj_msd_plan(gb, &pl) && pl.shift1 == my_shift1 && tot_n >= (1u << 20) &&
                    lay->abort == 0;
    // ---- per bucket: estimated load, and the capacity of every source's region
    u64 cap_me[2] = {0, 0}, cap_before[2] = {0, 0}, rowcap[2] = {0, 0}, load = 0;
    for (int g = 0; g < world; g++) {
        const u64 *p = ctl->post[g];
#pragma unroll
        for (int rel = 0; rel < 2; rel++) {
            const u64    n = p[PW_NR + rel], nsamp = p[8 + rel], k = p[16 + rel * XG_NB + b];
            const double scale = nsamp ? fmax(1.0, (double)n / (double)nsamp) : 1.0;
            const u64    cap   = n ? smj_cap1(k, scale, 0) : 0ull;
            if (g < me) cap_before[rel] += cap;
            if (g == me) cap_me[rel] = cap;
            rowcap[rel] += cap;
            load += (u64)((double)k * scale);
        }
    }
    u64 total_load, tot_r, tot_s;
    const u64 load_pre = block_excl_scan_512(load, s_tmp, &total_load);
    const u64 pr       = block_excl_scan_512(rowcap[0], s_tmp, &tot_r);
    const u64 ps       = block_excl_scan_512(rowcap[1], s_tmp, &tot_s);
    s_p[0][b] = pr;
    s_p[1][b] = ps;
    if (b == 0) {
        s_p[0][XG_NB] = tot_r;
        s_p[1][XG_NB] = tot_s;
    }
    // owners: contiguous runs of (about) equal estimated load; the bucket's midpoint decides
    u32 own = 0;
    if (total_load) own = (u32)min((u64)(world - 1), ((2 * load_pre + load) * (u64)world) / (2 * total_load));
    s_owner[b] = own;
    __syncthreads();
    {
        const u32 prev = b ? s_owner[b - 1] : 0xffffffffu;  // owners never decrease
        if (b == 0)
            for (u32 p = 0; p <= own; p++) s_first[p] = 0;
        else
            for (u32 p = prev + 1; p <= own; p++) s_first[p] = b;
    }
    __syncthreads();
    // ---- does every owner's layout fit the buffer it announced?
    if (b < world) {
        const u32 f0 = s_first[b], f1 = (b + 1 < world) ? s_first[b + 1] : XG_NB;
        const u64 *p = ctl->post[b];
        if (s_p[0][f1] - s_p[0][f0] > p[PW_CAPR] || s_p[1][f1] - s_p[1][f0] > p[PW_CAPS]) atomicOr(&s_status, ST_CAPACITY);
    }
    if (b == 0 && !ok) atomicOr(&s_status, lay->abort ? ST_TIMEOUT : ST_PLAN);
    __syncthreads();
    const bool go = s_status == 0;
    // ---- tables of pass A (this rank as a source) and segments of pass B (this rank as an owner)
    const u32 nseg = XG_NB * world;
#pragma unroll
    for (int rel = 0; rel < 2; rel++) {
        const u64 base  = s_p[rel][s_first[own]];
        const u64 mine  = s_p[rel][b] - base + cap_before[rel];
        const u64 abs0  = ((u64)(uintptr_t)peers->recv[rel][own] >> 3) + mine;
        lay->tab[rel][0][b] = go ? abs0 : 0ull;
        lay->tab[rel][1][b] = go ? abs0 : 0ull;
        lay->tab[rel][2][b] = go ? abs0 + cap_me[rel] : 0ull;  // limit 0: the scatter stores nothing
        u64 off = s_p[rel][b] - base;
        for (int g = 0; g < world; g++) {
            const u64 *p = ctl->post[g];
            const u64  n = p[PW_NR + rel], nsamp = p[8 + rel], k = p[16 + rel * XG_NB + b];
            const double scale = nsamp ? fmax(1.0, (double)n / (double)nsamp) : 1.0;
            const u64    cap   = n ? smj_cap1(k, scale, 0) : 0ull;
            const u32    i     = (u32)b * world + g;
            lay->seg[rel][i]            = off;
            lay->seg[rel][nseg + i]     = 0;  // filled by k_xg_segs
            lay->seg[rel][2 * nseg + i] = (u64)b;
            off += cap;
        }
    }
    lay->owner[b] = own;
    if (b == 0) {
        lay->status = s_status;
        for (int q = 0; q < 4; q++) lay->gbits[q] = gb[q];
        lay->n_recv[0] = lay->n_recv[1] = lay->sent[0] = lay->sent[1] = 0;
        lay->own_range[0] = s_first[me] * (u32)XG_NB;
        lay->own_range[1] = ((me + 1 < world) ? s_first[me + 1] : (u32)XG_NB) * (u32)XG_NB;
    }
}

// after pass A of `rel`: tell every owner how many items each of its regions received from this rank
__global__ void __launch_bounds__(XG_NB) k_xg_post_counts(const XgPeers *__restrict__ peers, int me, int rel,
                                                          XgLayout *__restrict__ lay, const u32 *__restrict__ over,
                                                          const u32 *__restrict__ keyflag, u64 epoch) {
    __shared__ u64 s_w[XG_NB / 32];
    XgCtl    *ctl = peers->ctl[blockIdx.x];
    const int b   = threadIdx.x;
    const u64 cur = lay->tab[rel][0][b], start = lay->tab[rel][1][b], lim = lay->tab[rel][2][b];
    const u64 c   = min(cur, lim) - start;
    ctl->cnt[rel][me][b] = c;
    if (b == 0) {
        ctl->cnt[rel][me][XG_NB]     = *over;
        ctl->cnt[rel][me][XG_NB + 1] = *keyflag;
    }
    if (blockIdx.x == 0) {  // statistics: items that left this GPU
        u64 v = (lay->owner[b] != (u32)me) ? c : 0ull;
        v     = warp_reduce_add(v);
        if ((b & 31) == 0) s_w[b >> 5] = v;
        __syncthreads();
        if (b == 0) {
            u64 t = 0;
            for (int w = 0; w < XG_NB / 32; w++) t += s_w[w];
            lay->sent[rel] = t;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (b == 0) st_release_sys(&ctl->flag[1 + rel][me], epoch);
}

// owner side: wait for the W count posts of `rel`, fill in the segment counts
__global__ void __launch_bounds__(XG_NB) k_xg_segs(XgCtl *ctl, int me, int world, int rel, u64 epoch,
                                                   XgLayout *__restrict__ lay) {
    __shared__ u64 s_w[XG_NB / 32];
    const int b = threadIdx.x;
    xg_wait(ctl, 1 + rel, epoch, world, &lay->abort);
    const u32 nseg = XG_NB * world;
    u64       tot  = 0;
    const bool mine = lay->owner[b] == (u32)me && lay->status == 0 && lay->abort == 0;
    for (int g = 0; g < world; g++) {
        const u64 c = mine ? ctl->cnt[rel][g][b] : 0ull;
        lay->seg[rel][nseg + (u32)b * world + g] = c;
        tot += c;
    }
    if (b < world) {
        if (ctl->cnt[rel][b][XG_NB]) atomicOr(&lay->status, ST_OVERFLOW);
        if (ctl->cnt[rel][b][XG_NB + 1]) atomicOr(&lay->status, ST_KEYRANGE);
    }
    tot = warp_reduce_add(tot);
    if ((b & 31) == 0) s_w[b >> 5] = tot;
    __syncthreads();
    if (b == 0) {
        u64 t = 0;
        for (int w = 0; w < XG_NB / 32; w++) t += s_w[w];
        lay->n_recv[rel] = t;
        if (lay->abort) atomicOr(&lay->status, ST_TIMEOUT);
    }
}

// this rank's partial result + status to everybody
__device__ __forceinline__ void xg_reduce(XgCtl *ctl, int me, int world, u64 epoch, u32 *abort, u64 *out);
// `world` CTAs: CTA p posts to rank p; with reduce_out != null CTA 0 then waits for everybody's
// result and reduces (one launch instead of two)
__global__ void __launch_bounds__(32) k_xg_post_result(const XgPeers *__restrict__ peers, int me, const u64 *__restrict__ acc,
                                                       const u64 *__restrict__ d_state, u32 lbits,
                                                       const XgLayout *__restrict__ lay, u64 epoch, XgCtl *my_ctl,
                                                       int world, u32 *abort, u64 *reduce_out) {
    XgCtl *ctl = peers->ctl[blockIdx.x];
    if (threadIdx.x == 0) {
        u32 status = lay ? lay->status : 0u;
        if (d_state) {  // heavy buckets the fused join does not cover (smj_bucket_join_valid)
            const u32 nbig = (u32)(d_state[6] & 0xffffffffu), unsupported = (u32)(d_state[6] >> 32);
            if (lbits > 10 ? nbig != 0 : unsupported != 0) status |= ST_HEAVY;
            if (d_state[10] | d_state[11]) status |= ST_OVERFLOW;  // a second-level slot overflowed
        }
        ctl->res[me][0] = acc[0];
        ctl->res[me][1] = acc[1];
        ctl->res[me][2] = status;
        ctl->res[me][3] = lay ? 8 * (lay->sent[0] + lay->sent[1]) : 0ull;
        __threadfence_system();
        st_release_sys(&ctl->flag[3][me], epoch);
    }
    if (reduce_out && blockIdx.x == 0) xg_reduce(my_ctl, me, world, epoch, abort, reduce_out);
}
__device__ __forceinline__ void xg_reduce(XgCtl *ctl, int me, int world, u64 epoch, u32 *abort, u64 *out) {
    xg_wait(ctl, 3, epoch, world, abort);
    if (threadIdx.x == 0) {
        u64 c = 0, x = 0, st = *abort ? ST_TIMEOUT : 0;
        for (int g = 0; g < world; g++) {
            c += ctl->res[g][0];
            x += ctl->res[g][1];
            st |= ctl->res[g][2];
        }
        out[0] = c;
        out[1] = x;
        out[2] = st;
        out[3] = ctl->res[me][3];
    }
}
// fallback: gather the first nwords of every rank's post into a dense matrix (after the wait)
__global__ void __launch_bounds__(256) k_xg_gather_posts(XgCtl *ctl, int world, u32 nwords, u64 epoch, u32 *abort,
                                                         u64 *out) {
    xg_wait(ctl, 0, epoch, world, abort);
    for (u32 i = threadIdx.x; i < nwords * (u32)world; i += blockDim.x) out[i] = ctl->post[i / nwords][i % nwords];
}
__global__ void __launch_bounds__(32) k_xg_barrier(XgCtl *ctl, int phase, int world, u64 epoch, u32 *abort) {
    xg_wait(ctl, phase, epoch, world, abort);
}

// ---------------------------------------------------------------------------------------
struct XgHandle {  // what a rank publishes (fits SMJ_GROUP_HANDLE_BYTES)
    cudaIpcMemHandle_t ipc[3];  // recv R, recv S, control block
    u64                ptr[3];  // the same as plain pointers (valid inside the owning process)
    int                pid, device, rank, world;
};
static_assert(sizeof(XgHandle) <= 256, "handle size");

struct XgGroup {
    SmjCtx  *ctx   = nullptr;
    int      rank  = 0, world = 1;
    u64      cap[2] = {0, 0};
    u64     *recv[2] = {nullptr, nullptr};
    XgCtl   *ctl   = nullptr;
    XgPeers  peers;            // host copy
    XgPeers *d_peers = nullptr;
    XgLayout *d_lay  = nullptr;
    void    *opened[3 * XG_MAXW];
    int      nopened = 0;
    u64      epoch   = 0;
    bool     connected = false;
};

int xg_fallback(XgGroup *g, const u64 *r, u64 nr, const u64 *s, u64 ns, const u64 *gbits, smj_join_stats_t *stats,
                u64 *sent_bytes);

}  // namespace

extern "C" {

int smjdev_group_create(int rank, int world, uint64_t cap_r, uint64_t cap_s, void **group) {
    SmjCtx *ctx = smj_ctx();
    if (!ctx) return -1;
    if (world < 1 || world > XG_MAXW || rank < 0 || rank >= world || !group)
        return smj_fail_msg("smjdev_group_create: bad rank / world (at most 16 ranks)");
    XgGroup *g = new XgGroup();
    g->ctx   = ctx;
    g->rank  = rank;
    g->world = world;
    g->cap[0] = std::max<u64>(cap_r, 16);
    g->cap[1] = std::max<u64>(cap_s, 16);
    memset(&g->peers, 0, sizeof(g->peers));
    // plain cudaMalloc: memory of the stream-ordered pool cannot be exported to other processes
    cudaError_t e = cudaMalloc((void **)&g->recv[0], g->cap[0] * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->recv[1], g->cap[1] * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->ctl, sizeof(XgCtl));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->d_peers, sizeof(XgPeers));
    if (e == cudaSuccess) e = cudaMalloc((void **)&g->d_lay, sizeof(XgLayout));
    if (e == cudaSuccess) e = cudaMemset(g->ctl, 0, sizeof(XgCtl));
    if (e == cudaSuccess) e = cudaMemset(g->d_lay, 0, sizeof(XgLayout));
    if (e != cudaSuccess) {
        smjdev_group_destroy(g);
        return smj_fail(e, "smjdev_group_create", __FILE__, __LINE__);
    }
    *group = g;
    return 0;
}

int smjdev_group_export(void *group, unsigned char *handle) {
    XgGroup *g = (XgGroup *)group;
    if (!g || !handle) return smj_fail_msg("smjdev_group_export: null argument");
    XgHandle h;
    memset(&h, 0, sizeof(h));
    void *ptrs[3] = {g->recv[0], g->recv[1], g->ctl};
    for (int i = 0; i < 3; i++) {
        SMJ_CHECK(cudaIpcGetMemHandle(&h.ipc[i], ptrs[i]));
        h.ptr[i] = (u64)(uintptr_t)ptrs[i];
    }
    h.pid    = (int)getpid();
    h.device = g->ctx->device;
    h.rank   = g->rank;
    h.world  = g->world;
    memset(handle, 0, SMJ_GROUP_HANDLE_BYTES);
    memcpy(handle, &h, sizeof(h));
    return 0;
}

// handles: world x SMJ_GROUP_HANDLE_BYTES, in rank order.  Ranks of the same process are reached
// through peer access (cudaDeviceEnablePeerAccess), ranks of other processes through CUDA IPC.
// The caller makes sure every rank has been created before anybody connects (a barrier).
int smjdev_group_connect(void *group, const unsigned char *handles) {
    XgGroup *g = (XgGroup *)group;
    if (!g || !handles) return smj_fail_msg("smjdev_group_connect: null argument");
    if (!smj_ctx()) return -1;
    for (int p = 0; p < g->world; p++) {
        XgHandle h;
        memcpy(&h, handles + (size_t)p * SMJ_GROUP_HANDLE_BYTES, sizeof(h));
        if (h.rank != p || h.world != g->world) return smj_fail_msg("smjdev_group_connect: handles out of order");
        void *ptr[3];
        if (p == g->rank) {
            ptr[0] = g->recv[0];
            ptr[1] = g->recv[1];
            ptr[2] = g->ctl;
        } else if (h.pid == (int)getpid()) {
            if (h.device != g->ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return smj_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
                cudaGetLastError();
            }
            for (int i = 0; i < 3; i++) ptr[i] = (void *)(uintptr_t)h.ptr[i];
        } else {
            for (int i = 0; i < 3; i++) {
                SMJ_CHECK(cudaIpcOpenMemHandle(&ptr[i], h.ipc[i], cudaIpcMemLazyEnablePeerAccess));
                g->opened[g->nopened++] = ptr[i];
            }
        }
        g->peers.recv[0][p] = (u64 *)ptr[0];
        g->peers.recv[1][p] = (u64 *)ptr[1];
        g->peers.ctl[p]     = (XgCtl *)ptr[2];
    }
    SMJ_CHECK(cudaMemcpy(g->d_peers, &g->peers, sizeof(XgPeers), cudaMemcpyHostToDevice));
    g->connected = true;
    return 0;
}

int smjdev_group_destroy(void *group) {
    XgGroup *g = (XgGroup *)group;
    if (!g) return 0;
    if (g->ctx) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cur != g->ctx->device) cudaSetDevice(g->ctx->device);
        cudaDeviceSynchronize();
    }
    for (int i = 0; i < g->nopened; i++) cudaIpcCloseMemHandle(g->opened[i]);
    cudaFree(g->recv[0]);
    cudaFree(g->recv[1]);
    cudaFree(g->ctl);
    cudaFree(g->d_peers);
    cudaFree(g->d_lay);
    cudaGetLastError();
    delete g;
    return 0;
}

// One m-way join of the group: every rank calls it with its own slices of R and S (device
// pointers); all ranks return the GLOBAL {matches, checksum}.  stats->path: SMJ_PATH_MSD_OPT when
// the fused exchange produced the result, SMJ_PATH_GENERAL when the exact repeat did (stats->repeats
// = 1).  *bytes_sent (may be NULL): item bytes this rank stored into other ranks' buffers.
int smjdev_group_join(void *group, const uint64_t *r_, uint64_t nr, const uint64_t *s_, uint64_t ns,
                      smj_join_stats_t *stats, uint64_t *bytes_sent) {
    XgGroup *g = (XgGroup *)group;
    if (!g || !g->connected || !stats) return smj_fail_msg("smjdev_group_join: group not connected");
    SmjCtx *ctx = smj_ctx();
    if (!ctx || ctx != g->ctx) return smj_fail_msg("smjdev_group_join: called from a thread bound to another device");
    const u64 *r = (const u64 *)r_, *s = (const u64 *)s_;
    memset(stats, 0, sizeof(*stats));
    const u64 launches0 = ctx->launches;
    const int W = g->world, me = g->rank;
    const u64 epoch = ++g->epoch;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    struct EvGuard {
        cudaEvent_t &a, &b;
        ~EvGuard() {
            if (a) cudaEventDestroy(a);
            if (b) cudaEventDestroy(b);
        }
    } evg{ev0, ev1};
    SMJ_CHECK(cudaEventCreate(&ev0));
    SMJ_CHECK(cudaEventCreate(&ev1));
    SMJ_CHECK(cudaEventRecord(ev0, ctx->stream));

    const u32  nb2 = 512u * 512u;
    SmjScratch sc(ctx);
    u64 *d_st = (u64 *)sc.get(smj_msd_sample_bytes(0));
    u32 *d_sub = (u32 *)sc.get(smj_msd_sample_bytes(1));
    u64 *d_state = (u64 *)sc.get(128), *acc = (u64 *)sc.get(16);
    u32 *d_list = (u32 *)sc.get((size_t)nb2 * 4);
    // (second-level output: sized for the optimistic slots, smj_msd_alloc2 >= capacity)
    const u64 alloc2[2] = {smj_msd_alloc2(g->cap[0]), smj_msd_alloc2(g->cap[1])};
    u64 *bR = (u64 *)sc.get(alloc2[0] * 8), *bS = (u64 *)sc.get(alloc2[1] * 8);
    u64 *offR = (u64 *)sc.get((size_t)(nb2 + 1) * 8), *offS = (u64 *)sc.get((size_t)(nb2 + 1) * 8);
    u64 *endR = (u64 *)sc.get((size_t)nb2 * 8), *endS = (u64 *)sc.get((size_t)nb2 * 8);
    if (!d_st || !d_sub || !d_state || !acc || !d_list || !bR || !bS || !offR || !offS || !endR || !endS)
        return smj_fail_msg("smjdev_group_join: out of device memory");
    SMJ_CHECK(cudaMemcpyAsync(d_state, MSD_STATE_INIT, 128, cudaMemcpyHostToDevice, ctx->stream));
    SMJ_CHECK(cudaMemsetAsync(acc, 0, 16, ctx->stream));
    // ---- 1. sample of the local slices -> local digit plan (host synchronisation 1 of 2)
    std::vector<u64> samp(SMJ_MSD_SAMPLE_WORDS, 0ull);
    MsdPlan pl;
    memset(&pl, 0, sizeof(pl));
    bool have_plan = false;
    int  rc        = 0;
    if (nr && ns) {
        rc = smj_msd_sample(ctx, r, nr, s, ns, 1u << 18, d_st, d_sub, samp.data());
        if (rc < 0) return rc;
        have_plan = rc == 0 && smj_msd_plan(samp.data(), &pl);
    }
    if (!have_plan) SMJ_CHECK(cudaMemsetAsync(d_st, 0, smj_msd_sample_bytes(0), ctx->stream));
    const u32 shift1 = have_plan ? pl.shift1 : 0u;
    const u32 krange = have_plan ? ((pl.hi == 31 ? 0xffffffffu : ((1u << (pl.hi + 1)) - 1)) & ~((1u << pl.lo) - 1)) : 0u;
    const u32 kmask = ~krange, kref = (u32)((samp[1] & samp[3]) >> 32) & kmask;
    u32 *flags = (u32 *)(d_state + 8);  // [0] R overflow, [2] S overflow, [12] R key, [14] S key (one u64 slot each)
    // ---- 2. + 3. all-gather of the samples through the mailboxes, layout
    SMJ_LAUNCH(ctx, k_xg_layout, W, XG_NB, 0, g->d_peers, g->ctl, me, W, epoch, shift1, g->d_lay, (const u64 *)d_st, nr, ns,
               g->cap[0], g->cap[1]);
    // ---- 4. pass A == the exchange, R then S (main stream); 5. owner side: pass B over the received
    // segments.  Pass B of R only needs every rank's R scatter: it runs on the second compute stream
    // while pass A of S is still pushing S over NVLink (SMJ_XG_OVERLAP=0: everything on one stream).
    const u64 *in[2] = {r, s};
    const u64  nn[2] = {nr, ns};
    const u32  nseg = (u32)(XG_NB * W);
    u64       *recv[2] = {g->recv[0], g->recv[1]}, *bb[2] = {bR, bS}, *off[2] = {offR, offS};
    const u32  shift2 = have_plan ? pl.shift2 : 32u;
    const char *e_tile = getenv("SMJ_XG_TILE"), *e_ovl = getenv("SMJ_XG_OVERLAP");
    const int  big_tile = e_tile ? (atoi(e_tile) >= 16384 ? 1 : atoi(e_tile) == 12288 ? 2 : 0) : (W > 1);
    const bool overlap  = (e_ovl ? atoi(e_ovl) != 0 : (W > 1)) && !ctx->profiling;
    // the exchange of S is NVLink-bound: it may leave part of the SMs to the owner side of R
    const char *e_grid = getenv("SMJ_XG_GRID");
    const int  xg_grid = e_grid ? atoi(e_grid) : ctx->sm_count * 5 / 8;  // measured at N = 2: 148 / 110 / 90 / 74 CTAs -> 2.19 / 2.01 / 1.98 / 2.05 ms
    cudaEvent_t ev_lay = smj_pool_event(ctx, 0), ev_aux = smj_pool_event(ctx, 1);
    if (!ev_lay || !ev_aux) return smj_fail_msg("smjdev_group_join: cudaEventCreate failed");
    cudaStream_t main_stream = ctx->stream;
    // Like the single-GPU path, the second level runs without a histogram (equal slots per bucket)
    // only for relations whose sample shows no duplicate keys; a slot that overflows raises the
    // status (level-2 flag of d_state) and the join is repeated exactly.  SMJ_XG_EXACT_B=1: always
    // histogram + scatter.
    u64       *ends[2] = {endR, endS};
    const bool no_opt2 = getenv("SMJ_XG_EXACT_B") != nullptr;
    const bool opt2[2] = {have_plan && !no_opt2 && samp[6] <= 16, have_plan && !no_opt2 && samp[7] <= 16};
    auto owner_side = [&](int rel) -> int {
        SMJ_LAUNCH(ctx, k_xg_segs, 1, XG_NB, 0, g->ctl, me, W, rel, epoch, g->d_lay);
        if (opt2[rel])
            return smj_msd_pass_b_opt_seg(ctx, recv[rel], g->cap[rel], (u32)W, g->d_lay->seg[rel], bb[rel], alloc2[rel],
                                          shift2, 1u << pl.lbits, off[rel], ends[rel], flags + 4 + 2 * rel);
        return smj_msd_pass_b_seg_dev(ctx, recv[rel], g->cap[rel], nseg, g->d_lay->seg[rel], &g->d_lay->n_recv[rel],
                                      bb[rel], shift2, off[rel]);
    };
    if (overlap) SMJ_CHECK(cudaEventRecord(ev_lay, main_stream));
    for (int rel = 0; rel < 2; rel++) {
        if (have_plan && nn[rel] &&
            (rc = smj_msd_pass_a_abs(ctx, in[rel], nn[rel], shift1, &g->d_lay->tab[rel][0][0], flags + 2 * rel, kmask, kref,
                                     flags + 12 + 2 * rel, big_tile, (rel == 1 && overlap) ? xg_grid : 0)))
            return rc;
        SMJ_LAUNCH(ctx, k_xg_post_counts, W, XG_NB, 0, g->d_peers, me, rel, g->d_lay, flags + 2 * rel,
                   flags + 12 + 2 * rel, epoch);
        if (rel == 0 && overlap) {
            // (scratch of this call was allocated on the main stream before ev_lay)
            SMJ_CHECK(cudaStreamWaitEvent(ctx->aux_stream, ev_lay, 0));
            ctx->stream = ctx->aux_stream;
            rc          = owner_side(0);
            cudaError_t e = cudaEventRecord(ev_aux, ctx->aux_stream);
            ctx->stream = main_stream;
            if (rc) return rc;
            SMJ_CHECK(e);
        }
    }
    if (overlap)
        SMJ_CHECK(cudaStreamWaitEvent(main_stream, ev_aux, 0));
    else if ((rc = owner_side(0)))
        return rc;
    if ((rc = owner_side(1))) return rc;
    const u32 lo_shift = have_plan ? pl.lo_shift : 32u, lbits = have_plan ? pl.lbits : 1u;
    const BucketSet R = {bR, offR, opt2[0] ? endR : offR + 1, g->cap[0]}, S = {bS, offS, opt2[1] ? endS : offS + 1, g->cap[1]};
    if ((rc = smj_bucket_collect(ctx, R, S, d_state, d_list))) return rc;
    if ((rc = smj_bucket_join(ctx, R, S, lo_shift, lbits, d_state, d_list, acc, g->d_lay->own_range))) return rc;
    SMJ_LAUNCH(ctx, k_xg_post_result, W, 32, 0, g->d_peers, me, acc, d_state, lbits, g->d_lay, epoch, g->ctl, W,
               &g->d_lay->abort, g->d_lay->out);
    u64 h[12];
    SMJ_CHECK(cudaMemcpyAsync(h, g->d_lay->out, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaMemcpyAsync(h + 4, g->d_lay->gbits, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaEventRecord(ev1, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));  // host synchronisation 2 of 2
    if (h[2] & ST_TIMEOUT) return smj_fail_msg("smjdev_group_join: a rank of the group did not answer within 4 s");
    if (bytes_sent) *bytes_sent = h[3];
    if (h[2] != 0 && getenv("SMJ_VERBOSE"))
        fprintf(stderr, "[smj_b200] group join, rank %d: status 0x%llx (1 plan, 2 capacity, 4 region overflow, 8 key range, "
                "16 heavy bucket) -> exact repeat\n", me, (unsigned long long)h[2]);
    if (h[2] != 0 || getenv("SMJ_GROUP_EXACT")) {
        // what the sample could not promise: all ranks saw the same status and repeat exactly
        sc.release();
        u64 sent = 0;
        rc = xg_fallback(g, r, nr, s, ns, h + 4, stats, &sent);
        stats->repeats  = 1;
        stats->launches = ctx->launches - launches0;
        if (bytes_sent) *bytes_sent = sent;
        return rc;
    }
    stats->matches  = h[0];
    stats->checksum = h[1];
    cudaEventElapsedTime(&stats->ms_total, ev0, ev1);
    stats->launches  = ctx->launches - launches0;
    stats->hbm_bytes = sc.bytes;
    stats->path      = SMJ_PATH_MSD_OPT;
    return 0;
}

}  // extern "C"

namespace {
// The exact repeat: range partition by equal-width splitters over the sampled key range with exact
// counts, the counts all-gathered through the mailboxes, scatter into the owners' buffers, general
// single-GPU join on what arrived.  Fails (on every rank alike) when a receive buffer is too small.
int xg_fallback(XgGroup *g, const u64 *r, u64 nr, const u64 *s, u64 ns, const u64 *gbits, smj_join_stats_t *stats,
                u64 *sent_bytes) {
    SmjCtx   *ctx = g->ctx;
    const int W = g->world, me = g->rank;
    u64 epoch = ++g->epoch;
    const u32 vor  = (u32)((gbits[0] | gbits[2]) >> 32);
    u32       bits = 1;
    while (bits < 32 && (vor >> bits)) bits++;
    const u64 kmax = (1ull << bits) - 1;
    uint32_t  sp[XG_MAXW];
    for (int i = 1; i < W; i++) sp[i - 1] = (uint32_t)((i * kmax) / W + 1);
    SmjScratch sc(ctx);
    u64 *d_cnt = (u64 *)sc.get(2 * XG_MAXW * 8), *d_mat = (u64 *)sc.get((size_t)W * 2 * XG_MAXW * 8);
    u64 *d_res = (u64 *)sc.get(64);
    if (!d_cnt || !d_mat || !d_res) return smj_fail_msg("smjdev_group_join: out of device memory");
    SMJ_CHECK(cudaMemsetAsync(d_cnt, 0, 2 * XG_MAXW * 8, ctx->stream));
    int rc;
    if ((rc = smjdev_range_count((const uint64_t *)r, nr, sp, (uint32_t)W, (uint64_t *)d_cnt))) return rc;
    if ((rc = smjdev_range_count((const uint64_t *)s, ns, sp, (uint32_t)W, (uint64_t *)(d_cnt + XG_MAXW)))) return rc;
    SMJ_LAUNCH(ctx, k_xg_post, W, 256, 0, g->d_peers, me, d_cnt, 2 * XG_MAXW, offsetof(XgCtl, post) + (size_t)me * XG_POSTW * 8,
               0, epoch);
    SMJ_LAUNCH(ctx, k_xg_gather_posts, 1, 256, 0, g->ctl, W, 2 * XG_MAXW, epoch, &g->d_lay->abort, d_mat);
    std::vector<u64> mat((size_t)W * 2 * XG_MAXW);
    u32 h_abort = 0;
    SMJ_CHECK(cudaMemcpyAsync(mat.data(), d_mat, mat.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaMemcpyAsync(&h_abort, &g->d_lay->abort, 4, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    if (h_abort) return smj_fail_msg("smjdev_group_join: a rank of the group did not answer within 4 s");
    auto cnt = [&](int src, int rel, int dst) { return mat[((size_t)src * 2 + rel) * XG_MAXW + dst]; };
    // every rank checks every receiver: the verdict is the same everywhere, nobody scatters on an error
    for (int dst = 0; dst < W; dst++)
        for (int rel = 0; rel < 2; rel++) {
            u64 tot = 0;
            for (int src = 0; src < W; src++) tot += cnt(src, rel, dst);
            // capacities were posted in the first round; ranks of one group are created alike, so the
            // local capacity stands for all of them
            if (tot > g->cap[rel]) {
                char msg[200];
                snprintf(msg, sizeof(msg), "smjdev_group_join: rank %d would receive %llu items of relation %c, its buffer "
                         "holds %llu (skewed keys: create the group with larger capacities)", dst,
                         (unsigned long long)tot, rel ? 'S' : 'R', (unsigned long long)g->cap[rel]);
                return smj_fail_msg(msg);
            }
        }
    uint64_t *dptr[2][XG_MAXW];
    uint64_t  doff[2][XG_MAXW], nrecv[2] = {0, 0}, sent = 0;
    for (int rel = 0; rel < 2; rel++)
        for (int dst = 0; dst < W; dst++) {
            dptr[rel][dst] = (uint64_t *)g->peers.recv[rel][dst];
            u64 o = 0;
            for (int src = 0; src < me; src++) o += cnt(src, rel, dst);
            doff[rel][dst] = o;
            if (dst != me) sent += 8 * cnt(me, rel, dst);
        }
    for (int rel = 0; rel < 2; rel++)
        for (int src = 0; src < W; src++) nrecv[rel] += cnt(src, rel, me);
    if ((rc = smjdev_range_scatter((const uint64_t *)r, nr, sp, (uint32_t)W, dptr[0], doff[0]))) return rc;
    if ((rc = smjdev_range_scatter((const uint64_t *)s, ns, sp, (uint32_t)W, dptr[1], doff[1]))) return rc;
    SMJ_LAUNCH(ctx, k_xg_post, W, 256, 0, g->d_peers, me, d_cnt, 0u, offsetof(XgCtl, res), 1, epoch);
    SMJ_LAUNCH(ctx, k_xg_barrier, 1, 32, 0, g->ctl, 1, W, epoch, &g->d_lay->abort);
    smj_join_stats_t st;
    if ((rc = smjdev_join(SMJ_ALGO_MWAY, (const uint64_t *)g->recv[0], nrecv[0], (const uint64_t *)g->recv[1], nrecv[1], 1, &st)))
        return rc;
    const u64 part[2] = {st.matches, st.checksum};
    SMJ_CHECK(cudaMemcpyAsync(d_res, part, 16, cudaMemcpyHostToDevice, ctx->stream));
    SMJ_LAUNCH(ctx, k_xg_post_result, W, 32, 0, g->d_peers, me, d_res, (const u64 *)nullptr, 0u, (const XgLayout *)nullptr,
               epoch, g->ctl, W, &g->d_lay->abort, g->d_lay->out);
    u64 h[4];
    SMJ_CHECK(cudaMemcpyAsync(h, g->d_lay->out, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SMJ_CHECK(cudaStreamSynchronize(ctx->stream));
    if (h[2] & ST_TIMEOUT) return smj_fail_msg("smjdev_group_join: a rank of the group did not answer within 4 s");
    *stats           = st;
    stats->matches   = h[0];
    stats->checksum  = h[1];
    stats->path = SMJ_PATH_GENERAL;
    *sent_bytes = sent;
    return 0;
}
}  // namespace

// ---------------------------------------------------------------------------------------
// One process, G GPUs: what sortmergejoin_multiway() does when SMJ_NGPUS > 1.  Like the reference's
// sortmergejoin_initrun (src/joins/joincommon.c:29-212) it cuts R and S into G contiguous chunks
// (n / G each, the last one takes the remainder, :105-139), starts one host thread per GPU, joins
// the threads and adds up -- except that a "thread" drives a B200 and the chunks meet over NVLink.
// The group (buffers, peer mappings) is kept between calls.
// ---------------------------------------------------------------------------------------
namespace {
struct MultiState {
    std::mutex           mu;
    int                  ngpus = 0;
    u64                  cap[2] = {0, 0};
    std::vector<void *>  groups;
};
MultiState g_multi;

void multi_destroy_locked() {
    for (size_t i = 0; i < g_multi.groups.size(); i++)
        if (g_multi.groups[i]) {
            std::thread t([i] {
                smjdev_init((int)i);
                smjdev_group_destroy(g_multi.groups[i]);
            });
            t.join();
        }
    g_multi.groups.clear();
    g_multi.ngpus = 0;
}
}  // namespace

extern "C" int smjdev_join_multi(int ngpus, const uint64_t *r_, uint64_t nr, const uint64_t *s_, uint64_t ns,
                                 smj_join_stats_t *stats) {
    const u64 *r = (const u64 *)r_, *s = (const u64 *)s_;
    if (!stats) return smj_fail_msg("smjdev_join_multi: stats must not be NULL");
    const int ndev = smjdev_device_count();
    if (ngpus < 1 || ngpus > XG_MAXW) return smj_fail_msg("smjdev_join_multi: 1..16 GPUs");
    if (ngpus > ndev) return smj_fail_msg("smjdev_join_multi: more GPUs requested than the box has");
    std::lock_guard<std::mutex> lk(g_multi.mu);
    const int W = ngpus;
    // receive capacity per GPU: the whole relation when that is cheap (no key distribution can
    // overflow it), else the even share plus room for imbalance
    u64 want[2];
    const u64 nn[2] = {nr, ns};
    for (int rel = 0; rel < 2; rel++)
        want[rel] = nn[rel] * 8 <= (4ull << 30) ? nn[rel] + 65536 : (nn[rel] / W) * 3 / 2 + (1u << 20);
    std::vector<int> rcs(W, 0);
    std::vector<std::string> errs(W);
    auto run_all = [&](auto fn) {
        std::vector<std::thread> th;
        for (int gidx = 0; gidx < W; gidx++)
            th.emplace_back([&, gidx] {
                rcs[gidx] = smjdev_init(gidx);
                if (!rcs[gidx]) rcs[gidx] = fn(gidx);
                if (rcs[gidx]) errs[gidx] = smjdev_last_error();
            });
        for (auto &t : th) t.join();
        for (int gidx = 0; gidx < W; gidx++)
            if (rcs[gidx]) {
                smj_fail_msg(errs[gidx].c_str());
                return rcs[gidx];
            }
        return 0;
    };
    int rc;
    if (g_multi.ngpus != W || g_multi.cap[0] < want[0] || g_multi.cap[1] < want[1]) {
        multi_destroy_locked();
        g_multi.groups.assign(W, nullptr);
        std::vector<unsigned char> handles((size_t)W * SMJ_GROUP_HANDLE_BYTES);
        if ((rc = run_all([&](int gidx) {
                 int e = smjdev_group_create(gidx, W, want[0], want[1], &g_multi.groups[gidx]);
                 return e ? e : smjdev_group_export(g_multi.groups[gidx], handles.data() + (size_t)gidx * SMJ_GROUP_HANDLE_BYTES);
             }))) {
            multi_destroy_locked();
            return rc;
        }
        if ((rc = run_all([&](int gidx) { return smjdev_group_connect(g_multi.groups[gidx], handles.data()); }))) {
            multi_destroy_locked();
            return rc;
        }
        g_multi.ngpus  = W;
        g_multi.cap[0] = want[0];
        g_multi.cap[1] = want[1];
    }
    const int kind_r = smjdev_pointer_kind(r), kind_s = smjdev_pointer_kind(s);
    std::vector<smj_join_stats_t> st(W);
    std::vector<float>            h2d(W, 0.f);
    rc = run_all([&](int gidx) {
        SmjCtx *ctx = smj_ctx();
        if (!ctx) return -1;
        // chunks like src/joins/joincommon.c:105-139
        const u64 pr = nr / W, ps = ns / W;
        const u64 r0 = pr * gidx, s0 = ps * gidx;
        const u64 cr = gidx == W - 1 ? nr - r0 : pr, cs = gidx == W - 1 ? ns - s0 : ps;
        SmjScratch sc(ctx);
        const u64 *dr = r + r0, *ds = s + s0;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx->stream);
        auto stage = [&](const u64 *src, u64 n, int kind, const u64 **out) -> int {
            // device memory of THIS gpu is used as it is; anything else (host memory, another GPU's
            // memory) is copied into local scratch first
            cudaPointerAttributes attr;
            bool local = false;
            if (kind == 2 && cudaPointerGetAttributes(&attr, src) == cudaSuccess) local = attr.device == ctx->device;
            cudaGetLastError();
            if (local || n == 0) return 0;
            u64 *d = (u64 *)sc.get(n * 8);
            if (!d) return smj_fail_msg("smjdev_join_multi: out of device memory");
            *out = d;
            if (kind == 2 || kind == 3) {
                SMJ_CHECK(cudaMemcpyAsync(d, src, n * 8, cudaMemcpyDefault, ctx->stream));
                return 0;
            }
            return smjdev_upload(d, src, n * 8);
        };
        int e = stage(r + r0, cr, kind_r, &dr);
        if (!e) e = stage(s + s0, cs, kind_s, &ds);
        cudaEventRecord(e1, ctx->stream);
        if (!e) e = smjdev_group_join(g_multi.groups[gidx], (const uint64_t *)dr, cr, (const uint64_t *)ds, cs, &st[gidx], nullptr);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&h2d[gidx], e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return e;
    });
    if (rc) return rc;
    *stats = st[0];  // every rank holds the global result
    for (int gidx = 0; gidx < W; gidx++) {
        stats->ms_total = std::max(stats->ms_total, st[gidx].ms_total);
        stats->ms_h2d   = std::max(stats->ms_h2d, h2d[gidx]);
        if (gidx) stats->launches += st[gidx].launches;
    }
    return 0;
}

// The layout rule of k_xg_layout on the host, for W posts given as a dense array (world x
// SMJ_MSD_SAMPLE_WORDS words: [0..3] sampled OR/AND, [8..9] sampled items, [10..11] sizes, [12] the
// digit the rank used, [13..14] capacities, [16..1040) sampled histograms).  No GPU needed: the
// CPU tests replay the exchange with it across processes.  Outputs for rank `me`:
//   owner[512]; start[2][512] / cap[2][512]: this rank's region of every bucket inside the owner's
//   buffer; segstart[2][512 * world]: start of the region of (bucket b, source g) at index b * world + g
//   (relative to the owner's buffer); need[world][2]: items every rank's buffers must hold.
// Returns the status word (0 = the fused exchange may run).
extern "C" uint32_t smjdev_group_layout_host(int world, int me, const uint64_t *posts, uint32_t my_shift1,
                                             uint32_t *owner, uint64_t *start, uint64_t *cap, uint64_t *segstart,
                                             uint64_t *need) {
    auto post = [&](int g) { return posts + (size_t)g * XG_POSTW; };
    u64  gb[4] = {0ull, ~0ull, 0ull, ~0ull}, tot_n = 0;
    bool same = true;
    for (int g = 0; g < world; g++) {
        const uint64_t *p = post(g);
        gb[0] |= p[0];
        gb[1] &= p[1];
        gb[2] |= p[2];
        gb[3] &= p[3];
        tot_n += p[PW_NR] + p[PW_NS];
        same = same && p[PW_SHIFT1] == (u64)my_shift1;
    }
    MsdPlan pl;
    u32 status = 0;
    if (!(my_shift1 != 0 && same && smj_msd_plan(gb, &pl) && pl.shift1 == my_shift1 && tot_n >= (1u << 20))) status |= ST_PLAN;
    auto capof = [&](int g, int rel, int b) -> u64 {
        const uint64_t *p = post(g);
        const u64    n = p[PW_NR + rel], nsamp = p[8 + rel], k = p[16 + rel * XG_NB + b];
        const double scale = nsamp ? fmax(1.0, (double)n / (double)nsamp) : 1.0;
        return n ? smj_cap1(k, scale, 0) : 0ull;
    };
    std::vector<u64> load(XG_NB, 0), pre(XG_NB + 1, 0), P[2] = {std::vector<u64>(XG_NB + 1, 0), std::vector<u64>(XG_NB + 1, 0)};
    for (int b = 0; b < XG_NB; b++) {
        for (int g = 0; g < world; g++)
            for (int rel = 0; rel < 2; rel++) {
                const uint64_t *p = post(g);
                const u64    n = p[PW_NR + rel], nsamp = p[8 + rel], k = p[16 + rel * XG_NB + b];
                const double scale = nsamp ? fmax(1.0, (double)n / (double)nsamp) : 1.0;
                (void)n;
                load[b] += (u64)((double)k * scale);
                P[rel][b + 1] += capof(g, rel, b);
            }
        pre[b + 1] = pre[b] + load[b];
        for (int rel = 0; rel < 2; rel++) P[rel][b + 1] += P[rel][b];
    }
    const u64 total = pre[XG_NB];
    std::vector<u32> first(world + 1, XG_NB);
    for (int b = 0; b < XG_NB; b++) {
        u32 own = 0;
        if (total) own = (u32)std::min<u64>((u64)(world - 1), ((2 * pre[b] + load[b]) * (u64)world) / (2 * total));
        owner[b] = own;
        const u32 prev = b ? owner[b - 1] : 0xffffffffu;
        if (b == 0)
            for (u32 p = 0; p <= own; p++) first[p] = 0;
        else
            for (u32 p = prev + 1; p <= own; p++) first[p] = (u32)b;
    }
    for (int p = 0; p < world; p++) {
        const u32 f0 = first[p], f1 = (p + 1 < world) ? first[p + 1] : (u32)XG_NB;
        for (int rel = 0; rel < 2; rel++) {
            need[p * 2 + rel] = P[rel][f1] - P[rel][f0];
            if (need[p * 2 + rel] > post(p)[PW_CAPR + rel]) status |= ST_CAPACITY;
        }
    }
    for (int rel = 0; rel < 2; rel++)
        for (int b = 0; b < XG_NB; b++) {
            u64 off = P[rel][b] - P[rel][first[owner[b]]];
            for (int g = 0; g < world; g++) {
                const u64 c = capof(g, rel, b);
                segstart[(size_t)rel * XG_NB * world + (size_t)b * world + g] = off;
                if (g == me) {
                    start[rel * XG_NB + b] = off;
                    cap[rel * XG_NB + b]   = c;
                }
                off += c;
            }
        }
    return status;
}

extern "C" int smjdev_join_multi_release(void) {
    std::lock_guard<std::mutex> lk(g_multi.mu);
    multi_destroy_locked();
    return 0;
}
