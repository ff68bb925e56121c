/**
 * smj_device.h -- the thin C-ABI layer over the hand-written sm_100a kernels.
 *
 * Plain pointers and sizes only; no CUDA or torch types appear in any signature (a CUDA
 * stream is passed as void*).  The reference-facing entry points of smj_b200.h are written in
 * C on top of this layer.  Every function returns 0 on success and a negative value on
 * failure (message via smjdev_last_error()); there is NO CPU fallback anywhere below this
 * line: if no CUDA device is usable every call fails.
 *
 * "item" = one 8-byte tuple viewed as uint64 = key<<32 | (uint32)payload
 * (reference src/types.h:51-54).  All `uint64_t *` data arguments below are DEVICE pointers
 * unless the comment says "host".
 */
#ifndef SMJ_DEVICE_H
#define SMJ_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ runtime ---------- */
int         smjdev_device_count(void);
int         smjdev_init(int device);          /* idempotent; device<0 -> $SMJ_DEVICE, $LOCAL_RANK or 0 */
int         smjdev_current_device(void);
int         smjdev_set_stream(void *cuda_stream); /* run on the caller's stream; NULL = legacy default stream */
int         smjdev_use_own_stream(void);          /* back to the library's own non-blocking stream (the default) */
void       *smjdev_get_stream(void);
int         smjdev_sync(void);
const char *smjdev_last_error(void);
uint64_t    smjdev_launch_count(void);        /* kernels launched by this library so far */
int         smjdev_sm_count(void);

/* ------------------------------------------------------------------ memory ----------- */
void *smjdev_malloc(size_t bytes);            /* HBM, stream-ordered pool */
int   smjdev_free(void *dptr);
void *smjdev_host_alloc(size_t bytes);        /* pinned host memory */
int   smjdev_host_free(void *hptr);
int   smjdev_memcpy_h2d(void *dst, const void *src, size_t bytes); /* async on the stream */
int   smjdev_memcpy_d2h(void *dst, const void *src, size_t bytes); /* async on the stream */
int   smjdev_memcpy_d2d(void *dst, const void *src, size_t bytes); /* async on the stream */
int   smjdev_memset(void *dst, int value, size_t bytes);
/* Transfers for ANY host memory: pinned memory moves with one DMA; ordinary malloc'ed (pageable)
 * memory, which a plain cudaMemcpy moves at ~11 GB/s, goes through a ring of pinned bounce buffers
 * filled / drained by several host threads ($SMJ_STAGE_THREADS, default min(8, cores/2)).
 * upload is asynchronous on the library's stream; download has finished when it returns. */
int   smjdev_upload(void *dst, const void *src, size_t bytes);
int   smjdev_download(void *dst, const void *src, size_t bytes);
/* 0 = ordinary (pageable) host memory, 1 = pinned host, 2 = device, 3 = managed */
int   smjdev_pointer_kind(const void *p);

/* ------------------------------------------------------------------ timing ----------- */
int smjdev_event_create(void **ev);
int smjdev_event_record(void *ev);            /* on the library's stream */
int smjdev_event_elapsed_ms(void *start, void *stop, float *ms); /* synchronises on stop */
int smjdev_event_destroy(void *ev);

/* Per-kernel timing for roofline reports: between begin and end every kernel launch of the
 * library is bracketed by CUDA events on the launch stream.  end() writes one text line per
 * kernel, "name;launches;total_ms;total_items", into out (NUL terminated, truncated to cap). */
int smjdev_profile_begin(void);
int smjdev_profile_end(char *out, size_t cap);

/* ------------------------------------------------------------------ sort ------------- */
/* Ascending sort of n items (replaces src/avxsort/avxsort.c:128-209 avxsort_aligned and
 * the in-partition sort of src/joins/sortmergejoin_multiway.c:388-460).
 * buf and tmp are two n-item device buffers, both clobbered; *sorted is set to whichever
 * of the two holds the result.  key_only != 0 orders by the key (upper 32 bits) only and
 * keeps the input order among equal keys (stable) -- enough for the merge-join and cheaper.
 * algo: 0 = auto, 1 = LSD one-sweep radix passes over the varying bits (the only sort so far). */
int smjdev_sort(uint64_t *buf, uint64_t *tmp, uint64_t n, int key_only, int algo,
                uint64_t **sorted);

/* In-place ascending sort of n int32 values (src/scalarsort/scalarsort.c:91-99); `data` may be a
 * host or a device pointer.  Implemented by widening to items and one key-only sort. */
int smjdev_sort_int32(int32_t *data, uint64_t n);

/* ------------------------------------------------------------------ partition -------- */
/* Stable radix partition (replaces src/partition/partition.c:94-287).
 * bucket(x) = ((key(x) - key_minus_one) >> shift) & (2^nbits - 1)     (partition.c:29)
 * padded != 0: every partition starts at a multiple of 8 tuples (partition.c:183-187),
 *              `out` must hold n + 8*2^nbits items; otherwise partitions are dense.
 * tmp: n-item scratch, needed only when nbits > 8 (two stable passes); may be NULL else.
 * hist: device array of 2^nbits uint64 that receives the partition sizes.
 * offs: device array of 2^nbits uint64 that receives each partition's start (in items). */
int smjdev_partition(const uint64_t *in, uint64_t *out, uint64_t *tmp, uint64_t n,
                     uint32_t nbits, uint32_t shift, int key_minus_one, int padded,
                     uint64_t *hist, uint64_t *offs);
/* histogram only (first half of partition.c:388-420 baseline_histogram_memcpy) */
int smjdev_histogram(const uint64_t *in, uint64_t n, uint32_t nbits, uint32_t shift,
                     int key_minus_one, uint64_t *hist);

/* ------------------------------------------------------------------ merge ------------ */
/* 2-way merge-path merge (replaces src/merge/merge.c:28-235 and the mergeN_varlen kernels
 * of src/avxsort/avxsort_core.h:390-902).  mode 0: compare whole 64-bit word (AVX
 * semantics); mode 1: compare key only, take from B on ties (scalar_merge_tuples,
 * merge.c:78-87).  Inputs are NOT modified. */
int smjdev_merge(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb,
                 uint64_t *out, int mode);
/* k-way merge of sorted runs (replaces src/merge/avx_multiwaymerge.c:199-338):
 * ceil(log2 k) rounds of batched pairwise merge-path merges, all pairs of a round in one
 * launch.  runs/lens are HOST arrays (k device pointers / k lengths); out and tmp are
 * device buffers of sum(lens) items; the result always lands in out. */
int smjdev_multiway_merge(uint64_t *out, uint64_t *tmp, const uint64_t *const *runs,
                          const uint64_t *lens, uint32_t nruns, int mode);

/* ------------------------------------------------------------------ merge join ------- */
/* Merge-join scan of two key-sorted device arrays (replaces src/joins/joincommon.c:239-312).
 * count    = sum over keys of |run_R| * |run_S|
 * checksum = sum over matched pairs of A(r)*B(s) mod 2^64 (DESIGN.md "checksum")
 * Results are written to HOST variables; the call synchronises. */
int smjdev_merge_join(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                      uint64_t *count, uint64_t *checksum);
/* semi != 0: count the S tuples that have a partner in R (what joincommon.c:325-394
 * merge_join_interpolation counts); the checksum is then undefined */
int smjdev_merge_join_ex(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns, int semi,
                         uint64_t *count, uint64_t *checksum);
/* asynchronous variant: accumulates into a device pair acc[0]=count, acc[1]=checksum */
int smjdev_merge_join_acc(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                          uint64_t *acc);

/* ------------------------------------------------------------------ join pipelines --- */
enum { SMJ_ALGO_MWAY = 0, SMJ_ALGO_MPASS = 1, SMJ_ALGO_MPSM = 2 };

typedef struct smj_join_stats_t {
    uint64_t matches;
    uint64_t checksum;
    float    ms_h2d;       /* host->device copies (0 when inputs were device resident) */
    float    ms_partition; /* range partition passes                                    */
    float    ms_sort;      /* in-partition / run sort                                   */
    float    ms_merge;     /* run merging (m-pass, m-way with nruns>1)                  */
    float    ms_join;      /* merge-join scan                                           */
    float    ms_total;     /* first kernel -> result on host                            */
    uint64_t launches;     /* kernels launched by this call                             */
    uint64_t hbm_bytes;    /* peak scratch bytes allocated by this call                 */
    uint32_t path;         /* SMJ_PATH_*: which pipeline produced the result            */
    uint32_t repeats;      /* optimistic partition passes that overflowed and were
                              repeated exactly (histogram first)                        */
} smj_join_stats_t;
/* m-way join: general = LSD sort of R and S + merge-path scan; msd = two-level MSD range
 * partition + fused bucket join with histogram passes; msd_opt = the same with at least one
 * histogram-free (sampled-capacity) partition pass.  m-pass / mpsm report SMJ_PATH_GENERAL. */
enum { SMJ_PATH_GENERAL = 0, SMJ_PATH_MSD = 1, SMJ_PATH_MSD_OPT = 2 };

/* Join of two relations in device memory or in PINNED host memory (left unmodified).  A pinned
 * host relation is uploaded in 128 MiB chunks on a separate copy stream; the m-way path runs its
 * first partitioning pass on every chunk as it arrives, so partitioning hides behind the PCIe
 * transfer (stats->ms_h2d = duration of the copies).
 * nruns: number of sorted runs per relation for m-pass / mpsm (>=1; m-way ignores it). */
int smjdev_join(int algo, const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                int nruns, smj_join_stats_t *stats);

/* The join with its result written out (the reference's JOIN_MATERIALIZE build,
 * src/joins/joincommon.c:256-278, appends <S.key, S.payload> per matching pair to per-thread
 * chained buffers).  r/s: device arrays, left unmodified.  out: device array of `capacity` 64-bit
 * items or NULL (count only).  *count (host) receives the number of matching pairs; at most
 * `capacity` of them are written, in unspecified order.  Synchronises. */
enum {
    SMJ_OUT_S_TUPLE  = 0, /* the reference's layout: the S tuple, once per matching R tuple */
    SMJ_OUT_RID_PAIR = 1  /* R.payload << 32 | S.payload (row-id pair)                      */
};
int smjdev_join_materialize(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                            int layout, uint64_t *out, uint64_t capacity, uint64_t *count);

/* Pass A into a local buffer laid out bucket-major; hist512_dev is the DEVICE histogram
 * smjdev_msd_hist produced (no host round trip: the pass can run while the ranks all-gather). */
int smjdev_msd_scatter_local(const uint64_t *in, uint64_t n, uint32_t shift1,
                             const uint64_t *hist512_dev, uint64_t *out);
int smjdev_join_level1(const uint64_t *r, const uint64_t *rcounts512, const uint64_t *s,
                       const uint64_t *scounts512, uint32_t shift1, uint32_t lo_bit,
                       smj_join_stats_t *stats);
/* Same, for first-level buckets that arrive in SEGMENTS (copy-engine exchange: every source rank
 * contributes one contiguous segment per bucket).  rseg / sseg: HOST arrays {start[nseg],
 * count[nseg], bucket[nseg]} (items, items, bucket index 0..511), nseg <= 8192. */
int smjdev_join_level1_seg(const uint64_t *r, uint32_t nseg_r, const uint64_t *rseg, const uint64_t *s,
                           uint32_t nseg_s, const uint64_t *sseg, uint32_t shift1, uint32_t lo_bit,
                           smj_join_stats_t *stats);
/* Bulk exchange on the copy engines, ordered with the compute stream: begin() makes the copy stream
 * wait for everything enqueued so far, copy() enqueues one contiguous copy (dst / src may be peer
 * buffers opened with smjdev_ipc_open), end() makes the compute stream wait for the copies. */
/* Measurement helper (tools/p2p_store_bw.py): SM-store copy of n items, `width` (8 or 16) bytes per
 * lane, in runs of `run` items separated by `gap` items in dst (dst may be a peer buffer). */
int smjdev_debug_copy(const uint64_t *src, uint64_t *dst, uint64_t n, int width, uint32_t run,
                      uint32_t gap);
int smjdev_exchange_begin(void);
int smjdev_exchange_copy(void *dst, const void *src, size_t bytes);
int smjdev_exchange_end(void);

/* Multi-GPU building blocks (one process per GPU; the exchange itself is done by the
 * caller with NCCL).
 * Range-partition n items into nparts (<= 32) destinations by splitters[0..nparts-2] (ascending
 * keys, HOST array of uint32): dest(x) = number of splitters <= key(x).  Items of one
 * destination are contiguous in out; counts (device, nparts uint64) receives the sizes. */
int smjdev_range_partition(const uint64_t *in, uint64_t *out, uint64_t n,
                           const uint32_t *splitters, uint32_t nparts, uint64_t *counts);

/* Fused range partition + exchange (one process per GPU): smjdev_range_count gives the number
 * of items per destination (device array of nparts uint64); after the ranks have exchanged their
 * counts, smjdev_range_scatter stores every item straight into the receive buffer of the GPU
 * that owns its key range.  dest_ptrs (HOST array of nparts device pointers) may point into peer
 * GPUs' memory mapped with smjdev_ipc_open (stores travel over NVLink); dest_offsets (HOST array)
 * is where this rank's slice starts inside each buffer, in items.  Order inside a destination is
 * unspecified.  The caller synchronises all ranks before reading the receive buffers. */
int smjdev_range_count(const uint64_t *in, uint64_t n, const uint32_t *splitters, uint32_t nparts,
                       uint64_t *counts);
int smjdev_range_scatter(const uint64_t *in, uint64_t n, const uint32_t *splitters, uint32_t nparts,
                         uint64_t *const *dest_ptrs, const uint64_t *dest_offsets);
/* The same idea one level deeper (what DistributedJoin.step_msd uses): the FIRST pass of the m-way
 * join's two-level MSD partition is the exchange.
 *   smjdev_sample_key_bits: {OR_R, AND_R, OR_S, AND_S} of a 2^16-item sample (host array of 4; syncs):
 *                           a guess of the varying key bits, to be verified with the exact values
 *   smjdev_msd_hist:        sizes of the 512 first-level buckets (digit = 9 bits at bit `shift1` of
 *                           the item) into hist512 (device), exact OR / AND accumulated into bits2
 *                           (device pair, caller initialises {0, ~0}; may be NULL)
 *   smjdev_msd_scatter:     first-level scatter with caller-supplied destinations: bucket d goes
 *                           to dest_ptrs512[d] (device pointer, possibly a peer GPU's buffer)
 *                           starting at item dest_offsets512[d] (HOST arrays of 512)
 *   smjdev_join_level1:     joins two relations that are already partitioned by the first-level
 *                           digit (buckets back to back, sizes in the HOST arrays rcounts512 /
 *                           scounts512): second-level partition + fused bucket join.  lo_bit =
 *                           lowest varying KEY bit (0..31). */
int smjdev_sample_key_bits(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                           uint64_t *host4);
int smjdev_msd_hist(const uint64_t *in, uint64_t n, uint32_t shift1, uint64_t *hist512, uint64_t *bits2);
int smjdev_msd_scatter(const uint64_t *in, uint64_t n, uint32_t shift1, uint64_t *const *dest_ptrs512,
                       const uint64_t *dest_offsets512);
/* (smjdev_join_level1 is declared after smj_join_stats_t below) */
/* device buffers that other processes can map (cudaMalloc + CUDA IPC handles, 64 bytes each) */
int smjdev_ipc_alloc(size_t bytes, void **dptr);
int smjdev_ipc_free(void *dptr);
int smjdev_ipc_export(void *dptr, unsigned char *handle64);
int smjdev_ipc_open(const unsigned char *handle64, void **peer_ptr);
int smjdev_ipc_close(void *peer_ptr);

/* ------------------------------------------------------------------ KEY_8B ----------- */
/* 16-byte tuples {int64 payload; int64 key} (the reference's KEY_8B build, src/types.h:23-29).
 * sort: by signed key, equal keys in input order; in != out.  join: count and checksum (64-bit
 * analogue of the 8-byte checksum, see csrc/smj_key8b.cu) of R join S; presorted != 0 when both
 * inputs are already key-sorted (merge_join).  Device pointers. */
int smjdev_sort_key8b(const void *in, void *out, uint64_t n);
int smjdev_join_key8b(const void *r, uint64_t nr, const void *s, uint64_t ns, int presorted,
                      smj_join_stats_t *stats);

/* ------------------------------------------------------------------ join groups ------ */
/* The m-way join over the G GPUs of one box (csrc/smj_mgpu.cu).  A GROUP has one member ("rank") per
 * GPU; a member lives in the host thread that is bound to its GPU (smjdev_init) -- one process per
 * GPU, or one thread per GPU inside one process.  The first partitioning pass of the join is the
 * exchange: every rank's scatter kernel stores each item into the receive buffer of the GPU that
 * owns its key bucket over NVLink; samples, counts and results travel through mailboxes in peer
 * memory, so that no collective library and no host round trip sits between the sample and the
 * result.  Replaces the fan-out inside sortmergejoin_initrun (src/joins/joincommon.c:105-173) and the
 * partition pull of src/joins/sortmergejoin_multiway.c:492-517.
 *   create   allocates this rank's receive buffers (cap_r / cap_s items) and control block
 *   export   SMJ_GROUP_HANDLE_BYTES describing them (CUDA IPC handles + plain pointers)
 *   connect  handles of ALL ranks in rank order (world x SMJ_GROUP_HANDLE_BYTES); call it on every
 *            rank after every rank has been created (ranks of one process: peer access; other
 *            processes: CUDA IPC)
 *   join     collective: every rank passes its slices of R and S (device pointers) and receives the
 *            GLOBAL {matches, checksum}; *bytes_sent = item bytes this rank stored into other GPUs
 *   destroy  collective as far as the peers' mappings go: call it on every rank */
#define SMJ_GROUP_HANDLE_BYTES 256
int smjdev_group_create(int rank, int world, uint64_t cap_r, uint64_t cap_s, void **group);
int smjdev_group_export(void *group, unsigned char *handle);
int smjdev_group_connect(void *group, const unsigned char *handles);
int smjdev_group_join(void *group, const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                      smj_join_stats_t *stats, uint64_t *bytes_sent);
int smjdev_group_destroy(void *group);
/* One process driving `ngpus` GPUs (what sortmergejoin_multiway() does under SMJ_NGPUS=G): R and S --
 * host pointers (pinned or pageable) or device pointers -- are cut into G contiguous chunks like
 * src/joins/joincommon.c:105-139, one host thread per GPU uploads and joins its chunk through a group
 * that is kept between calls (smjdev_join_multi_release drops it). */
int smjdev_join_multi(int ngpus, const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                      smj_join_stats_t *stats);
int smjdev_join_multi_release(void);
/* The group's layout rule as a plain host function (no GPU needed; the CPU tests replay the exchange
 * with it): posts = world x 1040 words as the ranks publish them ([0..3] sampled OR_R AND_R OR_S AND_S,
 * [8..9] sampled items of R / S, [10..11] |R| |S| of the rank, [12] the first-level digit shift the rank
 * derived, [13..14] its buffer capacities, [16..528) / [528..1040) sampled first-level histograms).
 * Outputs for rank `me`: owner[512]; start[2][512], cap[2][512] = its region of every bucket inside
 * the owner's buffer; segstart[2][512 * world] = region start of (bucket b, source g) at b * world + g;
 * need[world][2] = items the layout needs in every rank's buffers.  Returns the status word: 0 = go. */
uint32_t smjdev_group_layout_host(int world, int me, const uint64_t *posts, uint32_t my_shift1,
                                  uint32_t *owner, uint64_t *start, uint64_t *cap, uint64_t *segstart,
                                  uint64_t *need);

/* ------------------------------------------------------------------ verification ----- */
/* Device-side property checks for arrays that are too large to bring back to the host.
 * check_order: stats[0] = number of positions i with items[i] < items[i-1] (whole word, or key
 * only), stats[1] = sum of the items mod 2^64, stats[2] = xor of the items (multiset
 * fingerprints: equal before and after a sort).
 * check_partition: for the output of smjdev_partition (offs/hist device arrays, `used` = slots in
 * use): stats[0] = items whose bucket is not the partition that holds them, stats[1] = payload
 * descents inside partitions (0 for a stable scatter of rows numbered by payload), stats[2] =
 * items found, stats[3] = their sum mod 2^64. */
int smjdev_check_order(const uint64_t *items, uint64_t n, int key_only, uint64_t *stats);
int smjdev_check_partition(const uint64_t *out, const uint64_t *offs, const uint64_t *hist,
                           uint32_t nbits, uint32_t shift, int key_minus_one, uint64_t used,
                           uint64_t *stats);

/* ------------------------------------------------------------------ data generators -- */
/* Deterministic on-device relation generators (replace src/datagen/generator.c:83-93, 234-350
 * and :458-488 for the benchmark workloads; the reference's own are seeded from time()).
 * Rows [first, first+count) of the relation are written to out[0..count); payload = row index.
 * unique: keys are a pseudo-random permutation of 1..n (Feistel network + cycle walking).
 * fk:     keys are uniformly drawn from 1..maxkey; a key with bits 20..30 all set (a NaN when the
 *         word is read as a double) gets bit 20 cleared like generator.c:107-121 avoid_NaN does. */
int smjdev_gen_unique(uint64_t *out, uint64_t n, uint64_t first, uint64_t count, uint64_t seed);
int smjdev_gen_fk(uint64_t *out, uint64_t first, uint64_t count, uint32_t maxkey, uint64_t seed);
/* fk_from: rows [first, first + count) of an n-row relation made of the rows of src (device array of
 * src_n items) repeated n / src_n times (+ remainder) and shuffled: create_relation_fk_from_pk,
 * src/datagen/generator.c:458-488 (what `--full-range` gives S).  Row i = src[pi(i) mod src_n]. */
int smjdev_gen_fk_from(uint64_t *out, uint64_t n, uint64_t first, uint64_t count, const uint64_t *src,
                       uint64_t src_n, uint64_t seed);
/* zipf: src/datagen/genzipf.c:29-159 (create_relation_zipf, generator.c:517-534).  Keys are drawn from
 * the alphabet 1..A with P(rank i) ~ i^-theta through the reference's cumulative table and binary
 * search; the alphabet is the permutation smjdev_gen_unique(n = A, seed = alphabet_seed) produces
 * (so rank i is the key of R's row i when R was generated with that seed); payload = row index.
 * The reference draws from glibc's sequential rand(); here row i uses 31 bits of a counter-based
 * hash of (seed, i) -- same distribution, rows independent.  smjdev_zipf_lut builds the cumulative
 * table (A doubles; tmp = A/256 + A/65536 + 4 doubles of scratch) once for several calls;
 * smjdev_gen_zipf(lut = NULL) builds and drops its own. */
int smjdev_zipf_lut(double *lut, uint32_t A, double theta, double *tmp);
int smjdev_gen_zipf(uint64_t *out, uint64_t first, uint64_t count, uint32_t A, double theta,
                    uint64_t seed, uint64_t alphabet_seed, const double *lut);

#ifdef __cplusplus
}
#endif
#endif /* SMJ_DEVICE_H */
