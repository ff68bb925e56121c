// smj_internal.cuh -- functions shared between the .cu files of the library (not part of the ABI).
#pragma once
#include "smj_common.cuh"

// ---- smj_radix.cu
int smj_lsd_sort(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, u64 **sorted);
int smj_lsd_sort_bits(SmjCtx *ctx, const u64 *in, u64 *a, u64 *b, u64 n, int key_only, const u64 *h_bits,
                      u64 **sorted);
int smj_item_bits(SmjCtx *ctx, const u64 *in, u64 n, u64 *h_bits);
int smj_multiway_merge(SmjCtx *ctx, u64 *out, u64 *tmp, const u64 *const *runs, const u64 *lens,
                       u32 nruns, int mode);
int smj_msd_partition(SmjCtx *ctx, const u64 *in, u64 n, u64 *a, u64 *b, u32 shift1, u32 shift2,
                      u64 *offsets, u64 *exact_bits);
int smj_msd_pass_b(SmjCtx *ctx, const u64 *a, u64 n, const u64 *d_counts, const u64 *d_starts, u64 *b,
                   u32 shift2, u64 *offsets);
int smj_msd_pass_b_seg(SmjCtx *ctx, const u64 *a, u64 n, u32 nseg, const u64 *d_seg, u64 *b, u32 shift2,
                       u64 *offsets);
int smj_msd_hist1(SmjCtx *ctx, const u64 *in, u64 n, u32 shift1, u64 *hist, u64 *exact_bits);
int smj_msd_scatter_to(SmjCtx *ctx, const u64 *in, u64 n, u32 shift1, u64 *const *dest_ptrs,
                       const u64 *dest_offsets);
int smj_msd_scatter_local(SmjCtx *ctx, const u64 *in, u64 n, u32 shift1, const u64 *d_hist512, u64 *out);
int smj_key_bits_sampled(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, u64 *or_and_host);
int smj_msd_sample(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, u32 max_sectors, u64 *d_st,
                   u32 *d_sub, u64 *host);
size_t smj_msd_sample_bytes(int which);
u64 smj_msd_caps1(const u64 *hist, u64 nsamp, u64 n, int shrink);
int smj_msd_caps1_tables(SmjCtx *ctx, const u64 *d_st, u64 nr, u64 ns, u64 alloc_r, u64 alloc_s, int shrink,
                         u64 *tab_r, u64 *tab_s, u32 *d_flags);
u64 smj_msd_alloc2(u64 n);
int smj_msd_pass_a_opt(SmjCtx *ctx, const u64 *in, u64 n, u64 *a, u32 shift1, u64 *d_tab, u32 *d_flag,
                       u32 kmask, u32 kref, u32 *d_kflag);
int smj_msd_pass_b_opt(SmjCtx *ctx, const u64 *a, u64 n, const u64 *d_tab, u64 *b, u64 alloc2, u32 shift2,
                       u32 dense_slot, u64 *starts, u64 *ends, u64 *d_counts, u32 *d_flag, u64 *d_nfixed,
                       int shrink);
int smj_msd_counts_opt(SmjCtx *ctx, const u64 *d_tab, u64 *d_counts);

int smj_msd_pass_b_seg_dev(SmjCtx *ctx, const u64 *a, u64 n_upper, u32 nseg, const u64 *d_seg, const u64 *d_n,
                           u64 *b, u32 shift2, u64 *offsets);
int smj_msd_pass_a_abs(SmjCtx *ctx, const u64 *in, u64 n, u32 shift1, u64 *d_tab, u32 *d_flag, u32 kmask,
                       u32 kref, u32 *d_kflag, int big_tile, int max_grid);
int smj_msd_pass_b_opt_seg(SmjCtx *ctx, const u64 *a, u64 n_upper, u32 world, const u64 *d_seg, u64 *b, u64 alloc2,
                           u32 shift2, u32 dense_slot, u64 *starts, u64 *ends, u32 *d_flag);
constexpr u64 SMJ_SMALL_SORT_MAX = 16384;  // the reference's in-cache block (src/avxsort/avxsort.c:17)

// ---- smj_join.cu
// Buckets of one relation after the second partitioning level: bucket q is items[starts[q] ..
// ends[q]).  An exact pass B leaves them back to back (ends = starts + 1 of one offsets array), an
// optimistic one in fixed slots.
struct BucketSet {
    const u64 *items, *starts, *ends;
    u64        n;
};
int  smj_bucket_collect(SmjCtx *ctx, const BucketSet &R, const BucketSet &S, u64 *d_state, u32 *d_list);
int  smj_bucket_join(SmjCtx *ctx, const BucketSet &R, const BucketSet &S, u32 lo_shift, u32 lbits, u64 *d_state,
                     u32 *d_list, u64 *acc, const u32 *d_range = nullptr /* device: first and one-past-last bucket in use */);
bool smj_bucket_join_valid(const u64 *h_state, u32 lbits);
extern const u64 MSD_STATE_INIT[16];
int smj_merge_join_acc(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *s, u64 ns, u64 *acc, int semi = 0);
// ---- smj_mpsm.cu: all sorted runs of S against the sorted R in one pass over R (*used = 0: not applicable)
int smj_mpsm_probe(SmjCtx *ctx, const u64 *r, u64 nr, const u64 *const *runs, const u64 *lens, u32 nruns, u64 *acc,
                   int *used);
