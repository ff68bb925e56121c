/*
 * smj_host_api.c -- the reference-facing C entry points (include/smj_b200.h), written in C like
 * the reference's own host code, on top of the thin C-ABI device layer (include/smj_device.h).
 *
 * This file contains NO compute: it validates arguments the way the reference does, moves the
 * caller's buffers to HBM when they are host pointers, calls the CUDA kernels, copies results
 * back and reproduces the reference's ownership / pointer-swap / printing conventions.
 * There is no CPU fallback: if the device layer fails the call is fatal (smj_fatal).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "../../include/smj_b200.h"

/* ------------------------------------------------------------------------------------------
 * error handling: the reference prints and exit()s on its fatal paths
 * (src/joins/joincommon.c:79-82, 92-102, 160-163)
 * ---------------------------------------------------------------------------------------- */
static void smj_fatal(const char *where) {
    fprintf(stdout, "[ERROR] %s: %s\n", where, smjdev_last_error());
    fflush(stdout);
    fprintf(stderr, "[ERROR] %s: %s\n", where, smjdev_last_error());
    fflush(stderr);
    if (!getenv("SMJ_NO_EXIT")) exit(EXIT_FAILURE);
}

static int quiet(void) { return getenv("SMJ_QUIET") != NULL; }

/* a caller buffer that is usable by kernels: either the caller's own device pointer or a
 * temporary HBM copy of a host buffer */
typedef struct {
    uint64_t *d;
    int       owned;
} devbuf_t;

static int is_device_ptr(const void *p) {
    int k = smjdev_pointer_kind(p);
    return k == 2 || k == 3;
}

/* returns 0 on success; copy != 0 uploads the contents.  copy == 2 (joins): a pinned host buffer is
 * handed to the device layer as it is: smjdev_join uploads it chunk by chunk on its copy stream and
 * starts partitioning while the transfer is still running. */
static int stage_in(devbuf_t *b, const void *p, uint64_t n, int copy) {
    b->d     = NULL;
    b->owned = 0;
    if (n == 0) return 0;
    if (is_device_ptr(p) || (copy == 2 && smjdev_pointer_kind(p) == 1 && !getenv("SMJ_NO_OVERLAP"))) {
        b->d = (uint64_t *)p;
        return 0;
    }
    b->d = (uint64_t *)smjdev_malloc(n * sizeof(uint64_t));
    if (!b->d) return -1;
    b->owned = 1;
    if (copy && smjdev_upload(b->d, p, n * sizeof(uint64_t))) return -1;
    return 0;
}

static void stage_free(devbuf_t *b) {
    if (b->owned && b->d) smjdev_free(b->d);
    b->d     = NULL;
    b->owned = 0;
}

/* ------------------------------------------------------------------------------------------
 * joins
 * ---------------------------------------------------------------------------------------- */
static __thread smj_join_stats_t g_last_stats;
static __thread int              g_have_stats = 0;

int smj_last_join_stats(smj_join_stats_t *out) {
    if (!g_have_stats || !out) return -1;
    *out = g_last_stats;
    return 0;
}

static double usec_between(struct timeval *a, struct timeval *b) {
    return (double)((b->tv_sec * 1000000L + b->tv_usec) - (a->tv_sec * 1000000L + a->tv_usec));
}

/* src/joins/joincommon.c:214-227 */
void print_timing(uint64_t numtuples, struct timeval *start, struct timeval *end, FILE *out) {
    double diff_usec = usec_between(start, end);
    fprintf(out, "NUM-TUPLES = %lld TOTAL-TIME-USECS = %.4lf ", (long long)numtuples, diff_usec);
    fprintf(out, "TUPLES-PER-SECOND = ");
    fflush(out);
    fprintf(out, "%.4lf ", (numtuples / (diff_usec / 1000000L)));
    fflush(out);
}

static result_t *join_common(int algo, relation_t *relR, relation_t *relS, joinconfig_t *cfg,
                             const char *name) {
    struct timeval   start, end;
    devbuf_t         dr, ds;
    smj_join_stats_t st;
    void            *e0 = NULL, *e1 = NULL;
    int              nthreads = cfg->NTHREADS > 0 ? cfg->NTHREADS : 1;
    float            ms_h2d   = 0.f;

    int failed;
    gettimeofday(&start, NULL);
    memset(&dr, 0, sizeof(dr));
    memset(&ds, 0, sizeof(ds));
    /* SMJ_NGPUS=G: like the reference fans out over NTHREADS workers inside this call
     * (src/joins/joincommon.c:105-173), the m-way join fans out over G GPUs: one host thread per
     * GPU takes a contiguous chunk of R and S, the chunks meet over NVLink (csrc/smj_mgpu.cu) */
    const char *ng = getenv("SMJ_NGPUS");
    if (algo == SMJ_ALGO_MWAY && ng && atoi(ng) > 1 && relR->num_tuples && relS->num_tuples) {
        if (smjdev_join_multi(atoi(ng), (const uint64_t *)relR->tuples, relR->num_tuples,
                              (const uint64_t *)relS->tuples, relS->num_tuples, &st)) {
            smj_fatal(name);
            return NULL;
        }
        goto have_result;
    }
    failed = smjdev_init(-1) || smjdev_event_create(&e0) || smjdev_event_create(&e1) ||
             smjdev_event_record(e0) || stage_in(&dr, relR->tuples, relR->num_tuples, 2) ||
             stage_in(&ds, relS->tuples, relS->num_tuples, 2) || smjdev_event_record(e1) ||
             /* m-pass / mpsm work on NTHREADS sorted runs per relation, like the reference's threads */
             smjdev_join(algo, dr.d, relR->num_tuples, ds.d, relS->num_tuples, nthreads, &st);
    if (!failed && (dr.owned || ds.owned)) smjdev_event_elapsed_ms(e0, e1, &ms_h2d);
    stage_free(&dr);
    stage_free(&ds);
    smjdev_event_destroy(e0);
    smjdev_event_destroy(e1);
    if (failed) {
        smj_fatal(name); /* exits unless SMJ_NO_EXIT */
        return NULL;
    }
    if (dr.owned || ds.owned) st.ms_h2d += ms_h2d; /* (pinned relations: measured by smjdev_join) */
have_result:
    gettimeofday(&end, NULL);

    g_last_stats = st;
    g_have_stats = 1;

    /* result ownership as src/joins/joincommon.c:109-112, 168-173 */
    result_t *res = (result_t *)malloc(sizeof(result_t));
    if (res) res->resultlist = (threadresult_t *)calloc((size_t)nthreads, sizeof(threadresult_t));
    if (!res || !res->resultlist) { /* src/joins/joincommon.c:79-82: allocation failures are fatal */
        free(res);
        perror("[ERROR] result_t allocation failed");
        if (!getenv("SMJ_NO_EXIT")) exit(EXIT_FAILURE);
        return NULL;
    }
    for (int i = 0; i < nthreads; i++) res->resultlist[i].threadid = (uint32_t)i;
    res->resultlist[0].nresults = (int64_t)st.matches;
    res->totalresults           = (int64_t)st.matches;
    res->nthreads               = nthreads;

    if (!quiet()) {
        /* same lines as src/joins/joincommon.c:176-196; values are nanoseconds of GPU time */
        unsigned long long part = (unsigned long long)(st.ms_partition * 1e6);
        unsigned long long sort = part + (unsigned long long)(st.ms_sort * 1e6);
        unsigned long long mrg  = sort + (unsigned long long)(st.ms_merge * 1e6);
        unsigned long long join = mrg + (unsigned long long)(st.ms_join * 1e6);
        fprintf(stdout, "Total, Partitioning, Sort, First-Merge, Merge, Join\n");
        fprintf(stdout, "%llu, %llu, %llu, %llu, %llu, %llu\n", join, part, sort, sort, mrg, join);
        fprintf(stdout, "Perstage: ");
        fflush(stdout);
        fprintf(stderr, "%llu, %llu, %llu, %llu, %llu, ", part, sort - part, 0ull, mrg - sort,
                join - mrg);
        fflush(stderr);
        fprintf(stdout, "\n");
        fflush(stdout);
        print_timing(relS->num_tuples, &start, &end, stderr);
    }
    return res;
}

/* src/joins/sortmergejoin_multiway.c:50-61 */
result_t *sortmergejoin_multiway(relation_t *relR, relation_t *relS, joinconfig_t *joincfg) {
    if ((joincfg->NTHREADS & (joincfg->NTHREADS - 1)) != 0) {
        fprintf(stdout, "[ERROR] m-way sort-merge join runs with a power of 2 #threads.\n");
        return 0;
    }
    return join_common(SMJ_ALGO_MWAY, relR, relS, joincfg, "sortmergejoin_multiway");
}

/* src/joins/sortmergejoin_multipass.c:51-62 */
result_t *sortmergejoin_multipass(relation_t *relR, relation_t *relS, joinconfig_t *joincfg) {
    if ((joincfg->NTHREADS & (joincfg->NTHREADS - 1)) != 0) {
        fprintf(stdout, "[ERROR] m-pass sort-merge join runs with a power of 2 #threads.\n");
        return 0;
    }
    return join_common(SMJ_ALGO_MPASS, relR, relS, joincfg, "sortmergejoin_multipass");
}

/* src/joins/sortmergejoin_mpsm.c:38-45 is a stub (prints a warning and exit(0)s); this is a
 * working P-MPSM with the same signature. */
result_t *sortmergejoin_mpsm(relation_t *relR, relation_t *relS, joinconfig_t *joincfg) {
    return join_common(SMJ_ALGO_MPSM, relR, relS, joincfg, "sortmergejoin_mpsm");
}

/* JOIN_MATERIALIZE (src/joins/joincommon.c:256-278): count first, then emit into a device buffer
 * of exactly that size and bring it home */
result_t *sortmergejoin_multiway_materialize(relation_t *relR, relation_t *relS,
                                             joinconfig_t *joincfg) {
    devbuf_t  dr, ds;
    uint64_t  count = 0, written = 0;
    uint64_t *d_out = NULL;
    tuple_t  *h_out = NULL;
    int       nthreads = joincfg->NTHREADS > 0 ? joincfg->NTHREADS : 1;
    int       failed;
    if ((joincfg->NTHREADS & (joincfg->NTHREADS - 1)) != 0) {
        fprintf(stdout, "[ERROR] m-way sort-merge join runs with a power of 2 #threads.\n");
        return 0;
    }
    memset(&dr, 0, sizeof(dr));
    memset(&ds, 0, sizeof(ds));
    failed = smjdev_init(-1) || stage_in(&dr, relR->tuples, relR->num_tuples, 1) ||
             stage_in(&ds, relS->tuples, relS->num_tuples, 1) ||
             smjdev_join_materialize(dr.d, relR->num_tuples, ds.d, relS->num_tuples, SMJ_OUT_S_TUPLE, NULL, 0,
                                     &count);
    if (!failed && count) {
        d_out  = (uint64_t *)smjdev_malloc(count * sizeof(uint64_t));
        h_out  = (tuple_t *)malloc(count * sizeof(tuple_t));
        failed = !d_out || !h_out ||
                 smjdev_join_materialize(dr.d, relR->num_tuples, ds.d, relS->num_tuples, SMJ_OUT_S_TUPLE, d_out,
                                         count, &written) ||
                 written != count || smjdev_download(h_out, d_out, count * sizeof(uint64_t));
    }
    if (d_out) smjdev_free(d_out);
    stage_free(&dr);
    stage_free(&ds);
    if (failed) {
        free(h_out);
        smj_fatal("sortmergejoin_multiway_materialize");
        return NULL;
    }
    result_t *res   = (result_t *)malloc(sizeof(result_t));
    res->resultlist = (threadresult_t *)calloc((size_t)nthreads, sizeof(threadresult_t));
    for (int i = 0; i < nthreads; i++) res->resultlist[i].threadid = (uint32_t)i;
    res->resultlist[0].nresults = (int64_t)count;
    res->resultlist[0].results  = h_out;
    res->totalresults           = (int64_t)count;
    res->nthreads               = nthreads;
    return res;
}

result_t *sortmergejoin_initrun(relation_t *relR, relation_t *relS, joinconfig_t *joincfg,
                                void *(*jointhread)(void *)) {
    (void)jointhread;
    return join_common(SMJ_ALGO_MWAY, relR, relS, joincfg, "sortmergejoin_initrun");
}

static uint64_t merge_join_common(tuple_t *r, tuple_t *s, uint64_t nr, uint64_t ns, int semi,
                                  const char *name) {
    devbuf_t dr, ds;
    uint64_t count = 0, cks = 0;
    int      failed;
    if (nr == 0 || ns == 0) return 0;
    memset(&dr, 0, sizeof(dr));
    memset(&ds, 0, sizeof(ds));
    failed = stage_in(&dr, r, nr, 1) || stage_in(&ds, s, ns, 1) ||
             smjdev_merge_join_ex(dr.d, nr, ds.d, ns, semi, &count, &cks);
    stage_free(&dr);
    stage_free(&ds);
    if (failed) {
        smj_fatal(name);
        return 0;
    }
    return count;
}

/* src/joins/joincommon.c:239-312 */
uint64_t merge_join(tuple_t *rtuples, tuple_t *stuples, const uint64_t numR, const uint64_t numS,
                    void *output) {
    (void)output;
    return merge_join_common(rtuples, stuples, numR, numS, 0, "merge_join");
}

/* src/joins/joincommon.c:325-394: after the interpolated start it advances S on every match and
 * never R, i.e. it counts the S tuples that have a partner in R (== the join size when R keys
 * are unique, which the reference assumes). */
uint64_t merge_join_interpolation(tuple_t *rtuples, tuple_t *stuples, const uint64_t numR,
                                  const uint64_t numS, void *output) {
    (void)output;
    return merge_join_common(rtuples, stuples, numR, numS, 1, "merge_join_interpolation");
}

/* src/joins/joincommon.c:397-486 (8-byte tuple build): key order check on a HOST array */
int is_sorted_helper(int64_t *items, uint64_t nitems) {
    intkey_t curr   = 0;
    int      warned = 0;
    tuple_t *tuples = (tuple_t *)items;
    for (uint64_t i = 0; i < nitems; i++) {
        if (tuples[i].key == curr) {
            if (!warned) {
                printf("[WARN ] Equal items, still ok... item[%llu].key=%d is equal to "
                       "item[%lld].key=%d\n",
                       (unsigned long long)i, tuples[i].key, (long long)i - 1, curr);
                warned = 1;
            }
        } else if (tuples[i].key < curr) {
            printf("[ERROR] item[%llu].key=%d is less than item[%lld].key=%d\n",
                   (unsigned long long)i, tuples[i].key, (long long)i - 1, curr);
            return 0;
        }
        curr = tuples[i].key;
    }
    return 1;
}

/* src/joins/joincommon.c:489-501 */
void check_sorted(int64_t *R, int64_t *S, uint64_t nR, uint64_t nS, int my_tid) {
    printf("%d-thread -> R is %ssorted, size = %llu\n", my_tid, is_sorted_helper(R, nR) ? "" : "NOT ",
           (unsigned long long)nR);
    printf("%d-thread -> S is %ssorted, size = %llu\n", my_tid, is_sorted_helper(S, nS) ? "" : "NOT ",
           (unsigned long long)nS);
}

/* src/util/numa_shuffle.c:26-85.  RANDOM draws one permutation (here: a fixed LCG instead of the
 * reference's unseeded rand()); RING assumes a single NUMA region, i.e. the GPU box's NVSwitch. */
static enum numa_strategy_t g_numastrategy = NEXT;
static int                  g_shuffleorder[512];

void numa_shuffle_init(enum numa_strategy_t numastrategy, int nthreads) {
    g_numastrategy = numastrategy;
    if (numastrategy == RANDOM) {
        unsigned long long state = 88172645463325252ull;
        if (nthreads > 512) nthreads = 512;
        for (int s = 0; s < nthreads; s++) g_shuffleorder[s] = s;
        for (int i = nthreads - 1; i > 0; i--) {
            state     = state * 6364136223846793005ull + 1442695040888963407ull;
            int j     = (int)((state >> 33) % (unsigned long long)(i + 1));
            int tmp   = g_shuffleorder[i];
            g_shuffleorder[i] = g_shuffleorder[j];
            g_shuffleorder[j] = tmp;
        }
    }
}

int get_numa_shuffle_strategy(int my_tid, int nextidx, int nthreads) {
    if (g_numastrategy == RANDOM) return g_shuffleorder[nextidx];
    if (g_numastrategy == RING) return (my_tid + nthreads + nextidx) % nthreads;
    return (my_tid + nextidx) % nthreads;
}

/* ------------------------------------------------------------------------------------------
 * sort
 * ---------------------------------------------------------------------------------------- */
/* Which buffer the reference leaves the result in: a restatement of the POINTER bookkeeping only
 * of src/avxsort/avxsort_core.h:2056-2214 (avxsort_block_aligned), :2225-2324
 * (avxsort_rem_aligned), src/avxsort/avxsort.c:128-209 and avxsort_multiway.c:60-188.
 * A "side" is 0 for the buffer passed as *inputptr and 1 for the one passed as *outputptr. */
static int ilog2u(uint64_t v) {
    int l = 0;
    while (v >>= 1) l++;
    return l;
}
static int ceil_log2u(uint64_t v) {
    int l = ilog2u(v);
    return ((1ull << l) < v) ? l + 1 : l;
}
/* side holding a sorted power-of-two block that started on side 0 */
static int block_side(uint64_t bs) {
    int jend = ilog2u(bs) - 2;
    int j    = jend > 4 ? jend : 4;
    return j & 1;
}
/* pairwise merge rounds over chunk sides: merging chunk j with j+1 writes to the other side of j */
static int merge_rounds_side(int *side, uint64_t nchunks, int rounds, int until_one) {
    while ((until_one && nchunks > 1) || (!until_one && rounds-- > 0)) {
        uint64_t k = 0;
        for (uint64_t j = 0; j + 1 < nchunks; j += 2) side[k++] = !side[j];
        if (nchunks % 2) side[k++] = side[nchunks - 1];
        nchunks = k;
    }
    return side[0];
}
static int rem_side(uint32_t n) { /* avxsort_rem_aligned */
    int      side[8];
    uint32_t nxtpow = 8192, i = 0;
    while (n < nxtpow) nxtpow >>= 1;
    while (nxtpow > 128) {
        side[i++] = block_side(nxtpow);
        n -= nxtpow;
        while (n < nxtpow) nxtpow >>= 1;
    }
    if (n > 0) side[i++] = 0; /* std::sort in place on the input side */
    return merge_rounds_side(side, i, 0, 1);
}
static int avxsort_side(uint64_t nitems) { /* avxsort_aligned */
    const uint64_t BLOCKSIZE = 16384;
    uint64_t       nchunks   = nitems / BLOCKSIZE;
    uint32_t       rem       = (uint32_t)(nitems % BLOCKSIZE);
    int           *side      = (int *)malloc((nchunks + 2) * sizeof(int));
    for (uint64_t i = 0; i < nchunks; i++) side[i] = block_side(BLOCKSIZE);
    if (rem) side[nchunks++] = rem_side(rem);
    int rounds = ceil_log2u(nitems) - 14;
    int res    = merge_rounds_side(side, nchunks, rounds > 0 ? rounds : 0, 0);
    free(side);
    return res;
}
static int avxsortmultiway_side(uint64_t nitems) { /* avxsortmultiway_int64 */
    const uint64_t L3BLOCK = (20ull * 1024 * 1024) / 8;
    uint64_t       nblocks = nitems / L3BLOCK, rem = nitems % L3BLOCK;
    int           *side    = (int *)malloc((nblocks + 2) * sizeof(int));
    int            res;
    for (uint64_t i = 0; i < nblocks; i++) side[i] = avxsort_side(L3BLOCK);
    if (rem) side[nblocks++] = avxsort_side(rem);
    if (nblocks > 3) {
        if (nblocks % 2 == 1) { /* the odd last block is merged into its neighbour first */
            side[nblocks - 2] = !side[nblocks - 2];
            nblocks--;
        }
        res = !side[0]; /* multiway merge writes to the other side of block 0 */
    } else {
        int rounds = ceil_log2u(nitems) - ilog2u(L3BLOCK);
        res        = merge_rounds_side(side, nblocks, rounds > 0 ? rounds : 0, 0);
    }
    free(side);
    return res;
}
int smj_avxsort_placement(uint64_t nitems, int multiway) {
    if (nitems == 0) return 1; /* early return: pointers untouched, *outputptr is still "output" */
    return multiway ? avxsortmultiway_side(nitems) : avxsort_side(nitems);
}

/* placement: 0/1 = force that side for host buffers */
static void sort_common(uint64_t **inputptr, uint64_t **outputptr, uint64_t n, int key_only,
                        int placement, const char *name) {
    uint64_t *in = *inputptr, *out = *outputptr, *sorted = NULL;
    if (n == 0) return;
    if (smjdev_init(-1)) {
        smj_fatal(name);
        return;
    }
    if (is_device_ptr(in) && is_device_ptr(out)) {
        if (smjdev_sort(in, out, n, key_only, 0, &sorted) || smjdev_sync()) {
            smj_fatal(name);
            return;
        }
        *outputptr = sorted;
        *inputptr  = (sorted == in) ? out : in;
        return;
    }
    if (is_device_ptr(in) || is_device_ptr(out)) {
        fprintf(stdout, "[ERROR] %s: input and output must both be host or both be device\n", name);
        if (!getenv("SMJ_NO_EXIT")) exit(EXIT_FAILURE);
        return;
    }
    uint64_t *da = (uint64_t *)smjdev_malloc(n * 8), *db = (uint64_t *)smjdev_malloc(n * 8);
    uint64_t *dst = placement ? out : in;
    int failed = !da || !db || smjdev_upload(da, in, n * 8) ||
                 smjdev_sort(da, db, n, key_only, 0, &sorted) || smjdev_download(dst, sorted, n * 8);
    smjdev_free(da);
    smjdev_free(db);
    if (failed) {
        smj_fatal(name);
        return;
    }
    *outputptr = dst;
    *inputptr  = placement ? in : out;
}

/* src/avxsort/avxsort.c:212-245 */
void avxsort_tuples(tuple_t **inputptr, tuple_t **outputptr, uint64_t nitems) {
    sort_common((uint64_t **)inputptr, (uint64_t **)outputptr, nitems, 0,
                smj_avxsort_placement(nitems, 0), "avxsort_tuples");
}
void avxsort_int64(int64_t **inputptr, int64_t **outputptr, uint64_t nitems) {
    sort_common((uint64_t **)inputptr, (uint64_t **)outputptr, nitems, 0,
                smj_avxsort_placement(nitems, 0), "avxsort_int64");
}
/* src/avxsort/avxsort.c:247-250: "\todo: implement" -- an empty function in the reference */
void avxsort_int32(int32_t **inputptr, int32_t **outputptr, uint64_t nitems) {
    (void)inputptr;
    (void)outputptr;
    (void)nitems;
}
/* src/avxsort/avxsort_multiway.c:60-199 */
void avxsortmultiway_tuples(tuple_t **inputptr, tuple_t **outputptr, uint64_t nitems) {
    sort_common((uint64_t **)inputptr, (uint64_t **)outputptr, nitems, 0,
                smj_avxsort_placement(nitems, 1), "avxsortmultiway_tuples");
}
void avxsortmultiway_int64(int64_t **inputptr, int64_t **outputptr, uint64_t nitems) {
    sort_common((uint64_t **)inputptr, (uint64_t **)outputptr, nitems, 0,
                smj_avxsort_placement(nitems, 1), "avxsortmultiway_int64");
}
/* src/scalarsort/scalarsort.c:41-51, 102-111: in-place sort of the input buffer on the key only
 * (order among equal keys unspecified there; here it is the stable order), pointers swapped. */
void scalarsort_tuples(tuple_t **inputptr, tuple_t **outputptr, uint64_t nitems) {
    if (nitems == 0) {
        tuple_t *in = *inputptr;
        *inputptr   = *outputptr;
        *outputptr  = in;
        return;
    }
    sort_common((uint64_t **)inputptr, (uint64_t **)outputptr, nitems, 1, 0, "scalarsort_tuples");
}
void scalarsort_int64(int64_t **inputptr, int64_t **outputptr, uint64_t nitems) {
    if (nitems == 0) {
        int64_t *in = *inputptr;
        *inputptr   = *outputptr;
        *outputptr  = in;
        return;
    }
    sort_common((uint64_t **)inputptr, (uint64_t **)outputptr, nitems, 0, 0, "scalarsort_int64");
}
/* src/scalarsort/scalarsort.c:24-31, 91-99: sorts a plain int32 array in place */
void scalarsort_int32(int32_t **inputptr, int32_t **outputptr, uint64_t nitems) {
    int32_t *in = *inputptr;
    if (nitems) {
        if (smjdev_init(-1) || smjdev_sort_int32(in, nitems)) {
            smj_fatal("scalarsort_int32");
            return;
        }
    }
    *inputptr  = *outputptr;
    *outputptr = in;
}

/* ------------------------------------------------------------------------------------------
 * partition
 * ---------------------------------------------------------------------------------------- */
/* returns 0 on success; hist_h / offs_h: host arrays of 2^nbits uint64 */
static int partition_common(relation_t *input, relation_t *output, uint32_t nbits, uint32_t shift,
                            int padded, uint64_t *hist_h, uint64_t *offs_h) {
    const uint64_t n      = input->num_tuples;
    const uint64_t fanout = 1ull << nbits;
    const uint64_t cap    = n + (padded ? 8 * fanout : 0);
    devbuf_t       din, dout;
    uint64_t      *dtmp = NULL, *dhist = NULL, *doffs = NULL;
    int            rc   = -1;
    memset(&din, 0, sizeof(din));
    memset(&dout, 0, sizeof(dout));
    if (smjdev_init(-1)) return -1;
    if (stage_in(&din, input->tuples, n, 1) || stage_in(&dout, output->tuples, cap, 0)) goto done;
    dhist = (uint64_t *)smjdev_malloc(fanout * 8);
    doffs = (uint64_t *)smjdev_malloc(fanout * 8);
    if (nbits > 8 && n) dtmp = (uint64_t *)smjdev_malloc(n * 8);
    if (!dhist || !doffs || (nbits > 8 && n && !dtmp)) goto done;
    /* the gaps between padded partitions are never written by the kernels: define them */
    if (dout.owned && padded && smjdev_memset(dout.d, 0, cap * 8)) goto done;
    if (smjdev_partition(din.d, dout.d, dtmp, n, nbits, shift, 1, padded, dhist, doffs)) goto done;
    if (smjdev_memcpy_d2h(hist_h, dhist, fanout * 8) || smjdev_memcpy_d2h(offs_h, doffs, fanout * 8) ||
        smjdev_sync())
        goto done;
    if (dout.owned && n) {
        uint64_t used = offs_h[fanout - 1] + hist_h[fanout - 1];
        if (smjdev_download(output->tuples, dout.d, used * 8)) goto done;
    }
    rc = 0;
done:
    stage_free(&din);
    stage_free(&dout);
    smjdev_free(dtmp);
    smjdev_free(dhist);
    smjdev_free(doffs);
    return rc;
}

static void partition_api(relation_t **partitions, relation_t *input, relation_t *output,
                          uint32_t nbits, uint32_t shift, int padded, const char *name) {
    const uint64_t fanout = 1ull << nbits;
    uint64_t      *hist   = (uint64_t *)malloc(fanout * 8), *offs = (uint64_t *)malloc(fanout * 8);
    if (partition_common(input, output, nbits, shift, padded, hist, offs)) {
        free(hist);
        free(offs);
        smj_fatal(name);
        return;
    }
    for (uint64_t i = 0; i < fanout; i++) {
        partitions[i]->num_tuples = hist[i];
        partitions[i]->tuples     = output->tuples + offs[i];
    }
    free(hist);
    free(offs);
}

/* src/partition/partition.c:94-149: dense, hist[] (caller zeroed) receives the counts */
void radix_cluster(relation_t *outRel, relation_t *inRel, int32_t *hist, int R, int D) {
    const uint64_t fanout = 1ull << D;
    uint64_t      *h = (uint64_t *)malloc(fanout * 8), *o = (uint64_t *)malloc(fanout * 8);
    if (partition_common(inRel, outRel, (uint32_t)D, (uint32_t)R, 0, h, o)) {
        free(h);
        free(o);
        smj_fatal("radix_cluster");
        return;
    }
    for (uint64_t i = 0; i < fanout; i++) hist[i] += (int32_t)h[i];
    free(h);
    free(o);
}
/* src/partition/partition.c:302-327 */
void partition_relation(relation_t **partitions, relation_t *input, relation_t *output,
                        int radixbits, int shiftbits) {
    partition_api(partitions, input, output, (uint32_t)radixbits, (uint32_t)shiftbits, 0,
                  "partition_relation");
}
/* src/partition/partition.c:330-354 */
void partition_relation_optimized(relation_t **partitions, relation_t *input, relation_t *output,
                                  uint32_t nbits, uint32_t shiftbits) {
    partition_api(partitions, input, output, nbits, shiftbits, 1, "partition_relation_optimized");
}
/* src/partition/partition.c:357-385: same result as the variant above (tests/check_partitioning.c
 * :158-177 asserts exactly that); the V2 buffer-slot trick is a CPU cache detail */
void partition_relation_optimized_V2(relation_t **partitions, relation_t *input,
                                     relation_t *output, uint32_t nbits, uint32_t shiftbits) {
    partition_api(partitions, input, output, nbits, shiftbits, 1,
                  "partition_relation_optimized_V2");
}
/* src/partition/partition.c:388-436: histogram of the low nbits + plain copy; partitions[] is
 * not filled in by the reference either */
void histogram_memcpy_bench(relation_t **partitions, relation_t *input, relation_t *output,
                            uint32_t nbits) {
    const uint64_t n = input->num_tuples;
    devbuf_t       din, dout;
    uint64_t      *dhist = NULL;
    int            failed;
    (void)partitions;
    memset(&din, 0, sizeof(din));
    memset(&dout, 0, sizeof(dout));
    failed = smjdev_init(-1) || stage_in(&din, input->tuples, n, 1) ||
             stage_in(&dout, output->tuples, n, 0) ||
             !(dhist = (uint64_t *)smjdev_malloc((1ull << nbits) * 8)) ||
             smjdev_histogram(din.d, n, nbits, 0, 1, dhist) ||
             (n && smjdev_memcpy_d2d(dout.d, din.d, n * 8)) ||
             (dout.owned && n && smjdev_download(output->tuples, dout.d, n * 8)) || smjdev_sync();
    stage_free(&din);
    stage_free(&dout);
    smjdev_free(dhist);
    if (failed) smj_fatal("histogram_memcpy_bench");
}

/* ------------------------------------------------------------------------------------------
 * merge
 * ---------------------------------------------------------------------------------------- */
static uint64_t merge_common(const void *a, const void *b, void *out, uint64_t lenA, uint64_t lenB,
                             int mode, const char *name) {
    devbuf_t da, db, dout;
    int      failed;
    if (lenA + lenB == 0) return 0;
    memset(&da, 0, sizeof(da));
    memset(&db, 0, sizeof(db));
    memset(&dout, 0, sizeof(dout));
    failed = smjdev_init(-1) || stage_in(&da, a, lenA, 1) || stage_in(&db, b, lenB, 1) ||
             stage_in(&dout, out, lenA + lenB, 0) ||
             smjdev_merge(da.d, lenA, db.d, lenB, dout.d, mode) ||
             (dout.owned && smjdev_download(out, dout.d, (lenA + lenB) * 8)) || smjdev_sync();
    stage_free(&da);
    stage_free(&db);
    stage_free(&dout);
    if (failed) {
        smj_fatal(name);
        return 0;
    }
    return lenA + lenB;
}

/* src/merge/merge.c:106-171, 175-235: whole-word compare */
uint64_t avx_merge_tuples(tuple_t *const inA, tuple_t *const inB, tuple_t *const outp,
                          const uint64_t lenA, const uint64_t lenB) {
    return merge_common(inA, inB, outp, lenA, lenB, 0, "avx_merge_tuples");
}
uint64_t avx_merge_int64(int64_t *const inpA, int64_t *const inpB, int64_t *const out,
                         const uint64_t lenA, const uint64_t lenB) {
    return merge_common(inpA, inpB, out, lenA, lenB, 0, "avx_merge_int64");
}
/* src/merge/merge.c:28-64: whole-word compare (ties are bit-identical words) */
uint64_t scalar_merge_int64(int64_t *const inpA, int64_t *const inpB, int64_t *const out,
                            const uint64_t lenA, const uint64_t lenB) {
    return merge_common(inpA, inpB, out, lenA, lenB, 0, "scalar_merge_int64");
}
/* src/merge/merge.c:66-103: key-only compare, B first on ties */
uint64_t scalar_merge_tuples(tuple_t *const inpA, tuple_t *const inpB, tuple_t *const out,
                             const uint64_t lenA, const uint64_t lenB) {
    return merge_common(inpA, inpB, out, lenA, lenB, 1, "scalar_merge_tuples");
}

static uint64_t multiway_common(tuple_t *output, relation_t **parts, uint32_t nparts, int mode,
                                const char *name) {
    uint64_t         total = 0;
    devbuf_t        *runs  = (devbuf_t *)calloc(nparts ? nparts : 1, sizeof(devbuf_t));
    const uint64_t **ptrs  = (const uint64_t **)calloc(nparts ? nparts : 1, sizeof(uint64_t *));
    uint64_t        *lens  = (uint64_t *)calloc(nparts ? nparts : 1, sizeof(uint64_t));
    devbuf_t         dout  = {NULL, 0};
    uint64_t        *dtmp  = NULL;
    int              fail  = smjdev_init(-1);
    for (uint32_t i = 0; i < nparts && !fail; i++) {
        lens[i] = parts[i]->num_tuples;
        total += lens[i];
        fail    = stage_in(&runs[i], parts[i]->tuples, lens[i], 1);
        ptrs[i] = runs[i].d;
    }
    if (!fail && total) {
        fail = stage_in(&dout, output, total, 0);
        if (!fail) fail = !(dtmp = (uint64_t *)smjdev_malloc(total * 8));
        if (!fail) fail = smjdev_multiway_merge(dout.d, dtmp, ptrs, lens, nparts, mode);
        if (!fail && dout.owned) fail = smjdev_download(output, dout.d, total * 8);
        if (!fail) fail = smjdev_sync();
    }
    stage_free(&dout);
    smjdev_free(dtmp);
    for (uint32_t i = 0; i < nparts; i++) stage_free(&runs[i]);
    free(runs);
    free(ptrs);
    free(lens);
    if (fail) {
        smj_fatal(name);
        return 0;
    }
    /* the reference consumes its inputs (src/merge/avx_multiwaymerge.c:268-272) */
    for (uint32_t i = 0; i < nparts; i++) {
        parts[i]->tuples += parts[i]->num_tuples;
        parts[i]->num_tuples = 0;
    }
    return total;
}

/* src/merge/avx_multiwaymerge.c:199-338 */
uint64_t avx_multiway_merge(tuple_t *output, relation_t **parts, uint32_t nparts,
                            tuple_t *fifobuffer, uint32_t bufntuples) {
    (void)fifobuffer;
    (void)bufntuples;
    return multiway_common(output, parts, nparts, 0, "avx_multiway_merge");
}
/* src/merge/scalar_multiwaymerge.c:131, 263, 395: key-only compare */
uint64_t scalar_multiway_merge(tuple_t *output, relation_t **parts, uint32_t nparts,
                               tuple_t *fifobuffer, uint32_t bufntuples) {
    (void)fifobuffer;
    (void)bufntuples;
    return multiway_common(output, parts, nparts, 1, "scalar_multiway_merge");
}
uint64_t scalar_multiway_merge_modulo(tuple_t *output, relation_t **parts, uint32_t nparts,
                                      tuple_t *fifobuffer, uint32_t bufntuples) {
    (void)fifobuffer;
    (void)bufntuples;
    return multiway_common(output, parts, nparts, 1, "scalar_multiway_merge_modulo");
}
uint64_t scalar_multiway_merge_bitand(tuple_t *output, relation_t **parts, uint32_t nparts,
                                      tuple_t *fifobuffer, uint32_t bufntuples) {
    (void)fifobuffer;
    (void)bufntuples;
    return multiway_common(output, parts, nparts, 1, "scalar_multiway_merge_bitand");
}

/* ------------------------------------------------------------------------------------------
 * KEY_8B build: 16-byte tuples {int64 payload; int64 key} (src/types.h:23-29).  Such a build of the
 * reference runs scalar sort and merge only (src/main.c:871-877); the three join entry points, the
 * key sort and merge_join are provided here under a _key8b suffix (include/smj_b200_key8b.h maps
 * the reference's names onto them when KEY_8B is defined).
 * ---------------------------------------------------------------------------------------- */
#include "../../include/smj_b200_key8b.h"

static int stage16(devbuf_t *b, const void *p, uint64_t n, int copy) {
    b->d = NULL;
    b->owned = 0;
    if (n == 0) return 0;
    if (is_device_ptr(p)) {
        b->d = (uint64_t *)p;
        return 0;
    }
    b->d = (uint64_t *)smjdev_malloc(n * 16);
    if (!b->d) return -1;
    b->owned = 1;
    return copy ? smjdev_upload(b->d, p, n * 16) : 0;
}

static result_t *join16_common(relation16_t *relR, relation16_t *relS, joinconfig_t *cfg, const char *name) {
    struct timeval   start, end;
    devbuf_t         dr, ds;
    smj_join_stats_t st;
    int              nthreads = cfg->NTHREADS > 0 ? cfg->NTHREADS : 1;
    gettimeofday(&start, NULL);
    memset(&dr, 0, sizeof(dr));
    memset(&ds, 0, sizeof(ds));
    int failed = smjdev_init(-1) || stage16(&dr, relR->tuples, relR->num_tuples, 1) ||
                 stage16(&ds, relS->tuples, relS->num_tuples, 1) ||
                 smjdev_join_key8b(dr.d, relR->num_tuples, ds.d, relS->num_tuples, 0, &st);
    stage_free(&dr);
    stage_free(&ds);
    if (failed) {
        smj_fatal(name);
        return NULL;
    }
    gettimeofday(&end, NULL);
    g_last_stats = st;
    g_have_stats = 1;
    result_t *res = (result_t *)malloc(sizeof(result_t));
    if (res) res->resultlist = (threadresult_t *)calloc((size_t)nthreads, sizeof(threadresult_t));
    if (!res || !res->resultlist) {
        free(res);
        perror("[ERROR] result_t allocation failed");
        if (!getenv("SMJ_NO_EXIT")) exit(EXIT_FAILURE);
        return NULL;
    }
    for (int i = 0; i < nthreads; i++) res->resultlist[i].threadid = (uint32_t)i;
    res->resultlist[0].nresults = (int64_t)st.matches;
    res->totalresults           = (int64_t)st.matches;
    res->nthreads               = nthreads;
    if (!quiet()) print_timing(relS->num_tuples, &start, &end, stderr);
    return res;
}

result_t *sortmergejoin_multiway_key8b(relation16_t *relR, relation16_t *relS, joinconfig_t *joincfg) {
    if ((joincfg->NTHREADS & (joincfg->NTHREADS - 1)) != 0) {
        fprintf(stdout, "[ERROR] m-way sort-merge join runs with a power of 2 #threads.\n");
        return 0;
    }
    return join16_common(relR, relS, joincfg, "sortmergejoin_multiway_key8b");
}
result_t *sortmergejoin_multipass_key8b(relation16_t *relR, relation16_t *relS, joinconfig_t *joincfg) {
    if ((joincfg->NTHREADS & (joincfg->NTHREADS - 1)) != 0) {
        fprintf(stdout, "[ERROR] m-pass sort-merge join runs with a power of 2 #threads.\n");
        return 0;
    }
    return join16_common(relR, relS, joincfg, "sortmergejoin_multipass_key8b");
}
result_t *sortmergejoin_mpsm_key8b(relation16_t *relR, relation16_t *relS, joinconfig_t *joincfg) {
    return join16_common(relR, relS, joincfg, "sortmergejoin_mpsm_key8b");
}

/* src/scalarsort/scalarsort.c:41-51 for 16-byte tuples: the INPUT buffer ends up sorted on the key,
 * the pointers are swapped */
void scalarsort_tuples_key8b(tuple16_t **inputptr, tuple16_t **outputptr, uint64_t nitems) {
    tuple16_t *in = *inputptr, *out = *outputptr;
    if (nitems) {
        devbuf_t din, dout;
        memset(&din, 0, sizeof(din));
        memset(&dout, 0, sizeof(dout));
        int dev    = is_device_ptr(in);
        int failed = smjdev_init(-1) || stage16(&din, in, nitems, 1) ||
                     !(dout.d = (uint64_t *)smjdev_malloc(nitems * 16)) ||
                     smjdev_sort_key8b(din.d, dout.d, nitems) ||
                     (dev ? smjdev_memcpy_d2d(in, dout.d, nitems * 16) || smjdev_sync()
                          : smjdev_download(in, dout.d, nitems * 16));
        stage_free(&din);
        smjdev_free(dout.d);
        if (failed) {
            smj_fatal("scalarsort_tuples_key8b");
            return;
        }
    }
    *inputptr  = out;
    *outputptr = in;
}

/* src/joins/joincommon.c:239-312 for key-sorted 16-byte tuples */
uint64_t merge_join_key8b(tuple16_t *rtuples, tuple16_t *stuples, const uint64_t numR, const uint64_t numS,
                          void *output) {
    devbuf_t         dr, ds;
    smj_join_stats_t st;
    (void)output;
    if (numR == 0 || numS == 0) return 0;
    memset(&dr, 0, sizeof(dr));
    memset(&ds, 0, sizeof(ds));
    int failed = smjdev_init(-1) || stage16(&dr, rtuples, numR, 1) || stage16(&ds, stuples, numS, 1) ||
                 smjdev_join_key8b(dr.d, numR, ds.d, numS, 1, &st);
    stage_free(&dr);
    stage_free(&ds);
    if (failed) {
        smj_fatal("merge_join_key8b");
        return 0;
    }
    g_last_stats = st;
    g_have_stats = 1;
    return st.matches;
}
