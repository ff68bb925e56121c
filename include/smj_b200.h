/**
 * smj_b200.h -- drop-in C entry points of the B200 sort-merge join engine.
 *
 * Every function below has the name, signature, argument meaning and error behaviour of the
 * reference function it replaces (cited as file:line relative to the reference tree).  The
 * reference headers carry no `extern "C"` and its fork is compiled as C++, so ITS objects export
 * mangled names; this library exports the upstream-intended plain C names, which also lets both
 * be loaded into one process (how the parity tests run them side by side).
 *
 * All work is done by CUDA kernels for sm_100a reached through include/smj_device.h.  Pointers
 * in relation_t / the sort and merge arguments may be host pointers (reference behaviour: the
 * data is copied to HBM and back) or CUDA device pointers (data already resident).  There is no
 * CPU fallback: without a usable GPU the calls print an [ERROR] line and terminate the process
 * like the reference's own fatal paths (src/joins/joincommon.c:79-82), or return 0/NULL when
 * the environment variable SMJ_NO_EXIT is set.
 */
#ifndef SMJ_B200_H
#define SMJ_B200_H

#include <stdio.h>
#include <sys/time.h>

#include "smj_device.h"
#include "smj_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ joins ------------------
 * src/joins/sortmergejoin_multiway.h:37-38   (impl sortmergejoin_multiway.c:50-61)
 * src/joins/sortmergejoin_multipass.h:39-40  (impl sortmergejoin_multipass.c:51-62)
 * src/joins/sortmergejoin_mpsm.h:32-33       (reference impl is a stub that exit(0)s,
 *                                             sortmergejoin_mpsm.c:38-45; here it is a real P-MPSM)
 * Return a malloc()ed result_t (caller frees result->resultlist and result, src/main.c:629-632);
 * NTHREADS not a power of two -> "[ERROR] ..." on stdout and return 0.  Inputs are not modified.
 * Phase statistics are printed like src/joins/joincommon.c:176-196 (times in ns instead of TSC
 * cycles) unless SMJ_QUIET is set. */
result_t *sortmergejoin_multiway(relation_t *relR, relation_t *relS, joinconfig_t *joincfg);
result_t *sortmergejoin_multipass(relation_t *relR, relation_t *relS, joinconfig_t *joincfg);
result_t *sortmergejoin_mpsm(relation_t *relR, relation_t *relS, joinconfig_t *joincfg);

/* src/joins/joincommon.h:59-61.  The reference spawns `jointhread` on NTHREADS pinned pthreads
 * over a private arg_t; a GPU build has no such threads, so `jointhread` is ignored and the
 * m-way pipeline runs. */
result_t *sortmergejoin_initrun(relation_t *relR, relation_t *relS, joinconfig_t *joincfg,
                                void *(*jointhread)(void *));

/* The reference's JOIN_MATERIALIZE build of the same entry point (src/joins/joincommon.c:256-278
 * emits one tuple <S.key, S.payload> per matching pair into per-thread chained buffers; that build
 * needs a header the reference's tree lacks, tuple_buffer.h, so it is a separate function here
 * instead of a compile-time switch).  result->resultlist[0].results is a malloc()ed tuple_t array
 * of result->totalresults tuples in unspecified order; the caller frees it, resultlist and
 * result. */
result_t *sortmergejoin_multiway_materialize(relation_t *relR, relation_t *relS,
                                             joinconfig_t *joincfg);

/* Extension (not in the reference): statistics of the calling thread's last join, including the
 * checksum over matched (key, R-payload, S-payload) triples.  Returns 0 if there was a join. */
int smj_last_join_stats(smj_join_stats_t *out);

/* src/joins/joincommon.h:67-74 (impl joincommon.c:239-312 / :325-394).  Key-sorted inputs;
 * `output` is unused (JOIN_MATERIALIZE is broken in the reference, joincommon.c:22-24). */
uint64_t merge_join(tuple_t *rtuples, tuple_t *stuples, const uint64_t numR, const uint64_t numS,
                    void *output);
uint64_t merge_join_interpolation(tuple_t *rtuples, tuple_t *stuples, const uint64_t numR,
                                  const uint64_t numS, void *output);

/* src/joins/joincommon.h:63-65, :77-80 (host-side helpers, host pointers only) */
void print_timing(uint64_t numtuples, struct timeval *start, struct timeval *end, FILE *out);
int  is_sorted_helper(int64_t *items, uint64_t nitems);
void check_sorted(int64_t *R, int64_t *S, uint64_t nR, uint64_t nS, int my_tid);

/* src/util/numa_shuffle.h:26-43 (impl numa_shuffle.c:26-85).  Kept for API compatibility: on
 * NVSwitch every peer is one hop away, the exchange order does not matter. */
void numa_shuffle_init(enum numa_strategy_t numastrategy, int nthreads);
int  get_numa_shuffle_strategy(int my_tid, int i, int nthreads);

/* ------------------------------------------------------------------ sort -------------------
 * src/avxsort/avxsort.h:35-57 (impl avxsort.c:212-250), src/avxsort/avxsort_multiway.h:35-55
 * (impl avxsort_multiway.c:60-199), src/scalarsort/scalarsort.h:30-50 (impl scalarsort.c).
 * Two n-item buffers; on return *outputptr points at the buffer that holds the sorted items and
 * *inputptr at the other one; both buffers may be clobbered.  For host buffers the buffer chosen
 * is the one the reference would have chosen for this n (a pure function of n, restated in
 * smj_avxsort_placement). */
void avxsort_tuples(tuple_t **inputptr, tuple_t **outputptr, uint64_t nitems);
void avxsort_int64(int64_t **inputptr, int64_t **outputptr, uint64_t nitems);
void avxsort_int32(int32_t **inputptr, int32_t **outputptr, uint64_t nitems); /* reference: empty stub */
void avxsortmultiway_tuples(tuple_t **inputptr, tuple_t **outputptr, uint64_t nitems);
void avxsortmultiway_int64(int64_t **inputptr, int64_t **outputptr, uint64_t nitems);
void scalarsort_tuples(tuple_t **inputptr, tuple_t **outputptr, uint64_t nitems);
void scalarsort_int64(int64_t **inputptr, int64_t **outputptr, uint64_t nitems);
void scalarsort_int32(int32_t **inputptr, int32_t **outputptr, uint64_t nitems);

/* 1 if the reference leaves the result of avxsort_{tuples,int64} (multiway=0) or
 * avxsortmultiway_* (multiway=1) of n items in the buffer passed as *outputptr, 0 if in the one
 * passed as *inputptr. */
int smj_avxsort_placement(uint64_t nitems, int multiway);

/* ------------------------------------------------------------------ partition --------------
 * src/partition/partition.h:38-120 (impl partition.c:94-436).
 * bucket = ((key-1) & (((1<<D)-1)<<R)) >> R.  partitions[i] are caller-allocated relation_t that
 * receive {pointer into output->tuples, count}.  The _optimized variants start every partition
 * on a multiple of 8 tuples: output needs num_tuples + 8 * 2^nbits tuples of space. */
void radix_cluster(relation_t *outRel, relation_t *inRel, int32_t *hist, int R, int D);
void partition_relation(relation_t **partitions, relation_t *input, relation_t *output,
                        int radixbits, int shiftbits);
void partition_relation_optimized(relation_t **partitions, relation_t *input, relation_t *output,
                                  uint32_t nbits, uint32_t shiftbits);
void partition_relation_optimized_V2(relation_t **partitions, relation_t *input,
                                     relation_t *output, uint32_t nbits, uint32_t shiftbits);
void histogram_memcpy_bench(relation_t **partitions, relation_t *input, relation_t *output,
                            uint32_t nbits);

/* ------------------------------------------------------------------ merge ------------------
 * src/merge/merge.h:35-97 (impl merge.c:28-235): return lenA + lenB.
 * src/merge/avx_multiwaymerge.h:33-38 (impl avx_multiwaymerge.c:199-338) and
 * src/merge/scalar_multiwaymerge.h:31-78: return the number of merged tuples; parts[i] are
 * consumed (tuples advanced, num_tuples set to 0) like the reference does; fifobuffer unused. */
uint64_t avx_merge_tuples(tuple_t *const inA, tuple_t *const inB, tuple_t *const outp,
                          const uint64_t lenA, const uint64_t lenB);
uint64_t avx_merge_int64(int64_t *const inpA, int64_t *const inpB, int64_t *const out,
                         const uint64_t lenA, const uint64_t lenB);
uint64_t scalar_merge_int64(int64_t *const inpA, int64_t *const inpB, int64_t *const out,
                            const uint64_t lenA, const uint64_t lenB);
uint64_t scalar_merge_tuples(tuple_t *const inpA, tuple_t *const inpB, tuple_t *const out,
                             const uint64_t lenA, const uint64_t lenB);
uint64_t avx_multiway_merge(tuple_t *output, relation_t **parts, uint32_t nparts,
                            tuple_t *fifobuffer, uint32_t bufntuples);
uint64_t scalar_multiway_merge(tuple_t *output, relation_t **parts, uint32_t nparts,
                               tuple_t *fifobuffer, uint32_t bufntuples);
uint64_t scalar_multiway_merge_modulo(tuple_t *output, relation_t **parts, uint32_t nparts,
                                      tuple_t *fifobuffer, uint32_t bufntuples);
uint64_t scalar_multiway_merge_bitand(tuple_t *output, relation_t **parts, uint32_t nparts,
                                      tuple_t *fifobuffer, uint32_t bufntuples);

#ifdef __cplusplus
}
#endif
#endif /* SMJ_B200_H */
