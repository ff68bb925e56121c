/*
 * smj_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's sort-merge join path, used solely as the checker
 * in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  No
 * product code links, imports or calls anything in this directory.
 *
 * Parity pinning: the reference ships no golden vectors.  This restatement is
 * pinned against the reference itself, compiled unmodified from /root/reference by
 * oracle/Makefile (`make ref` -> oracle/_ref/libsmj_ref.so), in tests/test_oracle_vs_ref.py, and
 * against fixtures generated from that library (tests/golden/, tests/golden/make_golden.py).
 *
 * "item" = uint64 view of tuple_t {int32 payload; int32 key;} = key<<32 | (uint32)payload
 * (reference src/types.h:51-54).
 */
#ifndef SMJ_ORACLE_H
#define SMJ_ORACLE_H
#include <stdint.h>

/* sort: ascending on the 64-bit word (reference src/avxsort/avxsort.c:128-209 orders items as
 * IEEE doubles, identical for key >= 0 without the NaN exponent pattern) */
void orc_sort(uint64_t *items, uint64_t n);
/* stable sort on the key (upper 32 bits) only */
void orc_sort_key_stable(uint64_t *items, uint64_t n);
/* which buffer the reference returns the result in: 0 = *inputptr's, 1 = *outputptr's */
int orc_avxsort_placement(uint64_t n, int multiway);

/* partition: reference src/partition/partition.c:94-287.  padded: partition starts rounded up to
 * 8 tuples (:183-187).  hist/offs: 2^nbits entries.  Returns the number of output slots used. */
uint64_t orc_partition(const uint64_t *in, uint64_t n, uint32_t nbits, uint32_t shift, int padded,
                       uint64_t *out, uint64_t *hist, uint64_t *offs);

/* merge: mode 0 whole-word compare (avx_merge_*, scalar_merge_int64), mode 1 key-only compare
 * with B first on ties (scalar_merge_tuples, reference src/merge/merge.c:66-103) */
uint64_t orc_merge(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint64_t *out,
                   int mode);
/* k-way merge by pairwise merging (reference src/merge/avx_multiwaymerge.c:199-338: binary tree) */
uint64_t orc_multiway_merge(const uint64_t *const *runs, const uint64_t *lens, uint32_t nruns,
                            uint64_t *out);

/* merge join: reference src/joins/joincommon.c:239-312 restated literally (one match at a time);
 * the checksum is sum A(r)*B(s) mod 2^64 over the matches. */
void orc_merge_join(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                    uint64_t *count, uint64_t *checksum);
/* same result computed run by run (count += |R run|*|S run|) so that skewed inputs finish */
void orc_merge_join_runs(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                         uint64_t *count, uint64_t *checksum);
/* reference src/joins/joincommon.c:325-394 without the interpolated start */
uint64_t orc_merge_join_interpolation(const uint64_t *r, uint64_t nr, const uint64_t *s,
                                      uint64_t ns);

/* whole joins.  orc_join_bruteforce: sort both sides, merge join.
 * orc_join_mway: the reference's 4-phase structure (sortmergejoin_multiway.c:129-599) executed by
 * `nthreads` logical threads one after the other, fan-out `fanout`, with its bit-shift rule. */
void orc_join_bruteforce(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                         uint64_t *count, uint64_t *checksum);
int orc_join_mway(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns, int nthreads,
                  int fanout, uint64_t *count, uint64_t *checksum);

/* the join result written out: reference src/joins/joincommon.c:256-278 (JOIN_MATERIALIZE) emits
 * <S.key, S.payload> for every matching pair; layout 1 emits R.payload << 32 | S.payload instead.
 * Sorts copies of r and s, then the literal loop.  out may be NULL (count only); at most cap items
 * are written, in the loop's order.  Returns the number of matching pairs.
 * Pinned: oracle/Makefile compiles the unmodified reference with -DJOIN_MATERIALIZE (the header its
 * tree lacks, tuple_buffer.h, is supplied by oracle/shim/) and tests/test_oracle_vs_ref.py /
 * tests/golden/ref_golden_r2.npz compare the emitted tuples as multisets. */
uint64_t orc_join_materialize(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                              int layout, uint64_t *out, uint64_t cap);

/* deterministic generators (the reference's are seeded from time(), generator.c:136-139) */
void orc_gen_unique(uint64_t *out, uint64_t n, uint64_t seed);                  /* keys = perm of 1..n */
void orc_gen_fk(uint64_t *out, uint64_t n, uint32_t maxkey, uint64_t seed);     /* keys in 1..maxkey  */
/* Zipf: orc_gen_zipf_ref = src/datagen/genzipf.c:29-159 literally (glibc rand(), pow(); the caller
 * seeds with srand(); alphabet_out / lut_out may be NULL); orc_zipf_* / orc_gen_zipf = the device
 * generator's reproducible variant (smjdev_zipf_lut / smjdev_gen_zipf), bit for bit. */
void orc_gen_zipf_ref(uint64_t *out, uint32_t n, uint32_t A, double theta, uint32_t *alphabet_out,
                      double *lut_out);
double orc_zipf_weight(uint32_t i, double theta);
void orc_zipf_lut(double *lut, uint32_t A, double theta);
void orc_gen_zipf(uint64_t *out, uint64_t first, uint64_t count, uint32_t A, double theta, uint64_t seed,
                  uint64_t alphabet_seed, const double *lut);
void orc_gen_fk_from(uint64_t *out, uint64_t n, const uint64_t *src, uint64_t src_n, uint64_t seed);
/* KEY_8B build (16-byte tuples {int64 payload; int64 key}, src/types.h:23-29): stable sort on the
 * signed key; join count + checksum (64-bit analogue of the 8-byte checksum) */
void orc_sort16(void *items, uint64_t n);
void orc_join16(const void *r, uint64_t nr, const void *s, uint64_t ns, uint64_t *count, uint64_t *checksum);
uint64_t orc_cks_a(uint64_t r);
uint64_t orc_cks_b(uint64_t s);
#endif
