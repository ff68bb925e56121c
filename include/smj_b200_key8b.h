/**
 * smj_b200_key8b.h -- the KEY_8B build of the drop-in entry points: 16-byte tuples.
 *
 * The reference switches to 64-bit keys and payloads at compile time (src/types.h:23-29:
 * `#ifdef KEY_8B typedef int64_t intkey_t; typedef int64_t value_t;`), which makes tuple_t
 * {int64 payload; int64 key}; such a build has no AVX kernels and must be run with
 * --scalarsort --scalarmerge (src/main.c:871-877).  One library serves both builds: the 16-byte
 * entry points carry a _key8b suffix, and a caller compiled with -DKEY_8B against the reference's
 * names includes this header, which maps them:
 *
 *     cc -DKEY_8B -include include/smj_b200_key8b.h ... -lsmj_b200
 *
 * Provided: the three joins (count + checksum through smj_last_join_stats; all three run the same
 * sort + scan pipeline, csrc/smj_key8b.cu), the key sort of src/scalarsort/scalarsort.c:41-51 and
 * merge_join (src/joins/joincommon.c:239-312).  Pointers may be host or device pointers, as for
 * the 8-byte entry points.  Not provided for 16-byte tuples: the partition and merge APIs.
 */
#ifndef SMJ_B200_KEY8B_H
#define SMJ_B200_KEY8B_H

#include "smj_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tuple16_t {
    int64_t payload;
    int64_t key;
} tuple16_t; /* src/types.h:51-54 under KEY_8B */

typedef struct relation16_t {
    tuple16_t *tuples;
    uint64_t   num_tuples;
} relation16_t; /* src/types.h:61-64 */

result_t *sortmergejoin_multiway_key8b(relation16_t *relR, relation16_t *relS, joinconfig_t *joincfg);
result_t *sortmergejoin_multipass_key8b(relation16_t *relR, relation16_t *relS, joinconfig_t *joincfg);
result_t *sortmergejoin_mpsm_key8b(relation16_t *relR, relation16_t *relS, joinconfig_t *joincfg);
void      scalarsort_tuples_key8b(tuple16_t **inputptr, tuple16_t **outputptr, uint64_t nitems);
uint64_t  merge_join_key8b(tuple16_t *rtuples, tuple16_t *stuples, const uint64_t numR, const uint64_t numS,
                           void *output);

#ifdef KEY_8B
#define sortmergejoin_multiway(r, s, c) sortmergejoin_multiway_key8b((relation16_t *)(r), (relation16_t *)(s), (c))
#define sortmergejoin_multipass(r, s, c) sortmergejoin_multipass_key8b((relation16_t *)(r), (relation16_t *)(s), (c))
#define sortmergejoin_mpsm(r, s, c) sortmergejoin_mpsm_key8b((relation16_t *)(r), (relation16_t *)(s), (c))
#define scalarsort_tuples(i, o, n) scalarsort_tuples_key8b((tuple16_t **)(i), (tuple16_t **)(o), (n))
#define merge_join(r, s, nr, ns, out) merge_join_key8b((tuple16_t *)(r), (tuple16_t *)(s), (nr), (ns), (out))
#endif

#ifdef __cplusplus
}
#endif
#endif /* SMJ_B200_KEY8B_H */
