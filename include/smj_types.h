/**
 * smj_types.h -- ABI structs of the sort-merge join path.
 *
 * Layout-identical to the reference's src/types.h:51-98 (8-byte-tuple build, i.e. KEY_8B
 * undefined) and the padding macros of src/params.h:28-72, so a caller compiled against
 * the reference's headers can pass its structs straight in.
 *
 * A tuple is {int32 payload; int32 key;}.  On a little-endian machine the 64-bit word is
 * key<<32 | (uint32)payload, which is what every kernel in this library operates on
 * ("item").  Sort order everywhere is ascending on that unsigned 64-bit word, which equals
 * the reference's IEEE-double order (src/avxsort/avxcommon.h:82-83) for the supported key
 * domain: key >= 0 and key bits 20..30 not all ones (reference README:23-31).
 */
#ifndef SMJ_TYPES_H
#define SMJ_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t intkey_t; /* reference src/types.h:27 */
typedef int32_t value_t;  /* reference src/types.h:28 */

typedef struct tuple_t        tuple_t;
typedef struct relation_t     relation_t;
typedef struct result_t       result_t;
typedef struct threadresult_t threadresult_t;
typedef struct joinconfig_t   joinconfig_t;

/* reference src/types.h:51-54 */
struct tuple_t {
    value_t  payload;
    intkey_t key;
};

/* reference src/types.h:61-64.  `tuples` may be a host pointer (reference behaviour; the
 * library copies over PCIe) or a CUDA device pointer (data already resident in HBM). */
struct relation_t {
    tuple_t *tuples;
    uint64_t num_tuples;
};

/* reference src/types.h:68-72 */
struct threadresult_t {
    int64_t  nresults;
    void    *results;
    uint32_t threadid;
};

/* reference src/types.h:75-79.  malloc()ed by the join, free()d by the caller
 * (reference src/main.c:629-632). */
struct result_t {
    int64_t         totalresults;
    threadresult_t *resultlist;
    int             nthreads;
};

/* reference src/types.h:88 */
enum numa_strategy_t { RANDOM, RING, NEXT };

/* reference src/types.h:91-98 */
struct joinconfig_t {
    int NTHREADS;
    int PARTFANOUT;
    int SCALARSORT;
    int SCALARMERGE;
    int MWAYMERGEBUFFERSIZE;
    enum numa_strategy_t NUMASTRATEGY;
};

/* reference src/params.h:28-72 */
#ifndef CACHE_LINE_SIZE
#define CACHE_LINE_SIZE 64
#endif
#define TUPLESPERCACHELINE (CACHE_LINE_SIZE / sizeof(tuple_t))
#define ALIGN_NUMTUPLES(N) (((N) + TUPLESPERCACHELINE - 1) & ~(TUPLESPERCACHELINE - 1))
#define CACHELINEPADDING(PARTFANOUT) ((PARTFANOUT) * CACHE_LINE_SIZE / sizeof(tuple_t))
#define RELATION_PADDING(NTHREADS, PARTFANOUT) \
    ((NTHREADS) * CACHELINEPADDING(PARTFANOUT) * sizeof(tuple_t))

#ifdef __cplusplus
}
#endif
#endif /* SMJ_TYPES_H */
