/*
 * oracle/shim/dropin_prefix.h -- TEST INFRASTRUCTURE ONLY (forced include for `make -C oracle dropin`).
 *
 * The drop-in proof compiles the reference's OWN driver, src/main.c, unmodified and in place, and
 * links it against libsmj_b200.so instead of the reference's join / sort / merge objects.  The
 * reference's headers carry no `extern "C"` and the fork only compiles as C++, so
 * its main.c would look for mangled names; the library exports the upstream-intended plain C
 * names.  Declaring the five functions main.c takes from the library with C linkage BEFORE the
 * reference's headers are seen makes the later, unqualified declarations refer to the same
 * C-linkage functions ([dcl.link]: a redeclaration without a linkage specification keeps the
 * linkage of the earlier one).  This is the one-line `extern "C"` a maintainer would add to the
 * headers, applied from outside.
 */
#ifndef ORACLE_SHIM_DROPIN_PREFIX_H
#define ORACLE_SHIM_DROPIN_PREFIX_H
#include <stdio.h>
#include <sys/time.h>

#include "types.h" /* the reference's src/types.h */

#ifdef __cplusplus
extern "C" {
#endif
result_t *sortmergejoin_multiway(relation_t *relR, relation_t *relS, joinconfig_t *joincfg);  /* sortmergejoin_multiway.h:37-38 */
result_t *sortmergejoin_multipass(relation_t *relR, relation_t *relS, joinconfig_t *joincfg); /* sortmergejoin_multipass.h:39-40 */
result_t *sortmergejoin_mpsm(relation_t *relR, relation_t *relS, joinconfig_t *joincfg);      /* sortmergejoin_mpsm.h:32-33 */
void      print_timing(uint64_t numtuples, struct timeval *start, struct timeval *end, FILE *out); /* joincommon.h:63-65 */
void      numa_shuffle_init(enum numa_strategy_t numastrategy, int nthreads);                  /* numa_shuffle.h:26-33 */
#ifdef __cplusplus
}
#endif
#endif
