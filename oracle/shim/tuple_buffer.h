/*
 * oracle/shim/tuple_buffer.h -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference's JOIN_MATERIALIZE build includes "tuple_buffer.h" (src/joins/joincommon.c:22-24,
 * src/joins/sortmergejoin_multiway.c:34-36, src/main.c:340-342), a header its tree does not ship,
 * so that build never compiled.  This is a minimal stand-in with the three functions the
 * reference's call sites use -- chainedtuplebuffer_init() (sortmergejoin_multiway.c:565),
 * cb_next_writepos() (joincommon.c:268), chainedtuplebuffer_free() (main.c:624) -- so that
 * oracle/Makefile can compile the UNMODIFIED reference sources with -DJOIN_MATERIALIZE into
 * oracle/_ref/libsmj_ref_mat.so and the tuples its merge_join emits can be frozen into
 * tests/golden/.  A chained buffer is a list of fixed-size chunks; tuples keep emission order.
 * Nothing in the product includes this file.
 */
#ifndef ORACLE_SHIM_TUPLE_BUFFER_H
#define ORACLE_SHIM_TUPLE_BUFFER_H

#include <stdint.h>
#include <stdlib.h>

#include "types.h" /* the reference's tuple_t (src/types.h:51-54) */

#define SHIM_CHUNK_TUPLES 4096

typedef struct shim_chunk_t {
    tuple_t              tuples[SHIM_CHUNK_TUPLES];
    uint32_t             used;
    struct shim_chunk_t *next;
} shim_chunk_t;

typedef struct chainedtuplebuffer_t {
    shim_chunk_t *head, *tail;
    uint64_t      ntuples;
} chainedtuplebuffer_t;

static inline chainedtuplebuffer_t *chainedtuplebuffer_init(void) {
    chainedtuplebuffer_t *cb = (chainedtuplebuffer_t *)malloc(sizeof(chainedtuplebuffer_t));
    cb->head = cb->tail = (shim_chunk_t *)malloc(sizeof(shim_chunk_t));
    cb->head->used = 0;
    cb->head->next = NULL;
    cb->ntuples    = 0;
    return cb;
}

static inline tuple_t *cb_next_writepos(chainedtuplebuffer_t *cb) {
    if (cb->tail->used == SHIM_CHUNK_TUPLES) {
        shim_chunk_t *c = (shim_chunk_t *)malloc(sizeof(shim_chunk_t));
        c->used = 0;
        c->next = NULL;
        cb->tail->next = c;
        cb->tail       = c;
    }
    cb->ntuples++;
    return &cb->tail->tuples[cb->tail->used++];
}

static inline void chainedtuplebuffer_free(void *p) {
    chainedtuplebuffer_t *cb = (chainedtuplebuffer_t *)p;
    if (!cb) return;
    for (shim_chunk_t *c = cb->head; c;) {
        shim_chunk_t *n = c->next;
        free(c);
        c = n;
    }
    free(cb);
}

#endif
