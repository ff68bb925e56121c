/* oracle/shim/tuple_buffer_export.c -- TEST INFRASTRUCTURE ONLY (compiled as C++ with the reference).
 * Lets the Python tests read / free the chained buffers of oracle/shim/tuple_buffer.h. */
#include <string.h>

#include "tuple_buffer.h"

extern "C" uint64_t orc_shim_cb_count(void *p) { return p ? ((chainedtuplebuffer_t *)p)->ntuples : 0; }

/* copies the emitted tuples, in emission order, as 64-bit items; returns how many were copied */
extern "C" uint64_t orc_shim_cb_copy(void *p, uint64_t *out, uint64_t cap) {
    chainedtuplebuffer_t *cb = (chainedtuplebuffer_t *)p;
    uint64_t              k  = 0;
    if (!cb) return 0;
    for (shim_chunk_t *c = cb->head; c && k < cap; c = c->next) {
        uint64_t take = c->used;
        if (take > cap - k) take = cap - k;
        memcpy(out + k, c->tuples, take * sizeof(tuple_t));
        k += take;
    }
    return k;
}

extern "C" void orc_shim_cb_free(void *p) { chainedtuplebuffer_free(p); }
