/*
 * smj_oracle.c -- TEST INFRASTRUCTURE ONLY (see smj_oracle.h).
 *
 * Plain-C restatement of the reference's sort-merge join path.  Written for clarity, not speed;
 * every function cites the reference code it follows.  Pinned against the compiled reference
 * (oracle/_ref/libsmj_ref.so) by tests/test_oracle_vs_ref.py and tests/golden/.
 */
#include "smj_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t key_of(uint64_t x) { return (uint32_t)(x >> 32); }

/* ------------------------------------------------------------------ checksum ------------- */
uint64_t orc_cks_a(uint64_t r) {
    return (r >> 32) * 0x9E3779B97F4A7C15ull + (r & 0xffffffffull) * 0xC2B2AE3D27D4EB4Full + 1ull;
}
uint64_t orc_cks_b(uint64_t s) {
    return (s >> 32) * 0x165667B19E3779F9ull + (s & 0xffffffffull) * 0x27D4EB2F165667C5ull + 1ull;
}

/* ------------------------------------------------------------------ merge ---------------- */
/* src/merge/merge.c:28-103 (scalar) and :106-235 (AVX: whole-word compare, avxcommon.h:186) */
uint64_t orc_merge(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb, uint64_t *out,
                   int mode) {
    uint64_t i = 0, j = 0, k = 0;
    while (i < na && j < nb) {
        int take_a = mode == 0 ? (a[i] <= b[j]) : (key_of(a[i]) < key_of(b[j]));
        out[k++]   = take_a ? a[i++] : b[j++];
    }
    while (i < na) out[k++] = a[i++];
    while (j < nb) out[k++] = b[j++];
    return k;
}

/* ------------------------------------------------------------------ sort ----------------- */
/* src/avxsort/avxsort.c:128-209: sorted blocks, then 2-way merge passes between two buffers.
 * Restated as a bottom-up merge sort; the result is a pure function of the input multiset. */
static void merge_sort(uint64_t *items, uint64_t n, int mode) {
    if (n < 2) return;
    uint64_t *tmp = (uint64_t *)malloc(n * sizeof(uint64_t));
    uint64_t *src = items, *dst = tmp;
    for (uint64_t w = 1; w < n; w *= 2) {
        for (uint64_t lo = 0; lo < n; lo += 2 * w) {
            uint64_t mid = lo + w < n ? lo + w : n;
            uint64_t hi  = lo + 2 * w < n ? lo + 2 * w : n;
            /* mode 1 must stay stable: A (the earlier run) wins ties */
            uint64_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) {
                int take_a = mode == 0 ? (src[i] <= src[j]) : (key_of(src[i]) <= key_of(src[j]));
                dst[k++]   = take_a ? src[i++] : src[j++];
            }
            while (i < mid) dst[k++] = src[i++];
            while (j < hi) dst[k++] = src[j++];
        }
        uint64_t *t = src;
        src         = dst;
        dst         = t;
    }
    if (src != items) memcpy(items, src, n * sizeof(uint64_t));
    free(tmp);
}
void orc_sort(uint64_t *items, uint64_t n) { merge_sort(items, n, 0); }
void orc_sort_key_stable(uint64_t *items, uint64_t n) { merge_sort(items, n, 1); }

/* Pointer bookkeeping of the reference sorts, simulated on fake addresses:
 * input buffer = [IN, IN+n), output buffer = [OUT, OUT+n).  Returns 1 if the final *outputptr
 * lies in the output buffer. */
#define ORC_IN 0x1000000000ll
#define ORC_OUT 0x3000000000ll
typedef long long fptr; /* fake pointer, in items */

/* src/avxsort/avxsort_core.h:2056-2214 avxsort_block_aligned */
static void sim_block(fptr *inputptr, fptr *outputptr, int block_size) {
    fptr ptrs[2];
    int  logbsz = 0;
    while ((1 << (logbsz + 1)) <= block_size) logbsz++;
    ptrs[0]  = *inputptr;
    ptrs[1]  = *outputptr;
    int jend = logbsz - 2, j;
    /* j = 2 (merge4), j = 3 (merge8), then merge16 passes */
    for (j = 4; j < jend; j++) {
    }
    int  ptridx = j & 1;
    fptr inp = ptrs[ptridx], out = ptrs[ptridx ^ 1];
    /* 1.c merges inp -> out, 1.d merges out -> inp */
    *outputptr = inp;
    *inputptr  = out;
}
static void sim_swap(fptr *a, fptr *b) {
    fptr t = *a;
    *a     = *b;
    *b     = t;
}
static void sim_merge_rounds(fptr ptrs[][2], uint64_t *nchunks_io, int until_one, int rounds) {
    uint64_t nchunks = *nchunks_io;
    while (until_one ? nchunks > 1 : rounds-- > 0) {
        uint64_t k = 0;
        for (uint64_t j = 0; j + 1 < nchunks; j += 2) {
            fptr inpA = ptrs[j][0], out = ptrs[j][1];
            ptrs[k][0] = out;
            ptrs[k][1] = inpA;
            k++;
        }
        if (nchunks % 2) {
            ptrs[k][0] = ptrs[nchunks - 1][0];
            ptrs[k][1] = ptrs[nchunks - 1][1];
            k++;
        }
        nchunks = k;
    }
    *nchunks_io = nchunks;
}
/* src/avxsort/avxsort_core.h:2225-2324 avxsort_rem_aligned */
static void sim_rem(fptr *inputptr, fptr *outputptr, uint32_t nitems) {
    fptr     inp = *inputptr, out = *outputptr;
    fptr     ptrs[8][2];
    uint32_t n = nitems, pos = 0, nxtpow = 8192;
    uint64_t i = 0;
    while (n < nxtpow) nxtpow >>= 1;
    while (nxtpow > 128) {
        ptrs[i][0] = inp + pos;
        ptrs[i][1] = out + pos;
        sim_block(&ptrs[i][0], &ptrs[i][1], (int)nxtpow);
        pos += nxtpow;
        n -= nxtpow;
        sim_swap(&ptrs[i][0], &ptrs[i][1]);
        i++;
        while (n < nxtpow) nxtpow >>= 1;
    }
    if (n > 0) {
        ptrs[i][0] = inp + pos;
        ptrs[i][1] = out + pos;
        i++;
    }
    sim_merge_rounds(ptrs, &i, 1, 0);
    *outputptr = ptrs[0][0];
    *inputptr  = ptrs[0][1];
}
/* src/avxsort/avxsort.c:128-209 avxsort_aligned */
static void sim_avxsort(fptr *inputptr, fptr *outputptr, uint64_t nitems) {
    if (nitems == 0) return;
    const uint64_t BLOCKSIZE = 16384;
    fptr           input = *inputptr, output = *outputptr;
    uint64_t       nchunks = nitems / BLOCKSIZE, i;
    int            rem     = (int)(nitems % BLOCKSIZE);
    fptr(*ptrs)[2]         = (fptr(*)[2])malloc((nchunks + 1) * sizeof(fptr[2]));
    for (i = 0; i <= nchunks; i++) {
        ptrs[i][0] = input + (fptr)(i * BLOCKSIZE);
        ptrs[i][1] = output + (fptr)(i * BLOCKSIZE);
    }
    for (i = 0; i < nchunks; i++) {
        sim_block(&ptrs[i][0], &ptrs[i][1], (int)BLOCKSIZE);
        sim_swap(&ptrs[i][0], &ptrs[i][1]);
    }
    if (rem) {
        sim_rem(&ptrs[i][0], &ptrs[i][1], (uint32_t)rem);
        sim_swap(&ptrs[i][0], &ptrs[i][1]);
    }
    nchunks += (rem > 0);
    int logn   = (int)ceil(log2((double)nitems));
    int rounds = logn - 14;
    sim_merge_rounds(ptrs, &nchunks, 0, rounds > 0 ? rounds : 0);
    *outputptr = ptrs[0][0];
    *inputptr  = ptrs[0][1];
    free(ptrs);
}
/* src/avxsort/avxsort_multiway.c:60-188 avxsortmultiway_int64 */
static void sim_avxsortmultiway(fptr *inputptr, fptr *outputptr, uint64_t nitems) {
    const uint64_t L3BLOCKSIZE = (20ull * 1024 * 1024) / 8;
    fptr           input = *inputptr, output = *outputptr;
    uint64_t       nblocks = nitems / L3BLOCKSIZE, remsize = nitems % L3BLOCKSIZE, i;
    fptr(*ptrs)[2]         = (fptr(*)[2])malloc((nblocks + 1) * sizeof(fptr[2]));
    for (i = 0; i < nblocks; i++) {
        ptrs[i][0] = input + (fptr)(i * L3BLOCKSIZE);
        ptrs[i][1] = output + (fptr)(i * L3BLOCKSIZE);
    }
    for (i = 0; i < nblocks; i++) {
        sim_avxsort(&ptrs[i][0], &ptrs[i][1], L3BLOCKSIZE);
        sim_swap(&ptrs[i][0], &ptrs[i][1]);
    }
    if (remsize) {
        ptrs[i][0] = input + (fptr)(i * L3BLOCKSIZE);
        ptrs[i][1] = output + (fptr)(i * L3BLOCKSIZE);
        sim_avxsort(&ptrs[i][0], &ptrs[i][1], remsize);
        sim_swap(&ptrs[i][0], &ptrs[i][1]);
        nblocks++;
    }
    if (nblocks > 3) {
        if (nblocks % 2 == 1) {
            uint64_t nc1 = nblocks - 2;
            fptr     inpA = ptrs[nc1][0], out = ptrs[nc1][1];
            ptrs[nc1][0] = out;
            ptrs[nc1][1] = inpA;
            nblocks--;
        }
        *outputptr = ptrs[0][1];
        *inputptr  = ptrs[0][0];
    } else {
        int logn   = (int)ceil(log2((double)nitems));
        int rounds = logn - (int)log2((double)L3BLOCKSIZE);
        sim_merge_rounds(ptrs, &nblocks, 0, rounds > 0 ? rounds : 0);
        *outputptr = ptrs[0][0];
        *inputptr  = ptrs[0][1];
    }
    free(ptrs);
}
int orc_avxsort_placement(uint64_t n, int multiway) {
    fptr in = ORC_IN, out = ORC_OUT;
    if (n == 0) return 1;
    if (multiway)
        sim_avxsortmultiway(&in, &out, n);
    else
        sim_avxsort(&in, &out, n);
    return out >= ORC_OUT;
}

/* ------------------------------------------------------------------ partition ------------ */
/* src/partition/partition.c:29 HASH_BIT_MODULO, :94-149 radix_cluster, :153-219
 * radix_cluster_optimized (the write-combining buffers do not change WHERE tuples land) */
uint64_t orc_partition(const uint64_t *in, uint64_t n, uint32_t nbits, uint32_t shift, int padded,
                       uint64_t *out, uint64_t *hist, uint64_t *offs) {
    const uint32_t fanout = 1u << nbits;
    const uint32_t M      = (fanout - 1) << shift;
    uint64_t      *dst    = (uint64_t *)malloc((size_t)fanout * sizeof(uint64_t));
    uint64_t       offset = 0;
    memset(hist, 0, (size_t)fanout * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) hist[((key_of(in[i]) - 1u) & M) >> shift]++;
    for (uint32_t p = 0; p < fanout; p++) {
        dst[p] = offs[p] = offset;
        offset += padded ? ((hist[p] + 7) & ~7ull) : hist[p];
    }
    for (uint64_t i = 0; i < n; i++) out[dst[((key_of(in[i]) - 1u) & M) >> shift]++] = in[i];
    uint64_t used = n ? offs[fanout - 1] + hist[fanout - 1] : 0;
    free(dst);
    return used;
}

/* ------------------------------------------------------------------ multiway merge ------- */
/* src/merge/avx_multiwaymerge.c:199-338: a binary tree of 2-way merges over nparts runs */
uint64_t orc_multiway_merge(const uint64_t *const *runs, const uint64_t *lens, uint32_t nruns,
                            uint64_t *out) {
    uint64_t total = 0;
    for (uint32_t i = 0; i < nruns; i++) total += lens[i];
    if (total == 0) return 0;
    uint64_t *a = (uint64_t *)malloc(total * 8), *b = (uint64_t *)malloc(total * 8);
    uint64_t *start = (uint64_t *)malloc(((size_t)nruns + 1) * 8);
    uint64_t  off   = 0;
    for (uint32_t i = 0; i < nruns; i++) {
        start[i] = off;
        if (lens[i]) memcpy(a + off, runs[i], lens[i] * 8);
        off += lens[i];
    }
    start[nruns]  = off;
    uint32_t cnt  = nruns;
    uint64_t *src = a, *dst = b;
    while (cnt > 1) {
        uint32_t k = 0;
        for (uint32_t i = 0; i < cnt; i += 2) {
            uint64_t lo = start[i], mid = start[i + 1 <= cnt ? i + 1 : cnt];
            uint64_t hi = (i + 2 <= cnt) ? start[i + 2] : start[cnt];
            if (i + 1 >= cnt) mid = hi;
            orc_merge(src + lo, mid - lo, src + mid, hi - mid, dst + lo, 0);
            start[k++] = lo;
        }
        start[k] = total;
        cnt      = k;
        uint64_t *t = src;
        src         = dst;
        dst         = t;
    }
    memcpy(out, src, total * 8);
    free(a);
    free(b);
    free(start);
    return total;
}

/* ------------------------------------------------------------------ merge join ----------- */
/* src/joins/joincommon.c:239-312, literally: nested do/while over the two equal-key runs */
void orc_merge_join(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                    uint64_t *count, uint64_t *checksum) {
    uint64_t i = 0, j = 0, matches = 0, cks = 0;
    while (i < nr && j < ns) {
        int32_t rk = (int32_t)key_of(r[i]), sk = (int32_t)key_of(s[j]);
        if (rk < sk)
            i++;
        else if (rk > sk)
            j++;
        else {
            uint64_t jj;
            do {
                jj = j;
                do {
                    matches++;
                    cks += orc_cks_a(r[i]) * orc_cks_b(s[jj]);
                    jj++;
                } while (jj < ns && key_of(r[i]) == key_of(s[jj]));
                i++;
            } while (i < nr && key_of(r[i]) == key_of(s[j]));
            j = jj;
        }
    }
    *count    = matches;
    *checksum = cks;
}

void orc_merge_join_runs(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                         uint64_t *count, uint64_t *checksum) {
    uint64_t i = 0, j = 0, matches = 0, cks = 0;
    while (i < nr && j < ns) {
        uint32_t rk = key_of(r[i]), sk = key_of(s[j]);
        if (rk < sk)
            i++;
        else if (rk > sk)
            j++;
        else {
            uint64_t i2 = i, j2 = j, sa = 0, sb = 0;
            while (i2 < nr && key_of(r[i2]) == rk) sa += orc_cks_a(r[i2++]);
            while (j2 < ns && key_of(s[j2]) == rk) sb += orc_cks_b(s[j2++]);
            matches += (i2 - i) * (j2 - j);
            cks += sa * sb;
            i = i2;
            j = j2;
        }
    }
    *count    = matches;
    *checksum = cks;
}

/* src/joins/joincommon.c:367-391: on a match only S advances */
uint64_t orc_merge_join_interpolation(const uint64_t *r, uint64_t nr, const uint64_t *s,
                                      uint64_t ns) {
    uint64_t i = 0, j = 0, matches = 0;
    while (i < nr && j < ns) {
        int32_t rk = (int32_t)key_of(r[i]), sk = (int32_t)key_of(s[j]);
        if (rk < sk)
            i++;
        else if (rk > sk)
            j++;
        else {
            matches++;
            j++;
        }
    }
    return matches;
}

/* ------------------------------------------------------------------ joins ---------------- */
void orc_join_bruteforce(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                         uint64_t *count, uint64_t *checksum) {
    uint64_t *rs = (uint64_t *)malloc((nr ? nr : 1) * 8), *ss = (uint64_t *)malloc((ns ? ns : 1) * 8);
    memcpy(rs, r, nr * 8);
    memcpy(ss, s, ns * 8);
    orc_sort(rs, nr);
    orc_sort(ss, ns);
    orc_merge_join_runs(rs, nr, ss, ns, count, checksum);
    free(rs);
    free(ss);
}

/* reference src/joins/joincommon.c:239-312 with JOIN_MATERIALIZE (:256-278): the loop emits
 * <S.key, S.payload> once per matching (r, s) pair; layout 1 emits the row-id pair instead */
uint64_t orc_join_materialize(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns,
                              int layout, uint64_t *out, uint64_t cap) {
    uint64_t *rs = (uint64_t *)malloc((nr ? nr : 1) * 8), *ss = (uint64_t *)malloc((ns ? ns : 1) * 8);
    uint64_t  i = 0, j = 0, matches = 0;
    memcpy(rs, r, nr * 8);
    memcpy(ss, s, ns * 8);
    orc_sort(rs, nr);
    orc_sort(ss, ns);
    while (i < nr && j < ns) {
        const uint32_t kr = (uint32_t)(rs[i] >> 32), ks = (uint32_t)(ss[j] >> 32);
        if (kr < ks)
            i++;
        else if (kr > ks)
            j++;
        else {
            uint64_t jj;
            do {
                jj = j;
                do {
                    if (out && matches < cap)
                        out[matches] = layout == 0 ? ss[jj]
                                                   : ((rs[i] & 0xffffffffull) << 32) | (ss[jj] & 0xffffffffull);
                    matches++;
                    jj++;
                } while (jj < ns && (uint32_t)(ss[jj] >> 32) == kr);
                i++;
            } while (i < nr && (uint32_t)(rs[i] >> 32) == ks);
            j = jj;
        }
    }
    free(rs);
    free(ss);
    return matches;
}

/* The reference's m-way pipeline, its threads run one after the other:
 *   joincommon.c:105-139           chunking of R and S over the threads
 *   sortmergejoin_multiway.c:331-386 partitioning_phase (bit-shift rule :372-376)
 *   :388-460                       sorting_phase (each partition sorted on its own)
 *   :463-556                       multiwaymerge_phase (thread t merges partitions
 *                                  [t*F/T, (t+1)*F/T) of every thread; skipped when T == 1)
 *   :559-599                       mergejoin_phase */
int orc_join_mway(const uint64_t *r, uint64_t nr, const uint64_t *s, uint64_t ns, int nthreads,
                  int fanout, uint64_t *count, uint64_t *checksum) {
    if (nthreads < 1 || (nthreads & (nthreads - 1)) || fanout < nthreads || (fanout & (fanout - 1)))
        return -1;
    const int T = nthreads, F = fanout;
    int       nradixbits = 0;
    while ((1 << nradixbits) < F) nradixbits++;
    const uint64_t *rel[2] = {r, s};
    const uint64_t  num[2] = {nr, ns};
    /* parts[rel][thread][partition] -> sorted run */
    uint64_t **runs[2];
    uint64_t  *lens[2];
    uint64_t   perthr0 = nr / T;
    for (int x = 0; x < 2; x++) {
        runs[x] = (uint64_t **)calloc((size_t)T * F, sizeof(uint64_t *));
        lens[x] = (uint64_t *)calloc((size_t)T * F, sizeof(uint64_t));
        uint64_t per = num[x] / T;
        for (int t = 0; t < T; t++) {
            uint64_t cnt = (t == T - 1) ? num[x] - (uint64_t)t * per : per;
            /* the shift is derived from the thread's R chunk size (:372) */
            uint64_t numR_thr = (t == T - 1) ? nr - (uint64_t)t * perthr0 : perthr0;
            int bitshift = (int)ceil(log2((double)(numR_thr * (uint64_t)T))) - 1;
            bitshift     = (T == 1) ? bitshift - nradixbits + 1 : bitshift - nradixbits - 1;
            if (bitshift < 0) bitshift = 0;
            uint64_t *out  = (uint64_t *)malloc((cnt + 8ull * F + 1) * 8);
            uint64_t *hist = (uint64_t *)malloc((size_t)F * 8), *offs = (uint64_t *)malloc((size_t)F * 8);
            orc_partition(rel[x] + (uint64_t)t * per, cnt, (uint32_t)nradixbits, (uint32_t)bitshift, 1,
                          out, hist, offs);
            for (int p = 0; p < F; p++) {
                uint64_t *run = (uint64_t *)malloc((hist[p] ? hist[p] : 1) * 8);
                memcpy(run, out + offs[p], hist[p] * 8);
                orc_sort(run, hist[p]);
                runs[x][(size_t)t * F + p] = run;
                lens[x][(size_t)t * F + p] = hist[p];
            }
            free(out);
            free(hist);
            free(offs);
        }
    }
    uint64_t total = 0, cks = 0;
    if (T == 1) {
        for (int p = 0; p < F; p++) {
            uint64_t c, k;
            orc_merge_join_runs(runs[0][p], lens[0][p], runs[1][p], lens[1][p], &c, &k);
            total += c;
            cks += k;
        }
    } else {
        const int perthread = F / T;
        for (int t = 0; t < T; t++) {
            uint64_t *merged[2];
            uint64_t  mlen[2];
            for (int x = 0; x < 2; x++) {
                const uint64_t **pp = (const uint64_t **)malloc((size_t)F * sizeof(uint64_t *));
                uint64_t        *ll = (uint64_t *)malloc((size_t)F * 8);
                uint32_t         f  = 0;
                uint64_t         tot = 0;
                for (int j = t * perthread; j < (t + 1) * perthread; j++)
                    for (int i = 0; i < T; i++) {
                        int tid = (t + i) % T; /* NUMA_SHUFFLE NEXT, numa_shuffle.c:83 */
                        pp[f]   = runs[x][(size_t)tid * F + j];
                        ll[f]   = lens[x][(size_t)tid * F + j];
                        tot += ll[f++];
                    }
                merged[x] = (uint64_t *)malloc((tot ? tot : 1) * 8);
                mlen[x]   = orc_multiway_merge(pp, ll, f, merged[x]);
                free(pp);
                free(ll);
            }
            uint64_t c, k;
            orc_merge_join_runs(merged[0], mlen[0], merged[1], mlen[1], &c, &k);
            total += c;
            cks += k;
            free(merged[0]);
            free(merged[1]);
        }
    }
    for (int x = 0; x < 2; x++) {
        for (size_t i = 0; i < (size_t)T * F; i++) free(runs[x][i]);
        free(runs[x]);
        free(lens[x]);
    }
    *count    = total;
    *checksum = cks;
    return 0;
}

/* ------------------------------------------------------------------ generators ----------- */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* bijection of [0,n): 4-round Feistel network on 2*half bits + cycle walking */
static uint64_t permute(uint64_t i, uint64_t n, uint64_t seed) {
    int bits = 2;
    while ((1ull << bits) < n) bits++;
    const int      half = (bits + 1) / 2;
    const uint64_t mask = (1ull << half) - 1;
    uint64_t       x    = i;
    do {
        uint64_t l = x >> half, rr = x & mask;
        for (int round = 0; round < 4; round++) {
            uint64_t t = rr;
            rr         = l ^ (mix64(rr + seed * 4 + (uint64_t)round) & mask);
            l          = t;
        }
        x = (l << half) | rr;
    } while (x >= n);
    return x;
}
/* replaces src/datagen/generator.c:83-93 random_unique_gen: keys are a permutation of 1..n,
 * payload = row index (the reference leaves it uninitialised) */
void orc_gen_unique(uint64_t *out, uint64_t n, uint64_t seed) {
    for (uint64_t i = 0; i < n; i++) out[i] = ((permute(i, n, seed) + 1) << 32) | (i & 0xffffffffull);
}
/* replaces src/datagen/generator.c:458-488 (foreign keys uniformly drawn from 1..maxkey) */
void orc_gen_fk(uint64_t *out, uint64_t n, uint32_t maxkey, uint64_t seed) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t x = ((mix64(i + (seed << 32)) % maxkey + 1) << 32) | (i & 0xffffffffull);
        if ((x & (0x7FFull << 52)) == (0x7FFull << 52)) x &= ~(1ull << 52); /* generator.c:107-121 avoid_NaN */
        out[i] = x;
    }
}
/* replaces src/datagen/generator.c:458-488 create_relation_fk_from_pk: the rows of src repeated and shuffled */
void orc_gen_fk_from(uint64_t *out, uint64_t n, const uint64_t *src, uint64_t src_n, uint64_t seed) {
    for (uint64_t i = 0; i < n; i++) out[i] = src[permute(i, n, seed) % src_n];
}

/* ------------------------------------------------------------------ Zipf generator -------- */
/* (1) orc_gen_zipf_ref: src/datagen/genzipf.c:29-159 restated LITERALLY -- glibc rand() for the
 *     alphabet shuffle (:43-51) and the draws (:121), libm pow() and one running double sum for the
 *     cumulative table (:76-90), the binary search of :126-143.  Under the same srand() it must
 *     reproduce the compiled reference's gen_zipf bit for bit (tests/test_oracle_vs_ref.py).
 * (2) orc_zipf_weight / orc_zipf_lut / orc_gen_zipf: what the DEVICE generator computes
 *     (csrc/smj_datagen.cu): the same table and search, but rows are independent (31 bits of a
 *     counter-based hash instead of the sequential rand() stream, the Feistel permutation of
 *     orc_gen_unique as alphabet) and the table is reproducible: i^-theta from IEEE add / mul / div /
 *     fma only, summed in fixed blocks of 256 / 256 / rest.  The tests check (2) against (1): table
 *     values to 1e-13, drawn key frequencies by a chi-square test. */
#include <math.h>

void orc_gen_zipf_ref(uint64_t *out, uint32_t n, uint32_t A, double theta, uint32_t *alphabet_out,
                      double *lut_out) {
    uint32_t *alphabet = (uint32_t *)malloc((size_t)A * sizeof(uint32_t));
    double   *lut      = (double *)malloc((size_t)A * sizeof(double));
    for (uint32_t i = 0; i < A; i++) alphabet[i] = i + 1;
    for (uint32_t i = A - 1; i > 0; i--) {
        uint32_t k   = (uint32_t)((unsigned long)i * rand() / RAND_MAX);
        uint32_t tmp = alphabet[i];
        alphabet[i]  = alphabet[k];
        alphabet[k]  = tmp;
    }
    double scaling = 0.0, sum = 0.0;
    for (uint32_t i = 1; i <= A; i++) scaling += 1.0 / pow(i, theta);
    for (uint32_t i = 1; i <= A; i++) {
        sum += 1.0 / pow(i, theta);
        lut[i - 1] = sum / scaling;
    }
    for (uint32_t i = 0; i < n; i++) {
        double   r = ((double)rand()) / RAND_MAX;
        uint32_t left = 0, right = A - 1, pos;
        if (lut[0] >= r)
            pos = 0;
        else {
            while (right - left > 1) {
                uint32_t m = (left + right) / 2;
                if (lut[m] < r) left = m; else right = m;
            }
            pos = right;
        }
        /* the reference writes the key only (genzipf.c:147-150); the payload half is kept as it was */
        out[i] = ((uint64_t)alphabet[pos] << 32) | (out[i] & 0xffffffffull);
    }
    if (alphabet_out) memcpy(alphabet_out, alphabet, (size_t)A * sizeof(uint32_t));
    if (lut_out) memcpy(lut_out, lut, (size_t)A * sizeof(double));
    free(lut);
    free(alphabet);
}

/* i^-theta, reproducibly (every operation below is a correctly rounded IEEE operation; the file is
 * compiled with -ffp-contract=off so that nothing else is fused) */
double orc_zipf_weight(uint32_t i, double theta) {
    const double x = (double)i;
    uint64_t     bits;
    memcpy(&bits, &x, 8);
    int      e  = (int)((bits >> 52) & 0x7ff) - 1023;
    uint64_t mb = (bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    double   m;
    memcpy(&m, &mb, 8);
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    const double z = (m + -1.0) / (m + 1.0), z2 = z * z;
    double       p = 1.0 / 23.0;
    for (int k = 21; k >= 1; k -= 2) p = fma(p, z2, 1.0 / (double)k);
    const double ln_m = (2.0 * z) * p;
    const double ln_i = fma((double)e, 0.6931471805599453, ln_m);
    const double y    = -theta * ln_i;
    const double kf   = rint(y * 1.4426950408889634);
    double       r    = fma(-kf, 0.6931471803691238, y);
    r                 = fma(-kf, 1.9082149292705877e-10, r);
    static const double inv_fact[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                        1.0 / 40320.0,     1.0 / 5040.0,     1.0 / 720.0,     1.0 / 120.0,
                                        1.0 / 24.0,        1.0 / 6.0,        0.5,             1.0,
                                        1.0};
    double q = 1.0 / 6227020800.0;
    for (int k = 0; k < 13; k++) q = fma(q, r, inv_fact[k]);
    const uint64_t pb = (uint64_t)((int)kf + 1023) << 52;
    double         p2;
    memcpy(&p2, &pb, 8);
    return q * p2;
}

#define ORC_ZB 256
void orc_zipf_lut(double *lut, uint32_t A, double theta) {
    const uint32_t nblocks = (A + ORC_ZB - 1) / ORC_ZB, nsuper = (nblocks + ORC_ZB - 1) / ORC_ZB;
    double *t1 = (double *)malloc((size_t)nblocks * 8), *t2 = (double *)malloc((size_t)(nsuper + 1) * 8);
    for (uint32_t b = 0; b < nblocks; b++) {
        double sum = 0.0;
        for (uint64_t i = (uint64_t)b * ORC_ZB; i < A && i < (uint64_t)(b + 1) * ORC_ZB; i++) {
            sum    = sum + orc_zipf_weight((uint32_t)i + 1u, theta);
            lut[i] = sum;
        }
        t1[b] = sum;
    }
    for (uint32_t sb = 0; sb < nsuper; sb++) {
        double sum = 0.0;
        for (uint64_t b = (uint64_t)sb * ORC_ZB; b < nblocks && b < (uint64_t)(sb + 1) * ORC_ZB; b++) {
            const double v = t1[b];
            t1[b]          = sum;
            sum            = sum + v;
        }
        t2[sb] = sum;
    }
    double total = 0.0;
    for (uint32_t sb = 0; sb < nsuper; sb++) {
        const double v = t2[sb];
        t2[sb]         = total;
        total          = total + v;
    }
    for (uint64_t i = 0; i < A; i++) {
        const uint64_t b = i / ORC_ZB;
        lut[i]           = ((t2[b / ORC_ZB] + t1[b]) + lut[i]) / total;
    }
    free(t1);
    free(t2);
}

/* rows [first, first + count) of the device generator's Zipf relation (smjdev_gen_zipf) */
void orc_gen_zipf(uint64_t *out, uint64_t first, uint64_t count, uint32_t A, double theta, uint64_t seed,
                  uint64_t alphabet_seed, const double *lut_in) {
    double *own = NULL;
    if (!lut_in) {
        own = (double *)malloc((size_t)A * 8);
        orc_zipf_lut(own, A, theta);
        lut_in = own;
    }
    for (uint64_t i = 0; i < count; i++) {
        const uint64_t row = first + i;
        const double   r   = (double)(mix64(row + (seed << 32)) >> 33) / 2147483647.0;
        uint32_t       left = 0, right = A - 1, pos;
        if (lut_in[0] >= r)
            pos = 0;
        else {
            while (right - left > 1) {
                uint32_t m = (uint32_t)(((uint64_t)left + right) / 2);
                if (lut_in[m] < r) left = m; else right = m;
            }
            pos = right;
        }
        out[i] = ((permute(pos, A, alphabet_seed) + 1) << 32) | (row & 0xffffffffull);
    }
    free(own);
}

/* ------------------------------------------------------------------ KEY_8B (16-byte tuples) -- */
/* The reference's KEY_8B build (src/types.h:23-29: int64 key and payload; scalar sort and merge
 * only, src/main.c:871-877): std::sort on the key (scalarsort.c:35-51) and the merge_join loop of
 * joincommon.c:239-312 over 16-byte tuples.  The checksum is the 64-bit analogue of the 8-byte one:
 * A(r) = key * 0x9E3779B97F4A7C15 + payload * 0xC2B2AE3D27D4EB4F + 1, B(s) = key * 0x165667B19E3779F9 +
 * payload * 0x27D4EB2F165667C5 + 1 (wrapping), summed over matched pairs. */
typedef struct {
    int64_t payload, key;
} orc_tup16;

static void orc_sort16_rec(orc_tup16 *a, orc_tup16 *tmp, uint64_t n) { /* stable merge sort on the signed key */
    if (n < 2) return;
    const uint64_t h = n / 2;
    orc_sort16_rec(a, tmp, h);
    orc_sort16_rec(a + h, tmp, n - h);
    uint64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, n * sizeof(orc_tup16));
}
void orc_sort16(void *items, uint64_t n) {
    orc_tup16 *tmp = (orc_tup16 *)malloc((n ? n : 1) * sizeof(orc_tup16));
    orc_sort16_rec((orc_tup16 *)items, tmp, n);
    free(tmp);
}
void orc_join16(const void *r_, uint64_t nr, const void *s_, uint64_t ns, uint64_t *count, uint64_t *checksum) {
    orc_tup16 *r = (orc_tup16 *)malloc((nr ? nr : 1) * sizeof(orc_tup16));
    orc_tup16 *s = (orc_tup16 *)malloc((ns ? ns : 1) * sizeof(orc_tup16));
    memcpy(r, r_, nr * sizeof(orc_tup16));
    memcpy(s, s_, ns * sizeof(orc_tup16));
    orc_sort16(r, nr);
    orc_sort16(s, ns);
    uint64_t i = 0, j = 0, cnt = 0, cks = 0;
    while (i < nr && j < ns) {
        if (r[i].key < s[j].key) i++;
        else if (r[i].key > s[j].key) j++;
        else {
            const int64_t k = r[i].key;
            uint64_t      sa = 0, sb = 0, cr = 0, cs = 0;
            for (; i < nr && r[i].key == k; i++, cr++)
                sa += (uint64_t)r[i].key * 0x9E3779B97F4A7C15ull + (uint64_t)r[i].payload * 0xC2B2AE3D27D4EB4Full + 1ull;
            for (; j < ns && s[j].key == k; j++, cs++)
                sb += (uint64_t)s[j].key * 0x165667B19E3779F9ull + (uint64_t)s[j].payload * 0x27D4EB2F165667C5ull + 1ull;
            cnt += cr * cs;
            cks += sa * sb;
        }
    }
    *count = cnt;
    *checksum = cks;
    free(r);
    free(s);
}
