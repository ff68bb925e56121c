/*
 * sortmergejoins -- the reference's command line driver (src/main.c:447-906) on top of
 * libsmj_b200.so.  Same options for the parts that exist on a GPU build:
 *
 *   -a ALGO   m-way (default, src/main.c:414-420,469), m-pass, mpsm
 *   -n N      joinconfig NTHREADS: validated like the reference (power of two), number of sorted
 *             runs for m-pass / mpsm
 *   -r N -s N relation sizes (default 128000000, src/main.c:471-472)
 *   -x S -y S seeds of R and S (default 12345 / 54321, src/main.c:474-475)
 *   -f N      PARTFANOUT (accepted, must be a power of two >= NTHREADS, src/main.c:832-841)
 *   -u        S = foreign keys drawn uniformly from R's keys instead of a second permutation
 *   -d        keep the relations device resident (no host copy in the timed call)
 *
 * Relations are generated on the device with the library's deterministic generators (the
 * reference's are seeded from time(), src/datagen/generator.c:136-139), copied to ordinary host
 * memory like the reference's malloc'ed relations (unless -d), and joined through the same entry
 * point the reference calls at src/main.c:603.
 */
#define _GNU_SOURCE
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/smj_b200.h"

typedef result_t *(*join_fn)(relation_t *, relation_t *, joinconfig_t *);
static struct {
    const char *name;
    join_fn     fn;
} algos[] = {{"m-way", sortmergejoin_multiway},
             {"m-pass", sortmergejoin_multipass},
             {"mpsm", sortmergejoin_mpsm},
             {0, 0}};

static int make_relation(relation_t *rel, uint64_t n, uint64_t seed, int fk, uint32_t maxkey,
                         int device_resident) {
    uint64_t *d = (uint64_t *)smjdev_malloc(n * 8);
    if (!d) return -1;
    int rc = fk ? smjdev_gen_fk(d, 0, n, maxkey, seed) : smjdev_gen_unique(d, n, 0, n, seed);
    if (rc) return rc;
    rel->num_tuples = n;
    if (device_resident) {
        rel->tuples = (tuple_t *)d;
        return smjdev_sync();
    }
    /* like src/main.c:516-528, with the reference's padding so that either library can be used */
    rel->tuples = (tuple_t *)malloc(n * sizeof(tuple_t) + RELATION_PADDING(64, 1024));
    if (!rel->tuples) return -1;
    if (smjdev_memcpy_d2h(rel->tuples, d, n * 8) || smjdev_sync()) return -1;
    return smjdev_free(d);
}

int main(int argc, char **argv) {
    uint64_t     nr = 128000000ull, ns = 128000000ull, seed_r = 12345, seed_s = 54321;
    int          algo = 0, fk = 0, device_resident = 0, c;
    joinconfig_t cfg = {2, 128, 0, 0, 20 * 1024 * 1024, NEXT};
    while ((c = getopt(argc, argv, "a:n:r:s:x:y:f:udh")) != -1) {
        switch (c) {
            case 'a':
                for (algo = 0; algos[algo].name && strcmp(algos[algo].name, optarg); algo++) {
                }
                if (!algos[algo].name) {
                    printf("[ERROR] Join algorithm named `%s' does not exist!\n", optarg);
                    return EXIT_FAILURE;
                }
                break;
            case 'n': cfg.NTHREADS = atoi(optarg); break;
            case 'r': nr = strtoull(optarg, 0, 10); break;
            case 's': ns = strtoull(optarg, 0, 10); break;
            case 'x': seed_r = strtoull(optarg, 0, 10); break;
            case 'y': seed_s = strtoull(optarg, 0, 10); break;
            case 'f': cfg.PARTFANOUT = atoi(optarg); break;
            case 'u': fk = 1; break;
            case 'd': device_resident = 1; break;
            default:
                printf("usage: %s [-a m-way|m-pass|mpsm] [-n threads] [-r |R|] [-s |S|] [-x seedR] "
                       "[-y seedS] [-f fanout] [-u] [-d]\n", argv[0]);
                return c == 'h' ? 0 : EXIT_FAILURE;
        }
    }
    if ((cfg.PARTFANOUT & (cfg.PARTFANOUT - 1)) || cfg.PARTFANOUT < cfg.NTHREADS) {
        printf("[ERROR] Partitioning fan-out must be a power of 2 and >= number of threads.\n");
        return EXIT_FAILURE;
    }
    if (smjdev_init(-1)) {
        printf("[ERROR] %s\n", smjdev_last_error());
        return EXIT_FAILURE;
    }
    relation_t relR, relS;
    printf("[INFO ] Creating relation R with size = %.3lf MiB, #tuples = %llu : ",
           (double)sizeof(tuple_t) * nr / 1024.0 / 1024.0, (unsigned long long)nr);
    fflush(stdout);
    if (make_relation(&relR, nr, seed_r, 0, 0, device_resident)) {
        printf("[ERROR] %s\n", smjdev_last_error());
        return EXIT_FAILURE;
    }
    printf("OK \n[INFO ] Creating relation S with size = %.3lf MiB, #tuples = %llu : ",
           (double)sizeof(tuple_t) * ns / 1024.0 / 1024.0, (unsigned long long)ns);
    fflush(stdout);
    if (make_relation(&relS, ns, seed_s, fk || ns != nr, (uint32_t)nr, device_resident)) {
        printf("[ERROR] %s\n", smjdev_last_error());
        return EXIT_FAILURE;
    }
    printf("OK \n[INFO ] Running join algorithm %s ...\n", algos[algo].name);
    numa_shuffle_init(cfg.NUMASTRATEGY, cfg.NTHREADS);
    result_t *result = algos[algo].fn(&relR, &relS, &cfg);
    if (!result) return EXIT_FAILURE;
    smj_join_stats_t st;
    smj_last_join_stats(&st);
    printf("\n[INFO ] Results = %llu. DONE.\n", (unsigned long long)result->totalresults);
    printf("[INFO ] checksum = %llu, GPU time = %.3f ms, H2D = %.3f ms, kernels = %llu\n",
           (unsigned long long)st.checksum, st.ms_total, st.ms_h2d, (unsigned long long)st.launches);
    free(result->resultlist);
    free(result);
    if (device_resident) {
        smjdev_free(relR.tuples);
        smjdev_free(relS.tuples);
    } else {
        free(relR.tuples);
        free(relS.tuples);
    }
    return 0;
}
