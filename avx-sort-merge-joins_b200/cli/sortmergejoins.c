/*
 * sortmergejoins -- the reference's command line driver (src/main.c:447-906) on top of
 * libsmj_b200.so, with the reference's option table (src/main.c:716-745), so that its experiment
 * scripts (scripts/tput-scalability.sh:32,53, scalesize.sh, cardinality.sh) run unmodified:
 *
 *   -a --algo=<name>        m-way (default, src/main.c:414-420,469), m-pass, mpsm
 *   -n --nthreads=<N>       joinconfig NTHREADS: validated like the reference (power of two for
 *                           m-way / m-pass), number of sorted runs for m-pass / mpsm        [2]
 *   -r --r-size=<R>  -s --s-size=<S>   relation sizes                              [128000000]
 *   -x --r-seed=<x>  -y --s-seed=<y>   generator seeds                         [12345 / 54321]
 *   -z --skew=<z>           Zipf parameter of S (genzipf.c restated on the device)         [0.0]
 *   -f --partfanout=<F>     power of two >= nthreads (src/main.c:832-841,866-869)          [128]
 *   -S --numastrategy=<NEXT|RANDOM|RING>, -m --mwaybufsize=<bytes>, --scalarsort, --scalarmerge,
 *   --no-numa-localize, -p --perfconf=<P>, -o --perfout=<O>, --verbose, --brief
 *                           go into joinconfig_t / are accepted like the reference accepts them;
 *                           they select CPU cache / NUMA variants of the same result
 *   --non-unique            keys of R and S uniform in 1..|R| (create_relation_nonunique)
 *   --full-range            keys of R uniform in 1..INT_MAX, S = foreign keys drawn from R's rows
 *   -h --help, -v --version
 * extensions (not in the reference):
 *   -d --device-resident    generate the relations in HBM and keep them there (no H2D in the call)
 *   -u --fk                 S = foreign keys drawn uniformly from 1..|R| even when |S| == |R|
 *
 * Relations are generated on the device with the library's deterministic generators (the
 * reference's are seeded from time(), src/datagen/generator.c:136-139), copied to ordinary host
 * memory like the reference's malloc'ed relations (unless -d), and joined through the same entry
 * point the reference calls at src/main.c:603.  Output lines follow src/main.c:494-606.
 */
#define _GNU_SOURCE
#include <getopt.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>

#include "../../include/smj_b200.h"

typedef result_t *(*join_fn)(relation_t *, relation_t *, joinconfig_t *);
static struct {
    const char *name;
    join_fn     fn;
} algos[] = {{"m-way", sortmergejoin_multiway},
             {"m-pass", sortmergejoin_multipass},
             {"mpsm", sortmergejoin_mpsm},
             {0, 0}};

enum { GEN_UNIQUE, GEN_FK, GEN_ZIPF, GEN_FK_FROM_R };

static void print_help(const char *prog) {
    printf("Usage: %s [options]\n"
           "  -a --algo=<m-way|m-pass|mpsm>   join algorithm [m-way]\n"
           "  -n --nthreads=<N>   -r --r-size=<R>   -s --s-size=<S>   -x --r-seed=<x>   -y --s-seed=<y>\n"
           "  -z --skew=<z>       Zipf skew parameter for probe relation S [0.0]\n"
           "  -f --partfanout=<F> -S --numastrategy=<NEXT|RANDOM|RING> -m --mwaybufsize=<bytes>\n"
           "  --non-unique --full-range --scalarsort --scalarmerge --no-numa-localize --verbose --brief\n"
           "  -p --perfconf=<P> -o --perfout=<O> (accepted, unused)   -h --help   -v --version\n"
           "  -d --device-resident   -u --fk   (extensions of the GPU build)\n",
           prog);
}

/* rows of a relation generated on the device; rel->tuples is a device pointer (device_resident) or
 * ordinary host memory with the reference's padding (src/main.c:516-528) */
static int make_relation(relation_t *rel, uint64_t n, int kind, uint64_t seed, uint64_t maxkey, double skew,
                         uint64_t alphabet_seed, const relation_t *relR_dev, int device_resident, int nthreads,
                         int fanout) {
    uint64_t *d = (uint64_t *)smjdev_malloc(n * 8);
    int       rc;
    if (!d) return -1;
    switch (kind) {
        case GEN_FK: rc = smjdev_gen_fk(d, 0, n, (uint32_t)maxkey, seed); break;
        case GEN_ZIPF: rc = smjdev_gen_zipf(d, 0, n, (uint32_t)maxkey, skew, seed, alphabet_seed, NULL); break;
        case GEN_FK_FROM_R: rc = smjdev_gen_fk_from(d, n, 0, n, (const uint64_t *)relR_dev->tuples, relR_dev->num_tuples, seed); break;
        default: rc = smjdev_gen_unique(d, n, 0, n, seed); break;
    }
    if (rc) return rc;
    rel->num_tuples = n;
    if (device_resident) {
        rel->tuples = (tuple_t *)d;
        return smjdev_sync();
    }
    rel->tuples = (tuple_t *)malloc(n * sizeof(tuple_t) + RELATION_PADDING(nthreads, fanout));
    if (!rel->tuples) return -1;
    if (smjdev_download(rel->tuples, d, n * 8)) return -1;
    return smjdev_free(d);
}

int main(int argc, char **argv) {
    uint64_t     nr = 128000000ull, ns = 128000000ull, seed_r = 12345, seed_s = 54321;
    double       skew = 0.0;
    int          algo = 0, fk = 0, device_resident = 0, c;
    static int   verbose_flag, nonunique_flag, fullrange_flag, no_numa, scalarsort, scalarmerge;
    joinconfig_t cfg = {2, 128, 0, 0, 20 * 1024 * 1024, NEXT};
    static struct option long_options[] = {
        {"verbose", no_argument, &verbose_flag, 1},    {"brief", no_argument, &verbose_flag, 0},
        {"non-unique", no_argument, &nonunique_flag, 1}, {"full-range", no_argument, &fullrange_flag, 1},
        {"no-numa-localize", no_argument, &no_numa, 1},  {"scalarsort", no_argument, &scalarsort, 1},
        {"scalarmerge", no_argument, &scalarmerge, 1},   {"help", no_argument, 0, 'h'},
        {"version", no_argument, 0, 'v'},                {"algo", required_argument, 0, 'a'},
        {"nthreads", required_argument, 0, 'n'},         {"perfconf", required_argument, 0, 'p'},
        {"r-size", required_argument, 0, 'r'},           {"s-size", required_argument, 0, 's'},
        {"perfout", required_argument, 0, 'o'},          {"r-seed", required_argument, 0, 'x'},
        {"s-seed", required_argument, 0, 'y'},           {"skew", required_argument, 0, 'z'},
        {"partfanout", required_argument, 0, 'f'},       {"numastrategy", required_argument, 0, 'S'},
        {"mwaybufsize", required_argument, 0, 'm'},      {"device-resident", no_argument, 0, 'd'},
        {"fk", no_argument, 0, 'u'},                     {0, 0, 0, 0}};
    while (1) {
        int idx = 0;
        c = getopt_long(argc, argv, "a:n:p:r:s:o:x:y:z:hvf:m:S:du", long_options, &idx);
        if (c == -1) break;
        switch (c) {
            case 0: break; /* a flag option */
            case 'a':
                for (algo = 0; algos[algo].name && strcmp(algos[algo].name, optarg); algo++) {
                }
                if (!algos[algo].name) { /* src/main.c:776-781 */
                    printf("[ERROR] Join algorithm named `%s' does not exist!\n", optarg);
                    print_help(argv[0]);
                    exit(EXIT_SUCCESS);
                }
                break;
            case 'h':
            case '?': print_help(argv[0]); exit(EXIT_SUCCESS);
            case 'v': printf("sortmergejoins (B200 build of the sort-merge join engine)\n"); exit(EXIT_SUCCESS);
            case 'n': cfg.NTHREADS = atoi(optarg); break;
            case 'p':
            case 'o': break; /* Intel PCM configuration: no counterpart on a GPU (ncu instead) */
            case 'r': nr = strtoull(optarg, 0, 10); break;
            case 's': ns = strtoull(optarg, 0, 10); break;
            case 'x': seed_r = strtoull(optarg, 0, 10); break;
            case 'y': seed_s = strtoull(optarg, 0, 10); break;
            case 'z': skew = atof(optarg); break;
            case 'f':
                cfg.PARTFANOUT = atoi(optarg);
                if ((cfg.PARTFANOUT & (cfg.PARTFANOUT - 1)) != 0) { /* src/main.c:832-839 */
                    fprintf(stdout, "[ERROR] Partitioning fan-out must be a power of 2 (2^#radixbits).\n");
                    exit(0);
                }
                break;
            case 'm': cfg.MWAYMERGEBUFFERSIZE = atoi(optarg); break;
            case 'S':
                if (!strcmp(optarg, "NEXT")) cfg.NUMASTRATEGY = NEXT;
                else if (!strcmp(optarg, "RANDOM")) cfg.NUMASTRATEGY = RANDOM;
                else if (!strcmp(optarg, "RING")) cfg.NUMASTRATEGY = RING;
                else {
                    printf("Invalid NUMA-shuffle strategy. Options: NEXT, RANDOM, RING\n");
                    printf("Using RING as default.\n");
                }
                break;
            case 'd': device_resident = 1; break;
            case 'u': fk = 1; break;
            default: break;
        }
    }
    if (cfg.PARTFANOUT < cfg.NTHREADS) { /* src/main.c:866-869 */
        fprintf(stdout, "[ERROR] Partitioning fan-out must be >= num-threads.\n");
        exit(0);
    }
    cfg.SCALARSORT  = scalarsort;
    cfg.SCALARMERGE = scalarmerge;
    (void)verbose_flag;
    (void)no_numa;
    if (optind < argc) {
        printf("non-option arguments: ");
        while (optind < argc) printf("%s ", argv[optind++]);
        printf("\n");
    }
    if (smjdev_init(-1)) {
        printf("[ERROR] %s\n", smjdev_last_error());
        return EXIT_FAILURE;
    }
    relation_t     relR, relS, relR_dev;
    struct timeval start, end;
    printf("[INFO ] Creating relation R with size = %.3lf MiB, #tuples = %llu : ",
           (double)sizeof(tuple_t) * nr / 1024.0 / 1024.0, (unsigned long long)nr);
    fflush(stdout);
    gettimeofday(&start, NULL);
    /* src/main.c:540-551: full range -> keys in 1..INT_MAX, non-unique -> keys in 1..|R|, else a permutation */
    const int r_kind = (fullrange_flag || nonunique_flag) ? GEN_FK : GEN_UNIQUE;
    const uint64_t r_max = fullrange_flag ? (uint64_t)INT_MAX : nr;
    /* full range: S draws its keys from R's rows, so R is generated on the device first and kept there */
    if (make_relation(&relR_dev, nr, r_kind, seed_r, r_max, 0.0, 0, NULL, 1, cfg.NTHREADS, cfg.PARTFANOUT)) {
        printf("[ERROR] %s\n", smjdev_last_error());
        return EXIT_FAILURE;
    }
    if (device_resident) {
        relR = relR_dev;
    } else {
        relR.num_tuples = nr;
        relR.tuples     = (tuple_t *)malloc(nr * sizeof(tuple_t) + RELATION_PADDING(cfg.NTHREADS, cfg.PARTFANOUT));
        if (!relR.tuples || smjdev_download(relR.tuples, relR_dev.tuples, nr * 8)) {
            printf("[ERROR] %s\n", smjdev_last_error());
            return EXIT_FAILURE;
        }
    }
    gettimeofday(&end, NULL);
    printf("OK \n[INFO ] Time: ");
    print_timing(nr, &start, &end, stdout);
    printf("\n[INFO ] Creating relation S with size = %.3lf MiB, #tuples = %llu : ",
           (double)sizeof(tuple_t) * ns / 1024.0 / 1024.0, (unsigned long long)ns);
    fflush(stdout);
    gettimeofday(&start, NULL);
    /* src/main.c:560-586 */
    int s_kind;
    if (fullrange_flag) s_kind = GEN_FK_FROM_R;
    else if (nonunique_flag) s_kind = GEN_FK;
    else if (skew > 0) s_kind = GEN_ZIPF;
    else s_kind = (fk || ns != nr) ? GEN_FK : GEN_UNIQUE;
    if (make_relation(&relS, ns, s_kind, seed_s, nr, skew, seed_r, &relR_dev, device_resident, cfg.NTHREADS,
                      cfg.PARTFANOUT)) {
        printf("[ERROR] %s\n", smjdev_last_error());
        return EXIT_FAILURE;
    }
    if (!device_resident) smjdev_free(relR_dev.tuples);
    gettimeofday(&end, NULL);
    printf("OK \n[INFO ] Time: ");
    print_timing(ns, &start, &end, stdout);
    printf("\n[INFO ] Running join algorithm %s ...\n", algos[algo].name);
    numa_shuffle_init(cfg.NUMASTRATEGY, cfg.NTHREADS);
    result_t *result = algos[algo].fn(&relR, &relS, &cfg);
    if (result != NULL) {
        smj_join_stats_t st;
        smj_last_join_stats(&st);
        printf("\n[INFO ] Results = %llu. DONE.\n", (unsigned long long)result->totalresults);
        printf("[INFO ] checksum = %llu, GPU time = %.3f ms, H2D = %.3f ms, kernels = %llu\n",
               (unsigned long long)st.checksum, st.ms_total, st.ms_h2d, (unsigned long long)st.launches);
        free(result->resultlist);
        free(result);
    }
    if (device_resident) {
        smjdev_free(relR.tuples);
        smjdev_free(relS.tuples);
    } else {
        free(relR.tuples);
        free(relS.tuples);
    }
    return 0;
}
