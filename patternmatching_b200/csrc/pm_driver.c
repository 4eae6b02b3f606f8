/*
 * pm_driver.c -- measurement driver with the reference's command line and CSV schema, running the
 * GPU matchers through the C-ABI (include/pm_b200.h).
 *
 *   pm_driver -d FILE [-d FILE ...] -s FILE [-s FILE ...] -o FILE [-v]
 *
 * What it mirrors (yehonatan145/PatternMatching): the flags of parse_arguments (Core/src/parser.c:104-162,
 * Core/src/util.c:6-13), the flow of main (Core/src/main.c:7-24): build the structures from the merged
 * dictionaries, run every algorithm over every stream with `reset` per stream file, time only the matching
 * (measure.c:290-297), classify every position against a separate "reliable" exact instance
 * (measure.c:174-190, 300-303) and write one CSV row per algorithm with the reference's first six columns
 * (measure.c:352-364, 366-396).  The perf_event columns are replaced by throughput columns.
 * Reference bugs not reproduced: the pointer arrays are sized in pointers, the output file is
 * created with a mode and truncated.
 */
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../../include/pm_b200.h"

static int verbose = 0;
static const char* program_name = "pm_driver";

static void usage(void) {
    fprintf(stderr, "Usage: %s [OPTION]...\n", program_name);
    fprintf(stderr, "options:\n");
    fprintf(stderr, "  -d FILE               use FILE as one of the dictionary files (can be used many times).\n");
    fprintf(stderr, "  -s FILE               use FILE as one of the stream files (can be used many times).\n");
    fprintf(stderr, "  -o FILE               set FILE to be the output file.\n");
    fprintf(stderr, "  -v                    set verbose to true (print more information)\n");
}
static void fatal(const char* what) {
    fprintf(stderr, "%s: %s\n", what, pm_last_error());
    exit(EXIT_FAILURE);
}
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct { const char* name; int algo; } Algo;
static const Algo ALGOS[] = {
    {"B200 suffix-trie scan", PM_ALGO_SFX},
    {"B200 Aho-Corasick DFA", PM_ALGO_DFA},
    {"B200 Karp-Rabin stages", PM_ALGO_KR},
};
enum { N_ALGOS = 3, RELIABLE = PM_ALGO_DFA };
#define CHUNK ((size_t)64 << 20)

int main(int argc, char* argv[]) {
    program_name = argv[0];
    int opt, n_dict = 0, n_stream = 0, n_out = 0;
    opterr = 0;
    while ((opt = getopt(argc, argv, "d:s:o:v")) != -1) {
        if (opt == 'd') ++n_dict;
        else if (opt == 's') ++n_stream;
        else if (opt == 'o') ++n_out;
        else if (opt == 'v') verbose = 1;
        else {
            if (optopt == 'd' || optopt == 's' || optopt == 'o') fprintf(stderr, "Option -%c must have argument.\n\n", optopt);
            else fprintf(stderr, "Unknown option -%c.\n\n", optopt);
            usage();
            return EXIT_FAILURE;
        }
    }
    if (n_out != 1 || n_dict == 0 || n_stream == 0) {
        if (n_out > 1) fprintf(stderr, "Error: have more than one output file\n\n");
        usage();
        return EXIT_FAILURE;
    }
    char** dicts = (char**)calloc((size_t)n_dict, sizeof(char*));
    char** streams = (char**)calloc((size_t)n_stream, sizeof(char*));
    char* out_name = NULL;
    int di = 0, si = 0;
    optind = 1;
    while ((opt = getopt(argc, argv, "d:s:o:v")) != -1) {
        if (opt == 'd') dicts[di++] = optarg;
        else if (opt == 's') streams[si++] = optarg;
        else if (opt == 'o') out_name = optarg;
    }

    /* init_mps (mps.c:109-113): ingest the merged dictionaries, compile, upload */
    double t0 = now_s();
    pm_dict* dict = pm_dict_create();
    for (int i = 0; i < n_dict; ++i)
        if (pm_dict_add_file(dict, dicts[i])) {
            fprintf(stderr, "failed to open dictionary file %s: %s", dicts[i], pm_last_error());
            return EXIT_FAILURE;
        }
    if (pm_dict_compile(dict)) fatal("pm_dict_compile");
    const char* dev = getenv("PM_B200_DEVICE");
    pm_engine* eng = pm_engine_create(dict, dev ? atoi(dev) : 0);
    if (!eng) fatal("pm_engine_create");
    pm_dict_info info;
    pm_dict_get_info(dict, &info);
    if (verbose)
        printf("dictionaries: %u unique patterns (%llu lines, %llu rejected, %llu duplicates), %u AC states, built in %.2f s\n",
               info.n_patterns, (unsigned long long)info.n_lines, (unsigned long long)info.n_rejected,
               (unsigned long long)info.n_duplicates, info.n_ac_states, now_s() - t0);
    uint32_t* parent = (uint32_t*)calloc((size_t)info.n_patterns + 1, sizeof(uint32_t));
    for (uint32_t pid = 1; pid <= info.n_patterns; ++pid) pm_dict_pattern(dict, pid, NULL, NULL, NULL, &parent[pid], NULL, NULL);

    uint8_t* buf = (uint8_t*)pm_host_alloc(CHUNK);
    uint16_t* res = (uint16_t*)pm_host_alloc(CHUNK * 2);
    uint16_t* real = (uint16_t*)pm_host_alloc(CHUNK * 2);
    if (!buf || !res || !real) fatal("pm_host_alloc");

    double secs[N_ALGOS] = {0};
    uint64_t cnt[N_ALGOS][4];
    uint64_t bytes_total[N_ALGOS] = {0};
    size_t mem[N_ALGOS] = {0};
    memset(cnt, 0, sizeof(cnt));
    /* reliable pass results are recomputed per chunk: keep a second engine state?  The engine carries ONE
     * stream state, so each algorithm is run over the whole stream file with the reliable results streamed
     * from a second engine (== the separate "reliable" instance of mps.c:52-53). */
    pm_engine* reliable = pm_engine_create(dict, dev ? atoi(dev) : 0);
    if (!reliable) fatal("pm_engine_create");

    for (int a = 0; a < N_ALGOS; ++a) {
        if (verbose) { printf("Measuring algorithm %s...", ALGOS[a].name); fflush(stdout); }
        for (int s = 0; s < n_stream; ++s) {
            int fd = open(streams[s], O_RDONLY);
            if (fd == -1) {
                fprintf(stderr, "can't open stream file %s: %s\n", streams[s], strerror(errno));
                return EXIT_FAILURE;
            }
            pm_engine_reset(eng);        /* measure.c:274-275: reset both before every stream file */
            pm_engine_reset(reliable);
            for (;;) {
                size_t got = 0;
                while (got < CHUNK) {
                    ssize_t r = read(fd, buf + got, CHUNK - got);
                    if (r < 0) { fprintf(stderr, "can't read from stream file %s: %s\n", streams[s], strerror(errno)); return EXIT_FAILURE; }
                    if (r == 0) break;
                    got += (size_t)r;
                }
                if (!got) break;
                double b = now_s();
                if (pm_engine_scan_host(eng, ALGOS[a].algo, buf, got, res)) fatal("pm_engine_scan_host");
                secs[a] += now_s() - b;
                if (pm_engine_scan_host(reliable, RELIABLE, buf, got, real)) fatal("pm_engine_scan_host");
                for (size_t j = 0; j < got; ++j) {          /* measure.c:174-190 */
                    const uint32_t x = res[j], y = real[j];
                    if (x == y) { cnt[a][0]++; continue; }
                    uint32_t c = y;
                    while (c && c != x) c = parent[c];
                    if (x && c == x) cnt[a][1]++;             /* algo is an ancestor of real: partial */
                    else if (!x) cnt[a][2]++;                 /* false negative */
                    else cnt[a][3]++;                         /* false positive */
                }
                bytes_total[a] += got;
                if (got < CHUNK) break;
            }
            close(fd);
        }
        mem[a] = pm_engine_total_mem(eng);
        if (verbose) printf("Done\n");
    }

    if (verbose) printf("opening file %s to write results\n", out_name);
    FILE* f = fopen(out_name, "w");
    if (!f) {
        fprintf(stderr, "failed to open results file %s: %s\n", out_name, strerror(errno));
        return EXIT_FAILURE;
    }
    fprintf(f, "Algorithm,Time (in secs),Total Memory Used,False Positive Rate,False Negative Rate,Partial Success Rate,Stream Bytes,GB/s (host buffers)");
    for (int a = 0; a < N_ALGOS; ++a) {
        const long double sum = (long double)(cnt[a][0] + cnt[a][1] + cnt[a][2] + cnt[a][3]);
        fprintf(f, "\n%s,%.6f,%zu,%.6Lf,%.6Lf,%.6Lf,%llu,%.3f", ALGOS[a].name, secs[a], mem[a],
                sum ? cnt[a][3] / sum : 0.0L, sum ? cnt[a][2] / sum : 0.0L, sum ? cnt[a][1] / sum : 0.0L,
                (unsigned long long)bytes_total[a], secs[a] > 0 ? bytes_total[a] / secs[a] / 1e9 : 0.0);
    }
    fclose(f);
    pm_host_free(buf); pm_host_free(res); pm_host_free(real);
    pm_engine_free(eng); pm_engine_free(reliable); pm_dict_free(dict);
    free(parent); free(dicts); free(streams);
    return 0;
}
