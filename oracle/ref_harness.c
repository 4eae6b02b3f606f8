/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A thin driver around the UNMODIFIED reference sources of yehonatan145/PatternMatching,
 * compiled where they lie under /root/reference/Core/src by oracle/Makefile into
 * oracle/_ref/libpmref.so.  No reference source is copied into this repository; this file
 * only calls the reference's public entry points:
 *
 *   mps_table_setup()              Core/src/mps.c:120-124   (algorithm registry)
 *   patterns_tree_build()          Core/src/PatternsTree.c:469-475 (dict ingest, de-dup, ids)
 *   mps_table[a].create/add_pattern/compile/read_char/reset/total_mem   Core/src/mps.h:71-80
 *   is_pattern_suffix()            Core/src/PatternsTree.c:485-494
 *
 * It reproduces the reference driver's hot loop (Core/src/measure.c:292-294) and its success
 * classification (Core/src/measure.c:174-190) so that tests and bench.py's cpu_baseline /
 * `--impl reference` arm can run the real reference algorithms on the host cores.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/wait.h>

#include "conf.h" /* reference: Conf, pulls PatternsTree.h / mps.h / measure.h */

#define PMREF_MAX_ALGOS 3

typedef struct {
    uint32_t file, line, parent_file, parent_line;
    uint32_t len;
    uint64_t off; /* into g_bytes */
    pattern_id_t id;
} RefPattern;

static int g_setup_done = 0;
static void* g_obj[PMREF_MAX_ALGOS];
static int g_have[PMREF_MAX_ALGOS];
static RefPattern* g_pats = NULL;
static size_t g_npats = 0, g_cappats = 0;
static unsigned char* g_bytes = NULL;
static size_t g_nbytes = 0, g_capbytes = 0;
static size_t g_max_pat_len = 0;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* callback handed to patterns_tree_build: fan the pattern out to every selected algorithm,
 * like add_pattern_to_all_instances (Core/src/mps.c:64-77), and remember it for the tests. */
static void harness_add_pattern(void* ctx, char* pat, size_t len, pattern_id_t id) {
    (void)ctx;
    for (int a = 0; a < PMREF_MAX_ALGOS; ++a)
        if (g_have[a]) mps_table[a].add_pattern(g_obj[a], pat, len, id);
    if (g_npats == g_cappats) {
        g_cappats = g_cappats ? g_cappats * 2 : 1024;
        g_pats = (RefPattern*)realloc(g_pats, g_cappats * sizeof(RefPattern));
    }
    if (g_nbytes + len > g_capbytes) {
        while (g_nbytes + len > g_capbytes) g_capbytes = g_capbytes ? g_capbytes * 2 : (1u << 20);
        g_bytes = (unsigned char*)realloc(g_bytes, g_capbytes);
    }
    RefPattern* p = &g_pats[g_npats++];
    p->file = (uint32_t)id->pattern_id.file_number;
    p->line = (uint32_t)id->pattern_id.line_number;
    p->len = (uint32_t)len;
    p->off = g_nbytes;
    p->id = id;
    memcpy(g_bytes + g_nbytes, pat, len);
    g_nbytes += len;
}

/* algo_mask: bit0 = AC (mpac.c), bit1 = LMAC (mplmac.c), bit2 = MPBG (mpbg.c). */
int pmref_build(int n_dicts, const char** dict_paths, int algo_mask) {
    if (g_setup_done) return -1; /* the reference keeps global state and is not re-entrant */
    mps_table_setup();
    g_setup_done = 1;
    Conf conf;
    memset(&conf, 0, sizeof(conf));
    conf.n_dictionary_files = (size_t)n_dicts;
    conf.dictionary_files = (char**)malloc(sizeof(char*) * (size_t)n_dicts);
    for (int i = 0; i < n_dicts; ++i) conf.dictionary_files[i] = strdup(dict_paths[i]);
    for (int a = 0; a < PMREF_MAX_ALGOS; ++a) {
        g_have[a] = (algo_mask >> a) & 1;
        if (g_have[a]) g_obj[a] = mps_table[a].create();
    }
    patterns_tree_build(&conf, NULL, harness_add_pattern); /* return value is garbage */
    g_max_pat_len = conf.max_pat_len;
    /* parents are only final once the whole tree is converted */
    for (size_t i = 0; i < g_npats; ++i) {
        pattern_id_t par = g_pats[i].id->parent;
        g_pats[i].parent_file = par ? (uint32_t)par->pattern_id.file_number : 0xFFFFFFFFu;
        g_pats[i].parent_line = par ? (uint32_t)par->pattern_id.line_number : 0xFFFFFFFFu;
    }
    for (int a = 0; a < PMREF_MAX_ALGOS; ++a)
        if (g_have[a]) mps_table[a].compile(g_obj[a]);
    return 0;
}

/* number of entries of the reference's mps_table (a generic MpsElem host such as pm_driver -p walks the table) */
int mps_table_count(void) { return MPS_SIZE; }

size_t pmref_n_patterns(void) { return g_npats; }
size_t pmref_max_pat_len(void) { return g_max_pat_len; }
size_t pmref_total_mem(int algo) { return g_have[algo] ? mps_table[algo].total_mem(g_obj[algo]) : 0; }
const char* pmref_algo_name(int algo) { return g_setup_done ? mps_table[algo].name : ""; }

/* pattern i in the reference's add order: (file,line), parent (file,line) (0xFFFFFFFF = root), bytes */
int pmref_pattern(size_t i, uint32_t* file, uint32_t* line, uint32_t* pfile, uint32_t* pline,
                  uint32_t* len, const unsigned char** bytes) {
    if (i >= g_npats) return -1;
    *file = g_pats[i].file; *line = g_pats[i].line;
    *pfile = g_pats[i].parent_file; *pline = g_pats[i].parent_line;
    *len = g_pats[i].len; *bytes = g_bytes + g_pats[i].off;
    return 0;
}

void pmref_reset(int algo) { mps_table[algo].reset(g_obj[algo]); }

/* The reference hot loop (measure.c:292-294).  State continues from the previous call until
 * pmref_reset().  file_out/line_out may be NULL (timing only); "no match" = 0xFFFFFFFF.
 * Returns seconds spent in the read_char loop only (the region measure.c:290-297 times). */
double pmref_scan(int algo, const unsigned char* buf, size_t n, uint32_t* file_out, uint32_t* line_out) {
    pattern_id_t (*read_char_func)(void*, char) = mps_table[algo].read_char;
    void* obj = g_obj[algo];
    pattern_id_t* res = (pattern_id_t*)malloc(sizeof(pattern_id_t) * (n ? n : 1));
    double t0 = now_s();
    for (size_t j = 0; j < n; ++j) res[j] = read_char_func(obj, (char)buf[j]);
    double t1 = now_s();
    if (file_out && line_out) {
        for (size_t j = 0; j < n; ++j) {
            file_out[j] = res[j] ? (uint32_t)res[j]->pattern_id.file_number : 0xFFFFFFFFu;
            line_out[j] = res[j] ? (uint32_t)res[j]->pattern_id.line_number : 0xFFFFFFFFu;
        }
    }
    free(res);
    return t1 - t0;
}

#define FNV_OFF 1469598103934665603ULL
#define FNV_PRIME 1099511628211ULL

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* Order-independent (hence shardable / GPU-friendly) digest of one (pos, file, line) match:
 * splitmix64(pos ^ splitmix64(((file+1) << 32) | line)); digests are summed mod 2^64. */
static inline uint64_t match_digest(uint64_t pos, uint64_t file, uint64_t line) {
    return splitmix64(pos ^ splitmix64(((file + 1) << 32) | line));
}
/* hsum[0] = sum over longest matches only, hsum[1] = sum over all matches (with ancestors) */
static uint64_t g_hsum[2];
void pmref_last_hsum(uint64_t out[2]) { out[0] = g_hsum[0]; out[1] = g_hsum[1]; }

/* Summary of one scan from reset: positions with a match, matches incl. PatternsTree ancestors,
 * and the FNV-style checksum over (pos, file, line) triples defined in SURVEY.md Appendix A.
 * pos_base is added to the position (sharded scans); `skip` leading bytes are walked, not counted. */
static double scan_summary(int algo, const unsigned char* buf, size_t n, size_t skip, uint64_t pos_base,
                           uint64_t* positions, uint64_t* matches, uint64_t* checksum, int chain_hash) {
    pattern_id_t (*read_char_func)(void*, char) = mps_table[algo].read_char;
    void* obj = g_obj[algo];
    uint64_t h = chain_hash ? *checksum : FNV_OFF, np = 0, nm = 0, hs0 = 0, hs1 = 0;
    double t = 0;
    enum { CH = 1 << 16 };
    pattern_id_t* res = (pattern_id_t*)malloc(sizeof(pattern_id_t) * CH);
    for (size_t base = 0; base < n; base += CH) {
        size_t m = n - base < CH ? n - base : CH;
        double t0 = now_s();
        for (size_t j = 0; j < m; ++j) res[j] = read_char_func(obj, (char)buf[base + j]);
        t += now_s() - t0;
        for (size_t j = 0; j < m; ++j) {
            if (base + j < skip || !res[j]) continue;
            ++np;
            hs0 += match_digest(pos_base + (base + j - skip), res[j]->pattern_id.file_number, res[j]->pattern_id.line_number);
            for (pattern_id_t id = res[j]; id && id->pattern_id.file_number != (size_t)-1; id = id->parent) {
                uint64_t v[3] = { pos_base + (base + j - skip), id->pattern_id.file_number, id->pattern_id.line_number };
                for (int k = 0; k < 3; ++k) { h ^= v[k]; h *= FNV_PRIME; }
                hs1 += match_digest(v[0], v[1], v[2]);
                ++nm;
            }
        }
    }
    free(res);
    *positions = np; *matches = nm; *checksum = h;
    g_hsum[0] = hs0; g_hsum[1] = hs1;
    return t;
}

double pmref_summary(int algo, const unsigned char* buf, size_t n,
                     uint64_t* positions, uint64_t* matches, uint64_t* checksum) {
    mps_table[algo].reset(g_obj[algo]);
    return scan_summary(algo, buf, n, 0, 0, positions, matches, checksum, 0);
}

/* Success classification of `algo` against the reliable AC instance (algo 0), measure.c:174-190.
 * counts[4] = success, partial, false_neg, false_pos. */
void pmref_success(int algo, const unsigned char* buf, size_t n, uint64_t counts[4]) {
    mps_table[algo].reset(g_obj[algo]);
    mps_table[0].reset(g_obj[0]);
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    for (size_t j = 0; j < n; ++j) {
        pattern_id_t a = mps_table[algo].read_char(g_obj[algo], (char)buf[j]);
        pattern_id_t r = (algo == 0) ? a : mps_table[0].read_char(g_obj[0], (char)buf[j]);
        if (r == a) counts[0]++;
        else if (is_pattern_suffix(a, r)) counts[1]++;
        else if (a == null_pattern_id) counts[2]++;
        else counts[3]++;
    }
}

/* All-host-cores baseline: fork `workers` children after compile (the automaton is shared
 * copy-on-write); child w scans the contiguous shard [w*n/K, (w+1)*n/K) preceded by a left halo of
 * max_pat_len-1 bytes walked from reset (identical to the continuous scan).  Returns the
 * wall-clock seconds from the first fork to the last child's exit; per-shard summaries are combined
 * (positions and matches summed; shard checksums folded in shard order with FNV). */
double pmref_scan_parallel(int algo, const unsigned char* buf, size_t n, int workers,
                           uint64_t* positions, uint64_t* matches, uint64_t* checksum_fold,
                           double* max_loop_seconds) {
    if (workers < 1) workers = 1;
    typedef struct { uint64_t np, nm, h, hs0, hs1; double t; } Slot;
    Slot* slots = (Slot*)mmap(NULL, sizeof(Slot) * (size_t)workers, PROT_READ | PROT_WRITE,
                              MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (slots == MAP_FAILED) return -1.0;
    size_t halo = g_max_pat_len ? g_max_pat_len - 1 : 0;
    fflush(NULL);
    double t0 = now_s();
    for (int w = 0; w < workers; ++w) {
        pid_t pid = fork();
        if (pid == 0) {
            size_t lo = n * (size_t)w / (size_t)workers, hi = n * (size_t)(w + 1) / (size_t)workers;
            size_t start = lo > halo ? lo - halo : 0;
            mps_table[algo].reset(g_obj[algo]);
            Slot s; s.h = 0;
            s.t = scan_summary(algo, buf + start, hi - start, lo - start, lo, &s.np, &s.nm, &s.h, 0);
            s.hs0 = g_hsum[0]; s.hs1 = g_hsum[1];
            slots[w] = s;
            _exit(0);
        } else if (pid < 0) {
            return -1.0;
        }
    }
    for (int w = 0; w < workers; ++w) { int st; wait(&st); }
    double t1 = now_s();
    uint64_t np = 0, nm = 0, h = FNV_OFF; double tmax = 0;
    g_hsum[0] = g_hsum[1] = 0;
    for (int w = 0; w < workers; ++w) {
        np += slots[w].np; nm += slots[w].nm;
        g_hsum[0] += slots[w].hs0; g_hsum[1] += slots[w].hs1;
        h ^= slots[w].h; h *= FNV_PRIME;
        if (slots[w].t > tmax) tmax = slots[w].t;
    }
    munmap(slots, sizeof(Slot) * (size_t)workers);
    if (positions) *positions = np;
    if (matches) *matches = nm;
    if (checksum_fold) *checksum_fold = h;
    if (max_loop_seconds) *max_loop_seconds = tmax;
    return t1 - t0;
}
