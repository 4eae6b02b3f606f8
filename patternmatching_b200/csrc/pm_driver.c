/*
 * pm_driver.c -- measurement driver with the reference's command line and CSV schema.  It hosts matchers through
 * the reference's own plugin surface, MpsElem (Core/src/mps.h:71-80): the three B200 matchers of libpm_b200.so
 * (registered with mps_gpu*_register_into, driven through gpu_create / gpu_add_pattern / gpu_compile / gpu_reset /
 * gpu_read_block) and, with -p, every entry of another library's mps_table[] -- e.g. the reference's CPU
 * algorithms compiled unchanged -- driven with one read_char call per byte exactly like Core/src/measure.c:292-294.
 *
 *   pm_driver -d FILE [-d FILE ...] -s FILE [-s FILE ...] -o FILE [-v] [-p LIB.so] [-b BYTES] [-r gpu|plugin]
 *
 *   -p LIB.so   shared library that exports the reference's registry: `void mps_table_setup(void)`, `MpsElem mps_table[]`
 *               and `int mps_table_count(void)`; each entry becomes a CSV row beside the GPU rows
 *   -b BYTES    stream chunk per read (the reference: 100 KiB, measure.c:77; default here 16 MiB)
 *   -r WHICH    the "reliable" instance every row is classified against (mps.c:52-53): `plugin` = a second instance of
 *               the plugin's first entry (the reference's AC; default when -p is given), `gpu` = the B200 forward DFA
 *
 * What it mirrors (yehonatan145/PatternMatching): the flags of parse_arguments (Core/src/parser.c:104-162,
 * Core/src/util.c:6-13); the flow of main (Core/src/main.c:7-24): build every instance from the merged dictionaries
 * (init_mps, mps.c:109-113 -- add_pattern once per UNIQUE pattern, then compile), run every instance over every stream
 * with `reset` per stream file (measure.c:274-275), time only the matching (measure.c:290-297), classify every position
 * against the reliable instance (measure.c:174-190, 300-303) and write one CSV row per instance with the reference's
 * first six columns (measure.c:352-396).  The perf_event columns are replaced by throughput columns; time is wall
 * clock, not clock().  File reads, matching and classification of consecutive chunks overlap (three threads, a ring of
 * chunk slots).  Reference bugs not reproduced: pointer arrays sized in pointers, output file created with a
 * mode and truncated.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
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
    fprintf(stderr, "  -p LIB.so             also measure every algorithm of LIB.so's mps_table (MpsElem plugins)\n");
    fprintf(stderr, "  -b BYTES              stream bytes per chunk (default 16777216)\n");
    fprintf(stderr, "  -r gpu|plugin         which instance is the reliable one\n");
    fprintf(stderr, "  -g KIND:BYTES:FILE    write the seeded synthetic stream KIND (uniform, planted, almost, ab, ascii) of BYTES\n");
    fprintf(stderr, "                        bytes (a multiple of 4096) to FILE and exit; planted / almost use the -d dictionaries\n");
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

/* one row of the run: an MpsElem, its instance, and the batched entry when it has one */
typedef size_t (*read_block_fn)(void*, const char*, size_t, void**);
typedef struct {
    pm_mps_elem elem;
    read_block_fn read_block; /* NULL: loop read_char (measure.c:292-294) */
    void* obj;
} Row;

static void feed(const Row* r, const char* buf, size_t n, void** out) {
    if (r->read_block) { r->read_block(r->obj, buf, n, out); return; }
    void* (*read_char)(void*, char) = r->elem.read_char;
    void* obj = r->obj;
    for (size_t j = 0; j < n; ++j) out[j] = read_char(obj, buf[j]);
}

/* ---- chunk pipeline: reader -> matcher -> classifier -------------------------------------------------- */
enum { N_SLOTS = 3 };
typedef struct {
    char* buf;
    void** res;
    void** real;
    size_t len;
    int first_of_file; /* reset both instances before this chunk */
    int last;          /* no more chunks for this row */
} Slot;
typedef struct {
    pthread_mutex_t mu;
    pthread_cond_t cv;
    int ready[N_SLOTS + 1], head, tail;
} Chan;
static void chan_init(Chan* c) { pthread_mutex_init(&c->mu, NULL); pthread_cond_init(&c->cv, NULL); c->head = c->tail = 0; }
static void chan_put(Chan* c, int v) {
    pthread_mutex_lock(&c->mu);
    c->ready[c->tail] = v; c->tail = (c->tail + 1) % (N_SLOTS + 1);
    pthread_cond_signal(&c->cv);
    pthread_mutex_unlock(&c->mu);
}
static int chan_get(Chan* c) {
    pthread_mutex_lock(&c->mu);
    while (c->head == c->tail) pthread_cond_wait(&c->cv, &c->mu);
    const int v = c->ready[c->head]; c->head = (c->head + 1) % (N_SLOTS + 1);
    pthread_mutex_unlock(&c->mu);
    return v;
}

typedef struct {
    Slot slot[N_SLOTS];
    Chan free_q, read_q, scan_q;
    size_t chunk;
    char** streams; int n_stream;
    const Row* row; const Row* reliable;
    const uint32_t* parent; uint32_t n_pats;
    double secs; uint64_t bytes; uint64_t cnt[4];
} Run;

static void* reader_main(void* arg) {
    Run* R = (Run*)arg;
    for (int s = 0; s < R->n_stream; ++s) {
        int fd = open(R->streams[s], O_RDONLY);
        if (fd == -1) {
            fprintf(stderr, "can't open stream file %s: %s\n", R->streams[s], strerror(errno));
            exit(EXIT_FAILURE);
        }
        int first = 1;
        for (;;) {
            const int k = chan_get(&R->free_q);
            Slot* sl = &R->slot[k];
            size_t got = 0;
            while (got < R->chunk) {
                ssize_t r = read(fd, sl->buf + got, R->chunk - got);
                if (r < 0) { fprintf(stderr, "can't read from stream file %s: %s\n", R->streams[s], strerror(errno)); exit(EXIT_FAILURE); }
                if (r == 0) break;
                got += (size_t)r;
            }
            sl->len = got; sl->first_of_file = first; sl->last = 0;
            first = 0;
            chan_put(&R->read_q, k);
            if (got < R->chunk) break;
        }
        close(fd);
    }
    const int k = chan_get(&R->free_q);
    R->slot[k].len = 0; R->slot[k].first_of_file = 0; R->slot[k].last = 1;
    chan_put(&R->read_q, k);
    return NULL;
}

static void* classify_main(void* arg) {
    Run* R = (Run*)arg;
    for (;;) {
        const int k = chan_get(&R->scan_q);
        Slot* sl = &R->slot[k];
        const int last = sl->last;
        for (size_t j = 0; j < sl->len; ++j) { /* measure.c:174-190; ids are pids carried as pointers */
            const uintptr_t x = (uintptr_t)sl->res[j], y = (uintptr_t)sl->real[j];
            if (x == y) { R->cnt[0]++; continue; }
            uintptr_t c = y;
            while (c && c != x) c = c <= R->n_pats ? R->parent[c] : 0;
            if (x && c == x) R->cnt[1]++;   /* algo is an ancestor of real: partial success */
            else if (!x) R->cnt[2]++;       /* false negative */
            else R->cnt[3]++;               /* false positive */
        }
        chan_put(&R->free_q, k);
        if (last) return NULL;
    }
}

static void run_row(Run* R) {
    chan_init(&R->free_q); chan_init(&R->read_q); chan_init(&R->scan_q);
    for (int k = 0; k < N_SLOTS; ++k) chan_put(&R->free_q, k);
    R->secs = 0; R->bytes = 0; memset(R->cnt, 0, sizeof(R->cnt));
    pthread_t rd, cl;
    pthread_create(&rd, NULL, reader_main, R);
    pthread_create(&cl, NULL, classify_main, R);
    for (;;) {
        const int k = chan_get(&R->read_q);
        Slot* sl = &R->slot[k];
        if (sl->first_of_file) { /* measure.c:274-275: reset both before every stream file */
            R->reliable->elem.reset(R->reliable->obj);
            R->row->elem.reset(R->row->obj);
        }
        if (sl->len) {
            const double b = now_s();
            feed(R->row, sl->buf, sl->len, sl->res);           /* the timed region, measure.c:290-297 */
            R->secs += now_s() - b;
            feed(R->reliable, sl->buf, sl->len, sl->real);     /* measure.c:300-302 */
            R->bytes += sl->len;
        }
        const int last = sl->last;
        chan_put(&R->scan_q, k);
        if (last) break;
    }
    pthread_join(rd, NULL);
    pthread_join(cl, NULL);
}

int main(int argc, char* argv[]) {
    program_name = argv[0];
    int opt, n_dict = 0, n_stream = 0, n_out = 0;
    const char* plugin_path = NULL;
    const char* reliable_kind = NULL;
    size_t chunk = (size_t)16 << 20;
    opterr = 0;
    const char* gen_spec = NULL;
    while ((opt = getopt(argc, argv, "d:s:o:vp:b:r:g:")) != -1) {
        if (opt == 'g') { gen_spec = optarg; continue; }
        if (opt == 'd') ++n_dict;
        else if (opt == 's') ++n_stream;
        else if (opt == 'o') ++n_out;
        else if (opt == 'v') verbose = 1;
        else if (opt == 'p') plugin_path = optarg;
        else if (opt == 'b') chunk = (size_t)strtoull(optarg, NULL, 10);
        else if (opt == 'r') reliable_kind = optarg;
        else {
            if (optopt == 'd' || optopt == 's' || optopt == 'o' || optopt == 'p' || optopt == 'b' || optopt == 'r' || optopt == 'g')
                fprintf(stderr, "Option -%c must have argument.\n\n", optopt);
            else fprintf(stderr, "Unknown option -%c.\n\n", optopt);
            usage();
            return EXIT_FAILURE;
        }
    }
    if (gen_spec) {
        /* stream generator mode: the engine's seeded generators, written to a file */
        static const char* kinds[] = {"uniform", "planted", "almost", "ab", "ascii"};
        char kind_name[16] = {0};
        unsigned long long bytes = 0;
        char path[4096] = {0};
        if (sscanf(gen_spec, "%15[^:]:%llu:%4095s", kind_name, &bytes, path) != 3 || bytes == 0 || (bytes & 4095)) { usage(); return EXIT_FAILURE; }
        int kind = -1;
        for (int k = 0; k < 5; ++k) if (strcmp(kind_name, kinds[k]) == 0) kind = k;
        if (kind < 0 || ((kind == 1 || kind == 2) && n_dict == 0)) { usage(); return EXIT_FAILURE; }
        pm_dict* gd = pm_dict_create();
        optind = 1;
        while ((opt = getopt(argc, argv, "d:s:o:vp:b:r:g:")) != -1)
            if (opt == 'd' && pm_dict_add_file(gd, optarg)) fatal("pm_dict_add_file");
        if (n_dict == 0) pm_dict_add_pattern(gd, (const uint8_t*)"a", 1, 0, 1, 0);   /* the generators need a compiled dictionary */
        if (pm_dict_compile(gd)) fatal("pm_dict_compile");
        const char* dev = getenv("PM_B200_DEVICE");
        pm_engine* ge = pm_engine_create(gd, dev ? atoi(dev) : 0);
        if (!ge) fatal("pm_engine_create");
        const size_t piece = (size_t)64 << 20;
        uint8_t* buf = (uint8_t*)pm_host_alloc(piece);
        FILE* f = fopen(path, "wb");
        if (!buf || !f) { fprintf(stderr, "can't write %s\n", path); return EXIT_FAILURE; }
        for (unsigned long long o = 0; o < bytes; o += piece) {
            const size_t len = (size_t)(bytes - o < piece ? bytes - o : piece);
            if (pm_engine_generate_host(ge, kind, o, len, buf)) fatal("pm_engine_generate_host");
            if (fwrite(buf, 1, len, f) != len) { fprintf(stderr, "short write to %s\n", path); return EXIT_FAILURE; }
        }
        fclose(f);
        if (verbose) printf("wrote %llu bytes of the %s stream to %s\n", bytes, kind_name, path);
        pm_host_free(buf); pm_engine_free(ge); pm_dict_free(gd);
        return EXIT_SUCCESS;
    }
    if (n_out != 1 || n_dict == 0 || n_stream == 0 || chunk == 0) {
        if (n_out > 1) fprintf(stderr, "Error: have more than one output file\n\n");
        usage();
        return EXIT_FAILURE;
    }
    char** dicts = (char**)calloc((size_t)n_dict, sizeof(char*));
    char** streams = (char**)calloc((size_t)n_stream, sizeof(char*));
    char* out_name = NULL;
    int di = 0, si = 0;
    optind = 1;
    while ((opt = getopt(argc, argv, "d:s:o:vp:b:r:g:")) != -1) {
        if (opt == 'd') dicts[di++] = optarg;
        else if (opt == 's') streams[si++] = optarg;
        else if (opt == 'o') out_name = optarg;
    }

    /* ---- the rows: plugin entries first (the reference's order: its own algorithms), then the GPU matchers ---- */
    enum { MAX_ROWS = 16 };
    Row rows[MAX_ROWS];
    int n_rows = 0, n_plugin = 0;
    memset(rows, 0, sizeof(rows));
    if (plugin_path) {
        void* h = dlopen(plugin_path, RTLD_NOW | RTLD_LOCAL);
        if (!h) { fprintf(stderr, "can't load plugin %s: %s\n", plugin_path, dlerror()); return EXIT_FAILURE; }
        void (*setup)(void) = (void (*)(void))dlsym(h, "mps_table_setup");
        pm_mps_elem* table = (pm_mps_elem*)dlsym(h, "mps_table");
        int (*count)(void) = (int (*)(void))dlsym(h, "mps_table_count");
        if (!setup || !table || !count) {
            fprintf(stderr, "plugin %s must export mps_table_setup, mps_table and mps_table_count\n", plugin_path);
            return EXIT_FAILURE;
        }
        setup();
        n_plugin = count();
        for (int i = 0; i < n_plugin && n_rows < MAX_ROWS - 5; ++i) rows[n_rows++].elem = table[i];
    }
    mps_gpu_register_into(&rows[n_rows].elem);     rows[n_rows++].read_block = gpu_read_block;
    mps_gpu_dfa_register_into(&rows[n_rows].elem); rows[n_rows++].read_block = gpu_read_block;
    mps_gpu_kr_register_into(&rows[n_rows].elem);  rows[n_rows++].read_block = gpu_read_block;
    mps_gpu_mpbg_register_into(&rows[n_rows].elem); rows[n_rows++].read_block = gpu_read_block;   /* the reference MPBG's behaviour */
    Row reliable;
    memset(&reliable, 0, sizeof(reliable));
    const int reliable_plugin = reliable_kind ? strcmp(reliable_kind, "plugin") == 0 : n_plugin > 0;
    if (reliable_plugin && !n_plugin) { fprintf(stderr, "-r plugin needs -p\n\n"); usage(); return EXIT_FAILURE; }
    if (reliable_plugin) reliable.elem = rows[0].elem;  /* mps.c:52-53: a separate instance of MPS_AC */
    else { mps_gpu_dfa_register_into(&reliable.elem); reliable.read_block = gpu_read_block; }

    /* ---- init_mps (mps.c:109-113): one create per instance, add_pattern per UNIQUE pattern, compile ---- */
    double t0 = now_s();
    pm_dict* dict = pm_dict_create();
    for (int i = 0; i < n_dict; ++i)
        if (pm_dict_add_file(dict, dicts[i])) {
            fprintf(stderr, "failed to open dictionary file %s: %s", dicts[i], pm_last_error());
            return EXIT_FAILURE;
        }
    if (pm_dict_compile(dict)) fatal("pm_dict_compile");  /* de-dup, (file,line) ids, PatternsTree parents */
    pm_dict_info info;
    pm_dict_get_info(dict, &info);
    if (verbose)
        printf("dictionaries: %u unique patterns (%llu lines, %llu rejected, %llu duplicates), %u AC states, ingested in %.2f s\n",
               info.n_patterns, (unsigned long long)info.n_lines, (unsigned long long)info.n_rejected,
               (unsigned long long)info.n_duplicates, info.n_ac_states, now_s() - t0);
    uint32_t* parent = (uint32_t*)calloc((size_t)info.n_patterns + 1, sizeof(uint32_t));
    for (uint32_t pid = 1; pid <= info.n_patterns; ++pid) pm_dict_pattern(dict, pid, NULL, NULL, NULL, &parent[pid], NULL, NULL);
    char* scratch = (char*)malloc((size_t)info.max_pat_len + 1);
    for (int r = 0; r <= n_rows; ++r) {
        Row* row = r < n_rows ? &rows[r] : &reliable;
        t0 = now_s();
        row->obj = row->elem.create();
        for (uint32_t pid = 1; pid <= info.n_patterns; ++pid) {
            uint32_t len = 0; const uint8_t* bytes = NULL;
            pm_dict_pattern(dict, pid, NULL, NULL, NULL, NULL, &len, &bytes);
            memcpy(scratch, bytes, len);                      /* the callee must copy: the buffer is scratch */
            row->elem.add_pattern(row->obj, scratch, len, (void*)(uintptr_t)pid);
        }
        row->elem.compile(row->obj);
        if (verbose) printf("built %s%s in %.2f s\n", row->elem.name, r == n_rows ? " (reliable instance)" : "", now_s() - t0);
    }
    free(scratch);

    /* ---- measure_instances_stats (measure.c:324-333) ---- */
    Run R;
    memset(&R, 0, sizeof(R));
    R.chunk = chunk; R.streams = streams; R.n_stream = n_stream; R.reliable = &reliable; R.parent = parent; R.n_pats = info.n_patterns;
    for (int k = 0; k < N_SLOTS; ++k) {
        R.slot[k].buf = (char*)pm_host_alloc(chunk);
        R.slot[k].res = (void**)pm_host_alloc(chunk * sizeof(void*));
        R.slot[k].real = (void**)pm_host_alloc(chunk * sizeof(void*));
        if (!R.slot[k].buf || !R.slot[k].res || !R.slot[k].real) fatal("pm_host_alloc");
    }
    double secs[MAX_ROWS]; uint64_t cnt[MAX_ROWS][4], bytes_total[MAX_ROWS]; size_t mem[MAX_ROWS];
    for (int a = 0; a < n_rows; ++a) {
        if (verbose) { printf("Measuring algorithm %s...", rows[a].elem.name); fflush(stdout); }
        R.row = &rows[a];
        run_row(&R);
        secs[a] = R.secs; bytes_total[a] = R.bytes; memcpy(cnt[a], R.cnt, sizeof(R.cnt));
        mem[a] = rows[a].elem.total_mem(rows[a].obj);    /* measure.c:310 */
        if (verbose) printf("Done\n");
    }

    if (verbose) printf("opening file %s to write results\n", out_name);
    FILE* f = fopen(out_name, "w");
    if (!f) {
        fprintf(stderr, "failed to open results file %s: %s\n", out_name, strerror(errno));
        return EXIT_FAILURE;
    }
    fprintf(f, "Algorithm,Time (in secs),Total Memory Used,False Positive Rate,False Negative Rate,Partial Success Rate,Stream Bytes,GB/s (host buffers)");
    for (int a = 0; a < n_rows; ++a) {
        const long double sum = (long double)(cnt[a][0] + cnt[a][1] + cnt[a][2] + cnt[a][3]);
        fprintf(f, "\n%s,%.6f,%zu,%.6Lf,%.6Lf,%.6Lf,%llu,%.3f", rows[a].elem.name, secs[a], mem[a],
                sum ? cnt[a][3] / sum : 0.0L, sum ? cnt[a][2] / sum : 0.0L, sum ? cnt[a][1] / sum : 0.0L,
                (unsigned long long)bytes_total[a], secs[a] > 0 ? bytes_total[a] / secs[a] / 1e9 : 0.0);
    }
    fclose(f);
    for (int k = 0; k < N_SLOTS; ++k) { pm_host_free(R.slot[k].buf); pm_host_free(R.slot[k].res); pm_host_free(R.slot[k].real); }
    for (int a = n_plugin; a < n_rows; ++a) rows[a].elem.free(rows[a].obj);  /* the reference never frees its own (and ac_free is not re-entrant) */
    if (!reliable_plugin) reliable.elem.free(reliable.obj);
    pm_dict_free(dict);
    free(parent); free(dicts); free(streams);
    return 0;
}
