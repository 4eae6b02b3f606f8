/*
 * mps_gpu_shim.c -- the reference's algorithm-plugin surface (MpsElem, Core/src/mps.h:71-80) on top
 * of the C-ABI in include/pm_b200.h.  Plain C, no reference header needed: pattern_id_t is a
 * pointer there (Core/src/PatternsTree.h:104) and is carried as void* here.
 *
 * Call order is the reference's (Core/src/mps.c:44-54, 64-77, 84-96; Core/src/measure.c:274-275,
 * 292-294, 310): create, add_pattern per unique pattern, compile, then reset / read_char ... ;
 * errors print to stderr and exit(EXIT_FAILURE) like FatalExit() (Core/src/util.h:37-39).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pm_b200.h"

typedef struct {
    pm_dict* dict;
    pm_engine* eng;
    int algo;
    uint32_t seq;        /* add order: stands in for (file,line), which the plugin API does not pass */
    uint64_t* id_of_pid; /* pid -> pattern_id_t given to add_pattern (entry 0 = NULL = null_pattern_id) */
    uint32_t n_pids;
} GpuMps;

static void die(const char* what) {
    fprintf(stderr, "pm_b200: %s: %s\n", what, pm_last_error());
    exit(EXIT_FAILURE);
}

static void* create_with(int algo) {
    GpuMps* g = (GpuMps*)calloc(1, sizeof(GpuMps));
    if (!g) { perror("failed to allocate memory"); exit(EXIT_FAILURE); }
    g->dict = pm_dict_create();
    g->algo = algo;
    if (!g->dict) die("pm_dict_create");
    return g;
}

void* gpu_create(void) { return create_with(PM_ALGO_AUTO); }  /* exact; picks the kernel from a sample of the stream */
void* gpu_dfa_create(void) { return create_with(PM_ALGO_DFA); }
void* gpu_kr_create(void) { return create_with(PM_ALGO_KR); }
void* gpu_mpbg_create(void) { return create_with(PM_ALGO_MPBG); }   /* the reference's MPBG, position for position */

void gpu_add_pattern(void* obj, char* pat, size_t len, void* pattern_id) {
    GpuMps* g = (GpuMps*)obj;
    /* the caller's buffer is scratch and binary (Core/src/README.md:85-89): the bytes are copied */
    pm_dict_add_pattern(g->dict, (const uint8_t*)pat, len, 0, ++g->seq, (uint64_t)(uintptr_t)pattern_id);
}

void gpu_compile(void* obj) {
    GpuMps* g = (GpuMps*)obj;
    if (pm_dict_compile(g->dict)) die("pm_dict_compile");
    const char* dev = getenv("PM_B200_DEVICE");
    g->eng = pm_engine_create(g->dict, dev ? atoi(dev) : 0);
    if (!g->eng) die("pm_engine_create");
    if (pm_engine_prepare_host(g->eng)) die("pm_engine_prepare_host");   /* not inside the caller's timed loop */
    pm_dict_info info;
    pm_dict_get_info(g->dict, &info);
    g->n_pids = info.n_patterns;
    g->id_of_pid = (uint64_t*)calloc((size_t)g->n_pids + 1, sizeof(uint64_t));
    if (!g->id_of_pid) { perror("failed to allocate memory"); exit(EXIT_FAILURE); }
    for (uint32_t pid = 1; pid <= g->n_pids; ++pid) pm_dict_pattern(g->dict, pid, NULL, NULL, &g->id_of_pid[pid], NULL, NULL, NULL);
}

size_t gpu_read_block(void* obj, const char* buf, size_t n, void** out) {
    GpuMps* g = (GpuMps*)obj;
    /* pattern_id_t is a pointer (Core/src/PatternsTree.h:104): 8 bytes per position.  The engine translates pids to
     * the ids recorded in add_pattern on its staging threads, piece by piece while later pieces are still on the GPU. */
    if (pm_engine_scan_host_ids(g->eng, g->algo, (const uint8_t*)buf, n, g->id_of_pid, (size_t)g->n_pids + 1, (uint64_t*)out))
        die("pm_engine_scan_host_ids");
    return n;
}

void* gpu_read_char(void* obj, char c) {
    void* r = NULL;
    gpu_read_block(obj, &c, 1, &r);
    return r;
}

size_t gpu_total_mem(void* obj) {
    GpuMps* g = (GpuMps*)obj;
    return (g && g->eng) ? pm_engine_total_mem(g->eng) : 0;
}

void gpu_reset(void* obj) {
    GpuMps* g = (GpuMps*)obj;
    pm_engine_reset(g->eng);
}

void gpu_free(void* obj) {
    GpuMps* g = (GpuMps*)obj;
    if (!g) return;
    pm_engine_free(g->eng);
    pm_dict_free(g->dict);
    free(g->id_of_pid);
    free(g);
}

static void fill(pm_mps_elem* e, const char* name, void* (*create)(void)) {
    e->name = (char*)name;
    e->create = create;
    e->add_pattern = gpu_add_pattern;
    e->compile = gpu_compile;
    e->read_char = gpu_read_char;
    e->total_mem = gpu_total_mem;
    e->reset = gpu_reset;
    e->free = gpu_free;
}
/* what a mps_gpu_register() added to mps_table_setup (Core/src/mps.c:120-124) would call */
void mps_gpu_register_into(pm_mps_elem* slot) { fill(slot, "B200 exact dictionary scan", gpu_create); }
void mps_gpu_dfa_register_into(pm_mps_elem* slot) { fill(slot, "B200 Aho-Corasick DFA", gpu_dfa_create); }
void mps_gpu_kr_register_into(pm_mps_elem* slot) { fill(slot, "B200 Karp-Rabin stages", gpu_kr_create); }
void mps_gpu_mpbg_register_into(pm_mps_elem* slot) { fill(slot, "B200 MPBG (as shipped)", gpu_mpbg_create); }
