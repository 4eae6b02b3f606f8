// engine.cu -- the C-ABI of libpm_b200.so (include/pm_b200.h): dictionary objects, the device-resident
// engine, scan dispatch, the host-buffer pipeline.  No CPU fallback anywhere: every scan entry point
// needs a CUDA device and fails loudly without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pm_b200.h"
#include "aux_kernels.cuh"
#include "deep_scan.cuh"
#include "dfa_scan.cuh"
#include "dict.hpp"
#include "host_pool.hpp"
#include "kr_scan.cuh"
#include "pm_dev.cuh"
#include "sfx_scan.cuh"

namespace {
thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return -1;
}
#define CU(call)                                         \
    do {                                                 \
        cudaError_t e__ = (call);                        \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
    } while (0)
}  // namespace

struct pm_dict {
    pm::Dict d;
};

namespace {
template <class T>
cudaError_t upload(const std::vector<T>& v, T** dptr, size_t* total) {
    *dptr = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(dptr), bytes);
    if (e != cudaSuccess) return e;
    *total += bytes;
    if (!v.empty()) e = cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}
constexpr size_t kMaxCtas = 1024;  // upper bound of scan CTAs per launch (one per SM)
constexpr size_t kPatPad = 16;  // zero bytes in front of the device copy of the pattern text (8-byte windows)
constexpr size_t kSmallCall = size_t(256) << 10;   // host calls up to this many bytes take the single-launch path
constexpr size_t kPageableChunk = size_t(4) << 20; // pipeline piece when a buffer has to be staged by host threads
constexpr int kSlots = 4;   // pipeline slots of the host path: pieces on the GPU + pieces being unloaded by the host threads
constexpr size_t kQueueMaxPerCta = size_t(256) << 10;  // deferred-walk slots per scan CTA (2 MiB); beyond that walks finish inline

// Diagnostic / A-B switches, read ONCE when an engine is created.
struct EngineOpts {
    bool sfx_no_tex = false, sfx_no_l3 = false, dfa_no_fb = false, dfa_flat = false, dfa_deep = false, dfa_no_fused = false, kr_no_bulk = false;
    uint32_t l3_min = 4, l3_min_b = 4;
    size_t host_chunk = size_t(16) << 20;  // bytes per pipeline slot for pinned buffers (PM_HOST_CHUNK_MIB)
    int host_threads = 0;                  // staging threads (PM_HOST_THREADS; default: 3/4 of this rank's share of the host's cores, at most 16)
    int cpu_first = 0, cpu_count = 0;      // this rank's share of the CPUs (the workers are pinned inside it)
    // PM_HOST_IDS=device|host|<N>: where pids become 8-byte ids when the result buffer is page-locked: every piece on the
    // device (1), every piece by the host threads (0), or every N-th piece on the device and the rest by the host threads
    // (both resources at once: the PCIe link and the host's store bandwidth); -1 = chosen from the thread count
    int ids_device_every = -1;
    static EngineOpts from_env() {
        EngineOpts o;
        o.sfx_no_tex = getenv("PM_SFX_NO_TEX") != nullptr;
        o.sfx_no_l3 = getenv("PM_SFX_NO_L3") != nullptr;
        o.dfa_no_fb = getenv("PM_DFA_NO_FB") != nullptr;
        o.dfa_flat = getenv("PM_DFA_FLAT") != nullptr;
        o.dfa_deep = getenv("PM_DFA_DEEP") != nullptr;
        o.dfa_no_fused = getenv("PM_DFA_NO_FUSED") != nullptr;
        o.kr_no_bulk = getenv("PM_KR_NO_BULK") != nullptr;
        if (const char* v = getenv("PM_SFX_L3_MIN")) o.l3_min = o.l3_min_b = uint32_t(atoi(v));
        if (const char* v = getenv("PM_SFX_L3_MIN_B")) o.l3_min_b = uint32_t(atoi(v));
        if (const char* v = getenv("PM_HOST_CHUNK_MIB")) { const long m = atol(v); if (m >= 1 && m <= 1024) o.host_chunk = size_t(m) << 20; }
        // One process per GPU (torchrun sets LOCAL_WORLD_SIZE / LOCAL_RANK): the ranks of a box split its cores, and each
        // pins its workers inside its own share -- eight ranks that all take "half the cores" and pin them to the same
        // CPUs starve each other.  Within the share: three quarters of the cores (at most 16): on the 16-vCPU B200 box
        // 12 threads gave the best host-path numbers in same-box comparisons (8: 10.8-11.1, 12: 12.2-12.3, 16: 12.6 GB/s
        // for page-locked 1 GiB gpu_read_block calls, but 16 lost on 16 MiB calls: one per core competes with the CUDA
        // driver's own threads and the caller).
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned lws = 1, lrank = 0;
        if (const char* v = getenv("LOCAL_WORLD_SIZE")) lws = unsigned(std::max(1, atoi(v)));
        if (const char* v = getenv("LOCAL_RANK")) lrank = unsigned(std::max(0, atoi(v)));
        const unsigned share = std::max(1u, hw / lws);
        o.cpu_first = int((lrank % lws) * share);
        o.cpu_count = int(share);
        o.host_threads = int(std::max(1u, std::min<unsigned>(share >= 4 ? share * 3 / 4 : share, 16)));
        if (const char* v = getenv("PM_HOST_THREADS")) { const int t = atoi(v); if (t >= 1 && t <= 256) o.host_threads = t; }
        if (const char* v = getenv("PM_HOST_IDS")) o.ids_device_every = v[0] == 'd' ? 1 : v[0] == 'h' ? 0 : std::max(0, atoi(v));
        return o;
    }
};
}  // namespace

struct pm_engine {
    std::mutex mu;  // the entry points below are serialised per engine (one stream state, shared scratch buffers)
    const pm::Dict* dict = nullptr;
    int device = 0, n_sms = 0;
    size_t table_bytes = 0, scratch_bytes = 0;
    uint64_t launches = 0;
    EngineOpts opts;
    size_t halo = pm::kHalo;  // bytes of history that make a shard scan equal to the continuous one (>= max_pat_len - 1)
    // sfx tables
    uint16_t* d_root2 = nullptr;
    uint32_t *d_root1 = nullptr, *d_rows = nullptr, *d_row_best = nullptr;
    cudaTextureObject_t rows_tex = 0;
    uint8_t* d_cls = nullptr;
    // pattern tables
    uint32_t *d_pat_off = nullptr, *d_pat_len = nullptr;
    uint8_t* d_pat_bytes = nullptr;
    uint16_t *d_parent = nullptr, *d_chain = nullptr;
    uint32_t* d_tail_rec = nullptr;
    uint32_t* d_l3f = nullptr;
    uint32_t* d_anc_off = nullptr;
    uint16_t* d_anc_list = nullptr;
    uint64_t* d_pidhash = nullptr;
    uint16_t* d_hot_map = nullptr;            // summary fast path (PatTables::hot_*)
    uint64_t *d_hot_own = nullptr, *d_hot_anc = nullptr;
    pm::PatTables pt{};
    // dfa tables (device copies made on first use; the host tables belong to the dictionary, see Dict::build_dfa)
    uint32_t* d_delta = nullptr;
    uint16_t* d_longest = nullptr;
    uint8_t* d_dfa_cls = nullptr;
    uint32_t* d_fb_meta = nullptr;
    bool dfa_ready = false;
    // compact goto + failure automaton of the deep-match walker (device copies made on first use)
    uint16_t *d_deep_hot = nullptr, *d_deep_long = nullptr;
    uint32_t *d_deep_recs = nullptr, *d_deep_dense = nullptr;
    bool deep_ready = false;
    // kr tables: built for THIS engine's seed and owned by it
    pm::KrDevTables kr{};
    bool kr_ready = false;
    uint64_t kr_seed = 0xF1A90003ull, kr_built_seed = 0;
    size_t kr_bytes = 0;
    // scratch
    unsigned long long* d_acc = nullptr;  // 8 x u64
    unsigned long long* d_compact_counts = nullptr;  // per-CTA counts of pm_engine_compact, grown on demand
    size_t compact_cap = 0;
    uint32_t* d_flags = nullptr;          // sparse mode: one bit per position (pm_engine_scan_device_records), grown on demand
    size_t flags_words = 0;
    // deferred-walk queues of the sfx scan, one per pipeline slot (slot 0 also serves pm_engine_scan_device)
    uint64_t* d_queue[kSlots] = {};
    uint32_t* d_qcount = nullptr;         // kSlots x 2 x kMaxCtas counters
    size_t queue_cap[kSlots] = {};
    size_t last_ctas[kSlots] = {};        // scan CTAs of the last sfx launch that used the slot
    // slot-0 scratch is shared by successive pm_engine_scan_device calls whatever stream they are given: the next
    // call waits (on the device) for the previous one through this event
    cudaEvent_t scratch_free = nullptr;
    bool scratch_used = false;
    // PM_ALGO_AUTO: scratch result buffer for the sampling scans, decision cached per stream (until reset)
    uint16_t* d_sample_out = nullptr;
    int auto_choice = -1;
    bool auto_flat = false;
    // optional per-kernel timing of the sfx scan (bench.py's roofline): 3 events per profiled scan
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used = 0;
    // host pipeline (lazy)
    bool pipe_ready = false;
    uint8_t* d_in[kSlots] = {};
    uint16_t* d_out[kSlots] = {};
    uint8_t* h_in[kSlots] = {};
    uint16_t* h_out[kSlots] = {};
    cudaStream_t st[kSlots] = {};
    cudaEvent_t done[kSlots] = {};
    std::unique_ptr<pm::HostPool> pool;
    // pid -> caller's id translated on the device (page-locked result buffers): the table and per-slot id buffers
    unsigned long long* d_id_table = nullptr;
    const uint64_t* id_table_src = nullptr;   // host table last uploaded (re-uploaded when pointer or contents change)
    uint64_t id_table_sum = 0;
    size_t id_table_n = 0;
    unsigned long long* d_ids[kSlots] = {};
    size_t ids_cap = 0;                       // positions per d_ids buffer
    // record path of the host pipeline (lazy)
    uint64_t* d_rec[2] = {nullptr, nullptr};
    unsigned long long* d_rec_counts[2] = {nullptr, nullptr};
    unsigned long long* h_rec_total[2] = {nullptr, nullptr};
    uint32_t* d_rec_flags[2] = {nullptr, nullptr};   // sparse-mode bitmaps of the two pipeline slots
    // dictionaries of more than 65,535 patterns (pm::Dict::multi): one complete sub-engine per part; this object only
    // keeps the stream state, the host pipeline buffers and the merge scratch
    std::vector<pm_engine*> parts;
    uint32_t* d_glen = nullptr;               // pattern length by GLOBAL pid (the merge keeps the longer answer)
    uint16_t* d_tmp16 = nullptr;              // one part's dense answer for a slice
    size_t tmp16_cap = 0;
    uint32_t *d_out32 = nullptr, *h_out32 = nullptr;   // host path of multi-part engines: one piece of 32-bit results
    // stream state carried between host calls (== ac->current_state of the reference): the last `halo` bytes, right-aligned
    std::vector<uint8_t> h_hist;
    size_t hist_valid = 0;
};

namespace {

int ensure_dfa(pm_engine* e) {
    if (e->dfa_ready) return 0;
    const pm::Dict& d = *e->dict;
    d.build_dfa();  // at most once per dictionary, under the dictionary's own lock
    CU(upload(d.dfa.delta, &e->d_delta, &e->table_bytes));
    CU(upload(d.dfa.longest, &e->d_longest, &e->table_bytes));
    CU(upload(d.dfa.fb_meta, &e->d_fb_meta, &e->table_bytes));
    std::vector<uint8_t> cls(d.dfa.cls, d.dfa.cls + 256);
    CU(upload(cls, &e->d_dfa_cls, &e->table_bytes));
    e->dfa_ready = true;
    return 0;
}

int ensure_deep(pm_engine* e) {
    if (e->deep_ready) return 0;
    const pm::Dict& d = *e->dict;
    d.build_deep();
    if (!d.deep.usable) return fail("the automaton does not fit the deep-match layout");
    CU(upload(d.deep.hot_rows, &e->d_deep_hot, &e->table_bytes));
    CU(upload(d.deep.hot_longest, &e->d_deep_long, &e->table_bytes));
    CU(upload(d.deep.recs, &e->d_deep_recs, &e->table_bytes));
    CU(upload(d.deep.dense_rows, &e->d_deep_dense, &e->table_bytes));
    e->deep_ready = true;
    return 0;
}

int ensure_kr(pm_engine* e) {
    if (e->kr_ready && e->kr_built_seed == e->kr_seed) return 0;
    const pm::KrTables host = e->dict->build_kr(e->kr_seed);
    pm::kr_free_tables(&e->kr);
    e->table_bytes -= e->kr_bytes;
    e->kr_bytes = 0; e->kr_ready = false;
    cudaError_t ce = pm::kr_upload_tables(*e->dict, host, &e->kr, &e->kr_bytes);
    e->table_bytes += e->kr_bytes;
    if (ce != cudaSuccess) return cuda_fail(ce, "kr_upload_tables");
    e->kr_ready = true; e->kr_built_seed = e->kr_seed;
    return 0;
}

int ensure_pipe(pm_engine* e) {
    if (e->pipe_ready) return 0;
    const size_t chunk = e->opts.host_chunk;
    for (int b = 0; b < kSlots; ++b) {
        CU(cudaMalloc(reinterpret_cast<void**>(&e->d_in[b]), e->halo + chunk + 16));
        CU(cudaMalloc(reinterpret_cast<void**>(&e->d_out[b]), chunk * sizeof(uint16_t)));
        CU(cudaMallocHost(reinterpret_cast<void**>(&e->h_in[b]), e->halo + chunk));
        CU(cudaMallocHost(reinterpret_cast<void**>(&e->h_out[b]), chunk * sizeof(uint16_t)));
        CU(cudaStreamCreateWithFlags(&e->st[b], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&e->done[b], cudaEventDisableTiming));
        e->scratch_bytes += e->halo + chunk + 16 + chunk * sizeof(uint16_t);
    }
    e->pool.reset(new pm::HostPool(e->opts.host_threads, e->opts.cpu_first, e->opts.cpu_count));
    e->pipe_ready = true;
    return 0;
}

// both pipeline streams idle: nothing is in flight into a caller's buffer any more (called before an error return)
void quiesce(pm_engine* e) {
    for (int b = 0; b < kSlots; ++b) if (e->st[b]) cudaStreamSynchronize(e->st[b]);
}

bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int fill_sfx_params(pm_engine* e, pm::SfxParams* p, size_t n, int slot) {
    const pm::Dict& d = *e->dict;
    p->rows_tex = e->opts.sfx_no_tex ? 0 : e->rows_tex;
    p->root2 = e->d_root2; p->root1 = e->d_root1; p->rows = e->d_rows; p->row_best = e->d_row_best; p->cls = e->d_cls;
    p->cont_base = d.sfx.cont_base; p->row2_base = d.sfx.row2_base; p->log2_ncp = d.sfx.log2_ncp;
    p->l3f = e->opts.sfx_no_l3 ? nullptr : e->d_l3f; p->n_l3 = uint32_t(d.sfx.l3f.size());
    p->l3_min = e->opts.l3_min; p->l3_min_b = e->opts.l3_min_b;
    p->tail_rec = reinterpret_cast<const uint4*>(e->d_tail_rec); p->pat_bytes = e->d_pat_bytes + kPatPad;
    p->pat_len = e->d_pat_len; p->parent = e->d_parent;
    // Deferred-walk queue: one strip per scan CTA.  Random bytes defer ~1e-5 of the positions, C3 ~1.3e-3; strips
    // are sized for 1/64 of the positions, at least 4096 and at most kQueueMaxPerCta slots (2 MiB per CTA, 296 MiB
    // for the 16 GiB bench scan; counted in pm_engine_scratch_mem) -- a CTA that fills its strip finishes further
    // walks inline.  Growing the queue frees and allocates device memory, i.e. synchronises the device.
    const size_t ctas = std::max<size_t>(pm::sfx_scan_ctas(n, e->n_sms), 1);
    const size_t per_cta = std::min(kQueueMaxPerCta, std::max<size_t>(4096, n / 64 / ctas));
    const size_t want = ctas * per_cta;
    if (want > e->queue_cap[slot]) {
        if (e->d_queue[slot]) CU(cudaFree(e->d_queue[slot]));
        e->scratch_bytes -= e->queue_cap[slot] * sizeof(uint64_t);
        e->d_queue[slot] = nullptr; e->queue_cap[slot] = 0;
        CU(cudaMalloc(reinterpret_cast<void**>(&e->d_queue[slot]), want * sizeof(uint64_t)));
        e->queue_cap[slot] = want;
        e->scratch_bytes += want * sizeof(uint64_t);
    }
    p->queue = e->d_queue[slot]; p->qcount = e->d_qcount + size_t(slot) * 2 * kMaxCtas;
    p->q_per_cta = uint32_t(std::min<size_t>(e->queue_cap[slot] / ctas, 1u << 30));
    e->last_ctas[slot] = ctas;
    return 0;
}

int scan_device_impl(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid, uint16_t* d_out,
                     cudaStream_t st, int slot = 0);

// PM_ALGO_AUTO: run the backward scan over four 256 KiB windows of the stream and look at how many walks were
// still alive after level 4.  Uniform / planted / text traffic defers 1e-5 .. 2e-3 of the positions; inputs made
// of pattern prefixes or small-alphabet adversarial dictionaries defer 10..100% and belong to the forward DFA,
// whose work per byte is constant.  Synchronises `st` (one small D2H copy).
constexpr size_t kSampleWin = size_t(256) << 10;
int choose_algo(pm_engine* e, const uint8_t* d_stream, size_t n, size_t hist_valid, cudaStream_t st, int slot) {
    // four windows of 256 KiB; a shorter input (the host pipeline's pieces) is sampled whole, below 64 KiB not at all
    const size_t win = std::min(kSampleWin, n & ~size_t(4095));
    const int n_win = n >= 4 * kSampleWin ? 4 : 1;
    if (win < (size_t(64) << 10)) { e->auto_flat = false; return PM_ALGO_SFX; }
    if (!e->d_sample_out) {
        CU(cudaMalloc(reinterpret_cast<void**>(&e->d_sample_out), kSampleWin * sizeof(uint16_t)));
        e->scratch_bytes += kSampleWin * sizeof(uint16_t);
    }
    uint64_t deferred = 0;
    std::vector<uint32_t> counts(2 * kMaxCtas);
    for (int w = 0; w < n_win; ++w) {
        const size_t off = (n / 4 * size_t(w)) & ~size_t(4095);
        if (scan_device_impl(e, PM_ALGO_SFX, d_stream + off, win, std::min<size_t>(off + hist_valid, e->halo),
                             e->d_sample_out, st, slot)) return -1;
        const size_t ctas = pm::sfx_scan_ctas(win, e->n_sms);
        CU(cudaMemcpyAsync(counts.data(), e->d_qcount + size_t(slot) * 2 * kMaxCtas, 2 * ctas * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (size_t i = 0; i < 2 * ctas; ++i) deferred += counts[i];
    }
    const bool deep = deferred * 32 > size_t(n_win) * win;     // more than 1/32 of the positions walk past level 4
    if (!deep) { e->auto_flat = false; return PM_ALGO_SFX; }
    const pm::Dict& d = *e->dict;
    d.build_deep();   // cheap (no dense table): gives the forward trie's states per depth
    uint32_t hot_rows = 0, hot_long = 0, fb_count = 0;
    pm::dfa_plan_hot(d.deep.n_states, d.sfx.log2_ncp, d.deep.depth_count.data(), uint32_t(d.deep.depth_count.size()), &hot_rows, &hot_long, &fb_count);
    // deep walks and an automaton that does not fit shared memory: the compact-record walker (the dense DFA is never built)
    e->auto_flat = hot_rows < d.deep.n_states;
    return PM_ALGO_DFA;
}

int scan_device_impl(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid, uint16_t* d_out,
                     cudaStream_t st, int slot) {
    if (n == 0) return 0;
    if (!e->parts.empty())
        return fail("this dictionary has more than 65,535 patterns: 16-bit results cannot name them; use pm_engine_scan_device32");
    if (algo == PM_ALGO_MPBG) {
        // the reference's MPBG, position for position: exact scan, then every answer demoted to the longest pattern of
        // <= 8 bytes on its PatternsTree chain (short_only_kernel; pinned by tests/golden/ref_snort.json: mpbg_file/line)
        if (ensure_kr(e)) return -1;
        if (scan_device_impl(e, PM_ALGO_SFX, d_stream, n, hist_valid, d_out, st, slot)) return -1;
        cudaError_t ce = pm::kr_short_only_launch(e->kr, d_out, n, e->n_sms, st, &e->launches);
        if (ce != cudaSuccess) return cuda_fail(ce, "kr_short_only_launch");
        return 0;
    }
    bool force_flat = false;
    if (algo == PM_ALGO_AUTO && !e->dict->sfx.fits_u16) algo = PM_ALGO_SFX;   // routed to the forward walkers below
    if (algo == PM_ALGO_AUTO) {
        if (e->auto_choice < 0) {
            e->auto_choice = choose_algo(e, d_stream, n, hist_valid, st, slot);
            if (e->auto_choice < 0) return -1;
        }
        algo = e->auto_choice;
        force_flat = e->auto_flat;
    }
    if ((reinterpret_cast<uintptr_t>(d_stream) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15))
        return fail("pm_engine_scan_device: d_stream and d_out must be 16-byte aligned");
    const pm::Dict& d = *e->dict;
    bool kr_after = false;
    if (!d.sfx.fits_u16 && (algo == PM_ALGO_SFX || algo == PM_ALGO_KR)) {
        // no backward-scan tables for this dictionary (P + 2-byte continuations >= 65,536, or a pattern > 511 bytes):
        // the forward walkers serve it -- the compact-record walker when the automaton does not fit shared memory
        kr_after = algo == PM_ALGO_KR;
        if (kr_after && ensure_kr(e)) return -1;
        d.build_deep();
        uint32_t hot_rows = 0, hot_long = 0, fb_count = 0;
        pm::dfa_plan_hot(d.deep.n_states, d.sfx.log2_ncp, d.deep.depth_count.data(), uint32_t(d.deep.depth_count.size()), &hot_rows, &hot_long, &fb_count);
        force_flat = hot_rows < d.deep.n_states;
        algo = PM_ALGO_DFA;
    }
    if (algo == PM_ALGO_SFX) {
        pm::SfxParams p{};
        p.stream = d_stream; p.n = n; p.hist_valid = hist_valid; p.out = d_out;
        if (fill_sfx_params(e, &p, n, slot)) return -1;
        const bool ident = d.sfx.cls_identity;
        cudaEvent_t* ev = nullptr;
        if (e->profiling) {
            if (e->prof_used + 3 > e->prof_events.size()) {
                for (int k = 0; k < 3; ++k) {
                    cudaEvent_t x;
                    CU(cudaEventCreate(&x));
                    e->prof_events.push_back(x);
                }
            }
            ev = e->prof_events.data() + e->prof_used;
            e->prof_used += 3;
        }
        cudaError_t ce = pm::sfx_scan_launch(p, ident, e->n_sms, d.max_len, st, &e->launches, ev);
        if (ce != cudaSuccess) return cuda_fail(ce, "sfx_scan_launch");
        return 0;
    }
    if (algo == PM_ALGO_DFA && (force_flat || e->opts.dfa_deep) && !e->opts.dfa_flat) {
        // deep-match traffic on an automaton that does not fit shared memory: the compact goto + failure records
        e->dict->build_deep();
        if (e->dict->deep.usable) {
            if (ensure_deep(e)) return -1;
            pm::DeepParams p{};
            p.stream = d_stream; p.n = n; p.hist_valid = hist_valid; p.out = d_out;
            p.hot_rows = e->d_deep_hot; p.hot_longest = e->d_deep_long; p.n_hot = d.deep.n_hot;
            p.dense_end = d.deep.n_hot + d.deep.n_dense; p.n_small = d.deep.n_small;
            p.recs = e->d_deep_recs; p.dense_rows = e->d_deep_dense;
            p.warm = d.max_len ? d.max_len - 1 : 0;
            cudaError_t ce = pm::deep_scan_launch(p, e->n_sms, st, &e->launches);
            if (ce != cudaSuccess) return cuda_fail(ce, "deep_scan_launch");
            if (kr_after) {
                ce = pm::kr_scan_launch(e->kr, d_stream, n, hist_valid, d_out, e->pt, e->n_sms, st, &e->launches, !e->opts.kr_no_bulk);
                if (ce != cudaSuccess) return cuda_fail(ce, "kr_scan_launch");
            }
            return 0;
        }
    }
    if (algo == PM_ALGO_DFA) {
        if (ensure_dfa(e)) return -1;
        pm::DfaParams p{};
        p.stream = d_stream; p.n = n; p.hist_valid = hist_valid; p.out = d_out;
        p.delta = e->d_delta; p.longest = e->d_longest; p.cls = e->d_dfa_cls; p.log2_ncp = d.dfa.log2_ncp;
        p.warm = d.max_len ? d.max_len - 1 : 0;
        p.n_states = d.dfa.n_states; p.no_fused = e->opts.dfa_no_fused ? 1u : 0u;
        pm::dfa_plan_hot(d.dfa.n_states, d.dfa.log2_ncp, d.dfa.depth_count.data(), uint32_t(d.dfa.depth_count.size()),
                         &p.hot_rows, &p.hot_long, &p.fb_count);
        p.fb_meta = e->d_fb_meta;
        if (e->opts.dfa_no_fb) p.fb_count = 0;
        cudaError_t ce = pm::dfa_scan_launch(p, d.sfx.cls_identity, force_flat || e->opts.dfa_flat, e->n_sms, st, &e->launches);
        if (ce != cudaSuccess) return cuda_fail(ce, "dfa_scan_launch");
        if (kr_after) {
            ce = pm::kr_scan_launch(e->kr, d_stream, n, hist_valid, d_out, e->pt, e->n_sms, st, &e->launches, !e->opts.kr_no_bulk);
            if (ce != cudaSuccess) return cuda_fail(ce, "kr_scan_launch");
        }
        return 0;
    }
    if (algo == PM_ALGO_KR) {
        if (ensure_kr(e)) return -1;
        // The patterns of <= 8 bytes are matched exactly (bgps.c:459-464 does the same with its KMP): that part is the
        // exact scan, whose answer kr_scan_kernel reads only through short_of[] (the longest pattern of <= 8 bytes on
        // its PatternsTree chain); the patterns of > 8 bytes come from the fingerprints alone.  The exact scan's time is
        // part of every KR number reported.
        pm::SfxParams p{};
        p.stream = d_stream; p.n = n; p.hist_valid = hist_valid; p.out = d_out;
        if (fill_sfx_params(e, &p, n, slot)) return -1;
        cudaError_t ce = pm::sfx_scan_launch(p, d.sfx.cls_identity, e->n_sms, d.max_len, st, &e->launches);
        if (ce != cudaSuccess) return cuda_fail(ce, "sfx_scan_launch");
        ce = pm::kr_scan_launch(e->kr, d_stream, n, hist_valid, d_out, e->pt, e->n_sms, st, &e->launches, !e->opts.kr_no_bulk);
        if (ce != cudaSuccess) return cuda_fail(ce, "kr_scan_launch");
        return 0;
    }
    return fail("unknown algorithm id");
}

int ensure_compact_counts(pm_engine* e, size_t need) {
    if (need <= e->compact_cap) return 0;   // pooled: grows to the largest request seen and stays
    if (e->d_compact_counts) CU(cudaFree(e->d_compact_counts));
    e->scratch_bytes -= e->compact_cap * sizeof(unsigned long long);
    e->d_compact_counts = nullptr; e->compact_cap = 0;
    CU(cudaMalloc(reinterpret_cast<void**>(&e->d_compact_counts), need * sizeof(unsigned long long)));
    e->compact_cap = need;
    e->scratch_bytes += need * sizeof(unsigned long long);
    return 0;
}

// ---- 32-bit results: any number of patterns --------------------------------------------------------------------

constexpr size_t kSlice32 = size_t(64) << 20;   // positions per slice of a 32-bit scan (bounds the 16-bit scratch: 128 MiB)

int ensure_tmp16(pm_engine* e, size_t need) {
    if (need <= e->tmp16_cap) return 0;
    if (e->d_tmp16) { CU(cudaDeviceSynchronize()); CU(cudaFree(e->d_tmp16)); }
    e->scratch_bytes -= e->tmp16_cap * sizeof(uint16_t);
    e->d_tmp16 = nullptr; e->tmp16_cap = 0;
    CU(cudaMalloc(reinterpret_cast<void**>(&e->d_tmp16), need * sizeof(uint16_t)));
    e->tmp16_cap = need;
    e->scratch_bytes += need * sizeof(uint16_t);
    return 0;
}

// out32[i] = global pid of the longest pattern ending at i.  A single-part engine widens its 16-bit answer; a
// multi-part engine scans once per part and keeps the longer answer (pm::merge_parts_launch).
int scan_device32_impl(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid, uint32_t* d_out,
                       cudaStream_t st) {
    if (n == 0) return 0;
    if ((reinterpret_cast<uintptr_t>(d_stream) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15))
        return fail("pm_engine_scan_device32: d_stream and d_out must be 16-byte aligned");
    if (ensure_tmp16(e, std::min(n, kSlice32))) return -1;
    const pm::Dict& d = *e->dict;
    for (size_t off = 0; off < n; off += kSlice32) {
        const size_t len = std::min(kSlice32, n - off);
        const size_t hv = hist_valid + off;   // everything before the slice is readable
        if (e->parts.empty()) {
            if (scan_device_impl(e, algo, d_stream + off, len, hv, e->d_tmp16, st)) return -1;
            cudaError_t ce = pm::merge_parts_launch(d_out + off, e->d_tmp16, len, 0, nullptr, true, e->n_sms, st, &e->launches);
            if (ce != cudaSuccess) return cuda_fail(ce, "merge_parts_launch");
        } else {
            for (size_t k = 0; k < e->parts.size(); ++k) {
                pm_engine* part = e->parts[k];
                if (algo == PM_ALGO_KR) part->kr_seed = e->kr_seed;
                if (scan_device_impl(part, algo, d_stream + off, len, hv, e->d_tmp16, st)) return -1;
                cudaError_t ce = pm::merge_parts_launch(d_out + off, e->d_tmp16, len, d.part_first[k] - 1, e->d_glen, k == 0,
                                                        e->n_sms, st, &e->launches);
                if (ce != cudaSuccess) return cuda_fail(ce, "merge_parts_launch");
            }
        }
    }
    return 0;
}

// ---- host-buffer pipeline -------------------------------------------------------------------------------------

// what a host scan hands back for every position
struct HostSink {
    uint16_t* out16 = nullptr;          // dense pids, or
    uint64_t* out64 = nullptr;          // table[pid] (the plugin's pattern_id_t), 8 bytes per position
    const uint64_t* table = nullptr;
};

// remember the last `halo` bytes fed (right-aligned in h_hist): the stream state between calls
void carry_history(pm_engine* e, const uint8_t* stream, size_t n) {
    const size_t H = e->halo;
    if (n >= H) {
        memcpy(e->h_hist.data(), stream + n - H, H);
    } else if (n) {
        memmove(e->h_hist.data(), e->h_hist.data() + n, H - n);
        memcpy(e->h_hist.data() + H - n, stream, n);
    }
    e->hist_valid += n;
}

// Small calls: stage [history | bytes] in pinned memory, ONE H2D copy, ONE kernel that writes its results straight
// into mapped pinned memory, one stream synchronise (measured on the B200 box: 63 us per 100 KiB call, 14 us per
// read_char; with a D2H copy of the results instead of mapped stores: 77 us / 18 us).  Taken for PM_ALGO_SFX / PM_ALGO_AUTO (the walker is the backward
// scan's own bounded walk); an explicitly requested DFA or KR scan runs its own kernels whatever the size.
int scan_host_small(pm_engine* e, const uint8_t* stream, size_t n, const HostSink& sink) {
    const size_t H = e->halo;
    const size_t hv = std::min(e->hist_valid, H);
    uint8_t* h = e->h_in[0];
    memcpy(h + H - hv, e->h_hist.data() + H - hv, hv);
    memcpy(h + H, stream, n);
    CU(cudaMemcpyAsync(e->d_in[0] + H - hv, h + H - hv, hv + n, cudaMemcpyHostToDevice, e->st[0]));
    pm::SfxParams p{};
    p.stream = e->d_in[0] + H; p.n = n; p.hist_valid = hv; p.out = e->h_out[0];  // pinned host memory is device-addressable (UVA)
    if (fill_sfx_params(e, &p, n, 0)) return -1;
    cudaError_t ce = pm::sfx_walk_launch(p, e->st[0], &e->launches);
    if (ce != cudaSuccess) return cuda_fail(ce, "sfx_walk_launch");
    CU(cudaStreamSynchronize(e->st[0]));
    if (sink.out16) memcpy(sink.out16, e->h_out[0], n * sizeof(uint16_t));
    else if (n < (size_t(16) << 10)) pm::HostPool::expand_range(e->h_out[0], 0, n, sink.table, sink.out64);
    else {
        // 8 bytes per position: for the reference's 100 KiB chunks that is 800 KB per call -- on one thread a third of
        // the call's time; the pool's workers are still polling from the previous call when calls come back to back
        const uint16_t* res = e->h_out[0];
        const uint64_t* table = sink.table;
        uint64_t* dst = sink.out64;
        e->pool->run(n, size_t(8) << 10, [res, table, dst](size_t lo, size_t hi) { pm::HostPool::expand_range(res, lo, hi, table, dst); });
    }
    return 0;
}

// Host path of a multi-part engine: 8-byte ids only (16-bit pids cannot name its patterns).  Piece by piece and
// synchronous: copy in, one scan + merge per part, 32-bit global pids back, translated by the host threads.
int scan_host_ids_multi(pm_engine* e, int algo, const uint8_t* stream, size_t n, const HostSink& sink) {
    const size_t H = e->halo, chunk = e->opts.host_chunk;
    if (!e->d_out32) {
        CU(cudaMalloc(reinterpret_cast<void**>(&e->d_out32), chunk * sizeof(uint32_t)));
        CU(cudaMallocHost(reinterpret_cast<void**>(&e->h_out32), chunk * sizeof(uint32_t)));
        e->scratch_bytes += chunk * sizeof(uint32_t);
    }
    const bool in_pinned = is_pinned(stream);
    cudaStream_t st = e->st[0];
    for (size_t o = 0; o < n; o += chunk) {
        const size_t len = std::min(chunk, n - o);
        const size_t from_call = std::min(o, H), hist_total = std::min(e->hist_valid + o, H);
        uint8_t* din = e->d_in[0];
        if (from_call < H) CU(cudaMemcpyAsync(din, e->h_hist.data() + from_call, H - from_call, cudaMemcpyHostToDevice, st));
        const uint8_t* src = stream + o - from_call;
        if (!in_pinned) { e->pool->copy(e->h_in[0], src, from_call + len); src = e->h_in[0]; }
        CU(cudaMemcpyAsync(din + H - from_call, src, from_call + len, cudaMemcpyHostToDevice, st));
        if (scan_device32_impl(e, algo, din + H, len, hist_total, e->d_out32, st)) return -1;
        CU(cudaMemcpyAsync(e->h_out32, e->d_out32, len * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const uint32_t* res = e->h_out32;
        const uint64_t* table = sink.table;
        uint64_t* dst = sink.out64 + o;
        e->pool->run(len, size_t(64) << 10, [res, table, dst](size_t lo, size_t hi) { for (size_t j = lo; j < hi; ++j) dst[j] = table[res[j]]; });
    }
    carry_history(e, stream, n);
    return 0;
}

int scan_host_impl(pm_engine* e, int algo, const uint8_t* stream, size_t n, const HostSink& sink) {
    CU(cudaSetDevice(e->device));
    if (n == 0) return 0;
    if (ensure_pipe(e)) return -1;
    if (!e->parts.empty()) {
        if (!sink.out64) return fail("this dictionary has more than 65,535 patterns: 16-bit results cannot name them; use gpu_read_block / pm_engine_scan_host_ids / pm_engine_scan_device32");
        if (scan_host_ids_multi(e, algo, stream, n, sink)) { quiesce(e); return -1; }
        return 0;
    }
    if (algo == PM_ALGO_DFA && ensure_dfa(e)) return -1;
    if ((algo == PM_ALGO_KR || algo == PM_ALGO_MPBG) && ensure_kr(e)) return -1;
    if (n <= kSmallCall && (algo == PM_ALGO_SFX || algo == PM_ALGO_AUTO) && e->dict->sfx.fits_u16) {
        if (scan_host_small(e, stream, n, sink)) { quiesce(e); return -1; }
        carry_history(e, stream, n);
        return 0;
    }
    const size_t H = e->halo;
    const bool in_pinned = is_pinned(stream);
    const bool direct_out = sink.out16 && is_pinned(sink.out16);   // the D2H copy lands in the caller's buffer
    // 8-byte ids: translated by the host threads from the 2-byte pids (2 B per position over PCIe; the box's host writes
    // 107-131 GB/s of ids with 8-16 threads, scripts/microbench/host_mem.cpp), or -- into a page-locked buffer, when the
    // engine has a single host thread (or PM_HOST_IDS=device) -- translated on the device and sent by DMA (8 B per
    // position over PCIe: 6-7 GB/s of stream at most)
    const bool ids_pinned = sink.out64 && is_pinned(sink.out64) && (reinterpret_cast<uintptr_t>(sink.out64) & 15) == 0;
    // every dev_every-th piece is translated on the device (0 = none, 1 = all).  Default: none unless the engine has a
    // single host thread (eight ranks with 3 threads each on a 32-vCPU box: host 12.3 GB/s aggregate, device 6.4, every
    // 2nd / 3rd piece on the device 8.8 / 10.0 -- scripts/e2e_ranks.py).
    // Mixing the two was measured on the B200 box (1 GiB, page-locked buffers, 8 / 12 threads): host only 10.8 / 12.2 GB/s,
    // every 5th piece on the device 11.2 / 10.4, every 3rd 9.5 / 9.5, device only 6.3 / 6.3 -- the DMA writes of 8-byte
    // ids compete with the host threads for the same memory system, so the mix gains nothing.
    const size_t dev_every = !ids_pinned ? 0 : e->opts.ids_device_every >= 0 ? size_t(e->opts.ids_device_every)
                                             : (e->opts.host_threads < 2 ? 1 : 0);
    const bool device_ids = dev_every == 1;
    auto dev_piece = [&](size_t k) { return dev_every != 0 && k % dev_every == dev_every - 1; };
    if (dev_every != 0) {
        uint64_t sum = 0;
        const size_t tn = e->dict->pats.size() + 1;
        for (size_t i = 0; i < tn; ++i) sum = sum * 1099511628211ull + sink.table[i];
        if (!e->d_id_table || e->id_table_n != tn || e->id_table_src != sink.table || e->id_table_sum != sum) {
            if (e->d_id_table && e->id_table_n != tn) { CU(cudaFree(e->d_id_table)); e->d_id_table = nullptr; }
            if (!e->d_id_table) CU(cudaMalloc(reinterpret_cast<void**>(&e->d_id_table), tn * sizeof(uint64_t)));
            CU(cudaMemcpy(e->d_id_table, sink.table, tn * sizeof(uint64_t), cudaMemcpyHostToDevice));
            e->id_table_n = tn; e->id_table_src = sink.table; e->id_table_sum = sum;
        }
    }
    // What the host threads have to do per piece: stage its bytes into pinned memory (pageable stream) and unload its
    // results (translate to ids, or copy the pids into a pageable buffer).  Such calls are cut into at least ~8 pieces
    // (512 KiB .. 4 MiB) so that the exposed first stage-in and last unload stay a small part of the call.
    const bool do_stage = !in_pinned;
    const bool do_unload = !device_ids && (sink.out64 != nullptr || !direct_out);
    size_t chunk = e->opts.host_chunk;
    if (do_stage || do_unload) {
        chunk = std::min(chunk, kPageableChunk);
        while (chunk > (size_t(512) << 10) && n / chunk < 8) chunk >>= 1;
    }
    const size_t n_chunks = (n + chunk - 1) / chunk;
    if (dev_every != 0 && e->ids_cap < chunk) {   // per-slot id buffers of the pieces translated on the device
        quiesce(e);
        for (int b = 0; b < kSlots; ++b) {
            if (e->d_ids[b]) { CU(cudaFree(e->d_ids[b])); e->d_ids[b] = nullptr; }
            CU(cudaMalloc(reinterpret_cast<void**>(&e->d_ids[b]), chunk * sizeof(uint64_t)));
        }
        e->scratch_bytes += (chunk - e->ids_cap) * sizeof(uint64_t) * kSlots;
        e->ids_cap = chunk;
    }
    pm::HostPool& pool = *e->pool;
    // The pipeline: piece k is staged by the pool (job), then copied in / scanned / copied out on stream k % kSlots;
    // two pieces later its results have arrived and the pool unloads them (another job) while the GPU works on the next
    // pieces; its slot is reused by piece k + kSlots once that job is done.  Jobs are asynchronous (host_pool.hpp): the
    // calling thread only waits for the one it needs next, and works on it while it waits.
    constexpr size_t kLag = 2;   // pieces submitted to the GPU before the oldest one is unloaded
    static_assert(kLag < size_t(kSlots), "a slot's results are unloaded before the slot is reused");
    pm::HostPool::Ticket staged[kSlots], unloaded[kSlots];
    auto drain = [&]() {   // nothing of this call is in flight any more: streams idle, no job touches the caller's buffers
        quiesce(e);
        for (int b = 0; b < kSlots; ++b) { pool.wait(staged[b]); pool.wait(unloaded[b]); }
    };
    auto stage = [&](size_t k) {
        const size_t o = k * chunk, len = std::min(chunk, n - o), from_call = std::min(o, H);
        uint8_t* dst = e->h_in[k % kSlots];
        const uint8_t* src = stream + o - from_call;
        return pool.submit(from_call + len, size_t(256) << 10, [dst, src](size_t lo, size_t hi) { memcpy(dst + lo, src + lo, hi - lo); });
    };
    auto unload = [&](size_t k) {
        const size_t o = k * chunk, len = std::min(chunk, n - o);
        const uint16_t* res = e->h_out[k % kSlots];
        if (sink.out64) {
            uint64_t* dst = sink.out64 + o;
            const uint64_t* table = sink.table;
            return pool.submit(len, size_t(64) << 10, [res, table, dst](size_t lo, size_t hi) { pm::HostPool::expand_range(res, lo, hi, table, dst); });
        }
        uint16_t* dst = sink.out16 + o;
        return pool.submit(len, size_t(256) << 10, [res, dst](size_t lo, size_t hi) { memcpy(dst + lo, res + lo, (hi - lo) * sizeof(uint16_t)); });
    };
    auto submit = [&](size_t k) -> int {   // piece k is staged (or pinned in place): enqueue copy in, scan, copy out
        const int b = int(k % kSlots);
        const size_t o = k * chunk, len = std::min(chunk, n - o);
        // history in front of the chunk: from this call's own bytes when there are enough, else the carried tail
        const size_t from_call = std::min(o, H);
        const size_t hist_total = std::min(e->hist_valid + o, H);
        uint8_t* din = e->d_in[b];
        if (from_call < H)
            CU(cudaMemcpyAsync(din, e->h_hist.data() + from_call, H - from_call, cudaMemcpyHostToDevice, e->st[b]));
        const uint8_t* src = in_pinned ? stream + o - from_call : e->h_in[b];
        CU(cudaMemcpyAsync(din + H - from_call, src, from_call + len, cudaMemcpyHostToDevice, e->st[b]));
        if (scan_device_impl(e, algo, din + H, len, hist_total, e->d_out[b], e->st[b], b)) return -1;
        if (dev_piece(k)) {
            cudaError_t ce = pm::expand_ids_launch(e->d_out[b], len, e->d_id_table, e->d_ids[b], e->n_sms, e->st[b], &e->launches);
            if (ce != cudaSuccess) return cuda_fail(ce, "expand_ids_launch");
            CU(cudaMemcpyAsync(sink.out64 + o, e->d_ids[b], len * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->st[b]));
        } else {
            CU(cudaMemcpyAsync(direct_out ? sink.out16 + o : e->h_out[b], e->d_out[b], len * sizeof(uint16_t), cudaMemcpyDeviceToHost, e->st[b]));
        }
        CU(cudaEventRecord(e->done[b], e->st[b]));
        return 0;
    };
    for (size_t k = 0; k < n_chunks + kLag; ++k) {
        const int b = int(k % kSlots);
        // (a) stage piece k: its slot's pinned input was consumed by piece k - kSlots, whose completion was awaited at
        //     step k - kSlots + kLag
        if (k < n_chunks && do_stage) staged[b] = stage(k);
        // (b) piece k - kLag has left the device: hand its results to the pool
        if (k >= kLag) {
            const size_t ku = k - kLag;
            cudaError_t ce = cudaEventSynchronize(e->done[ku % kSlots]);
            if (ce != cudaSuccess) { drain(); return cuda_fail(ce, "cudaEventSynchronize"); }
            if (do_unload && !dev_piece(ku)) unloaded[ku % kSlots] = unload(ku);
        }
        // (c) piece k goes to the GPU once the previous user of its slot has been unloaded and its own bytes are staged
        if (k < n_chunks) {
            pool.wait(unloaded[b]); unloaded[b].reset();
            pool.wait(staged[b]); staged[b].reset();
            if (submit(k)) { drain(); return -1; }
        }
    }
    for (int b = 0; b < kSlots; ++b) pool.wait(unloaded[b]);
    carry_history(e, stream, n);
    return 0;
}

}  // namespace

extern "C" {

const char* pm_last_error(void) { return g_err.c_str(); }
int pm_version(void) { return 2; }

pm_dict* pm_dict_create(void) { return new (std::nothrow) pm_dict(); }
void pm_dict_free(pm_dict* d) { delete d; }
int pm_parse_pattern_line(const uint8_t* line, size_t n, uint8_t* out, size_t* out_len) {
    return pm::Dict::parse_line(line, n, out, out_len) ? 1 : 0;
}
int pm_dict_add_file(pm_dict* d, const char* path) {
    if (d->d.compiled) return fail("dictionary already compiled");
    if (d->d.add_file(path)) return fail(d->d.error);
    return 0;
}
int pm_dict_add_mem(pm_dict* d, const uint8_t* data, size_t n) {
    if (d->d.compiled) return fail("dictionary already compiled");
    return d->d.add_mem(data, n);
}
uint32_t pm_dict_add_pattern(pm_dict* d, const uint8_t* pat, size_t len, uint32_t file, uint32_t line, uint64_t user_id) {
    return d->d.add_pattern(pat, len, file, line, user_id);
}
int pm_dict_compile(pm_dict* d) {
    if (d->d.compile()) return fail(d->d.error);   // leaves `compiled` false when the dictionary is not usable
    return 0;
}
int pm_dict_save(const pm_dict* d, const char* path) {
    if (d->d.save(path)) return fail(std::string("pm_dict_save: cannot write ") + path);
    return 0;
}
pm_dict* pm_dict_load(const char* path) {
    pm_dict* d = new (std::nothrow) pm_dict();
    if (!d) return nullptr;
    if (d->d.load(path)) { fail(d->d.error); delete d; return nullptr; }
    return d;
}
// FNV-1a over the names' order, sizes and contents of the dictionary files: the cache key
static uint64_t hash_files(const char* const* paths, int n, bool* ok) {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t len) { const uint8_t* b = static_cast<const uint8_t*>(p); for (size_t i = 0; i < len; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    *ok = true;
    for (int i = 0; i < n; ++i) {
        FILE* f = fopen(paths[i], "rb");
        if (!f) { *ok = false; return 0; }
        uint8_t buf[1 << 16];
        size_t got; uint64_t total = 0;
        while ((got = fread(buf, 1, sizeof(buf), f)) > 0) { mix(buf, got); total += got; }
        fclose(f);
        mix(&total, 8); mix(&i, sizeof(i));
    }
    return h;
}
pm_dict* pm_dict_compile_files_cached(const char* const* paths, int n, const char* cache_dir) {
    bool ok = false;
    const uint64_t key = hash_files(paths, n, &ok);
    if (!ok) { fail("pm_dict_compile_files_cached: cannot read a dictionary file"); return nullptr; }
    char name[64];
    snprintf(name, sizeof(name), "/pmdict-%016llx.bin", (unsigned long long)key);
    const std::string file = std::string(cache_dir ? cache_dir : ".") + name;
    if (pm_dict* d = pm_dict_load(file.c_str())) return d;   // a missing, foreign, stale or damaged file just means "compile"
    pm_dict* d = pm_dict_create();
    for (int i = 0; i < n; ++i) if (pm_dict_add_file(d, paths[i])) { pm_dict_free(d); return nullptr; }
    if (pm_dict_compile(d)) { pm_dict_free(d); return nullptr; }
    d->d.save(file.c_str());  // best effort (written to a temporary name, then renamed): an unwritable cache directory only costs the next compile
    g_err.clear();
    return d;
}
int pm_dict_get_info(const pm_dict* d, pm_dict_info* info) {
    const pm::Dict& x = d->d;
    memset(info, 0, sizeof(*info));
    info->n_lines = x.n_lines; info->n_rejected = x.n_rejected; info->n_duplicates = x.n_dups;
    info->n_patterns = uint32_t(x.pats.size()); info->max_pat_len = x.max_len; info->total_pat_bytes = x.bytes.size();
    info->n_ac_states = x.n_ac_states; info->n_sfx_nodes = x.sfx.n_nodes; info->n_sfx_rows = x.sfx.n_rows;
    info->n_classes = x.sfx.n_classes; info->n_hot2_cont = x.sfx.n2_cont;
    info->table_bytes = x.sfx.root2.size() * 2 + x.sfx.rows.size() * 4 + x.sfx.row_best.size() * 4 + x.sfx.root1.size() * 4;
    return 0;
}
int pm_dict_pattern(const pm_dict* d, uint32_t pid, uint32_t* file, uint32_t* line, uint64_t* user_id,
                    uint32_t* parent_pid, uint32_t* len, const uint8_t** bytes) {
    if (pid == 0 || pid > d->d.pats.size()) return fail("pid out of range");
    const pm::Pattern& p = d->d.pats[pid - 1];
    if (file) *file = p.file;
    if (line) *line = p.line;
    if (user_id) *user_id = p.user;
    if (parent_pid) *parent_pid = p.parent;
    if (len) *len = p.len;
    if (bytes) *bytes = d->d.bytes.data() + p.off;
    return 0;
}
int pm_dict_table(const pm_dict* d, const char* name, const void** data, size_t* bytes) {
    const pm::Dict& x = d->d;
    if (!x.compiled) return fail("pm_dict_table: dictionary is not compiled");
    const std::string n = name ? name : "";
    auto give = [&](const auto& v) { *data = v.data(); *bytes = v.size() * sizeof(v[0]); return 0; };
    if (n == "sfx.root2") return give(x.sfx.root2);
    if (n == "sfx.rows") return give(x.sfx.rows);
    if (n.rfind("deep.", 0) == 0) {
        x.build_deep();
        if (!x.deep.usable) return fail("pm_dict_table: the automaton does not fit the deep-match layout");
        if (n == "deep.recs") return give(x.deep.recs);
        if (n == "deep.hot_rows") return give(x.deep.hot_rows);
        if (n == "deep.hot_longest") return give(x.deep.hot_longest);
        if (n == "deep.dense_rows") return give(x.deep.dense_rows);
    }
    return fail("pm_dict_table: unknown table " + n);
}
int pm_dict_is_pattern_suffix(const pm_dict* d, uint32_t first_pid, uint32_t second_pid) {
    return d->d.is_pattern_suffix(first_pid, second_pid) ? 1 : 0;
}

void* pm_host_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t ce = cudaMallocHost(&p, bytes ? bytes : 1);
    if (ce != cudaSuccess) { cuda_fail(ce, "cudaMallocHost"); return nullptr; }
    return p;
}
void pm_host_free(void* p) { if (p) cudaFreeHost(p); }
int pm_host_register(void* p, size_t bytes) {
    cudaError_t ce = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (ce == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaHostRegister");
    return 0;
}
int pm_host_unregister(void* p) {
    cudaError_t ce = cudaHostUnregister(p);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaHostUnregister");
    return 0;
}

static pm_engine* engine_create(const pm::Dict* dict, int device);
pm_engine* pm_engine_create(const pm_dict* dd, int device) {
    if (!dd || !dd->d.compiled) { fail("pm_engine_create: dictionary is not compiled"); return nullptr; }
    return engine_create(&dd->d, device);
}
static pm_engine* engine_create(const pm::Dict* dict, int device) {
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0) {
        g_err = std::string("pm_engine_create: no CUDA device (") + cudaGetErrorString(ce) + "); this engine has no CPU fallback";
        return nullptr;
    }
    if ((ce = cudaSetDevice(device)) != cudaSuccess) { cuda_fail(ce, "cudaSetDevice"); return nullptr; }
    pm_engine* e = new (std::nothrow) pm_engine();
    if (!e) return nullptr;
    e->dict = dict;
    e->device = device;
    e->opts = EngineOpts::from_env();
    cudaDeviceGetAttribute(&e->n_sms, cudaDevAttrMultiProcessorCount, device);
    const pm::Dict& d = *dict;
    e->halo = std::max<size_t>(pm::kHalo, (size_t(d.max_len ? d.max_len - 1 : 0) + 15) / 16 * 16);
    e->h_hist.assign(e->halo, 0);
    if (d.multi) {
        // more than 65,535 patterns: one complete sub-engine per part; this object keeps the stream state, the host
        // pipeline and the merge scratch (pattern lengths by global pid)
        std::vector<uint32_t> glen(d.pats.size() + 1, 0);
        for (size_t i = 0; i < d.pats.size(); ++i) glen[i + 1] = d.pats[i].len;
        bool ok = upload(glen, &e->d_glen, &e->table_bytes) == cudaSuccess &&
                  cudaEventCreateWithFlags(&e->scratch_free, cudaEventDisableTiming) == cudaSuccess;
        for (size_t k = 0; ok && k < d.parts.size(); ++k) {
            pm_engine* part = engine_create(d.parts[k].get(), device);
            if (!part) { ok = false; break; }
            e->parts.push_back(part);
        }
        if (!ok) { if (g_err.empty()) cuda_fail(cudaGetLastError(), "pm_engine_create"); pm_engine_free(e); return nullptr; }
        return e;
    }
    auto up = [&](auto& vec, auto** ptr) -> bool {
        cudaError_t r = upload(vec, ptr, &e->table_bytes);
        if (r != cudaSuccess) { cuda_fail(r, "table upload"); return false; }
        return true;
    };
    std::vector<uint8_t> cls(d.sfx.cls, d.sfx.cls + 256);
    const size_t P = d.pats.size();
    std::vector<uint32_t> off(P), len(P);
    std::vector<uint16_t> parent(P + 1, 0), chain(P + 1, 0);
    std::vector<uint64_t> pidhash(P + 1, 0);
    for (size_t i = 0; i < P; ++i) {
        off[i] = uint32_t(d.pats[i].off); len[i] = d.pats[i].len;
        parent[i + 1] = uint16_t(d.pats[i].parent); chain[i + 1] = uint16_t(d.pats[i].chain);
        pidhash[i + 1] = pm::splitmix64(((uint64_t(d.pats[i].file) + 1) << 32) | d.pats[i].line);
    }
    // summary fast path: the shortest patterns with at most one ancestor, up to kSummaryHotMax of them (and as many as fit
    // shared memory beside the 2-byte pid -> slot map)
    std::vector<uint16_t> hot_map(P + 1, 0xFFFFu);
    std::vector<uint64_t> hot_own, hot_anc;
    {
        std::vector<uint32_t> order;
        for (uint32_t pid = 1; pid <= P; ++pid) if (d.pats[pid - 1].chain <= 1) order.push_back(pid);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return d.pats[a - 1].len < d.pats[b - 1].len; });
        const size_t room = (size_t(200) << 10) > (P + 1) * 2 ? ((size_t(200) << 10) - (P + 1) * 2) / 16 : 0;
        const size_t n_hot = std::min<size_t>({order.size(), size_t(pm::kSummaryHotMax), room});
        for (size_t k = 0; k < n_hot; ++k) {
            const uint32_t pid = order[k], par = d.pats[pid - 1].parent;
            hot_map[pid] = uint16_t(k | (par ? 0x8000u : 0u));
            hot_own.push_back(pidhash[pid]);
            hot_anc.push_back(par ? pidhash[par] : 0);
        }
    }
    std::vector<uint8_t> padded_bytes(kPatPad + d.bytes.size() + 16, 0);
    std::copy(d.bytes.begin(), d.bytes.end(), padded_bytes.begin() + kPatPad);
    bool ok = up(d.sfx.root2, &e->d_root2) && up(d.sfx.root1, &e->d_root1) && up(d.sfx.rows, &e->d_rows) &&
              up(d.sfx.row_best, &e->d_row_best) && up(cls, &e->d_cls) && up(off, &e->d_pat_off) &&
              up(len, &e->d_pat_len) && up(padded_bytes, &e->d_pat_bytes) && up(parent, &e->d_parent) &&
              up(d.sfx.tail_rec, &e->d_tail_rec) && up(d.sfx.l3f, &e->d_l3f) &&
              up(d.anc_off, &e->d_anc_off) && up(d.anc_list, &e->d_anc_list) &&
              up(chain, &e->d_chain) && up(pidhash, &e->d_pidhash) &&
              up(hot_map, &e->d_hot_map) && up(hot_own, &e->d_hot_own) && up(hot_anc, &e->d_hot_anc);
    if (ok) {  // the rows table also as a linear texture: the scan kernel reads level 3 through the TEX pipe (sfx_scan.cu)
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = e->d_rows;
        rd.res.linear.desc = cudaCreateChannelDesc<unsigned int>(); rd.res.linear.sizeInBytes = d.sfx.rows.size() * sizeof(uint32_t);
        cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
        if (cudaCreateTextureObject(&e->rows_tex, &rd, &td, nullptr) != cudaSuccess) { e->rows_tex = 0; cudaGetLastError(); }
    }
    if (ok && cudaMalloc(reinterpret_cast<void**>(&e->d_acc), 8 * sizeof(unsigned long long)) != cudaSuccess) ok = false;
    if (ok && cudaMalloc(reinterpret_cast<void**>(&e->d_qcount), size_t(kSlots) * 2 * kMaxCtas * sizeof(uint32_t)) != cudaSuccess) ok = false;
    if (ok && cudaEventCreateWithFlags(&e->scratch_free, cudaEventDisableTiming) != cudaSuccess) ok = false;
    if (!ok) { if (g_err.empty()) cuda_fail(cudaGetLastError(), "pm_engine_create"); pm_engine_free(e); return nullptr; }
    e->pt.n_patterns = uint32_t(P);
    e->pt.off = e->d_pat_off; e->pt.len = e->d_pat_len; e->pt.bytes = e->d_pat_bytes + kPatPad;
    e->pt.parent = e->d_parent; e->pt.chain = e->d_chain; e->pt.pidhash = e->d_pidhash;
    e->pt.anc_off = e->d_anc_off; e->pt.anc_list = e->d_anc_list;
    e->pt.hot_map = e->d_hot_map; e->pt.hot_own = e->d_hot_own; e->pt.hot_anc = e->d_hot_anc; e->pt.n_hot = uint32_t(hot_own.size());
    return e;
}

void pm_engine_free(pm_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (pm_engine* part : e->parts) pm_engine_free(part);
    if (e->d_glen) cudaFree(e->d_glen);
    if (e->d_tmp16) cudaFree(e->d_tmp16);
    if (e->d_out32) cudaFree(e->d_out32);
    if (e->h_out32) cudaFreeHost(e->h_out32);
    e->pool.reset();
    if (e->rows_tex) cudaDestroyTextureObject(e->rows_tex);
    void* ptrs[] = {e->d_root2, e->d_root1, e->d_rows, e->d_row_best, e->d_cls, e->d_pat_off, e->d_pat_len, e->d_pat_bytes,
                    e->d_parent, e->d_chain, e->d_pidhash, e->d_hot_map, e->d_hot_own, e->d_hot_anc, e->d_tail_rec, e->d_l3f, e->d_anc_off, e->d_anc_list, e->d_sample_out, e->d_delta, e->d_longest, e->d_dfa_cls, e->d_fb_meta, e->d_deep_hot, e->d_deep_long, e->d_deep_recs, e->d_deep_dense, e->d_acc,
                    e->d_compact_counts, e->d_flags, e->d_id_table, e->d_qcount};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int b = 0; b < kSlots; ++b) {
        void* slot_ptrs[] = {e->d_ids[b], e->d_in[b], e->d_out[b], e->d_queue[b]};
        for (void* p : slot_ptrs) if (p) cudaFree(p);
        if (e->h_in[b]) cudaFreeHost(e->h_in[b]);
        if (e->h_out[b]) cudaFreeHost(e->h_out[b]);
        if (e->st[b]) cudaStreamDestroy(e->st[b]);
        if (e->done[b]) cudaEventDestroy(e->done[b]);
    }
    pm::kr_free_tables(&e->kr);
    for (cudaEvent_t x : e->prof_events) cudaEventDestroy(x);
    if (e->scratch_free) cudaEventDestroy(e->scratch_free);
    for (int b = 0; b < 2; ++b) {
        if (e->d_rec[b]) cudaFree(e->d_rec[b]);
        if (e->d_rec_counts[b]) cudaFree(e->d_rec_counts[b]);
        if (e->d_rec_flags[b]) cudaFree(e->d_rec_flags[b]);
        if (e->h_rec_total[b]) cudaFreeHost(e->h_rec_total[b]);
    }
    delete e;
}

size_t pm_engine_total_mem(const pm_engine* e) {
    if (!e) return 0;
    size_t sum = e->table_bytes;
    for (const pm_engine* part : e->parts) sum += part->table_bytes;
    return sum;
}
size_t pm_engine_scratch_mem(const pm_engine* e) { return e ? e->scratch_bytes : 0; }
int pm_engine_host_threads(const pm_engine* e) { return e ? e->opts.host_threads : 0; }
uint64_t pm_engine_launch_count(const pm_engine* e) { return e ? e->launches : 0; }
int pm_engine_auto_choice(const pm_engine* e) { return e->auto_choice < 0 ? -1 : (e->auto_choice == PM_ALGO_DFA && e->auto_flat ? 4 : e->auto_choice); }
uint64_t pm_engine_last_deferred(pm_engine* e) {
    std::lock_guard<std::mutex> lock(e->mu);
    const size_t ctas = std::min(e->last_ctas[0], kMaxCtas);   // only the CTAs of the last launch wrote their counters
    std::vector<uint32_t> c(2 * ctas + 1);
    cudaSetDevice(e->device);
    if (ctas == 0 || cudaMemcpy(c.data(), e->d_qcount, 2 * ctas * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
    uint64_t sum = 0;
    for (size_t i = 0; i < 2 * ctas; ++i) sum += c[i];
    return sum;
}
int pm_engine_set_profiling(pm_engine* e, int on) {
    std::lock_guard<std::mutex> lock(e->mu);
    e->profiling = on != 0;
    e->prof_used = 0;
    return 0;
}
int pm_engine_read_profile(pm_engine* e, uint32_t* n_scans, float* main_kernel_ms, float* total_ms) {
    std::lock_guard<std::mutex> lock(e->mu);
    CU(cudaSetDevice(e->device));
    CU(cudaDeviceSynchronize());
    float a = 0, b = 0;
    for (size_t i = 0; i + 3 <= e->prof_used; i += 3) {
        float x = 0, y = 0;
        CU(cudaEventElapsedTime(&x, e->prof_events[i], e->prof_events[i + 1]));
        CU(cudaEventElapsedTime(&y, e->prof_events[i], e->prof_events[i + 2]));
        a += x; b += y;
    }
    *n_scans = uint32_t(e->prof_used / 3);
    *main_kernel_ms = a; *total_ms = b;
    e->prof_used = 0;
    return 0;
}
int pm_engine_set_kr_seed(pm_engine* e, uint64_t seed) {
    std::lock_guard<std::mutex> lock(e->mu);
    e->kr_seed = seed;
    return 0;
}

int pm_engine_scan_device(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid,
                          uint16_t* d_out, void* cuda_stream) {
    std::lock_guard<std::mutex> lock(e->mu);
    CU(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (algo == PM_ALGO_AUTO) e->auto_choice = -1;  // device scans are independent calls: decide per call
    // successive scans share the deferred-walk queue and its counters: order them on the device even when the
    // caller alternates streams (no host synchronisation; a no-op when the stream is the same)
    if (e->scratch_used) CU(cudaStreamWaitEvent(st, e->scratch_free, 0));
    const int rc = scan_device_impl(e, algo, d_stream, n, hist_valid, d_out, st);
    CU(cudaEventRecord(e->scratch_free, st));
    e->scratch_used = true;
    return rc;
}

int pm_engine_scan_device32(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid,
                            uint32_t* d_out, void* cuda_stream) {
    std::lock_guard<std::mutex> lock(e->mu);
    CU(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (algo == PM_ALGO_AUTO) { e->auto_choice = -1; for (pm_engine* part : e->parts) part->auto_choice = -1; }
    if (e->scratch_used) CU(cudaStreamWaitEvent(st, e->scratch_free, 0));
    const int rc = scan_device32_impl(e, algo, d_stream, n, hist_valid, d_out, st);
    CU(cudaEventRecord(e->scratch_free, st));
    e->scratch_used = true;
    return rc;
}

void pm_engine_reset(pm_engine* e) {
    std::lock_guard<std::mutex> lock(e->mu);
    e->auto_choice = -1;  // a new stream: PM_ALGO_AUTO samples again
    for (pm_engine* part : e->parts) part->auto_choice = -1;
    e->hist_valid = 0;
    std::fill(e->h_hist.begin(), e->h_hist.end(), uint8_t(0));
}

int pm_engine_prepare_host(pm_engine* e) {
    std::lock_guard<std::mutex> lock(e->mu);
    CU(cudaSetDevice(e->device));
    return ensure_pipe(e);
}

int pm_engine_scan_host(pm_engine* e, int algo, const uint8_t* stream, size_t n, uint16_t* out) {
    std::lock_guard<std::mutex> lock(e->mu);
    HostSink sink;
    sink.out16 = out;
    return scan_host_impl(e, algo, stream, n, sink);
}

int pm_engine_scan_host_ids(pm_engine* e, int algo, const uint8_t* stream, size_t n, const uint64_t* id_of_pid,
                            size_t n_ids, uint64_t* out) {
    std::lock_guard<std::mutex> lock(e->mu);
    if (n_ids < e->dict->pats.size() + 1) return fail("pm_engine_scan_host_ids: id_of_pid needs P + 1 entries (entry 0 = no pattern)");
    HostSink sink;
    sink.out64 = out; sink.table = id_of_pid;
    return scan_host_impl(e, algo, stream, n, sink);
}

int pm_engine_scan_host_records(pm_engine* e, int algo, const uint8_t* stream, size_t n, uint32_t min_len,
                                uint64_t* records, size_t cap, uint64_t* n_records) {
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->parts.empty()) return fail("this dictionary has more than 65,535 patterns: only pm_engine_scan_device32, pm_engine_scan_host_ids and the plugin calls serve it");
    CU(cudaSetDevice(e->device));
    if (ensure_pipe(e)) return -1;
    if (algo == PM_ALGO_DFA && ensure_dfa(e)) return -1;
    if (algo == PM_ALGO_KR && ensure_kr(e)) return -1;
    const size_t chunk = e->opts.host_chunk, H = e->halo;
    const size_t blocks = pm::compact_blocks(chunk) + 1;
    for (int b = 0; b < 2; ++b) {   // each piece is checked on its own: a failed allocation leaves the others usable next time
        if (!e->d_rec[b]) CU(cudaMalloc(reinterpret_cast<void**>(&e->d_rec[b]), chunk * sizeof(uint64_t)));
        if (!e->d_rec_counts[b]) CU(cudaMalloc(reinterpret_cast<void**>(&e->d_rec_counts[b]), (blocks + 1) * sizeof(unsigned long long)));
        if (!e->h_rec_total[b]) CU(cudaMallocHost(reinterpret_cast<void**>(&e->h_rec_total[b]), sizeof(unsigned long long)));
        if (!e->d_rec_flags[b]) CU(cudaMalloc(reinterpret_cast<void**>(&e->d_rec_flags[b]), (chunk / 32 + 16) * 4));
    }
    const bool fused = algo == PM_ALGO_SFX && min_len >= 3 && e->dict->sfx.fits_u16;   // the scan kernel marks the qualifying positions itself
    const bool in_pinned = is_pinned(stream);
    const size_t n_chunks = (n + chunk - 1) / chunk;
    uint64_t produced = 0;
    auto finish = [&](size_t k) -> int {  // chunk k: fetch its record count, then exactly that many records
        const int b = int(k & 1);
        CU(cudaEventSynchronize(e->done[b]));
        const uint64_t cnt = *e->h_rec_total[b];
        const uint64_t room = produced < cap ? cap - produced : 0;
        const uint64_t take = cnt < room ? cnt : room;
        if (take) CU(cudaMemcpyAsync(records + produced, e->d_rec[b], take * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->st[b]));
        CU(cudaStreamSynchronize(e->st[b]));
        produced += cnt;
        return 0;
    };
    auto submit = [&](size_t k) -> int {
        const int b = int(k & 1);
        const size_t o = k * chunk, len = std::min(chunk, n - o);
        const size_t from_call = std::min(o, H);
        const size_t hist_total = std::min(e->hist_valid + o, H);
        uint8_t* din = e->d_in[b];
        if (from_call < H)
            CU(cudaMemcpyAsync(din, e->h_hist.data() + from_call, H - from_call, cudaMemcpyHostToDevice, e->st[b]));
        const uint8_t* src = stream + o - from_call;
        if (!in_pinned) {
            e->pool->copy(e->h_in[b], src, from_call + len);
            src = e->h_in[b];
        }
        CU(cudaMemcpyAsync(din + H - from_call, src, from_call + len, cudaMemcpyHostToDevice, e->st[b]));
        cudaError_t ce;
        if (fused) {
            pm::SfxParams p{};
            p.stream = din + H; p.n = len; p.hist_valid = hist_total; p.out = e->d_out[b];
            if (fill_sfx_params(e, &p, len, b)) return -1;
            p.flags = reinterpret_cast<uint8_t*>(e->d_rec_flags[b]); p.min_len = min_len;
            ce = pm::sfx_scan_launch(p, e->dict->sfx.cls_identity, e->n_sms, e->dict->max_len, e->st[b], &e->launches);
            if (ce == cudaSuccess)
                ce = pm::compact_bitmap_launch(e->d_rec_flags[b], e->d_out[b], len, e->hist_valid + o, e->d_rec_counts[b],
                                               e->d_rec_counts[b] + blocks, reinterpret_cast<unsigned long long*>(e->d_rec[b]),
                                               chunk, e->st[b], &e->launches);
        } else {
            if (scan_device_impl(e, algo, din + H, len, hist_total, e->d_out[b], e->st[b], b)) return -1;
            ce = pm::compact_launch(e->d_out[b], len, e->hist_valid + o, false, min_len, e->pt, e->d_rec_counts[b],
                                    e->d_rec_counts[b] + blocks, reinterpret_cast<unsigned long long*>(e->d_rec[b]),
                                    chunk, e->st[b], &e->launches);
        }
        if (ce != cudaSuccess) return cuda_fail(ce, "compact_launch");
        CU(cudaMemcpyAsync(e->h_rec_total[b], e->d_rec_counts[b] + blocks, sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->st[b]));
        CU(cudaEventRecord(e->done[b], e->st[b]));
        return 0;
    };
    for (size_t k = 0; k < n_chunks; ++k) {
        if (k >= 2 && finish(k - 2)) { quiesce(e); return -1; }
        if (submit(k)) { quiesce(e); return -1; }
    }
    for (size_t k = (n_chunks >= 2 ? n_chunks - 2 : 0); k < n_chunks; ++k)
        if (finish(k)) { quiesce(e); return -1; }
    carry_history(e, stream, n);
    *n_records = produced;
    return 0;
}

int pm_engine_summarize(pm_engine* e, const uint16_t* d_out, size_t n, uint64_t pos_base, uint64_t out4[4], void* cuda_stream) {
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->parts.empty()) return fail("this dictionary has more than 65,535 patterns: only pm_engine_scan_device32, pm_engine_scan_host_ids and the plugin calls serve it");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    cudaError_t ce = pm::summarize_launch(d_out, n, pos_base, e->pt, e->d_acc, e->n_sms, st, &e->launches);
    if (ce != cudaSuccess) return cuda_fail(ce, "summarize_launch");
    unsigned long long h[4];
    CU(cudaMemcpyAsync(h, e->d_acc, sizeof(h), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int k = 0; k < 4; ++k) out4[k] = h[k];
    return 0;
}

int pm_engine_classify(pm_engine* e, const uint16_t* d_algo, const uint16_t* d_real, size_t n, uint64_t counts4[4], void* cuda_stream) {
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->parts.empty()) return fail("this dictionary has more than 65,535 patterns: only pm_engine_scan_device32, pm_engine_scan_host_ids and the plugin calls serve it");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if ((reinterpret_cast<uintptr_t>(d_algo) & 15) || (reinterpret_cast<uintptr_t>(d_real) & 15))
        return fail("pm_engine_classify: both results must be 16-byte aligned");
    cudaError_t ce = pm::classify_launch(d_algo, d_real, n, e->pt, e->d_acc, e->n_sms, st, &e->launches);
    if (ce != cudaSuccess) return cuda_fail(ce, "classify_launch");
    unsigned long long h[4];
    CU(cudaMemcpyAsync(h, e->d_acc, sizeof(h), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int k = 0; k < 4; ++k) counts4[k] = h[k];
    return 0;
}

int pm_engine_compact(pm_engine* e, const uint16_t* d_out, size_t n, uint64_t pos_base, int expand_ancestors,
                      uint64_t* d_records, size_t cap, uint64_t* n_records, void* cuda_stream) {
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->parts.empty()) return fail("this dictionary has more than 65,535 patterns: only pm_engine_scan_device32, pm_engine_scan_host_ids and the plugin calls serve it");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (ensure_compact_counts(e, pm::compact_blocks(n) + 1)) return -1;
    cudaError_t ce = pm::compact_launch(d_out, n, pos_base, expand_ancestors != 0, 1, e->pt, e->d_compact_counts, e->d_acc + 4,
                                        reinterpret_cast<unsigned long long*>(d_records), cap, st, &e->launches);
    unsigned long long total = 0;
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(&total, e->d_acc + 4, sizeof(total), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) return cuda_fail(ce, "compact");
    *n_records = total;
    return 0;
}

int pm_engine_scan_device_records(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid,
                                  uint64_t pos_base, uint32_t min_len, uint16_t* d_out, uint64_t* d_records, size_t cap,
                                  uint64_t* n_records, void* cuda_stream) {
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->parts.empty()) return fail("this dictionary has more than 65,535 patterns: only pm_engine_scan_device32, pm_engine_scan_host_ids and the plugin calls serve it");
    CU(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    if (n == 0) { *n_records = 0; return 0; }
    if ((reinterpret_cast<uintptr_t>(d_stream) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15))
        return fail("pm_engine_scan_device_records: d_stream and d_out must be 16-byte aligned");
    if (e->scratch_used) CU(cudaStreamWaitEvent(st, e->scratch_free, 0));
    const bool fused = algo == PM_ALGO_SFX && min_len >= 3 && e->dict->sfx.fits_u16;
    int rc = 0;
    unsigned long long total = 0;
    cudaError_t ce = cudaSuccess;
    if (fused) {
        // the scan kernel itself marks the qualifying positions (one bit each); the compaction reads that bitmap and
        // only the flagged entries of the dense result
        const size_t words = (n + 31) / 32 + 16;
        if (words > e->flags_words) {
            if (e->d_flags) CU(cudaFree(e->d_flags));
            e->scratch_bytes -= e->flags_words * 4;
            e->d_flags = nullptr; e->flags_words = 0;
            CU(cudaMalloc(reinterpret_cast<void**>(&e->d_flags), words * 4));
            e->flags_words = words;
            e->scratch_bytes += words * 4;
        }
        if (ensure_compact_counts(e, pm::bitmap_blocks(n) + 1)) return -1;
        pm::SfxParams p{};
        p.stream = d_stream; p.n = n; p.hist_valid = hist_valid; p.out = d_out;
        if (fill_sfx_params(e, &p, n, 0)) return -1;
        p.flags = reinterpret_cast<uint8_t*>(e->d_flags); p.min_len = min_len;
        ce = pm::sfx_scan_launch(p, e->dict->sfx.cls_identity, e->n_sms, e->dict->max_len, st, &e->launches);
        if (ce == cudaSuccess)
            ce = pm::compact_bitmap_launch(e->d_flags, d_out, n, pos_base, e->d_compact_counts, e->d_acc + 4,
                                           reinterpret_cast<unsigned long long*>(d_records), cap, st, &e->launches);
    } else {
        if (ensure_compact_counts(e, pm::compact_blocks(n) + 1)) return -1;
        rc = scan_device_impl(e, algo, d_stream, n, hist_valid, d_out, st);
        if (rc == 0)
            ce = pm::compact_launch(d_out, n, pos_base, false, min_len ? min_len : 1, e->pt, e->d_compact_counts, e->d_acc + 4,
                                    reinterpret_cast<unsigned long long*>(d_records), cap, st, &e->launches);
    }
    if (rc == 0 && ce == cudaSuccess) ce = cudaMemcpyAsync(&total, e->d_acc + 4, sizeof(total), cudaMemcpyDeviceToHost, st);
    if (rc == 0 && ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaEventRecord(e->scratch_free, st);
    e->scratch_used = true;
    if (rc) return rc;
    if (ce != cudaSuccess) return cuda_fail(ce, "pm_engine_scan_device_records");
    *n_records = total;
    return 0;
}

int pm_engine_generate(pm_engine* e, int kind, uint64_t off, size_t n, uint8_t* d_dst, void* cuda_stream) {
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->parts.empty()) return fail("this dictionary has more than 65,535 patterns: only pm_engine_scan_device32, pm_engine_scan_host_ids and the plugin calls serve it");
    CU(cudaSetDevice(e->device));
    if ((off & 4095) || (n & 4095)) return fail("pm_engine_generate: off and n must be multiples of 4096");
    cudaError_t ce = pm::generate_launch(kind, off, n, d_dst, e->pt, static_cast<cudaStream_t>(cuda_stream), &e->launches);
    if (ce != cudaSuccess) return cuda_fail(ce, "generate_launch");
    return 0;
}

int pm_engine_generate_host(pm_engine* e, int kind, uint64_t off, size_t n, uint8_t* dst) {
    std::lock_guard<std::mutex> lock(e->mu);
    if (!e->parts.empty()) return fail("this dictionary has more than 65,535 patterns: only pm_engine_scan_device32, pm_engine_scan_host_ids and the plugin calls serve it");
    CU(cudaSetDevice(e->device));
    if ((off & 4095) || (n & 4095)) return fail("pm_engine_generate_host: off and n must be multiples of 4096");
    if (ensure_pipe(e)) return -1;
    const size_t chunk = e->opts.host_chunk & ~size_t(4095);
    for (size_t o = 0; o < n; o += chunk) {   // piece by piece through the pipeline's first device buffer
        const size_t len = std::min(chunk, n - o);
        cudaError_t ce = pm::generate_launch(kind, off + o, len, e->d_in[0], e->pt, e->st[0], &e->launches);
        if (ce != cudaSuccess) return cuda_fail(ce, "generate_launch");
        CU(cudaMemcpyAsync(dst + o, e->d_in[0], len, cudaMemcpyDeviceToHost, e->st[0]));
        CU(cudaStreamSynchronize(e->st[0]));
    }
    return 0;
}

int pm_engine_time_scan(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid, uint16_t* d_out,
                        int iters, float* ms_per_scan, void* cuda_stream) {
    std::lock_guard<std::mutex> lock(e->mu);
    CU(cudaSetDevice(e->device));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    cudaEvent_t a = nullptr, b = nullptr;
    int rc = -1;
    float ms = 0;
    cudaError_t ce = cudaEventCreate(&a);
    if (ce == cudaSuccess) ce = cudaEventCreate(&b);
    if (ce == cudaSuccess) ce = cudaEventRecord(a, st);
    if (ce == cudaSuccess) {
        rc = 0;
        for (int i = 0; i < iters && rc == 0; ++i) rc = scan_device_impl(e, algo, d_stream, n, hist_valid, d_out, st);
        if (rc == 0) {
            ce = cudaEventRecord(b, st);
            if (ce == cudaSuccess) ce = cudaEventSynchronize(b);
            if (ce == cudaSuccess) ce = cudaEventElapsedTime(&ms, a, b);
        }
    }
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    if (ce != cudaSuccess) return cuda_fail(ce, "pm_engine_time_scan");
    if (rc) return rc;
    *ms_per_scan = ms / float(iters > 0 ? iters : 1);
    return 0;
}

}  // extern "C"
