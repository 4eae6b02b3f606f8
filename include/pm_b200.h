/*
 * pm_b200.h -- C-ABI of the B200-native dictionary-matching engine (libpm_b200.so).
 *
 * Plain pointers and sizes only; no torch / C++ types.  Every entry point names the reference
 * interface it replaces (paths relative to the reference repository yehonatan145/PatternMatching).
 * The product has NO CPU fallback: every scan entry point fails (non-zero return / pm_last_error)
 * when no CUDA device is usable.
 *
 * Pattern identity.  The reference identifies a pattern by a PatternsTreeNode* whose payload is
 * PatternInternalID{file_number, line_number} of the FIRST occurrence of the byte string
 * (Core/src/PatternsTree.h:25-28, 90-106; PatternsTree.c:193-196, 274-283).  On the device a pattern
 * is a dense "pid": 0 = no pattern (null_pattern_id), 1..P in order of first occurrence, which is
 * the (file,line)-sorted order.  pm_dict_pattern() maps pid -> (file, line, user id).
 */
#ifndef PM_B200_H
#define PM_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pm_dict pm_dict;     /* host side: patterns + compiled tables */
typedef struct pm_engine pm_engine; /* device side: tables resident in HBM + scan state */

/* last error message of the calling thread ("" if none) */
const char* pm_last_error(void);
int pm_version(void);

/* ------------------------------------------------------------------------------------------------
 * Dictionary compilation on the host
 * ---------------------------------------------------------------------------------------------- */
/* replaces: ac_create (Core/src/mpac.c:236-247) + the PatternsTree ingest state */
pm_dict* pm_dict_create(void);
void pm_dict_free(pm_dict* d);
/* One .dict line -> raw bytes; grammar of parse_pattern_from_line (Core/src/parser.c:63-99) incl.
 * its rejection rules.  Returns 1 and the pattern in out[0..*out_len) (capacity >= n), 0 if rejected. */
int pm_parse_pattern_line(const uint8_t* line, size_t n, uint8_t* out, size_t* out_len);
/* replaces: fpt_fill_with_dict_file (Core/src/PatternsTree.c:260-291) for the next -d file;
 * file_number = number of previous pm_dict_add_file/_mem calls.  Returns 0, or -1 (cannot read). */
int pm_dict_add_file(pm_dict* d, const char* path);
int pm_dict_add_mem(pm_dict* d, const uint8_t* data, size_t n);
/* replaces: MpsElem.add_pattern (Core/src/mps.h:74; ac_add_pattern mpac.c:257-273).  The bytes are
 * copied (they are binary and may contain 0x00, Core/src/README.md:85-89).  `user_id` is an opaque
 * 64-bit value returned by pm_dict_pattern (the shim stores the pattern_id_t pointer there).
 * Returns the pid, or the pid of the identical pattern added earlier (de-dup keeps the first). */
uint32_t pm_dict_add_pattern(pm_dict* d, const uint8_t* pat, size_t len, uint32_t file, uint32_t line, uint64_t user_id);
/* replaces: MpsElem.compile (Core/src/mps.h:75; ac_compile mpac.c:282-291, add_failure_links :187-210)
 * and convert_fpt_to_patterns_tree (PatternsTree.c:416-428) for the suffix-parent relation. */
int pm_dict_compile(pm_dict* d);

/* Compiled-automaton cache (the reference rebuilds its structures on every run: 6-7 s PatternsTree,
 * PatternsTree.c:186-214, + 4-5 s ac_compile, mpac.c:282-291).  pm_dict_save writes a compiled dictionary to one
 * binary file (temporary name + rename: concurrent processes never see a partial file; versioned magic, trailing
 * checksum), pm_dict_load reads it back and verifies the checksum and every table size / index bound (NULL +
 * pm_last_error on a foreign, stale, truncated or damaged file); pm_dict_compile_files_cached keys the file by a
 * hash of the dictionary files' contents and order (<cache_dir>/pmdict-<hash>.bin): load when present and valid,
 * else ingest + compile + save. */
int pm_dict_save(const pm_dict* d, const char* path);
pm_dict* pm_dict_load(const char* path);
pm_dict* pm_dict_compile_files_cached(const char* const* paths, int n, const char* cache_dir);

typedef struct {
    uint64_t n_lines, n_rejected, n_duplicates; /* ingest accounting */
    uint32_t n_patterns;                        /* unique patterns P */
    uint32_t max_pat_len;                       /* conf->max_pat_len, PatternsTree.c:310 */
    uint64_t total_pat_bytes;
    uint32_t n_ac_states;                       /* forward trie states incl. root == reference n_states */
    uint32_t n_sfx_nodes;                       /* nodes of the reversed-pattern (suffix) trie incl. root */
    uint32_t n_sfx_rows;                        /* suffix-trie nodes that own a transition row */
    uint32_t n_classes;                         /* byte equivalence classes (alphabet compression) */
    uint32_t n_hot2_cont;                       /* 2-byte suffixes that continue below the shared-memory table */
    uint64_t table_bytes;                       /* bytes of all device tables */
} pm_dict_info;
int pm_dict_get_info(const pm_dict* d, pm_dict_info* info);
/* pid in 1..P -> identity; parent_pid = longest proper suffix that is itself a pattern (0 = none),
 * i.e. PatternsTreeNode.parent (Core/src/PatternsTree.h:90-94). */
int pm_dict_pattern(const pm_dict* d, uint32_t pid, uint32_t* file, uint32_t* line, uint64_t* user_id,
                    uint32_t* parent_pid, uint32_t* len, const uint8_t** bytes);
/* Read-only view of one compiled table, for inspection and for the tests' host-side emulation of the kernels' table
 * walks: "sfx.root2", "sfx.rows", "deep.recs", "deep.hot_rows", "deep.hot_longest", "deep.dense_rows" (layouts in
 * patternmatching_b200/csrc/dict.hpp).  The pointer stays valid until pm_dict_free. */
int pm_dict_table(const pm_dict* d, const char* name, const void** data, size_t* bytes);
/* replaces: is_pattern_suffix (Core/src/PatternsTree.c:485-494) on pids */
int pm_dict_is_pattern_suffix(const pm_dict* d, uint32_t first_pid, uint32_t second_pid);

/* ------------------------------------------------------------------------------------------------
 * Engine: tables in HBM, scan kernels
 * ---------------------------------------------------------------------------------------------- */
enum {
    PM_ALGO_SFX = 0, /* exact: per-position backward walk of the reversed-pattern trie (default) */
    PM_ALGO_DFA = 1, /* exact: per-thread forward walk of the flat Aho-Corasick DFA */
    PM_ALGO_KR  = 2, /* randomized: Karp-Rabin suffix-stage fingerprints (mpbg/bgps/kmprt style) */
    PM_ALGO_AUTO = 3, /* exact: SFX unless a sample of the stream shows deep walks (then DFA); see pm_engine_auto_choice */
    /* The reference's MPBG as shipped (Core/src/mpbg.c:132-145), position for position: its fingerprint stages never fire
     * (SURVEY Q5), so it reports the longest pattern of <= 8 bytes ending at each position (the patterns its exact KMP
     * path handles, Core/src/bgps.c:459-464).  Deterministic, pinned by tests/golden/ref_snort.json (mpbg_file / mpbg_line:
     * the reference's own per-position output).  PM_ALGO_KR is the repaired variant (patterns > 8 bytes by fingerprints). */
    PM_ALGO_MPBG = 5
};
enum {
    PM_STREAM_UNIFORM = 0, /* uniform bytes (splitmix64 counter generator) */
    PM_STREAM_PLANTED = 1, /* uniform background + one planted pattern per 4096-byte block */
    PM_STREAM_ALMOST  = 2, /* concatenated random-length pattern prefixes (write_first_lines.py semantics) */
    PM_STREAM_AB      = 3, /* {a,b} with P(a)=0.75 (adversarial small alphabet) */
    PM_STREAM_ASCII   = 4  /* uniform printable ASCII 0x20..0x7E (text-like traffic) */
};

/* Thread safety: a pm_dict is read-only once compiled, with one exception that is internally locked: the forward DFA
 * tables are built on first use (PM_ALGO_DFA / PM_ALGO_AUTO), at most once, under a dictionary-level mutex.  Karp-Rabin
 * tables are built per engine (per seed) and owned by the engine.  The pm_engine entry points are serialised per engine
 * by an internal mutex (an engine holds ONE stream state and shared scratch buffers) -- use one engine per concurrent
 * stream; engines of the same dictionary share nothing on the device but cost only 61 MB each.
 * Diagnostic environment switches are read once, when the engine is created.
 * Upload the compiled tables to CUDA device `device`.  NULL (and pm_last_error) when there is no
 * usable device: there is no CPU fallback. */
pm_engine* pm_engine_create(const pm_dict* d, int device);
void pm_engine_free(pm_engine* e);
/* replaces: MpsElem.total_mem (Core/src/mps.h:77): bytes of all device-resident tables */
size_t pm_engine_total_mem(const pm_engine* e);
/* device scratch that is not a table: deferred-walk queues (at most 2 MiB per SM and pipeline slot), the host
 * pipeline's device buffers, compaction counters.  Grows on demand, never shrinks. */
size_t pm_engine_scratch_mem(const pm_engine* e);
/* host threads used to stage pageable buffers and to translate pids (PM_HOST_THREADS; default = three quarters of this rank's share of the cores, at most 16) */
int pm_engine_host_threads(const pm_engine* e);
/* KR variant parameters: r is drawn from `seed` (the reference draws it from rand(), bgps.c:469-475) */
int pm_engine_set_kr_seed(pm_engine* e, uint64_t seed);

/* Scan n bytes that are already in device memory.  Asynchronous on `cuda_stream`, with these exceptions: the call
 * synchronises the device when it has to grow the deferred-walk queue (first call, or a larger n than ever before),
 * when it builds / uploads the DFA or KR tables (first use), and for PM_ALGO_AUTO (it reads back a sample).  An engine
 * owns ONE set of scan scratch: successive calls are ordered on the device by an internal event, so they may be given
 * different streams, but they never overlap -- use one engine per concurrent scan.
 *   d_stream   : device pointer, 16-byte aligned, first byte to report on
 *   hist_valid : how many bytes directly BEFORE d_stream belong to the same stream and are readable
 *                (0 = d_stream is the start of the stream / of the allocation).  With
 *                hist_valid >= max_pat_len-1 the result equals the continuous scan.
 *   d_out      : device pointer to n uint16 pids: d_out[i] = pid of the LONGEST pattern that is a
 *                suffix of stream[..i] (what read_char returns, Core/src/mps.h:41-42), 0 = none
 *   cuda_stream: a cudaStream_t (0 = default stream); the call is asynchronous on it.
 * replaces: the per-byte loop `algo_results[j] = read_char_func(obj, stream_buffer[j])`
 *           (Core/src/measure.c:292-294) over ac_read_char (Core/src/mpac.c:304-319). */
int pm_engine_scan_device(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid,
                          uint16_t* d_out, void* cuda_stream);

/* Same scan with 32-bit results: d_out[i] = pid of the longest pattern ending at i as a 32-bit number, for dictionaries of
 * ANY size.  A dictionary of more than 65,535 unique patterns (Core/src/mpac.c:257-291 accepts any number) is compiled
 * into parts of at most 49,152 patterns; the engine scans once per part and keeps the longer answer per position.  Such
 * an engine refuses the 16-bit entry points (pm_engine_scan_device, pm_engine_scan_host, records, summary) with a
 * message; pm_engine_scan_host_ids / gpu_read_block (8-byte ids) serve it like any other.  d_out 16-byte aligned. */
int pm_engine_scan_device32(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid,
                            uint32_t* d_out, void* cuda_stream);

/* Scan a HOST buffer: the call is cut into pieces that move through a four-slot pipeline (H2D copy, scan, D2H of the
 * dense uint16 result, each piece on its own CUDA stream); synchronous.  State is carried across calls exactly like
 * consecutive read_char calls until pm_engine_reset().  replaces: the chunk loop of measure_single_instance_stats
 * (measure.c:281-304).  Page-locked buffers (pm_host_alloc) are used in place; pageable ones are staged through the
 * engine's pinned buffers by its host threads (asynchronous jobs, 512 KiB .. 4 MiB pieces) while the transfers run.  Calls of <= 256 KiB (the reference's
 * 100 KiB chunks, measure.c:77; read_char) take a latency path: one H2D copy, one kernel launch writing into mapped
 * pinned memory, one synchronise. */
int pm_engine_scan_host(pm_engine* e, int algo, const uint8_t* stream, size_t n, uint16_t* out);
/* Same pipeline; out[i] = id_of_pid[pid of the longest pattern ending at i] -- 8 bytes per position, what the
 * reference's read_char returns (pattern_id_t is a pointer, Core/src/PatternsTree.h:104; Core/src/mps.h:41-42).
 * id_of_pid has n_ids >= P + 1 entries, entry 0 = the "no pattern" id.  The translation runs on the engine's host
 * threads piece by piece while later pieces are still on the GPU (2 B per position cross PCIe; 13-16 GB/s of stream on
 * the B200 box's host).  With PM_HOST_IDS=device -- or when the engine has a single host thread -- and a page-locked,
 * 16-byte aligned `out`, the translation happens on the device instead and the ids arrive by DMA (8 B per position
 * over PCIe: 6-7 GB/s at best, but no host thread works).  This is what gpu_read_block calls. */
int pm_engine_scan_host_ids(pm_engine* e, int algo, const uint8_t* stream, size_t n, const uint64_t* id_of_pid,
                            size_t n_ids, uint64_t* out);
/* Same pipeline, sparse result: only the positions whose longest match is a pattern of at least min_len bytes
 * come back, as position-sorted records (pos << 24 | pid), pos counted from the last pm_engine_reset().  The dense
 * result never crosses PCIe (2 B per stream byte down to 8 B per reported match).  At most `cap` records are
 * written; *n_records is the number found.  State is carried across calls like pm_engine_scan_host. */
int pm_engine_scan_host_records(pm_engine* e, int algo, const uint8_t* stream, size_t n, uint32_t min_len,
                                uint64_t* records, size_t cap, uint64_t* n_records);
/* Allocate the host pipeline (page-locked staging buffers, device buffers, CUDA streams, the host threads) now instead of
 * inside the first host call -- gpu_compile does, so that the reference's timed loop (measure.c:290-297) does not pay for it. */
int pm_engine_prepare_host(pm_engine* e);
/* replaces: MpsElem.reset (Core/src/mps.h:78; ac_reset mpac.c:339-342) */
void pm_engine_reset(pm_engine* e);

/* Device-side reduction of a dense result: positions with a match, matches including PatternsTree
 * ancestors, and the order-independent digest sums of (global position, file, line) defined in
 * oracle/pm_oracle.h (match_digest).  pos_base = global position of d_out[0].  Synchronous.
 * out4 = {positions, matches, hsum_longest, hsum_all}. */
int pm_engine_summarize(pm_engine* e, const uint16_t* d_out, size_t n, uint64_t pos_base, uint64_t out4[4],
                        void* cuda_stream);

/* replaces: measure_success_rate (Core/src/measure.c:174-190) on two dense results in device memory (both 16-byte
 * aligned): d_algo is what a matcher reported, d_real what the reliable one reported.  counts4 = {success, partial success
 * (the reported pattern is a PatternsTree ancestor of the real one), false negatives (nothing reported), false positives}.
 * Synchronous. */
int pm_engine_classify(pm_engine* e, const uint16_t* d_algo, const uint16_t* d_real, size_t n, uint64_t counts4[4], void* cuda_stream);

/* Compact a dense result into position-sorted (pos, pid) records (pos = pos_base + i, 40 bits;
 * pid 24 bits; record = pos << 24 | pid), longest match per position; with expand_ancestors != 0
 * one record per match incl. PatternsTree ancestors (longest first).  d_records has capacity `cap`
 * records; *n_records receives the number produced (may exceed cap: nothing is written past cap).
 * Synchronous. */
int pm_engine_compact(pm_engine* e, const uint16_t* d_out, size_t n, uint64_t pos_base, int expand_ancestors,
                      uint64_t* d_records, size_t cap, uint64_t* n_records, void* cuda_stream);

/* Sparse mode on device-resident data: scan and return only the positions whose longest match is a pattern of at least
 * min_len bytes, as position-sorted records (pos_base + i) << 24 | pid in d_records (capacity `cap`; *n_records = number
 * found, may exceed cap: nothing is written past cap).  The dense result is written to d_out as by pm_engine_scan_device.
 * PM_ALGO_SFX with min_len >= 3: the scan kernel itself marks the qualifying positions in a 1-bit-per-position bitmap
 * (row entries carry the pattern length, so the test is one compare per position) and the compaction reads that bitmap and
 * the flagged entries only -- the dense result is not read again.  Other algorithms / min_len < 3: dense scan followed by
 * the dense compaction.  Synchronous (returns the count). */
int pm_engine_scan_device_records(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid,
                                  uint64_t pos_base, uint32_t min_len, uint16_t* d_out, uint64_t* d_records, size_t cap,
                                  uint64_t* n_records, void* cuda_stream);

/* Seeded synthetic streams generated directly in HBM (definitions: SURVEY.md 8d, oracle/pm_oracle.c):
 * writes bytes [off, off+n) of stream `kind` to d_dst.  off and n multiples of 4096. */
int pm_engine_generate(pm_engine* e, int kind, uint64_t off, size_t n, uint8_t* d_dst, void* cuda_stream);

/* Same streams into a HOST buffer (generated on the device piece by piece and copied out): what a tool that writes
 * .stream files needs (pm_driver -g; the reference's Streams/write_first_lines.py:48-61 is unseeded Python 2). */
int pm_engine_generate_host(pm_engine* e, int kind, uint64_t off, size_t n, uint8_t* dst);

/* Timing helper used by bench.py: runs pm_engine_scan_device `iters` times on `cuda_stream`
 * bracketed by CUDA events recorded ON THAT STREAM and returns the mean milliseconds per scan. */
int pm_engine_time_scan(pm_engine* e, int algo, const uint8_t* d_stream, size_t n, size_t hist_valid,
                        uint16_t* d_out, int iters, float* ms_per_scan, void* cuda_stream);
/* Per-kernel timing of the exact (sfx) scan for the roofline: while profiling is on, every scan records CUDA
 * events on its own stream around the dominant kernel; pm_engine_read_profile synchronises the device and
 * returns the number of profiled scans with the summed milliseconds of the dominant kernel and of the whole
 * scans (dominant + deferred-walk + start-of-stream kernels), then clears the record. */
int pm_engine_set_profiling(pm_engine* e, int on);
int pm_engine_read_profile(pm_engine* e, uint32_t* n_scans, float* main_kernel_ms, float* total_ms);
/* what PM_ALGO_AUTO decided for the current stream: -1 undecided, PM_ALGO_SFX, PM_ALGO_DFA (hot rows), 4 = DFA flat */
int pm_engine_auto_choice(const pm_engine* e);
/* queue slots reserved for deferred deep walks by the last pm_engine_scan_device (sfx) call; synchronises */
uint64_t pm_engine_last_deferred(pm_engine* e);
/* Page-locked host buffers for pm_engine_scan_host (a pageable buffer is staged through internal ones). */
void* pm_host_alloc(size_t bytes);
void pm_host_free(void* p);
/* Page-lock a buffer the caller already owns and reuses for every chunk (the reference's stream_buffer / algo_results
 * arrays, Core/src/measure.c:243-245, when they are static or heap memory that lives as long as the matcher): after
 * this the host calls use it in place like pm_host_alloc memory.  The caller must unregister before freeing it. */
int pm_host_register(void* p, size_t bytes);
int pm_host_unregister(void* p);
/* number of kernel launches issued by this engine since creation (bench.py's gpu_launches) */
uint64_t pm_engine_launch_count(const pm_engine* e);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, the stream cut into contiguous shards that each read max_pat_len-1 bytes of history
 * (hist_valid) -- no exchange step in the scan (Core/src/measure.c:262-306 keeps state across chunks, a
 * warmed-up shard reports the same matches).  The only communication is the gather of the per-rank, position-sorted
 * record lists to one rank over NVLink: an NCCL all-gather of the counts, then one peer-to-peer copy per rank into the
 * root's CUDA-IPC-mapped staging buffer (grouped ncclSend / ncclRecv when the GPUs cannot map each other's memory, or with
 * PM_COMM_NO_P2P=1); rank order is position order, so the concatenation is sorted.  NCCL is dlopen'ed at first use
 * (libnccl.so.2 of the process), the library has no link-time dependency on it.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pm_comm pm_comm;
const char* pm_comm_last_error(void);
/* rank 0 creates the 128-byte NCCL unique id and hands it to the other ranks by any means (bench.py: torch.distributed) */
int pm_comm_unique_id(uint8_t id[128]);
pm_comm* pm_comm_create(const uint8_t id[128], int rank, int world, int device);
void pm_comm_free(pm_comm* c);
/* Every rank passes its n_local sorted records (device memory).  On return counts[0..world) (host, may be NULL) holds every
 * rank's count and *n_all their sum; on `root`, d_all (device, capacity `cap` records) receives the concatenation in rank
 * order -- enqueued on cuda_stream: synchronise it (or record an event) before reading.  The call itself synchronises
 * the stream once, after the all-gather of the counts (they size the receives). */
int pm_comm_gather_records(pm_comm* c, const uint64_t* d_local, uint64_t n_local, uint64_t* d_all, uint64_t cap,
                           uint64_t* counts, uint64_t* n_all, int root, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------
 * The reference plugin surface (Core/src/mps.h:71-80) -- implemented in mps_gpu_shim.c on top of
 * the calls above.  Types are spelled with void* / char* exactly as in MpsElem so that the
 * reference can register it unchanged.
 * ---------------------------------------------------------------------------------------------- */
void* gpu_create(void);                                                   /* MpsElem.create      */
void gpu_add_pattern(void* obj, char* pat, size_t len, void* pattern_id); /* MpsElem.add_pattern */
void gpu_compile(void* obj);                                              /* MpsElem.compile     */
void* gpu_read_char(void* obj, char c);                                   /* MpsElem.read_char   */
size_t gpu_total_mem(void* obj);                                          /* MpsElem.total_mem   */
void gpu_reset(void* obj);                                                /* MpsElem.reset       */
void gpu_free(void* obj);                                                 /* MpsElem.free        */
/* batched extension: out[j] = what read_char would have returned for buf[j]; returns n */
size_t gpu_read_block(void* obj, const char* buf, size_t n, void** out);
/* same object, other kernels behind it */
void* gpu_dfa_create(void);
void* gpu_kr_create(void);
void* gpu_mpbg_create(void);

/* Layout-compatible with MpsElem (Core/src/mps.h:71-80); pattern_id_t spelled void*. */
typedef struct {
    char* name;
    void* (*create)(void);
    void (*add_pattern)(void*, char*, size_t, void*);
    void (*compile)(void*);
    void* (*read_char)(void*, char);
    size_t (*total_mem)(void*);
    void (*reset)(void*);
    void (*free)(void*);
} pm_mps_elem;
/* fill one mps_table slot: the body of a mps_gpu_register() added to mps_table_setup
 * (Core/src/mps.c:120-124, recipe Core/src/README.md:119-130) */
void mps_gpu_register_into(pm_mps_elem* slot);     /* exact, suffix-trie scan   */
void mps_gpu_dfa_register_into(pm_mps_elem* slot); /* exact, forward DFA walker */
void mps_gpu_kr_register_into(pm_mps_elem* slot);  /* randomized Karp-Rabin     */
void mps_gpu_mpbg_register_into(pm_mps_elem* slot); /* PM_ALGO_MPBG: what the reference's MPBG reports */

#ifdef __cplusplus
}
#endif
#endif /* PM_B200_H */
