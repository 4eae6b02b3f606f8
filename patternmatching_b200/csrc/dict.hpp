// dict.hpp -- host-side dictionary compiler (product code, C++17, no CUDA).
//
// Turns the merged .dict patterns into the flat tables the sm_100a scan kernels read:
//   * ingest with the reference's line grammar, de-dup and (file,line) ids
//     (Core/src/parser.c:63-99, Core/src/PatternsTree.c:186-214, 260-291),
//   * the reversed-pattern ("suffix") trie with every row flattened to "final pid | continue",
//     including the 2-byte-suffix table that lives in shared memory,
//   * the forward Aho-Corasick DFA (Core/src/mpac.c:147-210 completed to a full DFA),
//   * the PatternsTree parent relation flattened to pid -> parent pid
//     (Core/src/PatternsTree.c:378-403, 485-494),
//   * the Karp-Rabin stage tables for the randomized variant
//     (Core/src/Fingerprint.c:29-42, Core/src/bgps.c:215-249).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace pm {

constexpr uint32_t kContFlag = 0x80000000u;  // row entry: continue to row (entry & 0x00FFFFFF)
constexpr uint32_t kTailFlag = 0x40000000u;  // row entry: the rest of the path is the single pattern (entry & 0xFFFF)
constexpr uint32_t kTailMinDepth = 4;        // tails start at nodes of at least this depth (levels 1-4 stay table-driven)
constexpr uint64_t kKrP = 2147483647ull;     // field size p = 2^31-1 (Core/src/mpbg.c:83)

struct Pattern {
    uint32_t file = 0, line = 0, len = 0;
    uint64_t off = 0;       // into Dict::bytes
    uint64_t user = 0;      // opaque id of the caller (pattern_id_t in the shim)
    uint32_t parent = 0;    // pid of the longest proper suffix that is a pattern, 0 = none
    uint32_t chain = 0;     // number of ancestors (length of the parent chain)
};

// Tables of the exact backward scan ("sfx"): for stream position i the kernel walks the trie of
// REVERSED patterns along c[i], c[i-1], ... until the path dies; every table entry already holds
// the answer for the case that the walk ends there.
struct SfxTables {
    uint32_t n_nodes = 0;        // incl. root
    uint32_t n_rows = 0;         // nodes that own a row: internal, and not inside a single-pattern tail; BFS order
    uint32_t n_tail_nodes = 0;   // internal nodes folded into tails (compared against the pattern text instead)
    uint32_t row2_base = 0;      // first row that belongs to a depth-2 node
    uint32_t n2_cont = 0;        // depth-2 nodes that own a row (continue codes of root2)
    uint32_t cont_base = 65536;  // root2 entries >= cont_base mean "continue at row2_base + (e - cont_base)"
    bool fits_u16 = false;       // P + 1 <= cont_base: dense uint16 results and root2 codes are possible
    uint32_t n_classes = 0;      // byte classes used by the rows (class 0 = "byte occurs in no pattern" if any)
    uint32_t log2_ncp = 0;       // row stride = 1 << log2_ncp >= n_classes
    uint8_t cls[256] = {0};
    bool cls_identity = false;   // cls[b] == b for every byte (all 256 byte values occur in patterns)
    std::vector<uint16_t> root2; // [c_i << 8 | c_{i-1}] -> final pid, or continue code
    // level-3 filter, one word per depth-2 row (index = continue code - cont_base): high half = deepest terminal
    // at the depth-2 node (the answer when c[i-2] leads nowhere), low half = 16-bit Bloom of the bytes that DO
    // lead somewhere (bit = byte & 15).  Lives in shared memory: only Bloom hits go to the rows in L2.
    std::vector<uint32_t> l3f;
    std::vector<uint32_t> root1; // [c_i] -> final pid | kContFlag+row (bounded walker at stream start)
    std::vector<uint32_t> rows;  // [row << log2_ncp | cls] -> final pid | kContFlag+row
    std::vector<uint32_t> row_best; // [row] -> deepest terminal pid on the path to the row's node
    // by pid (entry 0 unused): {offset of the pattern text in Dict::bytes, length, length of the shortest of
    // {the pattern and its PatternsTree ancestors} that is longer than the tail-start depth, pid of the longest
    // one that is not}: what a walk needs when it reaches "tail of pattern pid"
    std::vector<uint32_t> tail_rec;   // 4 x uint32 per pid
    std::vector<uint32_t> depth_hist; // nodes per depth
};

// Flat Aho-Corasick DFA of the forward trie, states in BFS order (hot states first).
struct DfaTables {
    uint32_t n_states = 0;
    uint32_t n_classes = 0, log2_ncp = 0;
    uint8_t cls[256] = {0};
    std::vector<uint32_t> delta;   // [state << log2_ncp | cls] -> next state
    std::vector<uint16_t> longest; // [state] -> pid of the longest pattern that is a suffix of the state string
    std::vector<uint32_t> depth_count; // states per depth
    // per state: 16-bit Bloom of the classes of its goto children (bit = class & 15) | failure state << 16 (0xFFFF =
    // no usable entry).  A state WITHOUT a goto child on c moves like its failure state: delta(s,c) = delta(fail(s),c).
    std::vector<uint32_t> fb_meta;
    bool built = false;
};

// Compact forward automaton for deep-match traffic (deep_scan.cu): the reference's goto + failure machine
// (Core/src/mpac.c:147-210) kept as goto + failure -- NOT completed to a dense DFA -- in one 32-byte record per
// state, 23 MB for snort+et instead of the dense table's 734 MB, so that it is L2-resident.
//   state numbering: the root and the depth-1 states first ("hot": their complete DFA rows live in shared memory as
//   u16), then every DENSE state (more than 6 children: a complete 256-entry DFA row in dense_rows, found from the id
//   alone: row = id - n_hot), then the remaining depth-2 states -- the ids below n_small, whose longest pids live in
//   shared memory -- then every deeper state in depth-first pre-order with children in byte order: the only child of
//   a single-child state is state + 1 (unless it is DENSE), so a run of single-child states ("chain") is a run of
//   consecutive ids and one record describes up to eight steps of it.
//   record of state s >= n_hot + n_dense (8 x u32):  w0 = failure state | kind << 24 | count << 26,  w1 = longest pid at s,
//     kind 0 BRANCH: 1 .. 6 goto edges, w2..w7 = child << 8 | byte (unused slots repeat the first edge); a miss follows the failure link
//     kind 1 CHAIN : count <= 8 steps: bytes of states s+1 .. s+count in w2,w3, their longest pids (u16) in w4..w7
//     kind 2 LEAF  : no goto edge: every byte follows the failure link
struct DeepTables {
    uint32_t n_states = 0, n_hot = 0, n_small = 0;   // hot rows point at states < n_small (fit u16)
    std::vector<uint16_t> hot_rows;      // [hot state << 8 | byte] -> next state (complete DFA transition)
    std::vector<uint16_t> hot_longest;   // [state < n_small] -> longest pid (hot, DENSE and depth-2 states)
    std::vector<uint32_t> recs;          // 8 words per state (records of the hot and DENSE states are unused)
    std::vector<uint32_t> dense_rows;    // [(state - n_hot) << 8 | byte] -> next state, 256 entries per DENSE state
    uint32_t n_dense = 0, n_chain = 0, n_branch = 0;
    std::vector<uint32_t> depth_count;   // states per depth (forward trie)
    bool usable = false;                 // false: the automaton does not fit this layout (ids >= 2^24, > 65535 pids, ...)
    bool built = false;
};

// Karp-Rabin suffix-stage tables (randomized variant).
struct KrTables {
    uint64_t seed = 0, r = 0;
    uint32_t n_long = 0;                 // patterns longer than 8 bytes
    uint32_t bucket_bits = 0;            // hash table of 8-byte-suffix fingerprints: 1 << bucket_bits slots
    std::vector<uint32_t> slot_fp;       // fp8 stored per slot (0xFFFFFFFF = empty), open addressing
    std::vector<uint32_t> slot_begin;    // [slot] -> first candidate in cand_*, candidates of a slot are contiguous
    std::vector<uint32_t> slot_count;
    std::vector<uint32_t> cand_pid;      // candidates sorted by decreasing length within a slot
    std::vector<uint32_t> cand_len;
    std::vector<uint32_t> cand_stage_off;// into stage_fp: fps of the suffixes of length 16,32,..,2^k and the full length
    std::vector<uint32_t> stage_fp;
    std::vector<uint32_t> bloom;         // 1 << bloom_bits bits: fp8 prefilter held in shared memory
    uint32_t bloom_bits = 0;
    bool built = false;
};

class Dict {
  public:
    Dict();
    // ---- ingest ----
    static bool parse_line(const uint8_t* line, size_t n, uint8_t* out, size_t* out_len);
    int add_file(const char* path);
    int add_mem(const uint8_t* data, size_t n);
    uint32_t add_pattern(const uint8_t* pat, size_t len, uint32_t file, uint32_t line, uint64_t user);
    int compile();
    // The forward DFA is built lazily (the table is large: n_states * 256 * 4 bytes), at most once and under a
    // dictionary-level lock: engines on different threads may ask for it at the same time.  After it returns the
    // tables are read-only like everything else in a compiled dictionary.
    void build_dfa() const;
    void build_deep() const;          // same contract as build_dfa(): lazily, once, under the dictionary's lock
    // Karp-Rabin tables for one seed: returned by value, owned by the caller (an engine) -- the dictionary itself is
    // not touched, so engines with different seeds share it safely.
    KrTables build_kr(uint64_t seed) const;
    bool is_pattern_suffix(uint32_t first, uint32_t second) const;
    // compiled-automaton cache: the compiled dictionary as one binary file
    int save(const char* path) const;
    int load(const char* path);

    // ---- data ----
    std::vector<Pattern> pats;        // pats[pid-1]
    std::vector<uint8_t> bytes;
    // output links flattened to pattern-id ranges: all patterns that end where pid ends (pid itself, then its
    // PatternsTree ancestors, longest first) are anc_list[anc_off[pid] .. anc_off[pid + 1])
    std::vector<uint32_t> anc_off;    // P + 2 entries (entry 0: the empty range of "no pattern")
    std::vector<uint16_t> anc_list;
    uint64_t n_lines = 0, n_rejected = 0, n_dups = 0;
    uint32_t n_files = 0, max_len = 0;
    uint32_t n_ac_states = 1;
    bool compiled = false;
    // More than 65,535 unique patterns: dense uint16 results cannot name them.  The dictionary is then cut into PARTS of
    // at most kPartPatterns patterns (consecutive pids), each a complete dictionary of its own with all its tables; an
    // engine scans the stream once per part and keeps, per position, the longer of the parts' answers (two different
    // patterns ending at the same position have different lengths) as a 32-bit global pid.  parents / chains below are
    // global in either case.  part k holds the global pids [part_first[k], part_first[k + 1]).
    static constexpr uint32_t kPartPatterns = 49152;
    bool multi = false;
    std::vector<std::unique_ptr<Dict>> parts;
    std::vector<uint32_t> part_first;
    SfxTables sfx;
    mutable DfaTables dfa;            // see build_dfa()
    mutable DeepTables deep;          // see build_deep()
    std::string error;

  private:
    int compile_parts();
    // forward trie (also the de-dup structure): hash of (state << 8 | byte) -> child
    struct Trie;
    Trie* fwd_;
    mutable std::mutex lazy_mu_;      // serialises build_dfa()
  public:
    ~Dict();
    Dict(const Dict&) = delete;
    Dict& operator=(const Dict&) = delete;
};

// Modular arithmetic in GF(2^31-1) shared by host table construction (and mirrored on the device).
inline uint64_t kr_mul(uint64_t a, uint64_t b) { return (a * b) % kKrP; }

// Bloom bitmap of the 8-byte-suffix fingerprints (2^19 bits): kKrBloomHashes multiplicative hashes per key.  With
// ~35 k keys the bitmap is 24% full and a random fingerprint passes all four tests with probability 0.3%; the
// kernel evaluates them one after the other, so the later ones are almost free.
constexpr int kKrBloomHashes = 4;
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t kr_bloom_bit(uint32_t fp8, int k) {
    const uint32_t c = k == 0 ? 0x9E3779B1u : k == 1 ? 0x85EBCA77u : k == 2 ? 0xC2B2AE3Du : 0x27D4EB2Fu;
    return (fp8 * c) >> 13;   // top 19 bits of the product
}
uint64_t kr_pow(uint64_t a, uint64_t e);
uint64_t kr_inv(uint64_t a);
uint64_t kr_fp(const uint8_t* s, size_t n, uint64_t r);  // sum s[i] r^i mod p, unsigned bytes
uint64_t splitmix64(uint64_t x);

}  // namespace pm
