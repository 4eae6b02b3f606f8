/*
 * pm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's dictionary-matching hot path
 * (yehonatan145/PatternMatching, Core/src).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library never does.
 *
 * Parity status: the exact path (parser, de-dup/ids, Aho-Corasick longest match, PatternsTree
 * parents) is PINNED against the reference itself (oracle/_ref/libpmref.so built from the
 * unmodified sources) and against the golden counts/checksums in tests/golden/.  The randomized
 * Karp-Rabin path is "parity unpinned" by design: the reference MPBG draws an unseeded r and,
 * because of two bugs, never reports a pattern longer than 8 bytes; the oracle
 * restates OUR seeded variant so that the GPU kernel can be checked bit-for-bit against it, and its
 * error rates are reported against exact ground truth beside the reference's.
 */
#ifndef PM_ORACLE_H
#define PM_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct PmOracle PmOracle;

/* ---- dictionary ingest (parser.c:63-99, PatternsTree.c:186-214, 260-291) ---- */
int pmo_parse_line(const uint8_t* line, size_t n, uint8_t* out, size_t* out_len); /* 1 = accepted */
PmOracle* pmo_create(void);
void pmo_free(PmOracle* o);
int pmo_add_dict_file(PmOracle* o, const char* path);                /* file_number = call order */
int pmo_add_dict_mem(PmOracle* o, const uint8_t* data, size_t n);    /* same, lines from memory */
int pmo_add_pattern(PmOracle* o, const uint8_t* pat, size_t len, uint32_t file, uint32_t line);
int pmo_compile(PmOracle* o);                                        /* mpac.c:282-291 */
size_t pmo_n_patterns(const PmOracle* o);
size_t pmo_n_states(const PmOracle* o);
size_t pmo_max_pat_len(const PmOracle* o);
size_t pmo_n_lines(const PmOracle* o);
size_t pmo_n_rejected(const PmOracle* o);
size_t pmo_n_duplicates(const PmOracle* o);
/* canonical pattern index = order of first occurrence = (file,line)-sorted order */
int pmo_pattern(const PmOracle* o, size_t idx, uint32_t* file, uint32_t* line, int32_t* parent_idx,
                uint32_t* len, const uint8_t** bytes);

/* ---- the hot path (mpac.c:304-319 driven by measure.c:292-294) ---- */
void pmo_reset(PmOracle* o);
/* longest_out[j] = canonical index of the longest pattern that is a suffix of the bytes fed since
 * reset, or -1; state is carried across calls like consecutive read_char calls. */
void pmo_scan(PmOracle* o, const uint8_t* buf, size_t n, int32_t* longest_out);
typedef struct {
    uint64_t positions;   /* positions with a match */
    uint64_t matches;     /* matches incl. PatternsTree ancestors */
    uint64_t fnv;         /* ordered FNV checksum of (pos,file,line) triples, SURVEY Appendix A */
    uint64_t hsum_longest;/* order-independent digest sum, longest matches */
    uint64_t hsum_all;    /* order-independent digest sum, all matches */
} PmoSummary;
/* scan buf from the CURRENT state; the first `skip` bytes are walked but not reported; reported
 * positions are pos_base + (j - skip). */
void pmo_summary(PmOracle* o, const uint8_t* buf, size_t n, size_t skip, uint64_t pos_base, PmoSummary* s);
/* measure.c:174-190: counts[4] = success, partial, false_neg, false_pos (indices, -1 = none) */
void pmo_classify(const PmOracle* o, const int32_t* algo, const int32_t* real, size_t n, uint64_t counts[4]);
int pmo_is_pattern_suffix(const PmOracle* o, int32_t first, int32_t second); /* PatternsTree.c:485-494 */

/* ---- exact single-pattern KMP (kmprt.c:255-285 semantics: report where the pattern ends) ---- */
size_t pmo_kmp_search(const uint8_t* pat, size_t n, const uint8_t* text, size_t m, size_t* ends, size_t cap);

/* ---- Karp-Rabin fingerprints (Fingerprint.c:29-42, field.h) with unsigned bytes ---- */
#define PMO_KR_P 2147483647ULL /* mpbg.c:83 */
uint64_t pmo_mulmod(uint64_t a, uint64_t b);
uint64_t pmo_powmod(uint64_t a, uint64_t e);
uint64_t pmo_invmod(uint64_t a);                       /* field.c:26-72 (extended Euclid) */
uint64_t pmo_fp(const uint8_t* seq, size_t n, uint64_t r); /* sum seq[i]*r^i mod p */
uint64_t pmo_kr_seed_r(uint64_t seed);
/* Our seeded randomized variant (DESIGN.md "KR variant"): longest pattern ending at each position
 * where patterns of <= 8 bytes are matched exactly and longer ones by suffix-stage fingerprints. */
void pmo_kr_scan(const PmOracle* o, uint64_t seed, const uint8_t* buf, size_t n, size_t hist,
                 int32_t* longest_out);

/* ---- seeded synthetic streams; all are pure functions of the absolute offset ---- */
uint64_t pmo_splitmix64(uint64_t x);
void pmo_gen_uniform(uint64_t off, size_t n, uint8_t* out);
void pmo_gen_planted(const PmOracle* o, uint64_t off, size_t n, uint8_t* out);
void pmo_gen_almost(const PmOracle* o, uint64_t off, size_t n, uint8_t* out);
void pmo_gen_ab(uint64_t off, size_t n, uint8_t* out);
void pmo_gen_ascii(uint64_t off, size_t n, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
