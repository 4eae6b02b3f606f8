/*
 * pm_oracle.c -- TEST INFRASTRUCTURE ONLY (see pm_oracle.h for who may load it).
 *
 * Plain-C restatement of the reference's hot path; each function names the reference lines it
 * follows (paths relative to /root/reference).  Nothing here is copied from the reference: the
 * data structures are our own (CSR trie instead of 2072-byte states), only the observable
 * behaviour is restated.  Pinned by tests/test_oracle_*.py against oracle/_ref (the reference
 * itself) and tests/golden/.  Randomized (KR) part: parity unpinned, see pm_oracle.h.
 */
#define _GNU_SOURCE
#include "pm_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* .dict line grammar -- Core/src/parser.c:63-99 (parse_pattern_from_line), :36-46              */
/* ------------------------------------------------------------------------------------------ */
static int hexval(int ch) {                                   /* parser.c:36-46 get_binary_val */
    if (ch >= '0' && ch <= '9') return ch - '0';
    if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
    if (ch >= 'A' && ch <= 'F') return ch - 'A' + 10;
    return -1;
}
/* The reference indexes line[pos] without checking pos < n while it skips spaces and reads
 * nibbles; what it finds there is the getline terminator ('\n' or '\0'), which is neither a space
 * nor a hex digit.  `at()` models that: anything at or beyond n reads as '\n'. */
static inline int at(const uint8_t* line, size_t n, size_t pos) { return pos < n ? line[pos] : '\n'; }

int pmo_parse_line(const uint8_t* line, size_t n, uint8_t* out, size_t* out_len) {
    size_t len = 0, pos = 0;
    *out_len = 0;
    if (n == 0) return 0;                                     /* parser.c:64-66 */
    while (pos < n) {
        if (line[pos] == '|') {                               /* parser.c:70 hex section */
            ++pos;
            while (pos < n && line[pos] != '|') {             /* parser.c:72 */
                while (at(line, n, pos) == ' ') ++pos;        /* skip_spaces before 1st nibble */
                int first = hexval(at(line, n, pos)); ++pos;
                while (at(line, n, pos) == ' ') ++pos;        /* skip_spaces before 2nd nibble */
                int second = hexval(at(line, n, pos)); ++pos;
                if (first < 0 || second < 0) return 0;        /* parser.c:79-82: whole line rejected (Q1) */
                out[len++] = (uint8_t)(first * 16 + second);
            }
            if (pos >= n) return 0;                           /* parser.c:85 unterminated section */
            ++pos;                                            /* closing bar */
        } else {
            out[len++] = line[pos++];                         /* parser.c:88 raw byte */
        }
    }
    *out_len = len;
    return len != 0;                                          /* PatternsTree.c:279 zero length skipped */
}

/* ------------------------------------------------------------------------------------------ */
/* The automaton                                                                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t file, line, len;
    uint64_t off;       /* into bytes */
    uint32_t state;     /* trie state where the pattern ends */
    int32_t parent;     /* PatternsTree parent (longest proper suffix that is a pattern) or -1 */
} OPat;

struct PmOracle {
    /* build-time trie: first-child / next-sibling */
    uint32_t *first_child, *next_sib, *fail, *slink;
    uint8_t* edge;          /* byte on the edge into the state */
    int32_t* term;          /* pattern index ending at the state, or -1 */
    uint32_t n_states, cap_states;
    /* compiled CSR goto */
    uint32_t *csr_off, *csr_child;
    uint8_t* csr_byte;
    uint32_t root_goto[256];
    int32_t* longest;       /* per state: pattern index of slink state, or -1 */
    OPat* pats; size_t n_pats, cap_pats;
    uint8_t* bytes; size_t n_bytes, cap_bytes;
    size_t max_len, n_lines, n_rejected, n_dups, n_files;
    uint32_t cur;           /* current state (ac->current_state, mpac.c:50-57) */
    int compiled;
};

static uint32_t new_state(PmOracle* o, uint8_t ch) {
    if (o->n_states == o->cap_states) {
        uint32_t c = o->cap_states ? o->cap_states * 2 : 4096;
        o->first_child = (uint32_t*)realloc(o->first_child, c * sizeof(uint32_t));
        o->next_sib = (uint32_t*)realloc(o->next_sib, c * sizeof(uint32_t));
        o->edge = (uint8_t*)realloc(o->edge, c);
        o->term = (int32_t*)realloc(o->term, c * sizeof(int32_t));
        o->cap_states = c;
    }
    uint32_t s = o->n_states++;
    o->first_child[s] = 0; o->next_sib[s] = 0; o->edge[s] = ch; o->term[s] = -1;
    return s;
}

PmOracle* pmo_create(void) {                                  /* mpac.c:236-247 ac_create */
    PmOracle* o = (PmOracle*)calloc(1, sizeof(PmOracle));
    new_state(o, 0); /* root */
    return o;
}

void pmo_free(PmOracle* o) {
    if (!o) return;
    free(o->first_child); free(o->next_sib); free(o->fail); free(o->slink); free(o->edge); free(o->term);
    free(o->csr_off); free(o->csr_child); free(o->csr_byte); free(o->longest); free(o->pats); free(o->bytes);
    free(o);
}

static uint32_t build_child(const PmOracle* o, uint32_t s, uint8_t ch) {
    for (uint32_t c = o->first_child[s]; c; c = o->next_sib[c])
        if (o->edge[c] == ch) return c;
    return 0;
}

/* Trie insert = mpac.c:257-273 (ac_add_pattern); an identical byte string that is already present
 * is ignored and keeps the FIRST id -- PatternsTree.c:193-196 (Q2). */
int pmo_add_pattern(PmOracle* o, const uint8_t* pat, size_t len, uint32_t file, uint32_t line) {
    if (o->compiled || len == 0) return -1;
    uint32_t s = 0;
    for (size_t i = 0; i < len; ++i) {
        uint32_t c = build_child(o, s, pat[i]);
        if (!c) {
            c = new_state(o, pat[i]);
            o->next_sib[c] = o->first_child[s];
            o->first_child[s] = c;
        }
        s = c;
    }
    if (o->term[s] >= 0) { o->n_dups++; return 1; }
    if (o->n_pats == o->cap_pats) {
        o->cap_pats = o->cap_pats ? o->cap_pats * 2 : 1024;
        o->pats = (OPat*)realloc(o->pats, o->cap_pats * sizeof(OPat));
    }
    if (o->n_bytes + len > o->cap_bytes) {
        while (o->n_bytes + len > o->cap_bytes) o->cap_bytes = o->cap_bytes ? o->cap_bytes * 2 : (1u << 20);
        o->bytes = (uint8_t*)realloc(o->bytes, o->cap_bytes);
    }
    OPat* p = &o->pats[o->n_pats];
    p->file = file; p->line = line; p->len = (uint32_t)len; p->off = o->n_bytes; p->state = s; p->parent = -1;
    memcpy(o->bytes + o->n_bytes, pat, len);
    o->n_bytes += len;
    o->term[s] = (int32_t)o->n_pats++;
    if (len > o->max_len) o->max_len = len;                   /* PatternsTree.c:285, :310 */
    return 0;
}

/* One dictionary file = PatternsTree.c:260-291 (fpt_fill_with_dict_file): every getline counts as a
 * line (1-based), a trailing '\n' is stripped, rejected/empty lines are skipped. */
int pmo_add_dict_mem(PmOracle* o, const uint8_t* data, size_t n) {
    uint32_t file = (uint32_t)o->n_files++;
    uint32_t line_no = 0;
    size_t cap = 1 << 16;
    uint8_t* tmp = (uint8_t*)malloc(cap);
    size_t pos = 0;
    while (pos < n) {
        const uint8_t* nl = (const uint8_t*)memchr(data + pos, '\n', n - pos);
        size_t len = nl ? (size_t)(nl - (data + pos)) : n - pos;
        ++line_no; o->n_lines++;
        if (len + 1 > cap) { cap = len * 2 + 16; tmp = (uint8_t*)realloc(tmp, cap); }
        size_t plen = 0;
        if (pmo_parse_line(data + pos, len, tmp, &plen)) pmo_add_pattern(o, tmp, plen, file, line_no);
        else if (len) o->n_rejected++;
        pos += len + (nl ? 1 : 0);
    }
    free(tmp);
    return 0;
}

int pmo_add_dict_file(PmOracle* o, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* data = (uint8_t*)malloc(sz > 0 ? (size_t)sz : 1);
    size_t got = fread(data, 1, (size_t)sz, f);
    fclose(f);
    int rc = pmo_add_dict_mem(o, data, got);
    free(data);
    return rc;
}

static uint32_t goto_of(const PmOracle* o, uint32_t s, uint8_t ch) {
    if (s == 0) return o->root_goto[ch];
    uint32_t lo = o->csr_off[s], hi = o->csr_off[s + 1];
    for (uint32_t k = lo; k < hi; ++k)
        if (o->csr_byte[k] == ch) return o->csr_child[k];
    return 0;
}

/* mpac.c:282-291 (ac_compile) -> :187-210 (add_failure_links, BFS) -> :172-180 (add_failure_to_state):
 * fail(child) = goto*(fail(parent), c); suffix_link(s) = s if s ends a pattern else
 * suffix_link(fail(s)).  read_char returns id[suffix_link[cur]]; we store that as longest[cur]. */
int pmo_compile(PmOracle* o) {
    if (o->compiled) return -1;
    uint32_t n = o->n_states;
    o->csr_off = (uint32_t*)calloc((size_t)n + 1, sizeof(uint32_t));
    o->csr_child = (uint32_t*)malloc((size_t)(n ? n : 1) * sizeof(uint32_t));
    o->csr_byte = (uint8_t*)malloc(n ? n : 1);
    uint32_t k = 0;
    for (uint32_t s = 0; s < n; ++s) {
        o->csr_off[s] = k;
        for (uint32_t c = o->first_child[s]; c; c = o->next_sib[c]) { o->csr_child[k] = c; o->csr_byte[k] = o->edge[c]; ++k; }
    }
    o->csr_off[n] = k;
    memset(o->root_goto, 0, sizeof(o->root_goto));
    for (uint32_t c = o->first_child[0]; c; c = o->next_sib[c]) o->root_goto[o->edge[c]] = c;

    o->fail = (uint32_t*)calloc(n, sizeof(uint32_t));
    o->slink = (uint32_t*)calloc(n, sizeof(uint32_t));
    o->longest = (int32_t*)malloc(n * sizeof(int32_t));
    uint32_t* queue = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint32_t qh = 0, qt = 0;
    o->fail[0] = 0; o->slink[0] = 0;                          /* mpac.c:191-192 */
    for (int ch = 0; ch < 256; ++ch) {                        /* mpac.c:193-200 first level */
        uint32_t c = o->root_goto[ch];
        if (c) { queue[qt++] = c; o->fail[c] = 0; o->slink[c] = o->term[c] >= 0 ? c : 0; }
    }
    while (qh < qt) {                                         /* mpac.c:201-209 */
        uint32_t s = queue[qh++];
        for (uint32_t e = o->csr_off[s]; e < o->csr_off[s + 1]; ++e) {
            uint32_t c = o->csr_child[e]; uint8_t ch = o->csr_byte[e];
            uint32_t fs = o->fail[s];                         /* mpac.c:172-180 */
            while (!goto_of(o, fs, ch) && fs) fs = o->fail[fs];
            o->fail[c] = goto_of(o, fs, ch);
            o->slink[c] = o->term[c] >= 0 ? c : o->slink[o->fail[c]];
            queue[qt++] = c;
        }
    }
    free(queue);
    for (uint32_t s = 0; s < n; ++s) o->longest[s] = o->term[o->slink[s]];   /* states[slink].id; root id = NULL */
    /* PatternsTree parent (PatternsTree.c:186-214 shape; Core/src/README.md:45-47): the longest
     * proper suffix of the pattern that is itself a pattern = id[suffix_link[fail[end state]]]. */
    for (size_t i = 0; i < o->n_pats; ++i) o->pats[i].parent = o->longest[o->fail[o->pats[i].state]];
    o->cur = 0;
    o->compiled = 1;
    return 0;
}

size_t pmo_n_patterns(const PmOracle* o) { return o->n_pats; }
size_t pmo_n_states(const PmOracle* o) { return o->n_states; }
size_t pmo_max_pat_len(const PmOracle* o) { return o->max_len; }
size_t pmo_n_lines(const PmOracle* o) { return o->n_lines; }
size_t pmo_n_rejected(const PmOracle* o) { return o->n_rejected; }
size_t pmo_n_duplicates(const PmOracle* o) { return o->n_dups; }

int pmo_pattern(const PmOracle* o, size_t idx, uint32_t* file, uint32_t* line, int32_t* parent_idx,
                uint32_t* len, const uint8_t** bytes) {
    if (idx >= o->n_pats) return -1;
    const OPat* p = &o->pats[idx];
    if (file) *file = p->file;
    if (line) *line = p->line;
    if (parent_idx) *parent_idx = p->parent;
    if (len) *len = p->len;
    if (bytes) *bytes = o->bytes + p->off;
    return 0;
}

void pmo_reset(PmOracle* o) { o->cur = 0; }                   /* mpac.c:339-342 */

/* mpac.c:304-319 (ac_read_char): follow failure links until a goto exists, step, report
 * id[suffix_link[state]]. */
static inline int32_t step(PmOracle* o, uint8_t uc) {
    uint32_t cur = o->cur, g;
    while (!(g = goto_of(o, cur, uc)) && cur) cur = o->fail[cur];
    o->cur = g ? g : cur;
    return o->longest[o->cur];
}

void pmo_scan(PmOracle* o, const uint8_t* buf, size_t n, int32_t* longest_out) {
    for (size_t j = 0; j < n; ++j) longest_out[j] = step(o, buf[j]);   /* measure.c:292-294 */
}

uint64_t pmo_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t match_digest(uint64_t pos, uint64_t file, uint64_t line) {
    return pmo_splitmix64(pos ^ pmo_splitmix64(((file + 1) << 32) | line));
}

void pmo_summary(PmOracle* o, const uint8_t* buf, size_t n, size_t skip, uint64_t pos_base, PmoSummary* s) {
    uint64_t h = 1469598103934665603ULL;
    memset(s, 0, sizeof(*s));
    for (size_t j = 0; j < n; ++j) {
        int32_t id = step(o, buf[j]);
        if (j < skip || id < 0) continue;
        uint64_t pos = pos_base + (j - skip);
        s->positions++;
        s->hsum_longest += match_digest(pos, o->pats[id].file, o->pats[id].line);
        for (; id >= 0; id = o->pats[id].parent) {            /* longest + its PatternsTree ancestors */
            uint64_t v[3] = { pos, o->pats[id].file, o->pats[id].line };
            for (int k = 0; k < 3; ++k) { h ^= v[k]; h *= 1099511628211ULL; }
            s->hsum_all += match_digest(pos, v[1], v[2]);
            s->matches++;
        }
    }
    s->fnv = h;
}

int pmo_is_pattern_suffix(const PmOracle* o, int32_t first, int32_t second) {   /* PatternsTree.c:485-494 */
    if (first < 0) return 0;
    for (int32_t cur = second; cur >= 0; cur = o->pats[cur].parent)
        if (cur == first) return 1;
    return 0;
}

void pmo_classify(const PmOracle* o, const int32_t* algo, const int32_t* real, size_t n, uint64_t counts[4]) {
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    for (size_t i = 0; i < n; ++i) {                          /* measure.c:174-190 */
        if (real[i] == algo[i]) counts[0]++;
        else if (pmo_is_pattern_suffix(o, algo[i], real[i])) counts[1]++;
        else if (algo[i] < 0) counts[2]++;
        else counts[3]++;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Exact single-pattern KMP: what kmp_read_char (kmprt.c:255-285) reports -- 1 exactly at the   */
/* positions where an occurrence ends -- using the failure table of kmprt.c:167-181.            */
/* ------------------------------------------------------------------------------------------ */
size_t pmo_kmp_search(const uint8_t* pat, size_t n, const uint8_t* text, size_t m, size_t* ends, size_t cap) {
    if (n == 0) return 0;
    size_t* f = (size_t*)malloc((n + 1) * sizeof(size_t));
    f[0] = 0; if (n >= 1) f[1] = 0;
    size_t pos = 2, cnd = 0;
    while (pos < n + 1) {
        if (pat[pos - 1] == pat[cnd]) f[pos++] = ++cnd;
        else if (cnd > 0) cnd = f[cnd];
        else f[pos++] = 0;
    }
    size_t q = 0, cnt = 0;
    for (size_t i = 0; i < m; ++i) {
        while (q > 0 && (q == n || pat[q] != text[i])) q = f[q];
        if (pat[q] == text[i]) ++q;
        if (q == n) { if (cnt < cap) ends[cnt] = i; ++cnt; }
    }
    free(f);
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* Karp-Rabin: GF(p), p = 2^31-1 (mpbg.c:83); fp(s) = sum s[i] r^i mod p (Fingerprint.c:29-42). */
/* Bytes are UNSIGNED here on both the pattern and the stream side (the reference sign-extends   */
/* `char` on the pattern side only -- SURVEY Q6 -- which we do not reproduce).                   */
/* ------------------------------------------------------------------------------------------ */
uint64_t pmo_mulmod(uint64_t a, uint64_t b) { return (a * b) % PMO_KR_P; }  /* field.h:61-70: p < 2^32 */
uint64_t pmo_powmod(uint64_t a, uint64_t e) {
    uint64_t r = 1; a %= PMO_KR_P;
    while (e) { if (e & 1) r = pmo_mulmod(r, a); a = pmo_mulmod(a, a); e >>= 1; }
    return r;
}
uint64_t pmo_invmod(uint64_t a) {                             /* field.c:26-72: extended Euclid */
    int64_t t = 0, nt = 1, r = (int64_t)PMO_KR_P, nr = (int64_t)(a % PMO_KR_P);
    while (nr) { int64_t q = r / nr, x; x = t - q * nt; t = nt; nt = x; x = r - q * nr; r = nr; nr = x; }
    return (uint64_t)(t < 0 ? t + (int64_t)PMO_KR_P : t);
}
uint64_t pmo_fp(const uint8_t* seq, size_t n, uint64_t r) {
    uint64_t ret = 0, rn = 1;
    for (size_t i = 0; i < n; ++i) { ret = (ret + seq[i] * rn) % PMO_KR_P; rn = (rn * r) % PMO_KR_P; }
    return ret;
}
uint64_t pmo_kr_seed_r(uint64_t seed) { return 1 + pmo_splitmix64(seed) % (PMO_KR_P - 1); }

/* Our randomized variant.  A pattern of length len > 8 is reported at i iff, for every stage length
 * l in {8,16,..,2^floor(log2 len)} U {len}, fp(stream[i-l+1..i]) == fp(last l bytes of the pattern)
 * (the prefix-doubling stages of bgps.c:215-249 turned around so that every stage window ENDS at i);
 * patterns of <= 8 bytes are matched exactly (bgps.c:459-464, BG_SHORT_PATTERN_LENGTH bgps.h:36).
 * Reported: the longest such pattern per position.  buf[0] is the start of the stream. */
typedef struct { uint32_t fp8; uint32_t pat; } KrEnt;
static int krent_cmp(const void* a, const void* b) {
    const KrEnt* x = (const KrEnt*)a; const KrEnt* y = (const KrEnt*)b;
    if (x->fp8 != y->fp8) return x->fp8 < y->fp8 ? -1 : 1;
    return x->pat < y->pat ? -1 : (x->pat > y->pat);
}
void pmo_kr_scan(const PmOracle* o, uint64_t seed, const uint8_t* buf, size_t n, size_t hist, int32_t* longest_out) {
    (void)hist;
    uint64_t r = pmo_kr_seed_r(seed);
    size_t nl = 0;
    KrEnt* ents = (KrEnt*)malloc((o->n_pats ? o->n_pats : 1) * sizeof(KrEnt));
    for (size_t i = 0; i < o->n_pats; ++i)
        if (o->pats[i].len > 8) {
            ents[nl].fp8 = (uint32_t)pmo_fp(o->bytes + o->pats[i].off + o->pats[i].len - 8, 8, r);
            ents[nl].pat = (uint32_t)i; ++nl;
        }
    qsort(ents, nl, sizeof(KrEnt), krent_cmp);
    PmOracle* oo = (PmOracle*)o; /* exact short matches come from the exact automaton */
    uint32_t saved = oo->cur; oo->cur = 0;
    for (size_t i = 0; i < n; ++i) {
        int32_t best = step(oo, buf[i]);
        while (best >= 0 && o->pats[best].len > 8) best = o->pats[best].parent;   /* longest pattern <= 8 bytes */
        uint32_t best_len = best >= 0 ? o->pats[best].len : 0;
        if (i >= 8) {
            uint32_t f8 = (uint32_t)pmo_fp(buf + i - 7, 8, r);
            size_t lo = 0, hi = nl;
            while (lo < hi) { size_t mid = (lo + hi) / 2; if (ents[mid].fp8 < f8) lo = mid + 1; else hi = mid; }
            for (size_t e = lo; e < nl && ents[e].fp8 == f8; ++e) {
                const OPat* p = &o->pats[ents[e].pat];
                if (p->len <= best_len || p->len > i + 1) continue;
                int ok = 1;
                for (uint32_t l = 16; l <= p->len && ok; l <<= 1)
                    ok = pmo_fp(buf + i + 1 - l, l, r) == pmo_fp(o->bytes + p->off + p->len - l, l, r);
                if (ok) ok = pmo_fp(buf + i + 1 - p->len, p->len, r) == pmo_fp(o->bytes + p->off, p->len, r);
                if (ok) { best = (int32_t)ents[e].pat; best_len = p->len; }
            }
        }
        longest_out[i] = best;
    }
    oo->cur = saved;
    free(ents);
}

/* ------------------------------------------------------------------------------------------ */
/* Seeded synthetic streams.  Every byte is a pure function of its absolute      */
/* offset, so any shard can be regenerated independently on the CPU and on the GPU.             */
/* ------------------------------------------------------------------------------------------ */
static inline uint8_t uniform_byte(uint64_t i) { return (uint8_t)(pmo_splitmix64(0x5EED0001ULL + (i >> 3)) >> (8 * (i & 7))); }

void pmo_gen_uniform(uint64_t off, size_t n, uint8_t* out) {
    for (size_t j = 0; j < n; ++j) out[j] = uniform_byte(off + j);
}

/* S-planted: uniform background; per 4096-byte block b one occurrence of pattern k = h mod P at
 * offset (h>>32) mod (4096-len+1), h = splitmix64(0xD1C70002 + b). */
void pmo_gen_planted(const PmOracle* o, uint64_t off, size_t n, uint8_t* out) {
    pmo_gen_uniform(off, n, out);
    if (!o->n_pats || !n) return;
    for (uint64_t b = off >> 12; b <= (off + n - 1) >> 12; ++b) {
        uint64_t h = pmo_splitmix64(0xD1C70002ULL + b);
        const OPat* p = &o->pats[h % o->n_pats];
        if (p->len > 4096) continue;
        uint64_t start = (b << 12) + (h >> 32) % (4096 - p->len + 1);
        for (uint32_t t = 0; t < p->len; ++t) {
            uint64_t pos = start + t;
            if (pos >= off && pos < off + n) out[pos - off] = o->bytes[p->off + t];
        }
    }
}

/* S-almost: every 4096-byte block is a concatenation of random-length prefixes (1..len) of random
 * patterns -- the semantics of Streams/write_first_lines.py:48-61, made seeded and block-local:
 * piece t of block b uses h = splitmix64(0xA1A50005 + 4096 b + t), pattern h mod P, length
 * 1 + (h>>32) mod len; the last piece is cut at the block end. */
void pmo_gen_almost(const PmOracle* o, uint64_t off, size_t n, uint8_t* out) {
    if (!n) return;
    if (!o->n_pats) { memset(out, 0, n); return; }
    for (uint64_t b = off >> 12; b <= (off + n - 1) >> 12; ++b) {
        uint32_t fill = 0;
        for (uint64_t t = 0; fill < 4096; ++t) {
            uint64_t h = pmo_splitmix64(0xA1A50005ULL + (b << 12) + t);
            const OPat* p = &o->pats[h % o->n_pats];
            uint32_t plen = 1 + (uint32_t)((h >> 32) % p->len);
            for (uint32_t u = 0; u < plen && fill < 4096; ++u, ++fill) {
                uint64_t pos = (b << 12) + fill;
                if (pos >= off && pos < off + n) out[pos - off] = o->bytes[p->off + u];
            }
        }
    }
}

/* S-ascii: printable ASCII (0x20..0x7E), uniform -- text-like traffic; byte i = 0x20 + (8-bit lane of
 * splitmix64(0xA5C11006 + i/8) * 95 >> 8). */
void pmo_gen_ascii(uint64_t off, size_t n, uint8_t* out) {
    for (size_t j = 0; j < n; ++j) {
        uint64_t i = off + j;
        uint32_t v = (uint8_t)(pmo_splitmix64(0xA5C11006ULL + (i >> 3)) >> (8 * (i & 7)));
        out[j] = (uint8_t)(0x20 + ((v * 95) >> 8));
    }
}

/* S-ab: bytes over {a,b} with P(a) = 192/256, seed 0xADE50004. */
void pmo_gen_ab(uint64_t off, size_t n, uint8_t* out) {
    for (size_t j = 0; j < n; ++j) {
        uint64_t i = off + j;
        uint8_t v = (uint8_t)(pmo_splitmix64(0xADE50004ULL + (i >> 3)) >> (8 * (i & 7)));
        out[j] = v < 192 ? 'a' : 'b';
    }
}
