// dict.cpp -- host dictionary compiler.  See dict.hpp for the reference lines each part restates.
#include "dict.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <unordered_map>

#include <unistd.h>

namespace pm {

uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t kr_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    a %= kKrP;
    while (e) {
        if (e & 1) r = kr_mul(r, a);
        a = kr_mul(a, a);
        e >>= 1;
    }
    return r;
}
uint64_t kr_inv(uint64_t a) { return kr_pow(a, kKrP - 2); }  // Fermat; same value as field.c's Euclid
uint64_t kr_fp(const uint8_t* s, size_t n, uint64_t r) {
    uint64_t acc = 0, rn = 1;
    for (size_t i = 0; i < n; ++i) {
        acc = (acc + s[i] * rn) % kKrP;
        rn = kr_mul(rn, r);
    }
    return acc;
}

// --------------------------------------------------------------------------------------------
// A byte trie built by hashing (node, byte) -> child, then renumbered breadth-first with children
// in byte order (deterministic, shallow nodes get the small ids).
// --------------------------------------------------------------------------------------------
struct Dict::Trie {
    std::unordered_map<uint64_t, uint32_t> edge;
    std::vector<uint32_t> term;  // pid ending at the node, 0 = none
    Trie() { term.push_back(0); edge.reserve(1 << 16); }
    uint32_t child(uint32_t node, uint8_t b) const {
        auto it = edge.find((uint64_t(node) << 8) | b);
        return it == edge.end() ? 0 : it->second;
    }
    uint32_t child_or_add(uint32_t node, uint8_t b) {
        auto ins = edge.emplace((uint64_t(node) << 8) | b, uint32_t(term.size()));
        if (ins.second) term.push_back(0);
        return ins.first->second;
    }
    size_t size() const { return term.size(); }
};

namespace {
// Breadth-first CSR view of a Trie.
struct Bfs {
    uint32_t n = 0;
    std::vector<uint32_t> off, child, parent, depth, term;
    std::vector<uint8_t> byte, in_byte;
    explicit Bfs(const std::unordered_map<uint64_t, uint32_t>& edge, const std::vector<uint32_t>& term_old) {
        n = uint32_t(term_old.size());
        // edges sorted by (old parent, byte)
        std::vector<std::pair<uint64_t, uint32_t>> es(edge.begin(), edge.end());
        std::sort(es.begin(), es.end());
        std::vector<uint32_t> old_off(n + 1, 0);
        for (auto& e : es) old_off[(e.first >> 8) + 1]++;
        for (uint32_t i = 0; i < n; ++i) old_off[i + 1] += old_off[i];
        std::vector<uint32_t> order;  // bfs position -> old id
        order.reserve(n);
        std::vector<uint32_t> new_id(n, 0);
        order.push_back(0);
        parent.assign(n, 0); depth.assign(n, 0); in_byte.assign(n, 0);
        for (size_t h = 0; h < order.size(); ++h) {
            uint32_t o = order[h];
            for (uint32_t k = old_off[o]; k < old_off[o + 1]; ++k) {
                uint32_t c = es[k].second;
                new_id[c] = uint32_t(order.size());
                parent[new_id[c]] = uint32_t(h);
                depth[new_id[c]] = depth[h] + 1;
                in_byte[new_id[c]] = uint8_t(es[k].first & 0xFF);
                order.push_back(c);
            }
        }
        off.assign(n + 1, 0); child.resize(n ? n - 1 : 0); byte.resize(n ? n - 1 : 0); term.resize(n);
        uint32_t k2 = 0;
        for (uint32_t h = 0; h < n; ++h) {
            uint32_t o = order[h];
            term[h] = term_old[o];
            off[h] = k2;
            for (uint32_t k = old_off[o]; k < old_off[o + 1]; ++k) {
                child[k2] = new_id[es[k].second];
                byte[k2] = uint8_t(es[k].first & 0xFF);
                ++k2;
            }
        }
        off[n] = k2;
    }
    bool internal(uint32_t v) const { return off[v + 1] > off[v]; }
};
}  // namespace

Dict::Dict() : fwd_(new Trie()) {}
Dict::~Dict() { delete fwd_; }

// Core/src/parser.c:63-99.  Reads at or past the end of the line see the getline terminator, which
// is neither ' ' nor a hex digit; a space right before the closing bar therefore rejects the line.
bool Dict::parse_line(const uint8_t* line, size_t n, uint8_t* out, size_t* out_len) {
    auto hex = [](int ch) -> int {
        if (ch >= '0' && ch <= '9') return ch - '0';
        if (ch >= 'a' && ch <= 'f') return ch - 'a' + 10;
        if (ch >= 'A' && ch <= 'F') return ch - 'A' + 10;
        return -1;
    };
    auto at = [&](size_t p) -> int { return p < n ? line[p] : '\n'; };
    size_t len = 0, pos = 0;
    *out_len = 0;
    while (pos < n) {
        if (line[pos] != '|') { out[len++] = line[pos++]; continue; }
        ++pos;
        while (pos < n && line[pos] != '|') {
            while (at(pos) == ' ') ++pos;
            int hi = hex(at(pos)); ++pos;
            while (at(pos) == ' ') ++pos;
            int lo = hex(at(pos)); ++pos;
            if (hi < 0 || lo < 0) return false;
            out[len++] = uint8_t(hi * 16 + lo);
        }
        if (pos >= n) return false;  // unterminated hex section
        ++pos;
    }
    *out_len = len;
    return len != 0;
}

uint32_t Dict::add_pattern(const uint8_t* pat, size_t len, uint32_t file, uint32_t line, uint64_t user) {
    if (compiled || len == 0) return 0;
    uint32_t s = 0;
    for (size_t i = 0; i < len; ++i) s = fwd_->child_or_add(s, pat[i]);
    if (fwd_->term[s]) { ++n_dups; return fwd_->term[s]; }  // PatternsTree.c:193-196: first occurrence wins
    Pattern p;
    p.file = file; p.line = line; p.len = uint32_t(len); p.off = bytes.size(); p.user = user;
    bytes.insert(bytes.end(), pat, pat + len);
    pats.push_back(p);
    fwd_->term[s] = uint32_t(pats.size());
    max_len = std::max<uint32_t>(max_len, uint32_t(len));
    return uint32_t(pats.size());
}

// PatternsTree.c:260-291: every line read counts (1-based), trailing '\n' stripped, rejected and
// empty lines skipped.
int Dict::add_mem(const uint8_t* data, size_t n) {
    uint32_t file = n_files++, line_no = 0;
    std::vector<uint8_t> tmp;
    size_t pos = 0;
    while (pos < n) {
        const uint8_t* nl = static_cast<const uint8_t*>(memchr(data + pos, '\n', n - pos));
        size_t len = nl ? size_t(nl - (data + pos)) : n - pos;
        ++line_no; ++n_lines;
        tmp.resize(len + 1);
        size_t plen = 0;
        if (parse_line(data + pos, len, tmp.data(), &plen)) add_pattern(tmp.data(), plen, file, line_no, 0);
        else if (len) ++n_rejected;
        pos += len + (nl ? 1 : 0);
    }
    return 0;
}

int Dict::add_file(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { error = std::string("cannot open dictionary file ") + path; return -1; }
    std::vector<uint8_t> data;
    uint8_t buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0) data.insert(data.end(), buf, buf + got);
    fclose(f);
    return add_mem(data.data(), data.size());
}

bool Dict::is_pattern_suffix(uint32_t first, uint32_t second) const {  // PatternsTree.c:485-494
    if (first == 0) return false;
    for (uint32_t cur = second; cur; cur = pats[cur - 1].parent)
        if (cur == first) return true;
    return false;
}

// More than 65,535 unique patterns (dict.hpp: parts): the PatternsTree relation has been computed over the whole
// dictionary; every part is compiled like a dictionary of its own.  mpac.c:257-291 accepts any number of patterns.
int Dict::compile_parts() {
    const uint32_t P = uint32_t(pats.size());
    for (uint32_t first = 1; first <= P; first += kPartPatterns) {
        std::unique_ptr<Dict> part(new Dict());
        const uint32_t last = std::min<uint64_t>(P, uint64_t(first) + kPartPatterns - 1);
        for (uint32_t pid = first; pid <= last; ++pid) {
            const Pattern& q = pats[pid - 1];
            part->add_pattern(bytes.data() + q.off, q.len, q.file, q.line, q.user);
        }
        if (part->compile()) { error = part->error; return -1; }
        part_first.push_back(first);
        parts.push_back(std::move(part));
    }
    part_first.push_back(P + 1);
    sfx.fits_u16 = false;
    multi = true;
    compiled = true;
    return 0;
}

int Dict::compile() {
    if (compiled) { error = "dictionary already compiled"; return -1; }
    n_ac_states = uint32_t(fwd_->size());
    const uint32_t P = uint32_t(pats.size());

    // ---- reversed-pattern trie ----
    Trie rev;
    rev.edge.reserve(bytes.size() * 2 + 16);
    for (uint32_t i = 0; i < P; ++i) {
        const uint8_t* b = bytes.data() + pats[i].off;
        uint32_t s = 0;
        for (uint32_t k = pats[i].len; k-- > 0;) s = rev.child_or_add(s, b[k]);
        rev.term[s] = i + 1;
    }
    Bfs t(rev.edge, rev.term);
    SfxTables& x = sfx;
    x.n_nodes = t.n;
    // best[v] = deepest terminal on the root->v path (v included)
    std::vector<uint32_t> best(t.n, 0);
    for (uint32_t v = 1; v < t.n; ++v) best[v] = t.term[v] ? t.term[v] : best[t.parent[v]];
    // PatternsTree parent = deepest terminal among the PROPER ancestors of the pattern's node
    for (uint32_t v = 1; v < t.n; ++v)
        if (t.term[v]) pats[t.term[v] - 1].parent = best[t.parent[v]];
    for (uint32_t i = 0; i < P; ++i) {  // parents are shorter, hence not necessarily earlier pids: iterate by chain
        uint32_t c = 0;
        for (uint32_t q = pats[i].parent; q; q = pats[q - 1].parent) ++c;
        pats[i].chain = c;
    }
    if (P > 65535) return compile_parts();
    anc_off.assign(size_t(P) + 2, 0);
    anc_list.clear();
    for (uint32_t pid = 1; pid <= P; ++pid) {
        anc_off[pid] = uint32_t(anc_list.size());
        for (uint32_t q = pid; q; q = pats[q - 1].parent) anc_list.push_back(uint16_t(q));
    }
    anc_off[P + 1] = uint32_t(anc_list.size());
    uint32_t maxd = 0;
    for (uint32_t v = 0; v < t.n; ++v) maxd = std::max(maxd, t.depth[v]);
    x.depth_hist.assign(maxd + 1, 0);
    for (uint32_t v = 0; v < t.n; ++v) x.depth_hist[t.depth[v]]++;

    // byte classes: every byte that occurs in a pattern gets its own class; the rest share class 0
    bool used[256] = {false};
    for (uint8_t b : bytes) used[b] = true;
    uint32_t n_used = 0;
    for (int b = 0; b < 256; ++b) n_used += used[b];
    uint32_t next_cls = (n_used == 256) ? 0 : 1;
    for (int b = 0; b < 256; ++b) x.cls[b] = used[b] ? uint8_t(next_cls++) : 0;
    x.n_classes = next_cls ? next_cls : 1;
    x.cls_identity = (n_used == 256);
    x.log2_ncp = 0;
    while ((1u << x.log2_ncp) < x.n_classes) ++x.log2_ncp;

    // Single-pattern tails: a node whose subtree is one path down to a leaf needs no rows -- the rest of the
    // walk is a comparison against the text of that leaf's pattern (every node on the path spells a suffix of
    // it, and the terminals on the path are exactly its PatternsTree ancestors).
    std::vector<uint8_t> single(t.n, 0);
    std::vector<uint32_t> leaf_pid(t.n, 0);
    for (uint32_t v = t.n; v-- > 0;) {
        const uint32_t nc = t.off[v + 1] - t.off[v];
        if (nc == 0) { single[v] = 1; leaf_pid[v] = t.term[v]; }
        else if (nc == 1 && single[t.child[t.off[v]]]) { single[v] = 1; leaf_pid[v] = leaf_pid[t.child[t.off[v]]]; }
    }
    auto in_tail = [&](uint32_t v) { return single[v] && t.depth[v] >= kTailMinDepth; };
    // rows: internal nodes of depth >= 1 that are not folded into a tail.  The rows of the depth-2 nodes sit at
    // the row indices [cont_base, cont_base + n2c) so that a root2 "continue" code IS the row index (the scan
    // kernel forms the table index with one byte permute and no offset); all other rows fill the indices below
    // cont_base in BFS order and, if there are more of them, go on after the depth-2 block.
    std::vector<uint32_t> row_of(t.n, 0xFFFFFFFFu);
    uint32_t n_other = 0, n2c = 0, n_tail = 0;
    for (uint32_t v = 1; v < t.n; ++v)
        if (t.internal(v)) {
            if (in_tail(v)) { ++n_tail; continue; }
            if (t.depth[v] == 2) ++n2c; else ++n_other;
        }
    // continue codes must lie above every pattern id and below 65536
    uint32_t cont_base = std::max<uint32_t>(P + 1, n_other);
    if (uint64_t(cont_base) + n2c > 65536) cont_base = n2c < 65536 ? 65536 - n2c : 0;
    {
        uint32_t next_other = 0, next_d2 = cont_base;
        for (uint32_t v = 1; v < t.n; ++v) {
            if (!t.internal(v) || in_tail(v)) continue;
            if (t.depth[v] == 2) { row_of[v] = next_d2++; continue; }
            if (next_other == cont_base) next_other = cont_base + n2c;   // jump over the depth-2 block
            row_of[v] = next_other++;
        }
    }
    const uint32_t n_rows = std::max(cont_base, n_other > cont_base ? n_other : 0u) + n2c;  // incl. unused rows below cont_base
    x.n_rows = n_rows; x.row2_base = cont_base; x.n2_cont = n2c; x.n_tail_nodes = n_tail;
    x.cont_base = cont_base;
    // the backward-scan tables need every pid and every 2-byte continue code in 16 bits, and walk depths in 9 bits
    x.fits_u16 = (n2c < 65536) && (uint64_t(P) + 1 <= x.cont_base) && n_rows < (1u << 24) && max_len <= 511;

    // a FINAL entry of rows / root1 carries, above the 16-bit pid, the pattern's length (capped at 255) in bits 16-23:
    // the scan kernel's sparse mode tests "length >= min_len" with one compare and no lookup; everything that stores a
    // result keeps the low 16 bits only.  (root2 entries are u16: their patterns have at most 2 bytes.)
    auto fin = [&](uint32_t pid) -> uint32_t { return pid ? pid | (std::min<uint32_t>(pats[pid - 1].len, 255u) << 16) : 0u; };
    auto entry_for_child = [&](uint32_t c) -> uint32_t {  // walk arrives at existing child c
        if (!t.internal(c)) return fin(best[c]);
        if (in_tail(c)) return kTailFlag | leaf_pid[c];
        return kContFlag | row_of[c];
    };
    const uint32_t ncp = 1u << x.log2_ncp;
    x.rows.assign(size_t(n_rows) * ncp, 0);
    x.row_best.assign(n_rows, 0);
    for (uint32_t v = 1; v < t.n; ++v) {
        if (!t.internal(v) || in_tail(v)) continue;
        uint32_t* row = x.rows.data() + size_t(row_of[v]) * ncp;
        for (uint32_t c = 0; c < ncp; ++c) row[c] = fin(best[v]);  // path dies here: answer is best(v)
        for (uint32_t k = t.off[v]; k < t.off[v + 1]; ++k) row[x.cls[t.byte[k]]] = entry_for_child(t.child[k]);
        x.row_best[row_of[v]] = best[v];
    }
    // tail records: for a leaf pattern q the tail starts at the shallowest node of depth >= kTailMinDepth on
    // its path whose subtree is a single path
    x.tail_rec.assign(size_t(P + 1) * 4, 0);
    for (uint32_t v = 1; v < t.n; ++v) {
        if (!in_tail(v) || in_tail(t.parent[v])) continue;     // v = first node of a tail
        const uint32_t q = leaf_pid[v], d = t.depth[v];
        uint32_t next_len = pats[q - 1].len;                     // shortest chain member longer than d
        for (uint32_t a = q; a && pats[a - 1].len > d; a = pats[a - 1].parent) next_len = pats[a - 1].len;
        uint32_t* r = x.tail_rec.data() + size_t(q) * 4;
        r[0] = uint32_t(pats[q - 1].off); r[1] = pats[q - 1].len; r[2] = next_len; r[3] = best[v];
    }
    x.l3f.assign(n2c, 0);
    for (uint32_t v = 1; v < t.n; ++v) {
        if (t.depth[v] != 2 || !t.internal(v)) continue;
        uint32_t bloom = 0;
        for (uint32_t k = t.off[v]; k < t.off[v + 1]; ++k) bloom |= 1u << (t.byte[k] & 15);
        x.l3f[row_of[v] - cont_base] = (best[v] << 16) | bloom;
    }
    x.root1.assign(256, 0);
    std::vector<uint32_t> d1(256, 0);
    for (uint32_t k = t.off[0]; k < t.off[1]; ++k) {
        d1[t.byte[k]] = t.child[k];
        x.root1[t.byte[k]] = entry_for_child(t.child[k]);
    }
    x.root2.assign(65536, 0);
    if (x.fits_u16) {
        for (uint32_t a = 0; a < 256; ++a) {
            uint32_t v = d1[a];
            if (!v) continue;  // c_i starts no reversed pattern: no match
            for (uint32_t b = 0; b < 256; ++b) x.root2[(a << 8) | b] = uint16_t(best[v]);
            for (uint32_t k = t.off[v]; k < t.off[v + 1]; ++k) {
                uint32_t c = t.child[k];
                uint32_t code = t.internal(c) ? row_of[c] : best[c];   // a continue code is the row index itself
                x.root2[(a << 8) | t.byte[k]] = uint16_t(code);
            }
        }
    }
    // !fits_u16 (pids fit, but P + #2-byte continuations >= 65,536, or a pattern longer than 511 bytes): the dictionary
    // is usable, the engine serves every exact scan with the forward walkers, which have no such limit (mpac.c has none)
    compiled = true;
    return 0;
}

// Core/src/mpac.c:147-210 (goto + failure + nearest-output link), completed to a DFA:
// delta(s,c) = goto(s,c) if present else delta(fail(s),c); longest(s) = id[suffix_link(s)].
void Dict::build_dfa() const {
    std::lock_guard<std::mutex> lock(lazy_mu_);
    if (dfa.built) return;
    if (fwd_->size() == 1 && !pats.empty()) {  // loaded from a cache file: the forward trie is rebuilt from the patterns
        for (uint32_t i = 0; i < pats.size(); ++i) {
            uint32_t st = 0;
            for (uint32_t k = 0; k < pats[i].len; ++k) st = fwd_->child_or_add(st, bytes[pats[i].off + k]);
            fwd_->term[st] = i + 1;
        }
    }
    Bfs t(fwd_->edge, fwd_->term);
    DfaTables& d = dfa;
    d.n_states = t.n;
    memcpy(d.cls, sfx.cls, 256);
    d.n_classes = sfx.n_classes; d.log2_ncp = sfx.log2_ncp;
    const uint32_t ncp = 1u << d.log2_ncp;
    d.delta.assign(size_t(t.n) * ncp, 0);
    d.longest.assign(t.n, 0);
    std::vector<uint32_t> fail(t.n, 0);
    uint32_t maxd = 0;
    for (uint32_t v = 0; v < t.n; ++v) maxd = std::max(maxd, t.depth[v]);
    d.depth_count.assign(maxd + 1, 0);
    for (uint32_t v = 0; v < t.n; ++v) d.depth_count[t.depth[v]]++;
    // BFS ids are already in breadth-first order: a state's failure target has a smaller depth,
    // hence a smaller id, and is complete when the state is reached.
    for (uint32_t s = 0; s < t.n; ++s) {
        uint32_t* row = d.delta.data() + size_t(s) * ncp;
        if (s == 0) {
            for (uint32_t c = 0; c < ncp; ++c) row[c] = 0;
        } else {
            const uint32_t* frow = d.delta.data() + size_t(fail[s]) * ncp;
            memcpy(row, frow, sizeof(uint32_t) * ncp);
            d.longest[s] = t.term[s] ? uint16_t(t.term[s]) : d.longest[fail[s]];
        }
        for (uint32_t k = t.off[s]; k < t.off[s + 1]; ++k) {
            uint32_t c = t.child[k], cl = d.cls[t.byte[k]];
            fail[c] = (s == 0) ? 0 : row[cl];  // delta(fail(s), byte) is still in row[cl] (copied from frow)
            row[cl] = c;
        }
    }
    d.fb_meta.assign(t.n, 0xFFFFu);
    for (uint32_t s = 1; s < t.n; ++s) {
        if (fail[s] >= 65536) continue;
        uint32_t bloom = 0;
        for (uint32_t k = t.off[s]; k < t.off[s + 1]; ++k) bloom |= 1u << (d.cls[t.byte[k]] & 15);
        d.fb_meta[s] = bloom | (fail[s] << 16);
    }
    d.built = true;
}

// Core/src/mpac.c:147-210 again, this time kept as goto + failure + nearest-output link and packed for the deep-match
// kernel (see DeepTables in dict.hpp).
void Dict::build_deep() const {
    std::lock_guard<std::mutex> lock(lazy_mu_);
    if (deep.built) return;
    if (fwd_->size() == 1 && !pats.empty()) {  // loaded from a cache file: rebuild the forward trie from the patterns
        for (uint32_t i = 0; i < pats.size(); ++i) {
            uint32_t st = 0;
            for (uint32_t k = 0; k < pats[i].len; ++k) st = fwd_->child_or_add(st, bytes[pats[i].off + k]);
            fwd_->term[st] = i + 1;
        }
    }
    Bfs t(fwd_->edge, fwd_->term);
    DeepTables& x = deep;
    x = DeepTables();
    x.built = true;
    x.n_states = t.n;
    {
        uint32_t maxd = 0;
        for (uint32_t v = 0; v < t.n; ++v) maxd = std::max(maxd, t.depth[v]);
        x.depth_count.assign(maxd + 1, 0);
        for (uint32_t v = 0; v < t.n; ++v) x.depth_count[t.depth[v]]++;
    }
    auto go = [&](uint32_t v, uint8_t b) -> uint32_t {  // goto(v, b) or 0xFFFFFFFF; children are sorted by byte
        uint32_t lo = t.off[v], hi = t.off[v + 1];
        while (lo < hi) {
            const uint32_t mid = (lo + hi) / 2;
            if (t.byte[mid] < b) lo = mid + 1; else hi = mid;
        }
        return (lo < t.off[v + 1] && t.byte[lo] == b) ? t.child[lo] : 0xFFFFFFFFu;
    };
    // failure links and longest pids in breadth-first order (ids ARE breadth-first)
    std::vector<uint32_t> fail(t.n, 0), longest(t.n, 0);
    for (uint32_t s = 0; s < t.n; ++s) {
        if (s) longest[s] = t.term[s] ? t.term[s] : longest[fail[s]];
        for (uint32_t k = t.off[s]; k < t.off[s + 1]; ++k) {
            const uint32_t c = t.child[k];
            const uint8_t b = t.byte[k];
            uint32_t f = 0;
            if (s) {
                uint32_t v = fail[s];
                for (;;) {
                    const uint32_t g = go(v, b);
                    if (g != 0xFFFFFFFFu) { f = g; break; }
                    if (!v) break;
                    v = fail[v];
                }
            }
            fail[c] = f;
        }
    }
    auto delta = [&](uint32_t s, uint8_t b) -> uint32_t {  // complete DFA transition
        for (;;) {
            const uint32_t g = go(s, b);
            if (g != 0xFFFFFFFFu) return g;
            if (!s) return 0;
            s = fail[s];
        }
    };
    // numbering: depth <= 1 ("hot"), then every DENSE state (more than 6 children, any depth >= 2), then the other depth-2
    // states -- all of these are "small" ids whose longest pid lives in shared memory -- then the rest depth-first
    auto n_children = [&](uint32_t v) { return t.off[v + 1] - t.off[v]; };
    auto is_dense = [&](uint32_t v) { return t.depth[v] >= 2 && n_children(v) > 6; };
    uint32_t n_le1 = 0;
    for (uint32_t v = 0; v < t.n; ++v) n_le1 += t.depth[v] <= 1;
    std::vector<uint32_t> nid(t.n, 0xFFFFFFFFu);
    uint32_t next = 0;
    for (uint32_t v = 0; v < n_le1; ++v) nid[v] = next++;   // breadth-first ids of depth <= 1 are already 0 .. n_le1-1
    for (uint32_t v = n_le1; v < t.n; ++v) if (is_dense(v)) nid[v] = next++;
    const uint32_t dense_end = next;
    std::vector<uint32_t> d2;
    for (uint32_t v = n_le1; v < t.n && t.depth[v] == 2; ++v) { d2.push_back(v); if (nid[v] == 0xFFFFFFFFu) nid[v] = next++; }
    const uint32_t n_small = next;
    {
        std::vector<uint32_t> stack;
        for (uint32_t v : d2) {      // subtrees of the depth-2 states, in order
            for (uint32_t k = t.off[v + 1]; k-- > t.off[v];) stack.push_back(t.child[k]);
            while (!stack.empty()) {
                const uint32_t u = stack.back();
                stack.pop_back();
                if (nid[u] == 0xFFFFFFFFu) nid[u] = next++;   // DENSE states already have their id
                for (uint32_t k = t.off[u + 1]; k-- > t.off[u];) stack.push_back(t.child[k]);   // reversed: smallest byte on top
            }
        }
    }
    std::vector<uint32_t> old_of(t.n, 0);
    for (uint32_t v = 0; v < t.n; ++v) old_of[nid[v]] = v;
    x.n_hot = n_le1; x.n_dense = dense_end - n_le1; x.n_small = n_small;
    if (n_small > 65535 || t.n >= (1u << 24) || pats.size() > 65535) return;   // usable stays false
    x.hot_rows.assign(size_t(n_le1) << 8, 0);
    x.hot_longest.assign(n_small, 0);
    for (uint32_t id = 0; id < n_small; ++id) x.hot_longest[id] = uint16_t(longest[old_of[id]]);
    for (uint32_t s = 0; s < n_le1; ++s)
        for (uint32_t b = 0; b < 256; ++b) x.hot_rows[(size_t(s) << 8) | b] = uint16_t(nid[delta(s, uint8_t(b))]);
    x.dense_rows.assign(size_t(x.n_dense) << 8, 0);
    for (uint32_t id = n_le1; id < dense_end; ++id) {
        const uint32_t v = old_of[id];
        for (uint32_t b = 0; b < 256; ++b) x.dense_rows[(size_t(id - n_le1) << 8) | b] = nid[delta(v, uint8_t(b))];
    }
    x.recs.assign(size_t(t.n) * 8, 0);
    for (uint32_t id = dense_end; id < t.n; ++id) {
        const uint32_t v = old_of[id];
        uint32_t* r = x.recs.data() + size_t(id) * 8;
        const uint32_t nc = n_children(v);
        r[1] = longest[v];
        uint32_t kind = 0, count = 0;
        if (nc == 1 && nid[t.child[t.off[v]]] == id + 1) {   // CHAIN: follow single children while they are the next id
            kind = 1;
            uint32_t u = v;
            uint64_t labels = 0;
            while (count < 8 && n_children(u) == 1 && nid[t.child[t.off[u]]] == id + count + 1) {
                const uint32_t c = t.child[t.off[u]];
                labels |= uint64_t(t.byte[t.off[u]]) << (8 * count);
                r[4 + count / 2] |= (longest[c] & 0xFFFFu) << (16 * (count & 1));
                ++count;
                u = c;
            }
            r[2] = uint32_t(labels); r[3] = uint32_t(labels >> 32);
            ++x.n_chain;
        } else if (nc == 0) {                          // LEAF: every byte follows the failure link
            kind = 2;
            ++x.n_branch;
        } else {                                       // BRANCH: 1 .. 6 goto edges; unused slots repeat the first edge
            for (uint32_t k = t.off[v]; k < t.off[v + 1]; ++k) r[2 + count++] = (nid[t.child[k]] << 8) | t.byte[k];
            for (uint32_t k = count; k < 6; ++k) r[2 + k] = r[2];
            ++x.n_branch;
        }
        r[0] = nid[fail[v]] | (kind << 24) | (count << 26);
    }
    x.usable = true;
}

// ---- compiled-automaton cache -----------------------------------------------------------------
// One binary file: magic (carries the layout version), the scalar block, the table vectors, and a trailing FNV-1a
// checksum over everything before it.  Written to a temporary name in the same directory and renamed into place, so
// that concurrent ranks (one process per GPU) never see a half-written file; load() trusts nothing: it verifies the
// checksum and every size / index bound the kernels rely on, and any mismatch means "compile instead".
namespace {
constexpr char kMagic[8] = {'P', 'M', 'B', '2', 'D', 'I', 'C', '4'};
struct Hasher {
    uint64_t h = 1469598103934665603ull;
    void mix(const void* p, size_t n) {
        const uint8_t* b = static_cast<const uint8_t*>(p);
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    }
};
struct Writer {
    FILE* f; Hasher hs; bool ok = true;
    void raw(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; hs.mix(p, n); }
    template <class T> void vec(const std::vector<T>& v) { const uint64_t n = v.size(); raw(&n, 8); raw(v.data(), n * sizeof(T)); }
};
struct Reader {
    FILE* f; Hasher hs; bool ok = true;
    void raw(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) ok = false; else hs.mix(p, n); }
    template <class T> void vec(std::vector<T>& v) {
        uint64_t n = 0;
        raw(&n, 8);
        if (!ok || n > (uint64_t(1) << 32)) { ok = false; return; }
        v.resize(n);
        raw(v.data(), n * sizeof(T));
    }
};
struct Scalars {
    uint64_t n_lines, n_rejected, n_dups;
    uint32_t n_files, max_len, n_ac_states;
    uint32_t n_nodes, n_rows, n_tail_nodes, row2_base, n2_cont, cont_base, fits_u16, n_classes, log2_ncp;
    uint8_t cls[256];
};
}  // namespace

int Dict::save(const char* path) const {
    if (!compiled || !sfx.fits_u16) return -1;   // only dictionaries with backward-scan tables are cached
    const std::string tmp = std::string(path) + ".tmp." + std::to_string(uint64_t(getpid()));
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return -1;
    Scalars sc{};
    sc.n_lines = n_lines; sc.n_rejected = n_rejected; sc.n_dups = n_dups;
    sc.n_files = n_files; sc.max_len = max_len; sc.n_ac_states = n_ac_states;
    sc.n_nodes = sfx.n_nodes; sc.n_rows = sfx.n_rows; sc.n_tail_nodes = sfx.n_tail_nodes; sc.row2_base = sfx.row2_base;
    sc.n2_cont = sfx.n2_cont; sc.cont_base = sfx.cont_base; sc.fits_u16 = sfx.fits_u16; sc.n_classes = sfx.n_classes;
    sc.log2_ncp = sfx.log2_ncp;
    memcpy(sc.cls, sfx.cls, 256);
    Writer w{f};
    w.raw(kMagic, 8); w.raw(&sc, sizeof(sc));
    w.vec(pats); w.vec(bytes); w.vec(anc_off); w.vec(anc_list); w.vec(sfx.root2); w.vec(sfx.l3f); w.vec(sfx.root1);
    w.vec(sfx.rows); w.vec(sfx.row_best); w.vec(sfx.tail_rec); w.vec(sfx.depth_hist);
    const uint64_t sum = w.hs.h;
    bool ok = w.ok && fwrite(&sum, 8, 1, f) == 1;
    ok = (fclose(f) == 0) && ok;
    if (ok && rename(tmp.c_str(), path) != 0) ok = false;
    if (!ok) remove(tmp.c_str());
    return ok ? 0 : -1;
}

int Dict::load(const char* path) {
    if (compiled || !pats.empty()) { error = "load needs an empty dictionary"; return -1; }
    FILE* f = fopen(path, "rb");
    if (!f) { error = std::string("cannot open ") + path; return -1; }
    char magic[8];
    Scalars sc{};
    Reader r{f};
    r.raw(magic, 8); r.raw(&sc, sizeof(sc));
    bool ok = r.ok && memcmp(magic, kMagic, 8) == 0;
    if (ok) {
        r.vec(pats); r.vec(bytes); r.vec(anc_off); r.vec(anc_list); r.vec(sfx.root2); r.vec(sfx.l3f); r.vec(sfx.root1);
        r.vec(sfx.rows); r.vec(sfx.row_best); r.vec(sfx.tail_rec); r.vec(sfx.depth_hist);
        uint64_t sum = 0;
        ok = r.ok && fread(&sum, 8, 1, f) == 1 && sum == r.hs.h;
    }
    fclose(f);
    // structural checks: every size and index the kernels use without looking
    const uint64_t P = pats.size();
    if (ok) {
        ok = sc.fits_u16 == 1 && sc.log2_ncp <= 8 && sc.n_classes >= 1 && sc.n_classes <= (1u << sc.log2_ncp) &&
             sfx.root2.size() == 65536 && sfx.root1.size() == 256 && sc.n_rows < (1u << 24) &&
             sfx.rows.size() == (size_t(sc.n_rows) << sc.log2_ncp) && sfx.row_best.size() == sc.n_rows &&
             sfx.tail_rec.size() == 4 * (P + 1) && sfx.l3f.size() == sc.n2_cont && anc_off.size() == P + 2 &&
             P + 1 <= sc.cont_base && uint64_t(sc.cont_base) + sc.n2_cont <= 65536 && sc.row2_base == sc.cont_base &&
             uint64_t(sc.cont_base) + sc.n2_cont <= sc.n_rows && sc.max_len >= (P ? 1u : 0u);
    }
    for (uint64_t i = 0; ok && i < P; ++i)
        ok = pats[i].len >= 1 && pats[i].len <= sc.max_len && pats[i].off + pats[i].len <= bytes.size() && pats[i].parent <= P;
    for (uint64_t i = 0; ok && i + 1 < anc_off.size(); ++i) ok = anc_off[i] <= anc_off[i + 1] && anc_off[i + 1] <= anc_list.size();
    for (size_t i = 0; ok && i < anc_list.size(); ++i) ok = anc_list[i] >= 1 && anc_list[i] <= P;
    auto entry_ok = [&](uint32_t v) {
        if (v & kContFlag) return (v & 0xFFFFFFu) < sc.n_rows;
        if (v & kTailFlag) return (v & 0xFFFFu) >= 1 && (v & 0xFFFFu) <= P;
        return (v & 0xFFFFu) <= P && (v >> 24) == 0 && ((v >> 16) == 0) == ((v & 0xFFFFu) == 0);
    };
    for (size_t i = 0; ok && i < sfx.rows.size(); ++i) ok = entry_ok(sfx.rows[i]);
    for (size_t i = 0; ok && i < sfx.root1.size(); ++i) ok = entry_ok(sfx.root1[i]);
    for (size_t i = 0; ok && i < sfx.row_best.size(); ++i) ok = sfx.row_best[i] <= P;
    for (size_t i = 0; ok && i < sfx.root2.size(); ++i) ok = sfx.root2[i] <= P || (sfx.root2[i] >= sc.cont_base && sfx.root2[i] < sc.cont_base + sc.n2_cont);
    for (uint64_t q = 1; ok && q <= P; ++q) {
        const uint32_t* t = sfx.tail_rec.data() + 4 * q;
        ok = uint64_t(t[0]) + t[1] <= bytes.size() && t[1] <= sc.max_len && t[3] <= P;
    }
    if (!ok) {
        pats.clear(); bytes.clear(); anc_off.clear(); anc_list.clear(); sfx = SfxTables();
        error = std::string("not a compiled dictionary of this version (or damaged): ") + path;
        return -1;
    }
    n_lines = sc.n_lines; n_rejected = sc.n_rejected; n_dups = sc.n_dups;
    n_files = sc.n_files; max_len = sc.max_len; n_ac_states = sc.n_ac_states;
    sfx.n_nodes = sc.n_nodes; sfx.n_rows = sc.n_rows; sfx.n_tail_nodes = sc.n_tail_nodes; sfx.row2_base = sc.row2_base;
    sfx.n2_cont = sc.n2_cont; sfx.cont_base = sc.cont_base; sfx.fits_u16 = sc.fits_u16 != 0; sfx.n_classes = sc.n_classes;
    sfx.log2_ncp = sc.log2_ncp;
    memcpy(sfx.cls, sc.cls, 256);
    sfx.cls_identity = true;
    for (int b = 0; b < 256; ++b) sfx.cls_identity = sfx.cls_identity && sfx.cls[b] == b;
    compiled = true;
    return 0;
}

KrTables Dict::build_kr(uint64_t seed) const {
    KrTables k;
    k.seed = seed;
    k.r = 1 + splitmix64(seed) % (kKrP - 1);
    struct Cand { uint32_t fp8, pid, len; };
    std::vector<Cand> cs;
    for (uint32_t i = 0; i < pats.size(); ++i)
        if (pats[i].len > 8)
            cs.push_back({uint32_t(kr_fp(bytes.data() + pats[i].off + pats[i].len - 8, 8, k.r)), i + 1, pats[i].len});
    k.n_long = uint32_t(cs.size());
    std::sort(cs.begin(), cs.end(), [](const Cand& a, const Cand& b) {
        if (a.fp8 != b.fp8) return a.fp8 < b.fp8;
        if (a.len != b.len) return a.len > b.len;  // longest candidate first
        return a.pid < b.pid;
    });
    uint32_t n_keys = 0;
    for (size_t i = 0; i < cs.size(); ++i) n_keys += (i == 0 || cs[i].fp8 != cs[i - 1].fp8);
    k.bucket_bits = 4;
    while ((1u << k.bucket_bits) < 2 * n_keys + 1) ++k.bucket_bits;
    const uint32_t mask = (1u << k.bucket_bits) - 1;
    k.slot_fp.assign(size_t(1) << k.bucket_bits, 0xFFFFFFFFu);
    k.slot_begin.assign(size_t(1) << k.bucket_bits, 0);
    k.slot_count.assign(size_t(1) << k.bucket_bits, 0);
    k.bloom_bits = 19;  // 64 KiB of shared memory
    k.bloom.assign((size_t(1) << k.bloom_bits) / 32, 0);
    for (size_t i = 0; i < cs.size(); ++i) {
        const Pattern& p = pats[cs[i].pid - 1];
        const uint8_t* b = bytes.data() + p.off;
        if (i == 0 || cs[i].fp8 != cs[i - 1].fp8) {
            uint32_t h = uint32_t(splitmix64(cs[i].fp8)) & mask;
            while (k.slot_fp[h] != 0xFFFFFFFFu) h = (h + 1) & mask;
            k.slot_fp[h] = cs[i].fp8;
            k.slot_begin[h] = uint32_t(i);
            for (int hk = 0; hk < kKrBloomHashes; ++hk) {
                const uint32_t bit = kr_bloom_bit(cs[i].fp8, hk);
                k.bloom[bit >> 5] |= 1u << (bit & 31);
            }
        }
        uint32_t h = uint32_t(splitmix64(cs[i].fp8)) & mask;
        while (k.slot_fp[h] != cs[i].fp8) h = (h + 1) & mask;
        k.slot_count[h]++;
        k.cand_pid.push_back(cs[i].pid);
        k.cand_len.push_back(p.len);
        k.cand_stage_off.push_back(uint32_t(k.stage_fp.size()));
        for (uint32_t l = 16; l <= p.len; l <<= 1) k.stage_fp.push_back(uint32_t(kr_fp(b + p.len - l, l, k.r)));
        k.stage_fp.push_back(uint32_t(kr_fp(b, p.len, k.r)));
    }
    k.built = true;
    return k;
}

}  // namespace pm
