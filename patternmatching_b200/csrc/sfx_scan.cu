// sfx_scan.cu -- exact dictionary scan for sm_100a, dense uint16 result.
//
// Semantics (bit-exact): out[i] = pid of the LONGEST dictionary pattern that is a suffix of
// stream[..i], 0 if none -- what ac_read_char returns per byte (Core/src/mpac.c:304-319) when driven
// by the loop at Core/src/measure.c:292-294.
//
// Method.  For position i the longest pattern ending at i is the deepest terminal on the path
// c[i], c[i-1], c[i-2], ... in the trie of REVERSED patterns.  The host compiler (dict.cpp) has
// flattened that trie so that every table entry is either the final answer or "continue at row r":
//   level 1+2 : root2[c[i] << 8 | c[i-1]]   65,536 x u16 = 128 KiB, resident in SHARED memory
//   level 3   : rows[row << log2_ncp | cls(c[i-2])]   u32, global memory (6.9 MB, L2-resident)
//   level 4   : same rows, second round of predicated loads for the ~1e-3 of positions still alive
//   level >=5 : same rows / pattern-text tails; real or planted matches, ~1e-5 of random positions --
//               DEFERRED to a work queue and finished by sfx_deep_kernel, so that a long dependent
//               chain never stalls a scanning warp
//   level 3   : rows[...], through the TEXTURE pipe (plain path) or behind a shared-memory Bloom word (filter path)
// All positions are independent, so there is no per-thread warm-up and no CTA-wide barrier after
// start-up.  A warp visits 512 consecutive positions at a time; a lane resolves 2 x 8 consecutive
// positions per visit, arranged so that the 16-byte result stores of a warp are fully coalesced, and
// gets its stream bytes with two 8-byte loads issued one visit ahead.  Kernels: sfx_scan_kernel (levels
// 1-4, every position), sfx_deep_kernel (the parked walks), sfx_edge_kernel (the ragged ends).
#include <type_traits>

#include "pm_dev.cuh"
#include "sfx_scan.cuh"

namespace pm {

namespace {

constexpr int kThreads = kSfxThreads;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = kSfxTile;                  // positions per warp visit
static_assert(kTile == 512, "a visit is 2 groups of 8 positions per lane");
constexpr uint32_t kCont = 0x80000000u;   // entry: continue at row (entry & 0xFFFFFF)
constexpr uint32_t kTail = 0x40000000u;   // entry: the rest of the path is the text of pattern (entry & 0xFFFF)
constexpr uint32_t kAlive = kCont | kTail;

// shared memory carve-up (bytes)
constexpr int kOffRoot2 = 0;                     // 131072
constexpr int kOffCls = 131072;                  // 256
constexpr int kOffQCnt = kOffCls + 256;          // 16 (three counters, padded)
constexpr int kOffL3 = kOffQCnt + 16;            // kSfxMaxL3 * 4
constexpr int kSmemBytes = kOffL3 + int(kSfxMaxL3) * 4;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kCont == 0x80000000u, "level4_group tests the sign bit");

// byte / 16-bit window extraction from the 12-byte register window W = {c[g-4..g-1], c[g..g+3], c[g+4..g+7]}
template <int O>
__device__ __forceinline__ uint32_t win_u8(const uint32_t (&W)[3]) {
    return __byte_perm(W[O >> 2], 0u, 0x4440u | uint32_t(O & 3));  // one PRMT, zero-extended
}
template <int O>
__device__ __forceinline__ uint32_t win_u16(const uint32_t (&W)[3]) {
    if constexpr ((O & 3) < 3) return __byte_perm(W[O >> 2], 0u, 0x4400u | uint32_t(((O & 3) + 1) << 4) | uint32_t(O & 3));
    else return __funnelshift_r(W[O >> 2], W[(O >> 2) + 1], 24) & 0xFFFFu;
}

// (e << 8) | byte O of the window, for a 16-bit entry e: one PRMT {byte O of W, e.b0, e.b1, e.b2 (= 0)}
template <int O>
__device__ __forceinline__ uint32_t win_row_index(const uint32_t (&W)[3], uint32_t e) {
    return __byte_perm(W[O >> 2], e, 0x6540u | uint32_t(O & 3));
}

// The 8 bytes that END at address a (a[-7..0]) as a little-endian word, a[0] in the top byte, from two aligned
// loads.  Bytes below `floor` are not read (they may lie outside the allocation) and come back as zero.
__device__ __forceinline__ uint64_t load8_ending_at(const uint8_t* a, const uint8_t* floor) {
    const uintptr_t ua = reinterpret_cast<uintptr_t>(a);
    const uint64_t* hi_p = reinterpret_cast<const uint64_t*>(ua & ~uintptr_t(7));
    const uint32_t sh = uint32_t(ua & 7) * 8;  // a[0] is byte (ua & 7) of the high word
    const uint64_t hi = __ldg(hi_p);
    uint64_t lo = 0;
    if (sh != 56 && reinterpret_cast<const uint8_t*>(hi_p) > floor) lo = __ldg(hi_p - 1);
    return sh == 56 ? hi : ((hi << (56 - sh)) | (lo >> (sh + 8)));
}

// Finish a walk that is alive after k consumed bytes, straight from global memory: follow rows while the
// entry says "continue"; once it says "tail of pattern q" the remaining path is the text of q, so compare the
// stream against it (8 bytes per step) and answer with the longest of {q and its PatternsTree ancestors} that
// fits the matched length (the terminals on a tail are exactly those patterns).  `ci` points at c[i] in the
// stream; avail = number of stream bytes that exist up to and including c[i].
__device__ __noinline__ uint32_t sfx_finish(const SfxParams& p, uint32_t v, uint64_t k, const uint8_t* __restrict__ ci,
                                            uint64_t avail) {
    while (v & kCont) {
        const uint32_t row = v & 0xFFFFFFu;
        if (k >= avail) return __ldg(p.row_best + row);
        v = __ldg(p.rows + ((size_t(row) << p.log2_ncp) | __ldg(p.cls + *(ci - k))));
        ++k;
    }
    if (v & kTail) {
        const uint32_t q = v & 0xFFFFu;
        const uint4 rec = __ldg(p.tail_rec + q);             // x = text offset, y = length, z = next terminal, w = best
        const uint32_t len = rec.y;
        const uint8_t* text = p.pat_bytes + rec.x;           // q's text; the k consumed bytes are its last k bytes
        const uint64_t lim = uint64_t(len) < avail ? uint64_t(len) : avail;
        const uint8_t* floor_s = ci - (avail - 1);
        uint64_t m = k;
        while (m < lim) {
            const uint64_t a = load8_ending_at(ci - m, floor_s);
            const uint64_t b = load8_ending_at(text + (len - 1 - m), p.pat_bytes);
            const uint64_t x = a ^ b;
            const uint64_t same = x ? uint64_t(__clzll((long long)x) >> 3) : 8;   // equal bytes from the top (= backwards)
            const uint64_t left = lim - m;
            m += same < left ? same : left;
            if (same < 8) break;
        }
        if (m < rec.z) return rec.w;                          // no further terminal reached: best at the tail start
        uint32_t cand = q;
        while (cand && uint64_t(__ldg(p.pat_len + cand - 1)) > m) cand = __ldg(p.parent + cand);
        return cand;
    }
    return v;
}

// sparse mode: is the longest match `pid` a pattern of at least min_len bytes?  (rare paths only: deferred walks, edges)
__device__ __forceinline__ bool is_long(const SfxParams& p, uint32_t pid) {
    pid &= 0xFFFFu;
    return pid != 0 && __ldg(p.pat_len + pid - 1) >= p.min_len;
}
__device__ __forceinline__ void flag_set(const SfxParams& p, uint64_t pos, bool on) {
    uint32_t* w = reinterpret_cast<uint32_t*>(p.flags) + (pos >> 5);
    const uint32_t bit = 1u << (pos & 31);
    if (on) atomicOr(w, bit); else atomicAnd(w, ~bit);
}

// a deferred walk's result: dense slot, and in sparse mode (kFlags) its flag bit (the scan kernel left it clear).
// The mode is a template parameter of the deep kernel: a run-time test of p.flags here cost the dense mode 552 bytes
// of register spills at 64 registers per thread (deep kernel 1.15 -> 1.77 ms per 16 GiB of planted traffic).
template <bool kFlags>
__device__ __forceinline__ void put_result(const SfxParams& p, uint64_t pos, uint32_t pid) {
    p.out[pos] = uint16_t(pid);
    if constexpr (kFlags) { if (is_long(p, pid)) flag_set(p, pos, true); }
}

// Where the main kernel finds its tables.  The 2-byte root table sits in shared memory (LSU pipe); the rows of
// levels 3 and 4 are read through the TEXTURE pipe when the table fits a linear texture (kTex): the kernel is bound
// by the LSU data pipe of the L1 (shared-memory gathers with ~3.5-way bank conflicts), and the TEX pipe is a
// second, otherwise idle, path into the same cache -- sparse predicated fetches cost it ~1.5 cycles each.
struct MainTabs {
    const uint16_t* s_root2;
    const uint32_t* s_l3;
    const uint8_t* s_cls;
    const uint32_t* rows;
    cudaTextureObject_t rows_tex;
    uint32_t cont_base, log2_ncp;   // a root2 entry >= cont_base is the index of the row to continue at
};
template <bool kTex>
__device__ __forceinline__ uint32_t fetch_row(const MainTabs& t, uint32_t idx) {  // idx = (row << log2_ncp) | class
    if constexpr (kTex) return tex1Dfetch<unsigned int>(t.rows_tex, int(idx));
    else return __ldg(t.rows + idx);
}

// Levels 1-3 for one group of 8 consecutive positions whose bytes sit in the register window W.
// Phase A: 8 shared-memory gathers from root2.  Phase B, plain path (!kL3): an
// entry that continues below the 2-byte table fetches its row entry with a PREDICATED load -- no branch and no use of
// the loaded value here, so all loads of a visit are in flight together.  Filter path (kL3): it first consults the
// level-3 filter word of its row in SHARED memory: unless the Bloom bit of c[i-2] is set the answer is the row's own
// best pid; only Bloom hits (true children 0.9%, false positives ~14% of the continuing entries) fetch the row entry.
template <bool kIdentCls, bool kL3, bool kTex>
__device__ __forceinline__ bool lookup_group(const MainTabs& t, const uint32_t (&W)[3], uint32_t (&e)[8]) {
    const uint16_t* s_root2 = t.s_root2;
    const uint32_t cont_base = t.cont_base;
    e[0] = s_root2[win_u16<3>(W)]; e[1] = s_root2[win_u16<4>(W)];
    e[2] = s_root2[win_u16<5>(W)]; e[3] = s_root2[win_u16<6>(W)];
    e[4] = s_root2[win_u16<7>(W)]; e[5] = s_root2[win_u16<8>(W)];
    e[6] = s_root2[win_u16<9>(W)]; e[7] = s_root2[win_u16<10>(W)];
    const bool first_cont = e[0] >= cont_base;  // sample for the adaptive choice of the level-3 path
    if constexpr (kL3) {
        // idx = (entry << 8) | c[i-2]: one byte permute; its low nibble picks the Bloom bit
        uint32_t idx[8];
        idx[0] = win_row_index<2>(W, e[0]); idx[1] = win_row_index<3>(W, e[1]);
        idx[2] = win_row_index<4>(W, e[2]); idx[3] = win_row_index<5>(W, e[3]);
        idx[4] = win_row_index<6>(W, e[4]); idx[5] = win_row_index<7>(W, e[5]);
        idx[6] = win_row_index<8>(W, e[6]); idx[7] = win_row_index<9>(W, e[7]);
        const uint32_t* l3_adj = t.s_l3 - cont_base;   // indexed by the root2 entry itself
        uint32_t f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            f[j] = 0;
            if (e[j] >= cont_base) f[j] = l3_adj[e[j]];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool cont = e[j] >= cont_base;
            const bool hit = ((f[j] >> (idx[j] & 15u)) & 1u) != 0;   // f == 0 when the entry does not continue
            uint32_t ix = idx[j];
            if constexpr (!kIdentCls) {
                if (hit) ix = (e[j] << t.log2_ncp) | t.s_cls[idx[j] & 0xFFu];
            }
            if (cont) e[j] = f[j] >> 16;
            if (hit) e[j] = fetch_row<false>(t, ix);   // ~1.5% of the lanes: not worth a TEX instruction
        }
    } else if constexpr (kIdentCls) {
        // 256 byte classes: the row index (e << 8 | c[i-2]) is ONE byte permute of the entry and the window word
        uint32_t idx[8];
        idx[0] = win_row_index<2>(W, e[0]); idx[1] = win_row_index<3>(W, e[1]);
        idx[2] = win_row_index<4>(W, e[2]); idx[3] = win_row_index<5>(W, e[3]);
        idx[4] = win_row_index<6>(W, e[4]); idx[5] = win_row_index<7>(W, e[5]);
        idx[6] = win_row_index<8>(W, e[6]); idx[7] = win_row_index<9>(W, e[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (e[j] >= cont_base) e[j] = fetch_row<kTex>(t, idx[j]);
    } else {
        uint32_t c2[8];
        c2[0] = win_u8<2>(W); c2[1] = win_u8<3>(W); c2[2] = win_u8<4>(W); c2[3] = win_u8<5>(W);
        c2[4] = win_u8<6>(W); c2[5] = win_u8<7>(W); c2[6] = win_u8<8>(W); c2[7] = win_u8<9>(W);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (e[j] >= cont_base) e[j] = fetch_row<kTex>(t, (e[j] << t.log2_ncp) | t.s_cls[c2[j]]);
        }
    }
    return first_cont;
}

// Level 4 for the entries of a group that are still "continue" after level 3: c[i-3] is byte 1+j of W.
template <bool kIdentCls, bool kTex>
__device__ __forceinline__ void level4_group(const MainTabs& t, const uint32_t (&W)[3], uint32_t (&e)[8]) {
    if constexpr (kIdentCls) {
        // (row << 8) | c[i-3] is one byte permute (it drops the flag byte of the entry); kCont is the sign bit
        uint32_t idx[8];
        idx[0] = win_row_index<1>(W, e[0]); idx[1] = win_row_index<2>(W, e[1]);
        idx[2] = win_row_index<3>(W, e[2]); idx[3] = win_row_index<4>(W, e[3]);
        idx[4] = win_row_index<5>(W, e[4]); idx[5] = win_row_index<6>(W, e[5]);
        idx[6] = win_row_index<7>(W, e[6]); idx[7] = win_row_index<8>(W, e[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (int32_t(e[j]) < 0) e[j] = fetch_row<kTex>(t, idx[j]);
    } else {
        uint32_t c3[8];
        c3[0] = win_u8<1>(W); c3[1] = win_u8<2>(W); c3[2] = win_u8<3>(W); c3[3] = win_u8<4>(W);
        c3[4] = win_u8<5>(W); c3[5] = win_u8<6>(W); c3[6] = win_u8<7>(W); c3[7] = win_u8<8>(W);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (int32_t(e[j]) < 0) e[j] = fetch_row<kTex>(t, ((e[j] & 0xFFFFFFu) << t.log2_ncp) | t.s_cls[c3[j]]);
        }
    }
}

__device__ __forceinline__ void store_group(uint16_t* out, uint64_t gpos, const uint32_t (&e)[8]) {
    uint4 r;
    r.x = __byte_perm(e[0], e[1], 0x5410); r.y = __byte_perm(e[2], e[3], 0x5410);  // low halves, one PRMT each
    r.z = __byte_perm(e[4], e[5], 0x5410); r.w = __byte_perm(e[6], e[7], 0x5410);
    __stcs(reinterpret_cast<uint4*>(out + gpos), r);  // write-once result: streaming store
}

// 8 stream bytes straight to registers: read-only path, not allocated in L1 (every byte is read exactly once)
__device__ __forceinline__ uint2 ldg_stream8(const uint8_t* ptr) {
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(ptr));
    return v;
}
__device__ __forceinline__ uint32_t ldg_stream4(const uint8_t* ptr) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(ptr));
    return v;
}

// One visit = 512 consecutive positions of one warp: lane l resolves positions 8l..8l+7 (group A, bytes in `a`) and
// 256+8l.. (group B, bytes in `b`), so that the two 16-byte result stores of a warp are fully coalesced.  The
// stream bytes go from global memory straight into registers, one visit ahead of their use: staging them in shared
// memory (bulk async copies into a per-warp double buffer, the first design of round 1) made every byte cross the
// LSU data pipe this kernel is bound by a second time, and cost ~56 warp instructions of bookkeeping per tile.
// kFlags: sparse mode -- besides the dense result, one bit per position: "the longest match has >= min_len bytes".
// A final entry of levels 3/4 carries the pattern length above the pid (dict.cpp), so the test is one compare; entries
// that ended in the 2-byte root table are patterns of <= 2 bytes and never qualify (min_len >= 3); deferred walks set
// their bit when they are finished.  A lane packs its 8 bits into one byte: 32 coalesced bytes per group and warp.
template <bool kIdentCls, bool kTex, bool kFlags>
__global__ void __launch_bounds__(kThreads, 1) sfx_scan_kernel(const SfxParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint16_t* s_root2 = reinterpret_cast<uint16_t*>(smem + kOffRoot2);
    uint8_t* s_cls = smem + kOffCls;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t* s_qcnt = reinterpret_cast<uint32_t*>(smem + kOffQCnt);
    uint32_t* s_l3 = reinterpret_cast<uint32_t*>(smem + kOffL3);
    MainTabs t;
    t.s_root2 = s_root2; t.s_l3 = s_l3; t.s_cls = s_cls; t.rows = p.rows;
    t.rows_tex = p.rows_tex;
    t.cont_base = p.cont_base; t.log2_ncp = p.log2_ncp;

    if (tid < 256) s_cls[tid] = p.cls[tid];
    if (tid == 0) { s_qcnt[0] = 0; s_qcnt[1] = 0; s_qcnt[2] = 0; }  // slots used / tail items / row items
    const bool have_l3 = p.l3f != nullptr;
    if (have_l3) for (uint32_t i = tid; i < p.n_l3; i += kThreads) s_l3[i] = __ldg(p.l3f + i);
    // Per-warp choice of the level-3 path, re-made every visit from a 32-position sample of the previous one.
    // The plain path (predicated texture fetches) waits for L2, the filter path (a shared-memory Bloom word first)
    // costs LSU wavefronts and ALU; on binary bytes (~10% of the positions continue below root2) the threshold
    // sends ~40% of the visits through the filter, which keeps the LSU pipe busy while the other warps wait for
    // their fetches; on text (~60% continue) all visits take it.
    bool use_l3a = false, use_l3b = false;

    const uint64_t gw = uint64_t(blockIdx.x) * kWarps + warp;   // global warp id
    const uint64_t G = uint64_t(gridDim.x) * kWarps;
    const uint64_t n_vis = p.n_tiles;                           // full visits; the ragged end belongs to sfx_edge_kernel
    const bool hist4 = p.hist_valid >= 4;

    // the visit's bytes: a = [8l, 8l+8), b = [256+8l, ..), h (lane 0) = the 4 bytes before the visit
    // Both pointers are carried from visit to visit (one 64-bit add each) instead of being rebuilt from the visit index.
    const uint8_t* in_ptr = p.stream + gw * uint64_t(kTile) + 8 * lane;   // this lane's bytes of the NEXT visit to load
    uint16_t* out_ptr = p.out + gw * uint64_t(kTile) + 8 * lane;          // this lane's results of the current visit
    const uint64_t step = G * uint64_t(kTile);
    auto load_visit = [&](bool first, uint2& a, uint2& b, uint32_t& h) {
        a = ldg_stream8(in_ptr);
        b = ldg_stream8(in_ptr + 256);
        h = 0;
        if (lane == 0 && (!first || hist4)) h = ldg_stream4(in_ptr - 4);
        in_ptr += step;
    };
    uint2 a = make_uint2(0, 0), b = make_uint2(0, 0);
    uint32_t h = 0;
    if (gw < n_vis) load_visit(gw == 0, a, b, h);

    // root2 -> shared memory (once per CTA; coalesced 16-byte loads, L2 hits after the first CTA)
    {
        const int4* src = reinterpret_cast<const int4*>(p.root2);
        int4* dst = reinterpret_cast<int4*>(s_root2);
        for (int i = tid; i < 131072 / 16; i += kThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();  // the only CTA-wide barrier before the end

    uint64_t* q_strip = p.queue + size_t(blockIdx.x) * p.q_per_cta;  // this CTA's strip of the deferred-walk queue
    const int ga = 8 * lane, gb = ga + 256;
#pragma unroll 1
    for (uint64_t v = gw; v < n_vis; v += G) {
        const uint64_t s0 = v * uint64_t(kTile);
        uint32_t WA[3], WB[3];
        WA[1] = a.x; WA[2] = a.y; WB[1] = b.x; WB[2] = b.y;
        WA[0] = __shfl_up_sync(0xFFFFFFFFu, a.y, 1);
        WB[0] = __shfl_up_sync(0xFFFFFFFFu, b.y, 1);
        const uint32_t a31 = __shfl_sync(0xFFFFFFFFu, a.y, 31);
        if (lane == 0) { WA[0] = h; WB[0] = a31; }
        if (v + G < n_vis) load_visit(false, a, b, h);   // next visit's bytes, in flight during this one

        uint32_t ea[8], eb[8];
        bool sample_cont;
        if (use_l3a) sample_cont = lookup_group<kIdentCls, true, kTex>(t, WA, ea);
        else sample_cont = lookup_group<kIdentCls, false, kTex>(t, WA, ea);
        if (use_l3b) lookup_group<kIdentCls, true, kTex>(t, WB, eb);
        else lookup_group<kIdentCls, false, kTex>(t, WB, eb);
        {
            const uint32_t cnt = uint32_t(__popc(__ballot_sync(0xFFFFFFFFu, sample_cont)));
            use_l3a = have_l3 && cnt >= p.l3_min;
            use_l3b = have_l3 && cnt >= p.l3_min_b;
        }
        const uint32_t anya = ea[0] | ea[1] | ea[2] | ea[3] | ea[4] | ea[5] | ea[6] | ea[7];
        const uint32_t anyb = eb[0] | eb[1] | eb[2] | eb[3] | eb[4] | eb[5] | eb[6] | eb[7];
        const bool alive_a = __any_sync(0xFFFFFFFFu, (anya & kAlive) != 0);
        const bool alive_b = __any_sync(0xFFFFFFFFu, (anyb & kAlive) != 0);
        if (alive_a || alive_b) {
            // Some walk of this visit is still alive after level 3 (~45% of the visits on random bytes, one
            // or two positions each).  Level 4 is taken here with one more round of predicated loads (c[i-3]
            // is still in the register window), per group of 256 positions; that ends ~99% of them.
            if (alive_a) level4_group<kIdentCls, false>(t, WA, ea);
            if (alive_b) level4_group<kIdentCls, false>(t, WB, eb);
            const uint32_t any5a = (ea[0] | ea[1] | ea[2] | ea[3] | ea[4] | ea[5] | ea[6] | ea[7]) & kAlive;
            const uint32_t any5b = (eb[0] | eb[1] | eb[2] | eb[3] | eb[4] | eb[5] | eb[6] | eb[7]) & kAlive;
            if ((any5a | any5b) != 0) {
                // Still alive after level 4 (planted / real matches, ~1e-5 of random positions): hand the walk to
                // the deep kernel.  Every CTA owns a strip of the queue and hands out its slots with a shared-memory
                // counter: no global atomics (one hot global counter cost 1.9 ms per GiB) and no warp-wide scans.
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (ea[j] & kAlive) {
                        const uint64_t pos = s0 + ga + j;
                        if (atomicAdd(s_qcnt, 1u) < p.q_per_cta) {
                            // "continue at row" items fill the strip from the front, "tail of pattern" items from the back
                            const bool tail = (ea[j] & kTail) != 0;
                            const uint32_t slot = tail ? p.q_per_cta - 1 - atomicAdd(s_qcnt + 1, 1u) : atomicAdd(s_qcnt + 2, 1u);
                            q_strip[slot] = (pos << 25) | (tail ? (4u << 16) | (ea[j] & 0xFFFFu) : (ea[j] & 0xFFFFFFu));
                            ea[j] = 0;  // placeholder; sfx_deep_kernel writes the result
                        } else {
                            ea[j] = sfx_finish(p, ea[j], 4, p.stream + pos, pos + p.hist_valid + 1) & 0xFFFFu;
                            if constexpr (kFlags) { if (is_long(p, ea[j])) ea[j] |= 255u << 16; }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (eb[j] & kAlive) {
                        const uint64_t pos = s0 + gb + j;
                        if (atomicAdd(s_qcnt, 1u) < p.q_per_cta) {
                            const bool tail = (eb[j] & kTail) != 0;
                            const uint32_t slot = tail ? p.q_per_cta - 1 - atomicAdd(s_qcnt + 1, 1u) : atomicAdd(s_qcnt + 2, 1u);
                            q_strip[slot] = (pos << 25) | (tail ? (4u << 16) | (eb[j] & 0xFFFFu) : (eb[j] & 0xFFFFFFu));
                            eb[j] = 0;
                        } else {
                            eb[j] = sfx_finish(p, eb[j], 4, p.stream + pos, pos + p.hist_valid + 1) & 0xFFFFu;
                            if constexpr (kFlags) { if (is_long(p, eb[j])) eb[j] |= 255u << 16; }
                        }
                    }
                }
            }
        }
        store_group(out_ptr, 0, ea);
        store_group(out_ptr, 256, eb);
        if constexpr (kFlags) {
            const uint32_t thr = p.min_len << 16;   // entry = length << 16 | pid for finals of levels 3/4, 0 for parked walks
            uint32_t fa = 0, fb = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (ea[j] >= thr) fa |= 1u << j;
                if (eb[j] >= thr) fb |= 1u << j;
            }
            uint8_t* fp = p.flags + (s0 >> 3) + lane;
            fp[0] = uint8_t(fa);
            fp[32] = uint8_t(fb);
        }
        out_ptr += step;
    }
    __syncthreads();  // every warp of the CTA has finished its visits
    if (tid == 0) { p.qcount[2 * blockIdx.x] = s_qcnt[2]; p.qcount[2 * blockIdx.x + 1] = s_qcnt[1]; }
}

// raw halves of load8_ending_at, so that issue and use can be separated
__device__ __forceinline__ void load8_issue(const uint8_t* a, const uint8_t* floor, uint64_t& hi, uint64_t& lo) {
    const uintptr_t ua = reinterpret_cast<uintptr_t>(a);
    const uint64_t* hi_p = reinterpret_cast<const uint64_t*>(ua & ~uintptr_t(7));
    hi = __ldg(hi_p);
    lo = 0;
    if ((ua & 7) != 7 && reinterpret_cast<const uint8_t*>(hi_p) > floor) lo = __ldg(hi_p - 1);
}
__device__ __forceinline__ uint64_t load8_merge(const uint8_t* a, uint64_t hi, uint64_t lo) {
    const uint32_t sh = uint32_t(reinterpret_cast<uintptr_t>(a) & 7) * 8;
    return sh == 56 ? hi : ((hi << (56 - sh)) | (lo >> (sh + 8)));
}

// Deferred walks (levels >= 5).  An item is a chain of dependent memory round trips (~900 cycles each: the stream
// bytes come from DRAM, rows / pattern text from L2), and nothing hides a round trip but other, independent, round trips.
// One CTA drains the strip of one scan CTA in two phases:
//   phase A, "continue at row" items (front of the strip; payload = row): 8 history bytes, then one row lookup per
//            byte; a walk that reaches a tail entry is appended to the tail items.  Walk lengths differ wildly, so
//            every lane keeps its own walks going and claims a new item whenever one ends;
//   phase B, "tail of pattern" items (back of the strip; payload = pid | depth << 16): the tail record, then 8-byte
//            compares of the stream against the pattern text, then (rarely) a few steps up the PatternsTree chain.
//            These are alike, so a WARP takes 2 x 32 consecutive items at a time and moves them through the same
//            steps together: all loads of a step are in flight at once, a finished lane is masked until the batch ends.
// The first generations gave every lane its own state machines (two per lane, items claimed one by one): no lane waited
// for another, but the lanes of a warp sat in different states and the hardware ran the states one after the other --
// ncu: 11.4 of 32 lanes active per issued instruction, 40 warp instructions per item.  Batches wait for their slowest
// item, yet issue a tenth of the instructions; items of a strip are neighbours in the stream, so their lengths are alike.
template <bool kIdentCls, bool kFlags>
__global__ void __launch_bounds__(1024) sfx_deep_kernel(const SfxParams p) {
    __shared__ uint32_t s_tails, s_next_a, s_next_b;   // tail items so far; next unclaimed item of each phase
    uint64_t* q_strip = p.queue + size_t(blockIdx.x) * p.q_per_cta;
    const uint32_t n_rows = p.qcount[2 * blockIdx.x];
    const uint32_t cap_tails = p.q_per_cta - n_rows;
    const uint8_t* const floor_s = p.stream - p.hist_valid;  // first readable stream byte
    if (threadIdx.x == 0) { s_tails = p.qcount[2 * blockIdx.x + 1]; s_next_a = 0; s_next_b = 0; }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;
    constexpr uint32_t kFull = 0xFFFFFFFFu;

    // ---- phase A ----
    // Row walks differ too much in length for batches (3.7 lookups on average, 16 for the longest of 32: a batch ran
    // with 7 of 32 lanes active): every lane keeps kA walks going and takes a new item the moment one ends.  One
    // iteration ISSUES the next load of every walk -- the item itself, 8 bytes of history, or a row entry -- and only
    // then CONSUMES them: one wait per iteration for all of them.  Items are claimed with one shared-memory atomic per
    // warp and iteration.
    {
        constexpr int kA = 2;
        enum : uint32_t { kNeedItem, kNeedHist, kLookup, kEnd };
        uint32_t st[kA], v[kA], left[kA], idx[kA], ld_c[kA];
        uint64_t pos[kA], k[kA], hist[kA], ld_a[kA], ld_b[kA];
#pragma unroll
        for (int u = 0; u < kA; ++u) { st[u] = kNeedItem; v[u] = left[u] = idx[u] = ld_c[u] = 0; pos[u] = k[u] = hist[u] = ld_a[u] = ld_b[u] = 0; }
        for (;;) {
            bool alive = false;
#pragma unroll
            for (int u = 0; u < kA; ++u) {
                const uint32_t need = __ballot_sync(kFull, st[u] == kNeedItem);
                if (need) {
                    const uint32_t leader = uint32_t(__ffs(int(need))) - 1u;
                    uint32_t first = 0;
                    if (lane == leader) first = atomicAdd(&s_next_a, uint32_t(__popc(need)));
                    first = __shfl_sync(kFull, first, int(leader));
                    if (st[u] == kNeedItem) {
                        idx[u] = first + uint32_t(__popc(need & ((1u << lane) - 1u)));
                        if (idx[u] >= n_rows) st[u] = kEnd;
                    }
                }
                alive = alive || st[u] != kEnd;
            }
            if (!__any_sync(kFull, alive)) break;
            // ---- issue ----
#pragma unroll
            for (int u = 0; u < kA; ++u) {
                if (st[u] == kNeedItem) {
                    ld_a[u] = q_strip[idx[u]];
                } else if (st[u] == kNeedHist) {
                    load8_issue(p.stream + pos[u] - k[u], floor_s, ld_a[u], ld_b[u]);
                } else if (st[u] == kLookup) {
                    uint32_t c = uint32_t(hist[u] >> 56);
                    hist[u] <<= 8; --left[u];
                    if constexpr (!kIdentCls) c = __ldg(p.cls + c);
                    ld_c[u] = __ldg(p.rows + ((size_t(v[u] & 0xFFFFFFu) << p.log2_ncp) | c));
                }
            }
            // ---- consume ----
#pragma unroll
            for (int u = 0; u < kA; ++u) {
                if (st[u] == kNeedItem) {
                    pos[u] = ld_a[u] >> 25;
                    v[u] = kCont | uint32_t(ld_a[u] & 0xFFFFFFu);
                    k[u] = 4;
                    st[u] = kNeedHist;
                } else if (st[u] == kNeedHist) {
                    hist[u] = load8_merge(p.stream + pos[u] - k[u], ld_a[u], ld_b[u]);   // c[pos-k] in the top byte
                    left[u] = 8;
                    st[u] = kLookup;
                } else if (st[u] == kLookup) {
                    v[u] = ld_c[u];
                    ++k[u];
                    if (v[u] & kTail) {  // hand over to phase B
                        const uint32_t t = atomicAdd(&s_tails, 1u);
                        if (t < cap_tails) q_strip[p.q_per_cta - 1 - t] = (pos[u] << 25) | (uint32_t(k[u]) << 16) | (v[u] & 0xFFFFu);
                        else put_result<kFlags>(p, pos[u], sfx_finish(p, v[u], k[u], p.stream + pos[u], pos[u] + p.hist_valid + 1));  // strip full
                        st[u] = kNeedItem;
                    } else if (!(v[u] & kCont)) {
                        put_result<kFlags>(p, pos[u], v[u]);
                        st[u] = kNeedItem;
                    } else if (left[u] == 0) {
                        st[u] = kNeedHist;
                    }
                }
                // the walk needs a byte that does not exist (start of the stream): the row's own best pattern
                if ((st[u] == kNeedHist || st[u] == kLookup) && k[u] >= pos[u] + p.hist_valid + 1) {
                    put_result<kFlags>(p, pos[u], __ldg(p.row_best + (v[u] & 0xFFFFFFu)));
                    st[u] = kNeedItem;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase B ----
    // kB batches of 32 items per warp at a time: the steps of the batches are independent, so their loads are in flight
    // together (the kernel is bound by the latency of its dependent round trips, not by instruction issue)
    constexpr int kB = 2;
    const uint32_t n_tails = min(s_tails, cap_tails);
    const uint64_t* q_back = q_strip + (p.q_per_cta - 1);  // item j sits at q_back[-j]
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&s_next_b, 32u * kB);
        base = __shfl_sync(kFull, base, 0);
        if (base >= n_tails) break;
        bool active[kB], cmp_done[kB];
        uint64_t pos[kB], avail[kB];
        uint32_t pid[kB], k[kB], lim[kB];
        uint4 rec[kB];
#pragma unroll
        for (int u = 0; u < kB; ++u) {
            const uint32_t idx = base + 32u * u + lane;
            active[u] = idx < n_tails;
            const uint64_t item = active[u] ? *(q_back - idx) : 0ull;
            pos[u] = item >> 25;
            pid[u] = uint32_t(item & 0xFFFFu);
            k[u] = uint32_t(item >> 16) & 0x1FFu;      // bytes matched so far (the last k bytes of the pattern)
        }
#pragma unroll
        for (int u = 0; u < kB; ++u) {
            rec[u] = make_uint4(0, 0, 0, 0);
            if (active[u]) rec[u] = __ldg(p.tail_rec + pid[u]);   // x = text offset, y = length, z = next terminal, w = best at the start
        }
#pragma unroll
        for (int u = 0; u < kB; ++u) {
            avail[u] = pos[u] + p.hist_valid + 1;
            lim[u] = uint64_t(rec[u].y) < avail[u] ? rec[u].y : uint32_t(avail[u]);
            cmp_done[u] = !active[u] || k[u] >= lim[u];
        }
        for (;;) {
            bool any = false;
#pragma unroll
            for (int u = 0; u < kB; ++u) any = any || !cmp_done[u];
            if (!__any_sync(kFull, any)) break;
            uint64_t a[kB], b[kB];
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                a[u] = b[u] = 0;
                if (!cmp_done[u]) {
                    a[u] = load8_ending_at(p.stream + pos[u] - k[u], floor_s);
                    b[u] = load8_ending_at(p.pat_bytes + rec[u].x + (rec[u].y - 1 - k[u]), p.pat_bytes);
                }
            }
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                if (!cmp_done[u]) {
                    const uint64_t x = a[u] ^ b[u];
                    const uint32_t same = x ? uint32_t(__clzll((long long)x) >> 3) : 8u;   // equal bytes from the top (= backwards)
                    const uint32_t left = lim[u] - k[u];
                    k[u] += same < left ? same : left;
                    if (same < 8 || k[u] >= lim[u]) cmp_done[u] = true;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kB; ++u) {
            if (active[u]) {
                uint32_t res, q = pid[u];
                if (k[u] < rec[u].z) res = rec[u].w;                  // no further terminal reached: best at the tail start
                else if (k[u] >= rec[u].y) res = q;                   // the whole pattern
                else {                                                // the longest pattern of the chain with length <= k
                    while (q && __ldg(p.pat_len + q - 1) > k[u]) q = __ldg(p.parent + q);
                    res = q;
                }
                put_result<kFlags>(p, pos[u], res);
            }
        }
    }
}

// The two ragged ends of a stream, by a bounded walk straight from global memory (avail(i) = i + hist_valid + 1
// bytes exist up to c[i]): the first `head` positions, whose look-back may cross the start of the stream, and the
// last `tail` positions, which do not fill a whole warp tile.
__global__ void sfx_edge_kernel(const SfxParams p, uint32_t head, uint32_t tail) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= head + tail) return;
    const uint64_t i = t < head ? uint64_t(t) : p.n - tail + (t - head);
    const uint64_t avail = i + p.hist_valid + 1;
    const uint32_t pid = sfx_finish(p, p.root1[p.stream[i]], 1, p.stream + i, avail) & 0xFFFFu;
    p.out[i] = uint16_t(pid);
    if (p.flags != nullptr) flag_set(p, i, is_long(p, pid));   // these positions are redone (head) or not covered (tail) by the scan kernel
}

}  // namespace

size_t sfx_smem_bytes() { return kSmemBytes; }

size_t sfx_scan_ctas(uint64_t n, int n_sms) {
    const uint64_t tiles = n / kTile;  // full tiles
    const uint64_t ctas = (tiles + kWarps - 1) / kWarps;
    return size_t(ctas < uint64_t(n_sms) ? (ctas ? ctas : 1) : uint64_t(n_sms));
}

cudaError_t sfx_scan_launch(const SfxParams& p_in, bool ident_cls, int n_sms, uint32_t max_pat_len, cudaStream_t st,
                            uint64_t* launches, cudaEvent_t* ev) {
    SfxParams p = p_in;
    if (p.n == 0) return cudaSuccess;
    p.n_tiles = p.n / kTile;
    if (p.n_l3 > kSfxMaxL3) p.l3f = nullptr;  // the filter does not fit beside root2: plain L2 lookups only
    const bool tex = p.rows_tex != 0;
    if (p.flags != nullptr && p.min_len < 3) return cudaErrorInvalidValue;   // sparse mode: patterns of >= 3 bytes only
    auto pick = [&](auto flags_tag) {
        constexpr bool F = decltype(flags_tag)::value;
        return ident_cls ? (tex ? sfx_scan_kernel<true, true, F> : sfx_scan_kernel<true, false, F>)
                         : (tex ? sfx_scan_kernel<false, true, F> : sfx_scan_kernel<false, false, F>);
    };
    auto kern = p.flags != nullptr ? pick(std::true_type{}) : pick(std::false_type{});
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    const uint32_t grid = uint32_t(sfx_scan_ctas(p.n, n_sms));
    if (ev) cudaEventRecord(ev[0], st);
    if (p.flags != nullptr) {   // the flag words the scan kernel does not write: those of the ragged end (or all, without a full visit)
        const uint64_t first = (p.n / kTile) * (kTile / 32), last = (p.n + 31) / 32;
        if (last > first) {
            e = cudaMemsetAsync(reinterpret_cast<uint32_t*>(p.flags) + first, 0, (last - first) * 4, st);
            if (e != cudaSuccess) return e;
        }
    }
    if (p.n_tiles > 0) {
        // shared memory actually needed (the rest of the 256 KB stays L1 / texture cache)
        const size_t smem = size_t(kOffL3) + (p.l3f ? size_t(p.n_l3) * 4 : 0) + 16;
        kern<<<grid, kThreads, smem, st>>>(p);
        if (ev) cudaEventRecord(ev[1], st);
        ++*launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        auto deep = p.flags != nullptr ? (ident_cls ? sfx_deep_kernel<true, true> : sfx_deep_kernel<false, true>)
                                       : (ident_cls ? sfx_deep_kernel<true, false> : sfx_deep_kernel<false, false>);
        deep<<<grid, 1024, 0, st>>>(p);  // CTA b drains the strip of scan CTA b
        ++*launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    } else {
        // fewer than one visit (the plugin's read_char / tiny read_block calls): the edge kernel alone
        e = cudaMemsetAsync(p.qcount, 0, 2 * sizeof(uint32_t), st);
        if (e != cudaSuccess) return e;
        if (ev) cudaEventRecord(ev[1], st);
    }
    // Start of a stream: visit 0 reads no bytes before the stream unless 4 bytes of history exist, and a deferred
    // walk is bounded by the bytes that exist; every position that could look back past the start of the readable
    // stream is redone by the bounded walker -- together with the ragged end that does not fill a tile.
    // (The second condition: with fewer than 4 bytes of history visit 0 does not load them at all -- hist4 in the scan
    // kernel -- so the first max_pat_len-1 positions are redone even when the history covers every pattern.)
    uint32_t head = 0;
    if (max_pat_len > 1 && (p.hist_valid < uint64_t(max_pat_len - 1) || p.hist_valid < 4)) {
        const uint64_t want = uint64_t(max_pat_len - 1);
        head = uint32_t(p.n < want ? p.n : want);
    }
    uint32_t tail = uint32_t(p.n - p.n_tiles * uint64_t(kTile));
    if (uint64_t(head) + tail > p.n) tail = uint32_t(p.n - head);
    if (head + tail) {
        sfx_edge_kernel<<<(head + tail + 127) / 128, 128, 0, st>>>(p, head, tail);
        ++*launches;
        e = cudaGetLastError();
    }
    if (ev) cudaEventRecord(ev[2], st);
    return e;
}

// Small calls (the plugin's read_char and the reference's 100 KiB chunks, measure.c:77): ONE launch of the bounded
// walker over every position instead of scan + deep + edge -- at these sizes the call is launch-latency bound.
cudaError_t sfx_walk_launch(const SfxParams& p, cudaStream_t st, uint64_t* launches) {
    if (p.n == 0) return cudaSuccess;
    if (p.n > (uint64_t(1) << 30)) return cudaErrorInvalidValue;
    sfx_edge_kernel<<<uint32_t((p.n + 127) / 128), 128, 0, st>>>(p, uint32_t(p.n), 0);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pm
