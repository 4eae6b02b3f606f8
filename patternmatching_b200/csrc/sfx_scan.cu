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
// All positions are independent, so there is no per-thread warm-up.  Every WARP runs its own
// software pipeline: lane 0 stages 1 KiB tiles plus a 352-byte left halo (>= max_pat_len-1,
// SURVEY Q8) with bulk async copies (TMA engine, SASS UBLKCP) into a private double buffer behind
// private mbarriers; there is no CTA-wide barrier after start-up, so warps hide each other's L2
// latency.  A lane resolves 2 x 8 consecutive positions per 512-byte visit, arranged so that the
// 16-byte result stores of a warp are fully coalesced.
#include "pm_dev.cuh"
#include "sfx_scan.cuh"

namespace pm {

namespace {

constexpr int kThreads = kSfxThreads;
constexpr int kWarps = kThreads / 32;
constexpr int kTile = kSfxTile;                  // bytes per warp tile
constexpr int kVisits = kTile / 512;             // 512 positions per warp visit
constexpr int kStages = kSfxStages;
constexpr int kMainHalo = 16;                    // staged left halo: levels 1-4 look back 3 bytes (bulk-copy granularity 16)
constexpr int kStageBuf = kMainHalo + kTile;     // staged bytes per stage
constexpr uint32_t kCont = 0x80000000u;   // entry: continue at row (entry & 0xFFFFFF)
constexpr uint32_t kTail = 0x40000000u;   // entry: the rest of the path is the text of pattern (entry & 0xFFFF)
constexpr uint32_t kAlive = kCont | kTail;

// shared memory carve-up (bytes)
constexpr int kOffRoot2 = 0;                     // 131072
constexpr int kOffCls = 131072;                  // 256
constexpr int kOffBar = kOffCls + 256;           // kWarps * kStages * 8
constexpr int kOffQCnt = kOffBar + kWarps * kStages * 8;           // 16 (one counter, padded)
constexpr int kOffL3 = kOffQCnt + 16;                             // kSfxMaxL3 * 4
constexpr int kOffStages = kOffL3 + int(kSfxMaxL3) * 4;
constexpr int kSmemBytes = kOffStages + kWarps * kStages * kStageBuf;
static_assert(kOffStages % 16 == 0 && kStageBuf % 16 == 0, "bulk copies need 16-byte alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(kCont == 0x80000000u, "level4_group tests the sign bit");
static_assert((kStages & (kStages - 1)) == 0, "kStages must be a power of two");

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

// Levels 1-3 for one group of 8 consecutive positions whose bytes sit in the register window W.
// Phase A: 8 shared-memory gathers from root2.  Phase B: an entry that continues below the 2-byte table
// (~10% on random bytes, ~60% on text) first consults the level-3 filter word of its row in SHARED memory:
// unless the Bloom bit of c[i-2] is set the answer is the row's own best pid; only Bloom hits (true children
// 0.9%, false positives ~14% of the continuing entries) fetch their row entry from L2, with PREDICATED loads --
// no branch and no use of the loaded value here, so all loads of a visit are in flight together.
template <bool kIdentCls, bool kL3>
__device__ __forceinline__ bool lookup_group(const uint16_t* s_root2, const uint32_t* s_l3, const uint32_t (&W)[3],
                                             uintptr_t rows_adj, uint32_t cont_base, uint32_t log2_ncp,
                                             const uint8_t* s_cls, uint32_t (&e)[8]) {
    e[0] = s_root2[win_u16<3>(W)]; e[1] = s_root2[win_u16<4>(W)];
    e[2] = s_root2[win_u16<5>(W)]; e[3] = s_root2[win_u16<6>(W)];
    e[4] = s_root2[win_u16<7>(W)]; e[5] = s_root2[win_u16<8>(W)];
    e[6] = s_root2[win_u16<9>(W)]; e[7] = s_root2[win_u16<10>(W)];
    const bool first_cont = e[0] >= cont_base;  // sample for the adaptive choice of the level-3 path
    uint32_t c2[8];
    c2[0] = win_u8<2>(W); c2[1] = win_u8<3>(W); c2[2] = win_u8<4>(W); c2[3] = win_u8<5>(W);
    c2[4] = win_u8<6>(W); c2[5] = win_u8<7>(W); c2[6] = win_u8<8>(W); c2[7] = win_u8<9>(W);
    if constexpr (kL3) {
        uint32_t f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            f[j] = 0;
            if (e[j] >= cont_base) f[j] = s_l3[e[j] - cont_base];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t c = c2[j];
            const bool cont = e[j] >= cont_base;
            const bool hit = ((f[j] >> (c & 15u)) & 1u) != 0;   // f == 0 when the entry does not continue
            if constexpr (!kIdentCls) c = s_cls[c];
            const uint32_t* addr = reinterpret_cast<const uint32_t*>(rows_adj + ((uintptr_t((e[j] << log2_ncp) | c)) << 2));
            if (cont) e[j] = f[j] >> 16;
            if (hit) e[j] = __ldg(addr);
        }
    } else if constexpr (kIdentCls) {
        // 256 byte classes: the row index (e << 8 | c[i-2]) is ONE byte permute of the entry and the window word
        uint32_t idx[8];
        idx[0] = win_row_index<2>(W, e[0]); idx[1] = win_row_index<3>(W, e[1]);
        idx[2] = win_row_index<4>(W, e[2]); idx[3] = win_row_index<5>(W, e[3]);
        idx[4] = win_row_index<6>(W, e[4]); idx[5] = win_row_index<7>(W, e[5]);
        idx[6] = win_row_index<8>(W, e[6]); idx[7] = win_row_index<9>(W, e[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // rows_adj = &rows[(row2_base - cont_base) << 8]: entry e addresses row (e - cont_base + row2_base)
            const uint32_t* addr = reinterpret_cast<const uint32_t*>(rows_adj + (uintptr_t(idx[j]) << 2));
            if (e[j] >= cont_base) e[j] = __ldg(addr);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t c = s_cls[c2[j]];
            const uint32_t* addr = reinterpret_cast<const uint32_t*>(rows_adj + ((uintptr_t((e[j] << log2_ncp) | c)) << 2));
            if (e[j] >= cont_base) e[j] = __ldg(addr);
        }
    }
    return first_cont;
}

// Level 4 for the entries of a group that are still "continue" after level 3: c[i-3] is byte 1+j of W.
template <bool kIdentCls>
__device__ __forceinline__ void level4_group(const uint32_t (&W)[3], const uint32_t* __restrict__ rows, uint32_t log2_ncp,
                                             const uint8_t* s_cls, uint32_t (&e)[8]) {
    if constexpr (kIdentCls) {
        // (row << 8) | c[i-3] is one byte permute (it drops the flag byte of the entry); kCont is the sign bit
        uint32_t idx[8];
        idx[0] = win_row_index<1>(W, e[0]); idx[1] = win_row_index<2>(W, e[1]);
        idx[2] = win_row_index<3>(W, e[2]); idx[3] = win_row_index<4>(W, e[3]);
        idx[4] = win_row_index<5>(W, e[4]); idx[5] = win_row_index<6>(W, e[5]);
        idx[6] = win_row_index<7>(W, e[6]); idx[7] = win_row_index<8>(W, e[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (int32_t(e[j]) < 0) e[j] = __ldg(rows + idx[j]);
    } else {
        uint32_t c3[8];
        c3[0] = win_u8<1>(W); c3[1] = win_u8<2>(W); c3[2] = win_u8<3>(W); c3[3] = win_u8<4>(W);
        c3[4] = win_u8<5>(W); c3[5] = win_u8<6>(W); c3[6] = win_u8<7>(W); c3[7] = win_u8<8>(W);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t c = s_cls[c3[j]];
            if (int32_t(e[j]) < 0) e[j] = __ldg(rows + ((size_t(e[j] & 0xFFFFFFu) << log2_ncp) | c));
        }
    }
}

__device__ __forceinline__ void store_group(uint16_t* out, uint64_t gpos, const uint32_t (&e)[8]) {
    uint4 r;
    r.x = __byte_perm(e[0], e[1], 0x5410); r.y = __byte_perm(e[2], e[3], 0x5410);  // low halves, one PRMT each
    r.z = __byte_perm(e[4], e[5], 0x5410); r.w = __byte_perm(e[6], e[7], 0x5410);
    __stcs(reinterpret_cast<uint4*>(out + gpos), r);  // write-once result: streaming store
}

template <bool kIdentCls>
__global__ void __launch_bounds__(kThreads, 1) sfx_scan_kernel(const SfxParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint16_t* s_root2 = reinterpret_cast<uint16_t*>(smem + kOffRoot2);
    uint8_t* s_cls = smem + kOffCls;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar) + warp * kStages;
    uint8_t* wbuf = smem + kOffStages + warp * (kStages * kStageBuf);
    const bool have_halo = p.hist_valid >= uint64_t(kMainHalo);
    const uint32_t cont_base = p.cont_base, log2_ncp = p.log2_ncp;
    const uintptr_t rows_adj = reinterpret_cast<uintptr_t>(p.rows) +
                               ((uintptr_t(p.row2_base) << log2_ncp) << 2) - ((uintptr_t(cont_base) << log2_ncp) << 2);

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    // halo of stage 0 zeroed: the very first tile of a stream without history reads it (the fix-up pass
    // owns the positions that would need real history)
    if (lane < kMainHalo / 4) reinterpret_cast<uint32_t*>(wbuf)[lane] = 0;
    if (tid < 256) s_cls[tid] = p.cls[tid];
    uint32_t* s_qcnt = reinterpret_cast<uint32_t*>(smem + kOffQCnt);
    uint32_t* s_l3 = reinterpret_cast<uint32_t*>(smem + kOffL3);
    if (tid == 0) { s_qcnt[0] = 0; s_qcnt[1] = 0; s_qcnt[2] = 0; }  // slots used / tail items / row items
    const bool have_l3 = p.l3f != nullptr;
    if (have_l3) for (uint32_t i = tid; i < p.n_l3; i += kThreads) s_l3[i] = __ldg(p.l3f + i);
    // Per-warp choice of the level-3 path, re-made every visit from a 32-position sample of the previous one:
    // on binary / random bytes ~10% of the positions continue below root2 and the plain predicated L2 lookups
    // are cheapest; on text ~60% continue, and the shared-memory filter (2.1x faster there) takes over.
    bool use_l3a = false, use_l3b = false;
    fence_proxy_async();
    __syncwarp();

    const uint64_t gw = uint64_t(blockIdx.x) * kWarps + warp;   // global warp id
    const uint64_t stride = uint64_t(gridDim.x) * kWarps * kTile;   // stream bytes between two tiles of a warp

    // lane 0 only: stage the (full) tile that starts at stream offset s0, with its left halo unless it is the
    // very first tile of a stream that comes without history
    const uint64_t n_main = p.n_tiles * uint64_t(kTile);   // the ragged end (< kTile bytes) belongs to sfx_edge_kernel
    auto issue_tile = [&](uint64_t s0, int s) {
        uint8_t* dst = wbuf + s * kStageBuf;
        if (s0 != 0 || have_halo) {
            mbar_arrive_expect_tx(&bars[s], kTile + kMainHalo);
            bulk_g2s(dst, p.stream + s0 - kMainHalo, kTile + kMainHalo, &bars[s]);
        } else {
            mbar_arrive_expect_tx(&bars[s], kTile);
            bulk_g2s(dst + kMainHalo, p.stream, kTile, &bars[s]);
        }
    };

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            const uint64_t s0 = gw * uint64_t(kTile) + uint64_t(s) * stride;
            if (s0 < n_main) issue_tile(s0, s);
        }
    }

    // root2 -> shared memory (once per CTA; coalesced 16-byte loads, L2 hits after the first CTA)
    {
        const int4* src = reinterpret_cast<const int4*>(p.root2);
        int4* dst = reinterpret_cast<int4*>(s_root2);
        for (int i = tid; i < 131072 / 16; i += kThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();  // the only CTA-wide barrier

    uint64_t* q_strip = p.queue + size_t(blockIdx.x) * p.q_per_cta;  // this CTA's strip of the deferred-walk queue
    uint32_t it = 0;
    for (uint64_t s0 = gw * uint64_t(kTile); s0 < n_main; s0 += stride, ++it) {
        const int s = it & (kStages - 1);
        uint8_t* stage = wbuf + s * kStageBuf;
        mbar_wait(&bars[s], (it / kStages) & 1);

#pragma unroll 1
        for (int v = 0; v < kVisits; ++v) {
            const int base_off = v * 512;
            const uint8_t* vb = stage + kMainHalo + base_off;
            // group A: positions base_off + 8*lane .. +8 ; group B: base_off + 256 + 8*lane .. +8
            const uint2 a = *reinterpret_cast<const uint2*>(vb + 8 * lane);
            const uint2 b = *reinterpret_cast<const uint2*>(vb + 256 + 8 * lane);
            uint32_t WA[3], WB[3];
            WA[1] = a.x; WA[2] = a.y; WB[1] = b.x; WB[2] = b.y;
            WA[0] = __shfl_up_sync(0xFFFFFFFFu, a.y, 1);
            WB[0] = __shfl_up_sync(0xFFFFFFFFu, b.y, 1);
            const uint32_t a31 = __shfl_sync(0xFFFFFFFFu, a.y, 31);
            if (lane == 0) {
                WA[0] = *reinterpret_cast<const uint32_t*>(vb - 4);
                WB[0] = a31;
            }
            const int ga = base_off + 8 * lane, gb = ga + 256;
            uint32_t ea[8], eb[8];
            bool sample_cont;
            if (use_l3a) sample_cont = lookup_group<kIdentCls, true>(s_root2, s_l3, WA, rows_adj, cont_base, log2_ncp, s_cls, ea);
            else sample_cont = lookup_group<kIdentCls, false>(s_root2, s_l3, WA, rows_adj, cont_base, log2_ncp, s_cls, ea);
            if (use_l3b) lookup_group<kIdentCls, true>(s_root2, s_l3, WB, rows_adj, cont_base, log2_ncp, s_cls, eb);
            else lookup_group<kIdentCls, false>(s_root2, s_l3, WB, rows_adj, cont_base, log2_ncp, s_cls, eb);
            {
                const uint32_t cnt = uint32_t(__popc(__ballot_sync(0xFFFFFFFFu, sample_cont)));
                use_l3a = have_l3 && cnt >= p.l3_min;
                use_l3b = have_l3 && cnt >= p.l3_min_b;
            }
            const uint32_t anya = ea[0] | ea[1] | ea[2] | ea[3] | ea[4] | ea[5] | ea[6] | ea[7];
            const uint32_t anyb = eb[0] | eb[1] | eb[2] | eb[3] | eb[4] | eb[5] | eb[6] | eb[7];
            if (__any_sync(0xFFFFFFFFu, ((anya | anyb) & kAlive) != 0)) {
                // Some walk of this visit is still alive after level 3 (~40% of the visits on random bytes, one
                // or two positions each).  Level 4 is taken here with one more round of predicated loads (c[i-3]
                // is still in the register window); that ends ~99% of them.
                level4_group<kIdentCls>(WA, p.rows, log2_ncp, s_cls, ea);
                level4_group<kIdentCls>(WB, p.rows, log2_ncp, s_cls, eb);
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
                                ea[j] = sfx_finish(p, ea[j], 4, p.stream + pos, pos + p.hist_valid + 1);
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
                                eb[j] = sfx_finish(p, eb[j], 4, p.stream + pos, pos + p.hist_valid + 1);
                            }
                        }
                    }
                }
            }
            store_group(p.out, s0 + ga, ea);
            store_group(p.out, s0 + gb, eb);
        }
        __syncwarp();  // every lane is done reading the stage before it is refilled
        const uint64_t sn = s0 + uint64_t(kStages) * stride;
        if (lane == 0 && sn < n_main) issue_tile(sn, s);
    }
    __syncthreads();  // every warp of the CTA has finished its tiles
    if (tid == 0) { p.qcount[2 * blockIdx.x] = s_qcnt[2]; p.qcount[2 * blockIdx.x + 1] = s_qcnt[1]; }
}

// Deferred walks (levels >= 5).  The items are independent but their dependent chains differ wildly in length
// (one lookup ... hundreds for a long repetitive pattern), so every lane pulls its next item the moment its current
// one ends.  One CTA drains the strip of one scan CTA in two phases, so that the lanes of a warp do the SAME kind of
// step and differ only in how many of them their item needs:
//   phase A, "continue at row" items (front of the strip; payload = row): row lookups, the stream bytes coming
//            from an 8-byte history register; a walk that reaches a tail entry is appended to the tail items;
//   phase B, "tail of pattern" items (back of the strip; payload = pid | depth << 16): 8-byte compares of the
//            stream against the pattern text, then (rarely) a few steps up the PatternsTree chain.
// A step is one dependent memory round trip and nothing hides it but the other warps, so each lane keeps a
// two-deep software pipeline: the item after next is in flight, and for the next item so are the loads that
// depend only on the item (its first 8 stream bytes, its tail record).
template <bool kIdentCls>
__global__ void __launch_bounds__(1024) sfx_deep_kernel(const SfxParams p) {
    __shared__ uint32_t s_tails;
    uint64_t* q_strip = p.queue + size_t(blockIdx.x) * p.q_per_cta;
    const uint32_t n_rows = p.qcount[2 * blockIdx.x];
    const uint32_t cap_tails = p.q_per_cta - n_rows;
    const uint8_t* const floor_s = p.stream - p.hist_valid;  // first readable stream byte
    if (threadIdx.x == 0) s_tails = p.qcount[2 * blockIdx.x + 1];
    __syncthreads();

    // ---- phase A ----
    {
        uint32_t q = threadIdx.x;
        bool v1 = q < n_rows, v2 = q + 1024 < n_rows;
        uint64_t it1 = v1 ? q_strip[q] : 0, it2 = v2 ? q_strip[q + 1024] : 0, h1 = 0;
        q += 2048;
        if (v1 && (it1 >> 25) + p.hist_valid >= 4) h1 = load8_ending_at(p.stream + (it1 >> 25) - 4, floor_s);
        bool busy = false;
        uint32_t v = 0, hist_left = 0;
        uint64_t pos = 0, k = 0, avail = 0, hist = 0;
        const uint8_t* ci = nullptr;
        for (;;) {
            if (!busy) {
                if (!v1) break;
                pos = it1 >> 25;
                v = kCont | uint32_t(it1 & 0xFFFFFFu);
                k = 4;
                ci = p.stream + pos;
                avail = pos + p.hist_valid + 1;
                hist = h1; hist_left = 8;
                busy = true;
                it1 = it2; v1 = v2;
                if (v1 && (it1 >> 25) + p.hist_valid >= 4) h1 = load8_ending_at(p.stream + (it1 >> 25) - 4, floor_s);
                v2 = q < n_rows;
                if (v2) it2 = q_strip[q];
                q += 1024;
            }
            const uint32_t row = v & 0xFFFFFFu;
            if (k >= avail) {  // start of the stream: no byte left
                p.out[pos] = uint16_t(__ldg(p.row_best + row));
                busy = false;
                continue;
            }
            if (hist_left == 0) { hist = load8_ending_at(ci - k, floor_s); hist_left = 8; }
            uint32_t c = uint32_t(hist >> 56);
            hist <<= 8; --hist_left;
            if constexpr (!kIdentCls) c = __ldg(p.cls + c);
            v = __ldg(p.rows + ((size_t(row) << p.log2_ncp) | c));
            ++k;
            if (v & kTail) {  // hand over to phase B
                const uint32_t t = atomicAdd(&s_tails, 1u);
                if (t < cap_tails) q_strip[p.q_per_cta - 1 - t] = (pos << 25) | (uint32_t(k) << 16) | (v & 0xFFFFu);
                else p.out[pos] = uint16_t(sfx_finish(p, v, k, ci, avail));  // strip full: finish here
                busy = false;
            } else if (!(v & kCont)) {
                p.out[pos] = uint16_t(v);
                busy = false;
            }
        }
    }
    __syncthreads();

    // ---- phase B ----
    {
        const uint32_t n_tails = min(s_tails, cap_tails);
        const uint64_t* q_back = q_strip + (p.q_per_cta - 1);  // item j sits at q_back[-j]
        uint32_t q = threadIdx.x;
        bool v1 = q < n_tails, v2 = q + 1024 < n_tails;
        uint64_t it1 = v1 ? *(q_back - q) : 0, it2 = v2 ? *(q_back - (q + 1024)) : 0, a1 = 0;
        q += 2048;
        uint4 rec1 = make_uint4(0, 0, 0, 0);
        if (v1) {
            rec1 = __ldg(p.tail_rec + uint32_t(it1 & 0xFFFFu));
            const uint64_t k1 = (it1 >> 16) & 0x1FFu;
            if ((it1 >> 25) + p.hist_valid >= k1) a1 = load8_ending_at(p.stream + (it1 >> 25) - k1, floor_s);
        }
        bool busy = false, have_a = false;
        uint32_t pid = 0, len = 0, next_term = 0, best_start = 0;
        uint64_t pos = 0, k = 0, lim = 0, a = 0;
        const uint8_t* ci = nullptr;
        const uint8_t* text = nullptr;
        for (;;) {
            if (!busy) {
                if (!v1) break;
                pos = it1 >> 25;
                pid = uint32_t(it1 & 0xFFFFu);
                k = (it1 >> 16) & 0x1FFu;                    // bytes matched so far (the last k bytes of the pattern)
                text = p.pat_bytes + rec1.x; len = rec1.y; next_term = rec1.z; best_start = rec1.w;
                ci = p.stream + pos;
                const uint64_t avail = pos + p.hist_valid + 1;
                lim = uint64_t(len) < avail ? uint64_t(len) : avail;
                a = a1; have_a = true;
                busy = true;
                it1 = it2; v1 = v2;
                if (v1) {
                    rec1 = __ldg(p.tail_rec + uint32_t(it1 & 0xFFFFu));
                    const uint64_t k1 = (it1 >> 16) & 0x1FFu;
                    if ((it1 >> 25) + p.hist_valid >= k1) a1 = load8_ending_at(p.stream + (it1 >> 25) - k1, floor_s);
                }
                v2 = q < n_tails;
                if (v2) it2 = *(q_back - q);
                q += 1024;
            }
            bool done = k >= lim;
            if (!done) {
                if (!have_a) a = load8_ending_at(ci - k, floor_s);
                have_a = false;
                const uint64_t b = load8_ending_at(text + (len - 1 - k), p.pat_bytes);
                const uint64_t x = a ^ b;
                const uint64_t same = x ? uint64_t(__clzll((long long)x) >> 3) : 8;   // equal bytes from the top (= backwards)
                const uint64_t left = lim - k;
                k += same < left ? same : left;
                done = same < 8 || k >= lim;
            }
            if (done) {
                uint32_t cand = best_start;
                if (k >= next_term) {  // some pattern of the chain fits: the longest one with length <= k
                    cand = pid;
                    while (cand && uint64_t(__ldg(p.pat_len + cand - 1)) > k) cand = __ldg(p.parent + cand);
                }
                p.out[pos] = uint16_t(cand);
                busy = false;
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
    p.out[i] = uint16_t(sfx_finish(p, p.root1[p.stream[i]], 1, p.stream + i, avail));
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
    auto kern = ident_cls ? sfx_scan_kernel<true> : sfx_scan_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    const uint32_t grid = uint32_t(sfx_scan_ctas(p.n, n_sms));
    if (ev) cudaEventRecord(ev[0], st);
    kern<<<grid, kThreads, kSmemBytes, st>>>(p);
    if (ev) cudaEventRecord(ev[1], st);
    ++*launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    {
        auto deep = ident_cls ? sfx_deep_kernel<true> : sfx_deep_kernel<false>;
        deep<<<grid, 1024, 0, st>>>(p);  // CTA b drains the strip of scan CTA b
        ++*launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    // Start of a stream: tile 0 is staged without its halo unless kMainHalo bytes of history exist, and a deferred
    // walk is bounded by the bytes that exist; every position that could look back past the start of the readable
    // stream is redone by the bounded walker -- together with the ragged end that does not fill a tile.
    uint32_t head = 0;
    if (max_pat_len > 1 && p.hist_valid < uint64_t(kHalo)) {
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

}  // namespace pm
