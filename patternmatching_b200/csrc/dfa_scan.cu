// dfa_scan.cu -- exact dictionary scan by per-thread forward walks of the flat Aho-Corasick DFA.
//
// Semantics identical to sfx_scan.cu: out[i] = pid of the longest pattern that is a suffix of
// stream[..i] (ac_read_char, Core/src/mpac.c:304-319).  The automaton is the reference's goto +
// failure + nearest-output-link machine (Core/src/mpac.c:147-210) completed on the host to a full
// DFA, so one table lookup per byte replaces the failure-link loop; worst-case work per byte is
// constant whatever the input (the backward scan's is O(match depth)).
//
// The stream is cut into per-thread segments; a thread first walks the max_pat_len-1 bytes before its
// segment from the root without reporting (after that many bytes the state reports exactly what a
// continuous scan reports -- SURVEY Q8) and then reports its own bytes.  Two variants:
//   hot  (default): states are numbered breadth-first, so the hot ones are the first ones: the transition
//         rows of the first `hot_rows` states (u16 entries) and the longest-pattern ids of the first
//         `hot_long` states live in SHARED memory (snort+et: the root and all 256 depth-1 states = 88% of the
//         steps on random bytes; a small-alphabet dictionary: the whole automaton).  The states of the NEXT
//         level keep a {Bloom of child classes, failure state} word there: without a goto child on c their
//         transition is the failure state's hot row, so 99% of the steps on random bytes stay in shared memory;
//         the rest is read from the flat u32 table in global memory.
//   flat (PM_DFA_FLAT=1): every lookup from global memory through L1/L2 at full occupancy -- better when
//         most steps are deep (pattern-prefix soup), where occupancy and L1 matter more than hot rows.
#include "dfa_scan.cuh"
#include "pm_dev.cuh"

namespace pm {
namespace {

constexpr int kThreads = 128;      // flat variant
constexpr int kHotThreads = 1024;  // hot variant: one CTA per SM, ~200 KB of hot tables

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): a lane's 32 stream bytes or its 16 results (32 bytes)
// move as ONE sector instead of two half-filled ones -- the per-lane segment I/O is a third of this kernel's
// LSU traffic.  Addresses must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const uint8_t* ptr, uint32_t (&w)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr));
}
__device__ __forceinline__ void stg256(uint16_t* ptr, const uint32_t (&r)[8]) {
    asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(ptr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

__global__ void __launch_bounds__(kThreads) dfa_flat_kernel(const DfaParams p) {
    const uint64_t seg = uint64_t(blockIdx.x) * kThreads + threadIdx.x;
    const uint64_t s0 = seg * uint64_t(p.seg);
    if (s0 >= p.n) return;
    const uint64_t s1 = min(p.n, s0 + uint64_t(p.seg));
    // warm-up start: max_pat_len-1 bytes back, never before the readable history
    int64_t w = int64_t(s0) - int64_t(p.warm);
    if (w < -int64_t(p.hist_valid)) w = -int64_t(p.hist_valid);
    const uint8_t* __restrict__ cls = p.cls;
    const uint32_t* __restrict__ delta = p.delta;
    const uint16_t* __restrict__ longest = p.longest;
    const uint32_t l2 = p.log2_ncp;
    uint32_t s = 0;
    // warm-up: walk, do not report (byte steps up to a 16-byte boundary, then vector loads)
    int64_t q0 = w;
    for (; q0 < int64_t(s0) && (q0 & 15); ++q0) s = __ldg(delta + ((size_t(s) << l2) | __ldg(cls + *(p.stream + q0))));
    for (; q0 < int64_t(s0); q0 += 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.stream + q0));
        const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t c = (ws[k >> 2] >> (8 * (k & 3))) & 0xFF;
            s = __ldg(delta + ((size_t(s) << l2) | __ldg(cls + c)));
        }
    }
    // own segment: 16 bytes in, 16 results (32 bytes) out per step
    uint64_t q = s0;
    if (p.wide) {
        for (; q + 32 <= s1; q += 32) {
            uint32_t ws[8];
            ldg256(p.stream + q, ws);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[8];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t c = (ws[half * 4 + (k >> 2)] >> (8 * (k & 3))) & 0xFF;
                    s = __ldg(delta + ((size_t(s) << l2) | __ldg(cls + c)));
                    const uint32_t o = __ldg(longest + s);
                    if (k & 1) r[k >> 1] |= o << 16; else r[k >> 1] = o;
                }
                stg256(p.out + q + half * 16, r);
            }
        }
    }
    for (; q + 16 <= s1; q += 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.stream + q));
        const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
        uint32_t r[8];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t c = (ws[k >> 2] >> (8 * (k & 3))) & 0xFF;
            s = __ldg(delta + ((size_t(s) << l2) | __ldg(cls + c)));
            const uint32_t o = __ldg(longest + s);
            if (k & 1) r[k >> 1] |= o << 16; else r[k >> 1] = o;
        }
        uint4* dst = reinterpret_cast<uint4*>(p.out + q);
        dst[0] = make_uint4(r[0], r[1], r[2], r[3]);
        dst[1] = make_uint4(r[4], r[5], r[6], r[7]);
    }
    for (; q < s1; ++q) {  // ragged end
        s = __ldg(delta + ((size_t(s) << l2) | __ldg(cls + p.stream[q])));
        p.out[q] = __ldg(longest + s);
    }
}


// One transition.  Hot states: their u16 row in shared memory.  States of the first level below the hot rows
// (snort+et: the depth-2 states, 11% of the steps on random bytes): their Bloom word says whether a goto child on
// c can exist; if not (93% of those steps) the transition is the failure state's -- a hot row again, so the step
// stays in shared memory; only Bloom hits and deeper states read the dense table in global memory.
template <bool kIdentCls>
__device__ __forceinline__ uint32_t dfa_step(uint32_t s, uint32_t c, const DfaParams& p, const uint16_t* s_hot,
                                             const uint32_t* s_fb, const uint8_t* s_cls) {
    if constexpr (!kIdentCls) c = s_cls[c];
    const bool hot = s < p.hot_rows;
    const uint32_t fbi = s - p.hot_rows;
    const bool fb = fbi < p.fb_count;              // false for hot states (the subtraction wraps)
    uint32_t m = 0xFFFFu;
    if (fb) m = s_fb[fbi];
    const bool via_fail = fb && !((m >> (c & 15u)) & 1u);
    const uint32_t row = via_fail ? (m >> 16) : s;
    const uint32_t idx = (row << p.log2_ncp) | c;
    return (hot || via_fail) ? uint32_t(s_hot[idx]) : __ldg(p.delta + idx);
}

__device__ __forceinline__ uint32_t dfa_longest(uint32_t s, const DfaParams& p, const uint16_t* s_long) {
    return s < p.hot_long ? uint32_t(s_long[s]) : uint32_t(__ldg(p.longest + s));
}

template <bool kIdentCls>
__global__ void __launch_bounds__(kHotThreads, 1) dfa_hot_kernel(const DfaParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* s_hot = reinterpret_cast<uint16_t*>(smem);
    uint32_t* s_fb = reinterpret_cast<uint32_t*>(s_hot + (size_t(p.hot_rows) << p.log2_ncp));   // 4-byte aligned: rows are >= 2 entries
    uint16_t* s_long = reinterpret_cast<uint16_t*>(s_fb + p.fb_count);
    uint8_t* s_cls = reinterpret_cast<uint8_t*>(s_long + p.hot_long);
    // hot tables -> shared memory (u32 global entries narrowed to u16: the host guarantees they fit)
    for (uint32_t i = threadIdx.x; i < (p.hot_rows << p.log2_ncp); i += kHotThreads) s_hot[i] = uint16_t(__ldg(p.delta + i));
    for (uint32_t i = threadIdx.x; i < p.fb_count; i += kHotThreads) s_fb[i] = __ldg(p.fb_meta + p.hot_rows + i);
    for (uint32_t i = threadIdx.x; i < p.hot_long; i += kHotThreads) s_long[i] = __ldg(p.longest + i);
    if (threadIdx.x < 256) s_cls[threadIdx.x] = p.cls[threadIdx.x];
    __syncthreads();

    const uint64_t n_seg = (p.n + p.seg - 1) / p.seg;
    for (uint64_t seg = uint64_t(blockIdx.x) * kHotThreads + threadIdx.x; seg < n_seg; seg += uint64_t(gridDim.x) * kHotThreads) {
        const uint64_t s0 = seg * uint64_t(p.seg);
        const uint64_t s1 = min(p.n, s0 + uint64_t(p.seg));
        // warm-up start: max_pat_len-1 bytes back, never before the readable history
        int64_t w = int64_t(s0) - int64_t(p.warm);
        if (w < -int64_t(p.hist_valid)) w = -int64_t(p.hist_valid);
        uint32_t s = 0;
        int64_t q0 = w;
        for (; q0 < int64_t(s0) && (q0 & 15); ++q0) s = dfa_step<kIdentCls>(s, *(p.stream + q0), p, s_hot, s_fb, s_cls);
        for (; q0 < int64_t(s0); q0 += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.stream + q0));
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) s = dfa_step<kIdentCls>(s, (ws[k >> 2] >> (8 * (k & 3))) & 0xFF, p, s_hot, s_fb, s_cls);
        }
        uint64_t q = s0;
        if (p.wide) {   // stream and result base 32-byte aligned (segments start at multiples of 4 KiB)
            for (; q + 32 <= s1; q += 32) {
                uint32_t ws[8];
                ldg256(p.stream + q, ws);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[8];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        s = dfa_step<kIdentCls>(s, (ws[half * 4 + (k >> 2)] >> (8 * (k & 3))) & 0xFF, p, s_hot, s_fb, s_cls);
                        const uint32_t o = dfa_longest(s, p, s_long);
                        if (k & 1) r[k >> 1] |= o << 16; else r[k >> 1] = o;
                    }
                    stg256(p.out + q + half * 16, r);
                }
            }
        }
        for (; q + 16 <= s1; q += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.stream + q));
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
            uint32_t r[8];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                s = dfa_step<kIdentCls>(s, (ws[k >> 2] >> (8 * (k & 3))) & 0xFF, p, s_hot, s_fb, s_cls);
                const uint32_t o = dfa_longest(s, p, s_long);
                if (k & 1) r[k >> 1] |= o << 16; else r[k >> 1] = o;
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out + q);
            __stcs(dst, make_uint4(r[0], r[1], r[2], r[3]));
            __stcs(dst + 1, make_uint4(r[4], r[5], r[6], r[7]));
        }
        for (; q < s1; ++q) {  // ragged end
            s = dfa_step<kIdentCls>(s, p.stream[q], p, s_hot, s_fb, s_cls);
            p.out[q] = uint16_t(dfa_longest(s, p, s_long));
        }
    }
}

// The WHOLE automaton in shared memory with the transition and the longest-pattern id of its target fused into one
// u32 entry: one gather per byte instead of two (small-alphabet / small dictionaries, config C5a).
template <bool kIdentCls>
__global__ void __launch_bounds__(kHotThreads, 1) dfa_small_kernel(const DfaParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem);
    const uint32_t n_pad = (p.n_states + 31) & ~31u;          // states per class plane
    uint8_t* s_cls = reinterpret_cast<uint8_t*>(s_tab + (size_t(n_pad) << p.log2_ncp));
    // Layout and bank swizzle, both free at run time.  The table is stored class-major, [class][state]: with few byte
    // classes a state-major entry's bank is (state & 7, class), and a two-letter stream uses 2 of the 4 class slots --
    // half of the banks never.  Class-major, the bank is the low five bits of the state id alone; those are XORed with a
    // fold of the higher bits, because the states an adversarial stream dwells in collide there: breadth-first ids of
    // a^k are 2^k - 1, all = 31 mod 32 (ncu on the plain layout: 11 wavefronts per gather, 81 % of the shared-memory
    // wavefronts conflict replays; state-major with a 3-bit swizzle: 7).  The swizzle is a bijection that keeps the
    // table size (rounded up to 32 states), and the table stores SWIZZLED targets: the walk never computes it.
    auto swz = [](uint32_t s) { return s ^ (((s >> 5) ^ (s >> 10) ^ (s >> 15)) & 31u); };
    const uint32_t ncp = 1u << p.log2_ncp;
    for (uint32_t i = threadIdx.x; i < (p.n_states << p.log2_ncp); i += kHotThreads) {
        const uint32_t nx = __ldg(p.delta + i);
        s_tab[(i & (ncp - 1)) * n_pad + swz(i >> p.log2_ncp)] = swz(nx) | (uint32_t(__ldg(p.longest + nx)) << 16);
    }
    if (threadIdx.x < 256) s_cls[threadIdx.x] = p.cls[threadIdx.x];
    __syncthreads();
    auto step = [&](uint32_t e, uint32_t c) -> uint32_t {
        if constexpr (!kIdentCls) c = s_cls[c];
        return s_tab[c * n_pad + (e & 0xFFFFu)];
    };
    const uint64_t n_seg = (p.n + p.seg - 1) / p.seg;
    for (uint64_t seg = uint64_t(blockIdx.x) * kHotThreads + threadIdx.x; seg < n_seg; seg += uint64_t(gridDim.x) * kHotThreads) {
        const uint64_t s0 = seg * uint64_t(p.seg);
        const uint64_t s1 = min(p.n, s0 + uint64_t(p.seg));
        int64_t w = int64_t(s0) - int64_t(p.warm);
        if (w < -int64_t(p.hist_valid)) w = -int64_t(p.hist_valid);
        uint32_t e = 0;   // state in the low half, its longest pid in the high half
        int64_t q0 = w;
        for (; q0 < int64_t(s0) && (q0 & 15); ++q0) e = step(e, *(p.stream + q0));
        for (; q0 < int64_t(s0); q0 += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.stream + q0));
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 16; ++k) e = step(e, (ws[k >> 2] >> (8 * (k & 3))) & 0xFF);
        }
        uint64_t q = s0;
        if (p.wide) {
            for (; q + 32 <= s1; q += 32) {
                uint32_t ws[8];
                ldg256(p.stream + q, ws);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t r[8];
#pragma unroll
                    for (int k = 0; k < 16; k += 2) {
                        e = step(e, (ws[half * 4 + (k >> 2)] >> (8 * (k & 3))) & 0xFF);
                        const uint32_t e2 = step(e, (ws[half * 4 + (k >> 2)] >> (8 * ((k + 1) & 3))) & 0xFF);
                        r[k >> 1] = __byte_perm(e, e2, 0x7632);   // {longest(e), longest(e2)}
                        e = e2;
                    }
                    stg256(p.out + q + half * 16, r);
                }
            }
        }
        for (; q < s1; ++q) {
            e = step(e, p.stream[q]);
            p.out[q] = uint16_t(e >> 16);
        }
    }
}

}  // namespace

void dfa_plan_hot(uint32_t n_states, uint32_t log2_ncp, const uint32_t* depth_count, uint32_t n_depths,
                  uint32_t* hot_rows, uint32_t* hot_long, uint32_t* fb_count) {
    // Hot rows: whole BFS levels while their transition targets (states of the next level) still fit u16 and
    // the rows fit the shared-memory budget.  The next level keeps one Bloom/failure word per state if those fit
    // beside the longest-ids of all states up to and including that level; hot longest-ids: as many leading
    // states as fit in the rest.
    const size_t budget = 200 * 1024, row_bytes = size_t(2) << log2_ncp;
    uint32_t rows = 0, upto = 0, levels = 0;
    for (uint32_t d = 0; d < n_depths; ++d) {
        const uint32_t level_end = upto + depth_count[d];                       // states of depth <= d
        const uint64_t next_end = uint64_t(level_end) + (d + 1 < n_depths ? depth_count[d + 1] : 0);
        if (next_end > 65536 || size_t(level_end) * row_bytes > budget - 8192) break;
        rows = level_end;
        upto = level_end;
        levels = d + 1;
    }
    *hot_rows = rows;
    size_t left = budget - size_t(rows) * row_bytes;
    *fb_count = 0;
    if (rows > 0 && levels < n_depths) {
        const size_t cnt = depth_count[levels];
        if (cnt * 4 + (size_t(rows) + cnt) * 2 <= left) { *fb_count = uint32_t(cnt); left -= cnt * 4; }
    }
    const uint64_t max_long = left / 2;
    *hot_long = uint32_t(n_states < max_long ? n_states : max_long);
}

cudaError_t dfa_scan_launch(const DfaParams& p_in, bool ident_cls, bool flat, int n_sms, cudaStream_t st, uint64_t* launches) {
    DfaParams p = p_in;
    if (p.n == 0) return cudaSuccess;
    p.wide = ((reinterpret_cast<uintptr_t>(p.stream) | reinterpret_cast<uintptr_t>(p.out)) & 31) == 0;
    if (flat || p.hot_rows == 0) {
        p.seg = 4096;
        const uint64_t segs = (p.n + p.seg - 1) / p.seg;
        const uint32_t grid = uint32_t((segs + kThreads - 1) / kThreads);
        dfa_flat_kernel<<<grid, kThreads, 0, st>>>(p);
        ++*launches;
        return cudaGetLastError();
    }
    // segment length: long enough to amortise the warm-up, short enough to give every SM work
    // up to 16 KiB, cut so that every lane of the persistent grid gets the same number of segments (262,144 segments
    // of 4 KiB on 151,552 lanes leave the second pass 73 % full) and the warm-up of max_pat_len-1 bytes per segment stays
    // small; multiples of 32 bytes (256-bit loads and stores)
    const uint64_t lanes = uint64_t(n_sms) * kHotThreads;
    uint64_t per_lane = (p.n + lanes * 16384 - 1) / (lanes * 16384);
    if (per_lane == 0) per_lane = 1;
    uint64_t seg64 = (p.n + lanes * per_lane - 1) / (lanes * per_lane);
    seg64 = (seg64 + 31) / 32 * 32;
    if (seg64 < 1024) seg64 = 1024;
    const uint32_t seg = uint32_t(seg64);
    p.seg = seg;
    // the whole automaton is hot and its fused u32 table fits: one gather per byte
    const size_t small_smem = (size_t((p.n_states + 31) & ~31u) << p.log2_ncp) * 4 + 256;
    if (p.hot_rows == p.n_states && p.n_states <= 65535 && small_smem <= 200 * 1024 && !p.no_fused) {
        auto kern = ident_cls ? dfa_small_kernel<true> : dfa_small_kernel<false>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(small_smem));
        if (e != cudaSuccess) return e;
        const uint64_t n_seg = (p.n + seg - 1) / seg;
        const uint64_t ctas = (n_seg + kHotThreads - 1) / kHotThreads;
        kern<<<uint32_t(ctas < uint64_t(n_sms) ? ctas : uint64_t(n_sms)), kHotThreads, small_smem, st>>>(p);
        ++*launches;
        return cudaGetLastError();
    }
    const size_t smem = (size_t(p.hot_rows) << p.log2_ncp) * 2 + size_t(p.fb_count) * 4 + size_t(p.hot_long) * 2 + 256;
    auto kern = ident_cls ? dfa_hot_kernel<true> : dfa_hot_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    const uint64_t n_seg = (p.n + seg - 1) / seg;
    const uint64_t ctas = (n_seg + kHotThreads - 1) / kHotThreads;
    const uint32_t grid = uint32_t(ctas < uint64_t(n_sms) ? ctas : uint64_t(n_sms));
    kern<<<grid, kHotThreads, smem, st>>>(p);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pm
