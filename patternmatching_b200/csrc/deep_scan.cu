// deep_scan.cu -- exact dictionary scan for DEEP-match traffic (pattern-prefix soup, long shared prefixes): per-thread
// forward walks of the reference's own goto + failure automaton (Core/src/mpac.c:147-210, 304-319), packed into one
// 32-byte record per state (dict.hpp: DeepTables; 23 MB for snort+et, L2-resident) instead of the 734 MB dense DFA
// that dfa_flat_kernel gathers from (one 32-byte DRAM sector per stream byte: 52 GB/s on C5b in round 1).
//
// What a forward walker pays for on such traffic is the DIVERGENT global gather: 32 lanes in 32 unrelated states
// cost 32 L1 wavefronts and 32 sectors per load instruction, whatever the table looks like.  The record layout makes
// one 32-byte fetch go a long way:
//   * states are numbered depth-first below depth 2, so a run of single-child states is a run of consecutive ids and
//     ONE record carries the bytes and longest-pattern ids of up to eight steps of the run (kind CHAIN): following a
//     pattern's text costs one fetch per eight bytes and a few register shifts per byte;
//   * a branching state's record holds its goto edges (<= 6; busier states point at a complete 256-entry row), its
//     failure link and its own longest-pattern id: arriving, reporting and leaving cost one fetch;
//   * the root and the depth-1 states keep complete DFA rows in shared memory (u16): a failure chain that reaches
//     them is resolved without another fetch.
// 0.62 record fetches per byte on the C5b stream (tests/test_host_compiler.py prints it), 0.13 on planted traffic.
//
// The walk is byte-synchronous per warp (all lanes take byte k of their own segment together, eight bytes per loop
// iteration from one 8-byte load, eight results out as one 16-byte store) and every byte is resolved in warp-wide
// rounds with one converged record fetch per round (see step()).
#include "deep_scan.cuh"
#include "pm_dev.cuh"

namespace pm {
namespace {

constexpr int kThreads = 1024;

__device__ __forceinline__ void ldg_rec(const uint32_t* ptr, uint32_t (&w)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr));
}

struct Walk {
    uint32_t s;       // current state
    uint32_t w[8];    // record of s (CHAIN: shifted as the run is consumed), valid when `have`
    bool have;        // w belongs to s
    bool head;        // CHAIN: nothing of the record has been consumed yet, so w[0] holds s's own failure link
};

struct Tabs {
    const uint16_t* s_hot;
    const uint16_t* s_long;
    const uint32_t* recs;
    const uint32_t* dense;
    uint32_t n_hot;
};

// One stream byte for every lane of the warp (lanes with active == false only keep the others company).  The walk of
// one byte may need several dependent record fetches (arrive at a cold state: its record for the longest id; failure
// chain: the record of every cold failure state tried).  The loop below runs in warp-wide ROUNDS: in each round all
// lanes that need a record fetch it with the SAME load instruction, then every unfinished lane advances as far as it
// can without another fetch.  (A per-lane `for (;;)` with the fetch inside it made the hardware run the lanes' loop
// iterations one after the other: 1-2 active lanes per LDG, every fetch latency exposed 32 times -- 50 GB/s.)
__device__ __forceinline__ uint32_t step(Walk& W, uint32_t c, bool active, const Tabs& t) {
    uint32_t o = 0;
    bool done = !active;
    bool arrived = false;   // s was just entered and is cold: only its longest id is missing
    for (;;) {
        if (!done && W.s >= t.n_hot && !W.have) {
            ldg_rec(t.recs + size_t(W.s) * 8, W.w);
            W.have = true; W.head = true;
        }
        if (!done) {
            if (arrived) { o = W.w[1]; done = true; }
            else if (W.s < t.n_hot) {                       // complete row in shared memory
                W.s = t.s_hot[(W.s << 8) | c];
                W.have = false;
                if (W.s < t.n_hot) { o = t.s_long[W.s]; done = true; } else arrived = true;
            } else {
                const uint32_t kind = (W.w[0] >> 24) & 3u, fail = W.w[0] & 0xFFFFFFu;
                if (kind == 1u) {                           // CHAIN: the next state of the run is s + 1
                    if ((W.w[2] & 0xFFu) == c) {
                        o = W.w[4] & 0xFFFFu;
                        ++W.s;
                        W.w[2] = __funnelshift_r(W.w[2], W.w[3], 8); W.w[3] >>= 8;
                        W.w[4] = __funnelshift_r(W.w[4], W.w[5], 16); W.w[5] = __funnelshift_r(W.w[5], W.w[6], 16);
                        W.w[6] = __funnelshift_r(W.w[6], W.w[7], 16); W.w[7] >>= 16;
                        W.w[0] -= 1u << 26;
                        W.head = false;
                        if ((W.w[0] >> 26) == 0) W.have = false;   // this record's part of the run is used up
                        done = true;
                    } else {
                        if (W.head) W.s = fail;             // failure transition; the byte is not consumed
                        W.have = false;                     // (inside the run: s's own record has its failure link)
                    }
                } else if (kind == 0u) {                    // BRANCH: goto edges in w[2..]
                    const uint32_t cnt = W.w[0] >> 26;
                    uint32_t next = 0xFFFFFFFFu;
#pragma unroll
                    for (int k = 0; k < 6; ++k)
                        if (uint32_t(k) < cnt && (W.w[2 + k] & 0xFFu) == c) next = W.w[2 + k] >> 8;
                    W.have = false;
                    if (next != 0xFFFFFFFFu) { W.s = next; arrived = true; }
                    else W.s = fail;
                } else {                                    // DENSE: a complete row
                    W.s = __ldg(t.dense + ((size_t(W.w[2]) << 8) | c));
                    W.have = false;
                    if (W.s < t.n_hot) { o = t.s_long[W.s]; done = true; } else arrived = true;
                }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, !done)) break;
    }
    return o;
}

__global__ void __launch_bounds__(kThreads, 1) deep_scan_kernel(const DeepParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* s_hot = reinterpret_cast<uint16_t*>(smem);                              // n_hot x 256
    uint16_t* s_long = s_hot + (size_t(p.n_hot) << 8);                                // n_hot
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.hot_rows);
        uint4* dst = reinterpret_cast<uint4*>(s_hot);
        for (uint32_t i = threadIdx.x; i < (p.n_hot << 8) / 8; i += kThreads) dst[i] = __ldg(src + i);
        for (uint32_t i = threadIdx.x; i < p.n_hot; i += kThreads) s_long[i] = __ldg(p.hot_longest + i);
    }
    __syncthreads();
    Tabs t;
    t.s_hot = s_hot; t.s_long = s_long; t.recs = p.recs; t.dense = p.dense_rows; t.n_hot = p.n_hot;
    const uint8_t* __restrict__ stream = p.stream;
    const int64_t lo = -int64_t(p.hist_valid), hi = int64_t(p.n);
    const int64_t warm8 = (int64_t(p.warm) + 7) & ~int64_t(7);   // warm-up in whole 8-byte blocks (more never hurts)

    // the segment loop is warp-uniform (step() is a warp-wide routine): lanes without a segment run along inactive
    const uint64_t seg_stride = uint64_t(gridDim.x) * kThreads;
    const uint64_t warp_first = uint64_t(blockIdx.x) * kThreads + (threadIdx.x & ~31u);
    for (uint64_t base = warp_first; base < p.n_seg; base += seg_stride) {
        const uint64_t seg = base + (threadIdx.x & 31u);
        const bool lane_valid = seg < p.n_seg;
        const int64_t s0 = int64_t(seg * uint64_t(p.seg));
        const int64_t s1 = lane_valid ? min(hi, s0 + int64_t(p.seg)) : s0;
        Walk W;
        W.s = 0; W.have = false; W.head = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) W.w[k] = 0;
#pragma unroll 1
        for (int64_t rel = -warm8; rel < int64_t(p.seg); rel += 8) {
            const int64_t q = s0 + rel;
            uint2 v = make_uint2(0u, 0u);
            const bool whole = lane_valid && q >= lo && q + 8 <= s1;
            if (whole) {
                v = __ldg(reinterpret_cast<const uint2*>(stream + q));
            } else if (lane_valid) {
                for (int k = 0; k < 8; ++k) {
                    const int64_t g = q + k;
                    if (g >= lo && g < s1) { if (k < 4) v.x |= uint32_t(stream[g]) << (8 * k); else v.y |= uint32_t(stream[g]) << (8 * (k - 4)); }
                }
            }
            uint32_t r[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const bool active = lane_valid && (whole || (q + k >= lo && q + k < s1));
                const uint32_t o = step(W, ((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 0xFFu, active, t);
                if (k & 1) r[k >> 1] |= o << 16; else r[k >> 1] = o;
            }
            if (rel >= 0) {
                if (whole) __stcs(reinterpret_cast<uint4*>(p.out + q), make_uint4(r[0], r[1], r[2], r[3]));
                else if (lane_valid)
                    for (int k = 0; k < 8; ++k)
                        if (q + k < s1) p.out[q + k] = uint16_t(r[k >> 1] >> (16 * (k & 1)));
            }
        }
    }
}

}  // namespace

size_t deep_smem_bytes(uint32_t n_hot) { return (size_t(n_hot) << 9) + size_t((n_hot + 7) & ~7u) * 2; }

cudaError_t deep_scan_launch(const DeepParams& p_in, int n_sms, cudaStream_t st, uint64_t* launches) {
    DeepParams p = p_in;
    if (p.n == 0) return cudaSuccess;
    const size_t smem = deep_smem_bytes(p.n_hot);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    // segments: about 4 KiB each (the warm-up is max_pat_len-1 bytes), cut so that the persistent grid's lanes get the
    // same number of them; multiples of 16 bytes (8-byte loads, 16-byte result stores)
    const uint64_t lanes = uint64_t(n_sms) * kThreads;
    uint64_t per_lane = (p.n + lanes * 4096 - 1) / (lanes * 4096);
    if (per_lane == 0) per_lane = 1;
    uint64_t seg = (p.n + lanes * per_lane - 1) / (lanes * per_lane);
    seg = (seg + 15) / 16 * 16;
    if (seg < 1024) seg = 1024;
    p.seg = uint32_t(seg);
    p.n_seg = (p.n + seg - 1) / seg;
    cudaError_t e = cudaFuncSetAttribute(deep_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    const uint64_t ctas = (p.n_seg + kThreads - 1) / kThreads;
    const uint32_t grid = uint32_t(ctas < uint64_t(n_sms) ? ctas : uint64_t(n_sms));
    deep_scan_kernel<<<grid, kThreads, smem, st>>>(p);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pm
