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
// iteration from one 8-byte load, eight results out as one 16-byte store): the per-byte code is short and the
// failure loop is the only data-dependent part.  A first version ran every lane as an independent state machine with
// a shifted byte window and results staged in shared memory; it executed ~12 warp instructions per stream byte and
// was issue-bound at 55 GB/s.
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

__device__ __forceinline__ void fetch(Walk& W, const Tabs& t) {
    ldg_rec(t.recs + size_t(W.s) * 8, W.w);
    W.have = true; W.head = true;
}

// one stream byte: returns the longest-pattern id of the state reached
__device__ __forceinline__ uint32_t step(Walk& W, uint32_t c, const Tabs& t) {
    for (;;) {
        if (W.s < t.n_hot) {   // complete row in shared memory
            W.s = t.s_hot[(W.s << 8) | c];
            if (W.s < t.n_hot) return t.s_long[W.s];
            fetch(W, t);       // a cold state's longest id lives in its record, which the next byte needs anyway
            return W.w[1];
        }
        if (!W.have) fetch(W, t);
        const uint32_t kind = (W.w[0] >> 24) & 3u;
        if (kind == 1u) {      // CHAIN: the next state of the run is s + 1
            if ((W.w[2] & 0xFFu) == c) {
                const uint32_t o = W.w[4] & 0xFFFFu;
                ++W.s;
                W.w[2] = __funnelshift_r(W.w[2], W.w[3], 8); W.w[3] >>= 8;
                W.w[4] = __funnelshift_r(W.w[4], W.w[5], 16); W.w[5] = __funnelshift_r(W.w[5], W.w[6], 16);
                W.w[6] = __funnelshift_r(W.w[6], W.w[7], 16); W.w[7] >>= 16;
                W.w[0] -= 1u << 26;
                W.head = false;
                if ((W.w[0] >> 26) == 0) W.have = false;   // run (or this record's part of it) used up
                return o;
            }
            if (W.head) { W.s = W.w[0] & 0xFFFFFFu; W.have = false; }   // failure transition; the byte is not consumed
            else W.have = false;                                         // inside the run: s's own record has its failure link
            continue;
        }
        if (kind == 0u) {      // BRANCH: goto edges in w[2..]
            const uint32_t cnt = W.w[0] >> 26;
            uint32_t next = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (uint32_t(k) < cnt && (W.w[2 + k] & 0xFFu) == c) next = W.w[2 + k] >> 8;
            if (next != 0xFFFFFFFFu) { W.s = next; fetch(W, t); return W.w[1]; }
            W.s = W.w[0] & 0xFFFFFFu; W.have = false;
            continue;
        }
        // DENSE: a complete row
        W.s = __ldg(t.dense + ((size_t(W.w[2]) << 8) | c));
        W.have = false;
        if (W.s < t.n_hot) return t.s_long[W.s];
        fetch(W, t);
        return W.w[1];
    }
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

    for (uint64_t seg = uint64_t(blockIdx.x) * kThreads + threadIdx.x; seg < p.n_seg; seg += uint64_t(gridDim.x) * kThreads) {
        const uint64_t s0 = seg * uint64_t(p.seg);
        const uint64_t s1 = min(p.n, s0 + uint64_t(p.seg));
        // warm-up: max_pat_len-1 bytes back (never before the readable history), walked but not reported
        int64_t q0 = int64_t(s0) - int64_t(p.warm);
        if (q0 < -int64_t(p.hist_valid)) q0 = -int64_t(p.hist_valid);
        Walk W;
        W.s = 0; W.have = false; W.head = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) W.w[k] = 0;
        for (; q0 < int64_t(s0) && (q0 & 7); ++q0) step(W, stream[q0], t);
#pragma unroll 1
        for (; q0 < int64_t(s0); q0 += 8) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(stream + q0));
#pragma unroll
            for (int k = 0; k < 8; ++k) step(W, ((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 0xFFu, t);
        }
        uint64_t q = s0;
#pragma unroll 1
        for (; q + 8 <= s1; q += 8) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(stream + q));
            uint32_t r[4];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t o = step(W, ((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 0xFFu, t);
                if (k & 1) r[k >> 1] |= o << 16; else r[k >> 1] = o;
            }
            __stcs(reinterpret_cast<uint4*>(p.out + q), make_uint4(r[0], r[1], r[2], r[3]));
        }
        for (; q < s1; ++q) p.out[q] = uint16_t(step(W, stream[q], t));   // ragged end of the stream
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
