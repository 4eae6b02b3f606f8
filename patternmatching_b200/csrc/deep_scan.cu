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
// Lanes are independent state machines advanced in warp-wide rounds (see the kernel); stream bytes arrive eight at a
// time in two registers that are shifted, results leave eight at a time as one 16-byte store.
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

// Every lane is its own state machine over its own segment, and the warp runs them in ROUNDS: in each round a lane
// performs at most one record fetch (all lanes that need one issue it with the same load instruction) and then
// advances as far as it can without another fetch -- at most one stream byte.  A lane whose byte is resolved takes its
// next byte at once; nobody waits for the lane with the longest failure chain.
//   Why this shape: resolving one byte takes 1.6 rounds on average on the C5b stream but 4.9 for the slowest of 32
//   lanes, so a byte-synchronous warp (all lanes on byte k together) wastes two thirds of its rounds; and a per-lane
//   `for (;;)` with the fetch inside it made the hardware run the lanes' iterations one after the other (1-2 active
//   lanes per LDG, every latency exposed 32 times).  Both were measured at 45-50 GB/s.
struct LaneIO {
    const uint8_t* in;      // stream + s0: positions are 32-bit offsets from the segment start (negative = warm-up)
    uint16_t* out;          // out + s0
    int32_t rel, rel_end;   // next offset to consume, end of the segment
    int32_t rel_lo;         // first readable offset (history limit)
    uint32_t in_lo, in_hi;  // the bytes at rel, rel+1, ... (shifted as they are consumed; refilled at multiples of 8)
    uint32_t r[4];          // results of the current group of 8 positions
};

// slow path of the refill: a group of 8 that is not entirely readable (start of the history, end of the stream) or a
// start that is not a multiple of 8
__device__ __noinline__ uint2 load8_slow(const uint8_t* in, int32_t rel, int32_t rel_lo, int32_t rel_end) {
    const int32_t ra = rel & ~7;
    uint32_t a = 0, b = 0;
    for (int k = 0; k < 8; ++k) {
        const int32_t g = ra + k;
        if (g >= rel_lo && g < rel_end) { if (k < 4) a |= uint32_t(in[g]) << (8 * k); else b |= uint32_t(in[g]) << (8 * (k - 4)); }
    }
    const uint32_t sk = uint32_t(rel - ra) * 8;
    if (sk >= 32) { a = b >> (sk - 32); b = 0; }
    else if (sk) { a = __funnelshift_r(a, b, sk); b >>= sk; }
    return make_uint2(a, b);
}

// the byte at io.rel has been resolved with result o: report it (not during the warm-up) and step to the next byte
__device__ __forceinline__ void emit_and_advance(LaneIO& io, uint32_t o) {
    const int32_t rel = io.rel;
    if (rel >= 0) {
        // results of a group of 8 accumulate in r[] by shifting: after 8 of them r[0..3] hold positions 0..7 in order
        io.r[0] = __funnelshift_r(io.r[0], io.r[1], 16); io.r[1] = __funnelshift_r(io.r[1], io.r[2], 16);
        io.r[2] = __funnelshift_r(io.r[2], io.r[3], 16); io.r[3] = (io.r[3] >> 16) | (o << 16);
        if ((rel & 7) == 7) __stcs(reinterpret_cast<uint4*>(io.out + (rel - 7)), make_uint4(io.r[0], io.r[1], io.r[2], io.r[3]));
    }
    io.rel = rel + 1;
    if ((io.rel & 7) == 0) {
        if (io.rel + 8 <= io.rel_end && io.rel >= io.rel_lo) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(io.in + io.rel));
            io.in_lo = v.x; io.in_hi = v.y;
        } else if (io.rel < io.rel_end) { const uint2 v = load8_slow(io.in, io.rel, io.rel_lo, io.rel_end); io.in_lo = v.x; io.in_hi = v.y; }
    } else {
        io.in_lo = __funnelshift_r(io.in_lo, io.in_hi, 8); io.in_hi >>= 8;
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
    const uint32_t n_hot = p.n_hot;

    for (uint64_t seg = uint64_t(blockIdx.x) * kThreads + threadIdx.x; seg < p.n_seg; seg += uint64_t(gridDim.x) * kThreads) {
        LaneIO io;
        const uint64_t s0 = seg * uint64_t(p.seg);
        io.in = p.stream + s0; io.out = p.out + s0;
        io.rel_end = int32_t(min(uint64_t(p.seg), p.n - s0));
        io.rel_lo = -int32_t(min(uint64_t(p.warm), s0 + p.hist_valid));   // warm-up: max_pat_len-1 bytes back, never before the readable history
        io.rel = io.rel_lo;
        io.r[0] = io.r[1] = io.r[2] = io.r[3] = 0;
        { const uint2 v = load8_slow(io.in, io.rel, io.rel_lo, io.rel_end); io.in_lo = v.x; io.in_hi = v.y; }
        Walk W;
        W.s = 0; W.have = false; W.head = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) W.w[k] = 0;
        bool arrived = false;   // s was just entered and is cold: the byte is consumed, only s's longest id is missing
#pragma unroll 1
        while (io.rel < io.rel_end) {
            // ---- one round: at most one record fetch, then as far as the lane gets without another ----
            if (W.s >= n_hot && !W.have) {
                ldg_rec(p.recs + size_t(W.s) * 8, W.w);
                W.have = true; W.head = true;
            }
            if (arrived) {                                  // the record just fetched completes the previous byte ...
                arrived = false;
                emit_and_advance(io, W.w[1]);
                if (io.rel >= io.rel_end) break;            // ... and serves the next one in the same round
            }
            const uint32_t c = io.in_lo & 0xFFu;
            if (W.s < n_hot) {                              // complete row in shared memory
                W.s = s_hot[(W.s << 8) | c];
                W.have = false;
                if (W.s < n_hot) emit_and_advance(io, s_long[W.s]); else arrived = true;
                continue;
            }
            const uint32_t kind = (W.w[0] >> 24) & 3u, fail = W.w[0] & 0xFFFFFFu;
            if (kind == 1u) {                               // CHAIN: the next state of the run is s + 1
                if ((W.w[2] & 0xFFu) == c) {
                    const uint32_t o = W.w[4] & 0xFFFFu;
                    ++W.s;
                    W.w[2] = __funnelshift_r(W.w[2], W.w[3], 8); W.w[3] >>= 8;
                    W.w[4] = __funnelshift_r(W.w[4], W.w[5], 16); W.w[5] = __funnelshift_r(W.w[5], W.w[6], 16);
                    W.w[6] = __funnelshift_r(W.w[6], W.w[7], 16); W.w[7] >>= 16;
                    W.w[0] -= 1u << 26;
                    W.head = false;
                    if ((W.w[0] >> 26) == 0) W.have = false;   // this record's part of the run is used up
                    emit_and_advance(io, o);
                } else {
                    if (W.head) W.s = fail;                 // failure transition; the byte is not consumed
                    W.have = false;                         // (inside the run: s's own record has its failure link)
                }
            } else if (kind == 0u) {                        // BRANCH: goto edges in w[2..]
                const uint32_t cnt = W.w[0] >> 26;
                uint32_t next = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    if (uint32_t(k) < cnt && (W.w[2 + k] & 0xFFu) == c) next = W.w[2 + k] >> 8;
                W.have = false;
                if (next != 0xFFFFFFFFu) { W.s = next; arrived = true; }
                else W.s = fail;
            } else {                                        // DENSE: a complete row
                W.s = __ldg(p.dense_rows + ((size_t(W.w[2]) << 8) | c));
                W.have = false;
                if (W.s < n_hot) emit_and_advance(io, s_long[W.s]); else arrived = true;
            }
        }
        if (arrived) {   // the last byte of the segment ended in a cold state: its longest id still has to be fetched
            ldg_rec(p.recs + size_t(W.s) * 8, W.w);
            emit_and_advance(io, W.w[1]);
        }
        // ragged end of the stream: the last group is not full
        const uint32_t rest = uint32_t(io.rel_end) & 7u;
        if (rest) {
            // the `rest` results sit in the top of r[]: shift them down to position 0
            for (uint32_t k = rest; k < 8; ++k) {
                io.r[0] = __funnelshift_r(io.r[0], io.r[1], 16); io.r[1] = __funnelshift_r(io.r[1], io.r[2], 16);
                io.r[2] = __funnelshift_r(io.r[2], io.r[3], 16); io.r[3] >>= 16;
            }
            const int32_t base = io.rel_end - int32_t(rest);
            for (uint32_t k = 0; k < rest; ++k) {
                io.out[base + k] = uint16_t(io.r[0]);
                io.r[0] = __funnelshift_r(io.r[0], io.r[1], 16); io.r[1] = __funnelshift_r(io.r[1], io.r[2], 16);
                io.r[2] = __funnelshift_r(io.r[2], io.r[3], 16); io.r[3] >>= 16;
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
