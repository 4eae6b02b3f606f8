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
//   * a branching state's record holds its goto edges (<= 6), its failure link and its own longest-pattern id:
//     arriving, reporting and leaving cost one fetch;
//   * busier states (DENSE) own a complete 256-entry DFA row and are recognised by their id alone: one 4-byte load per
//     byte, no record, no failure chain;
//   * the root and the depth-1 states keep complete DFA rows in shared memory (u16), and so do the longest-pattern ids
//     of every state a shared-memory row or a DENSE row usually leads to (depth <= 2 and DENSE states).
// 0.62 record fetches per byte on the C5b stream (tests/test_host_compiler.py prints it), 0.13 on planted traffic.
//
// Lanes are independent state machines advanced in warp-wide ROUNDS of one straight-line instruction sequence (see the
// kernel); stream bytes arrive eight at a time, one group ahead of their use, results leave eight at a time as one
// 16-byte store.
#include "deep_scan.cuh"
#include "pm_dev.cuh"

namespace pm {
namespace {

constexpr int kThreads = 1024;

__device__ __forceinline__ void ldg_rec(const uint32_t* ptr, uint32_t (&w)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr));
}

// The 8 stream bytes of group g (a multiple of 8, relative to the segment start): one aligned 8-byte load when the whole
// group is readable, byte by byte at the two ends of the readable range [lo, hi).
__device__ __forceinline__ uint2 load_group(const uint8_t* in, int32_t g, int32_t lo, int32_t hi) {
    if (g >= lo && g + 8 <= hi) return __ldg(reinterpret_cast<const uint2*>(in + g));
    uint32_t a = 0, b = 0;
    if (g + 8 > lo && g < hi) {
        for (int k = 0; k < 8; ++k) {
            const int32_t q = g + k;
            if (q >= lo && q < hi) { if (k < 4) a |= uint32_t(in[q]) << (8 * k); else b |= uint32_t(in[q]) << (8 * (k - 4)); }
        }
    }
    return make_uint2(a, b);
}

// Every lane walks its own segment; the warp advances all of them in ROUNDS.  One round, the same for every lane:
//   1. issue this round's table load -- the 32-byte record of a cold state that is not in registers yet, or the 4-byte
//      DENSE-row entry for the current byte -- every lane that needs one with the same instruction;
//   2. report the byte consumed in the previous round (its longest-pattern id is known by now: it came with that
//      round's transition, or it is word 1 of the record just fetched);
//   3. make one transition on the current byte: shared-memory row, DENSE entry, the next step of a CHAIN record, a
//      BRANCH record's goto edge -- or, without consuming the byte, a failure link.
// A round consumes at most one byte and contains at most one dependent global load; nobody waits for the lane with the
// longest failure chain (a byte-synchronous warp wastes two thirds of its rounds on the C5b stream: 1.6 rounds per byte
// on average, 4.9 for the slowest of 32 lanes), and the per-lane bookkeeping is one straight-line sequence instead of
// the nested, separately diverging paths of the first generations of this kernel (9.5 warp instructions per stream byte
// with 9.8 of 32 lanes active, profiles/r02_deep_kernel.md).
__global__ void __launch_bounds__(kThreads, 1) deep_scan_kernel(const DeepParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* s_hot = reinterpret_cast<uint16_t*>(smem);                              // n_hot x 256
    uint16_t* s_long = s_hot + (size_t(p.n_hot) << 8);                                // n_small
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.hot_rows);
        uint4* dst = reinterpret_cast<uint4*>(s_hot);
        for (uint32_t i = threadIdx.x; i < (p.n_hot << 8) / 8; i += kThreads) dst[i] = __ldg(src + i);
        for (uint32_t i = threadIdx.x; i < p.n_small; i += kThreads) s_long[i] = __ldg(p.hot_longest + i);
    }
    __syncthreads();
    const uint32_t n_hot = p.n_hot, dense_end = p.dense_end, n_small = p.n_small;

    for (uint64_t seg = uint64_t(blockIdx.x) * kThreads + threadIdx.x; seg < p.n_seg; seg += uint64_t(gridDim.x) * kThreads) {
        const uint64_t s0 = seg * uint64_t(p.seg);
        const uint8_t* in = p.stream + s0;
        uint16_t* out = p.out + s0;
        const int32_t rel_end = int32_t(min(uint64_t(p.seg), p.n - s0));
        const int32_t rel_lo = -int32_t(min(uint64_t(p.warm), s0 + p.hist_valid));   // warm-up: max_pat_len-1 bytes back, never before the readable history
        const int32_t rd_hi = int32_t(min(uint64_t(p.seg) + 16, p.n - s0));           // bytes after the segment are readable up to the end of the stream
        int32_t rel = rel_lo;        // next byte to consume
        int32_t orel = rel_lo;       // next position to report
        uint32_t cur_lo, cur_hi, nxt_lo, nxt_hi;   // bytes rel.. of the current group; the whole next group
        {
            const int32_t g = rel & ~7;
            const uint2 a = load_group(in, g, rel_lo, rd_hi), b = load_group(in, g + 8, rel_lo, rd_hi);
            const uint32_t sk = uint32_t(rel - g) * 8;
            cur_lo = a.x; cur_hi = a.y;
            if (sk >= 32) { cur_lo = cur_hi >> (sk - 32); cur_hi = 0; }
            else if (sk) { cur_lo = __funnelshift_r(cur_lo, cur_hi, sk); cur_hi >>= sk; }
            nxt_lo = b.x; nxt_hi = b.y;
        }
        uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;   // results of the current group of 8 positions, shifted in from the top
        uint32_t s = 0;              // current state
        uint32_t w[8];               // a record (CHAIN: shifted as the run is consumed; bit 30 of w[0]: part of it has been consumed)
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = 0;
        constexpr uint32_t kNone = 0xFFFFFFFFu;
        uint32_t hs = kNone;         // the state w describes (kNone: nothing usable in w)
        // The byte consumed in the previous round is reported one round later (orel == rel - 1 then): its longest id is
        // pend_val, or -- kNone -- word 1 of the record the next round fetches.  Everything a lane carries from round to
        // round is a number: flags cost the compiler a register shuffle each at every join of the diverging paths.
        uint32_t pend_val = 0;
#pragma unroll 1
        for (;;) {
            const bool active = rel < rel_end;
            const bool pend = orel < rel;
            if (!active && !pend) break;
            const uint32_t c = cur_lo & 0xFFu;
            // ---- 1. this round's table load ----
            const bool is_hot = s < n_hot;
            const bool is_dense = !is_hot && s < dense_end;
            const bool fetch = s >= dense_end && hs != s;
            uint32_t dv = 0;
            if (fetch) { ldg_rec(p.recs + size_t(s) * 8, w); hs = s; }
            if (is_dense && active) dv = __ldg(p.dense_rows + (((s - n_hot) << 8) | c));
            // ---- 2. report the previous byte ----
            if (pend) {
                const uint32_t o = pend_val == kNone ? w[1] : pend_val;
                r0 = __funnelshift_r(r0, r1, 16); r1 = __funnelshift_r(r1, r2, 16);
                r2 = __funnelshift_r(r2, r3, 16); r3 = (r3 >> 16) | (o << 16);
                if ((orel & 7) == 7 && orel >= 0) __stcs(reinterpret_cast<uint4*>(out + (orel - 7)), make_uint4(r0, r1, r2, r3));
                ++orel;
            }
            if (!active) continue;   // only the last report was left
            // ---- 3. one transition on c ----
            uint32_t ns, val = kNone;   // val: the longest id at ns when this round already knows it
            bool consumed = true;
            if (is_hot) {
                ns = s_hot[(s << 8) | c];
            } else if (is_dense) {
                ns = dv;
            } else {
                const uint32_t w0 = w[0], fail = w0 & 0xFFFFFFu;
                hs = kNone;
                if (w0 & (1u << 24)) {                          // CHAIN (kind 1): the next state of the run is s + 1
                    if ((w[2] & 0xFFu) == c) {
                        val = w[4] & 0xFFFFu;
                        ns = s + 1;
                        w[2] = __funnelshift_r(w[2], w[3], 8); w[3] >>= 8;
                        w[4] = __funnelshift_r(w[4], w[5], 16); w[5] = __funnelshift_r(w[5], w[6], 16);
                        w[6] = __funnelshift_r(w[6], w[7], 16); w[7] >>= 16;
                        w[0] = (w0 - (1u << 26)) | (1u << 30);
                        if (w[0] & (15u << 26)) hs = ns;        // steps left in this record: it now describes s + 1
                    } else {
                        ns = (w0 & (1u << 30)) ? s : fail;      // failure transition (inside the run: s's own record has its link)
                        consumed = false;
                    }
                } else {                                        // BRANCH (kind 0): goto edges in w[2..7]; LEAF (kind 2): none
                    uint32_t next = kNone;
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const uint32_t t = w[2 + k] ^ c;        // low byte zero = the edge's byte is c; the child sits above it
                        if ((t & 0xFFu) == 0) next = t >> 8;
                    }
                    consumed = !(w0 & (2u << 24)) && next != kNone;
                    ns = consumed ? next : fail;
                }
            }
            if (consumed) {
                if (val == kNone && ns < n_small) val = s_long[ns];
                pend_val = val;
                ++rel;
                if ((rel & 7) == 0) {
                    cur_lo = nxt_lo; cur_hi = nxt_hi;
                    // the group after the next one (used eight bytes from now); rel + 8 > rel_lo always holds here
                    if (rel + 16 <= rd_hi) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(in + rel + 8)); nxt_lo = v.x; nxt_hi = v.y; }
                    else { const uint2 v = load_group(in, rel + 8, rel_lo, rd_hi); nxt_lo = v.x; nxt_hi = v.y; }
                } else {
                    cur_lo = __funnelshift_r(cur_lo, cur_hi, 8); cur_hi >>= 8;
                }
            }
            s = ns;
        }
        // ragged end of the stream: the last group is not full; its `rest` results sit in the top of r
        const uint32_t rest = uint32_t(rel_end) & 7u;
        if (rest) {
            for (uint32_t k = rest; k < 8; ++k) {
                r0 = __funnelshift_r(r0, r1, 16); r1 = __funnelshift_r(r1, r2, 16);
                r2 = __funnelshift_r(r2, r3, 16); r3 >>= 16;
            }
            const int32_t base = rel_end - int32_t(rest);
            for (uint32_t k = 0; k < rest; ++k) {
                out[base + k] = uint16_t(r0);
                r0 = __funnelshift_r(r0, r1, 16); r1 = __funnelshift_r(r1, r2, 16);
                r2 = __funnelshift_r(r2, r3, 16); r3 >>= 16;
            }
        }
    }
}

}  // namespace

size_t deep_smem_bytes(uint32_t n_hot, uint32_t n_small) { return (size_t(n_hot) << 9) + size_t((n_small + 7) & ~7u) * 2; }

cudaError_t deep_scan_launch(const DeepParams& p_in, int n_sms, cudaStream_t st, uint64_t* launches) {
    DeepParams p = p_in;
    if (p.n == 0) return cudaSuccess;
    const size_t smem = deep_smem_bytes(p.n_hot, p.n_small);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    // segments: up to 16 KiB each (every segment pays a warm-up of max_pat_len-1 bytes: 346 of 4 KiB were 8 % of the
    // walk), cut so that the persistent grid's lanes get the same number of them; multiples of 16 bytes (8-byte loads,
    // 16-byte result stores)
    const uint64_t lanes = uint64_t(n_sms) * kThreads;
    uint64_t per_lane = (p.n + lanes * 16384 - 1) / (lanes * 16384);
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
