// deep_scan.cu -- exact dictionary scan for DEEP-match traffic (pattern-prefix soup, long shared prefixes): per-thread
// forward walks of the reference's own goto + failure automaton (Core/src/mpac.c:147-210, 304-319), packed into one
// 32-byte record per state (dict.hpp: DeepTables; 23 MB for snort+et, L2-resident) instead of the 734 MB dense DFA
// that dfa_flat_kernel gathers from.
//
// What binds a forward walker on such traffic is the number of DIVERGENT global gathers: 32 lanes in 32 unrelated
// states cost 32 L1 wavefronts per load instruction, whatever the table looks like (profiles/r01: dfa_flat 52 GB/s,
// ~70 wavefronts per 32 bytes).  So the unit here is "one 32-byte record fetch per lane and loop iteration, and as
// many stream bytes as possible out of it":
//   * states are numbered depth-first below depth 2, so a run of single-child states is a run of consecutive ids and
//     ONE record carries the bytes and longest-pattern ids of up to eight steps of the run (kind CHAIN);
//   * a branching state's record holds its goto edges (<= 6; busier states point at a complete 256-entry row), its
//     failure link and its own longest-pattern id -- arriving, reporting and leaving cost one fetch;
//   * the root and the depth-1 states keep complete DFA rows in shared memory (u16): a failure chain that reaches
//     them is resolved without another fetch.
// Every lane is an independent state machine over its own segment (max_pat_len-1 warm-up bytes first); a
// loop iteration issues at most one record fetch per lane and lanes in hot states take several shared-memory steps
// meanwhile.  Stream bytes sit in a 16-byte register window that is shifted, results are staged per lane in shared
// memory and leave as 32-byte stores.
#include "deep_scan.cuh"
#include "pm_dev.cuh"

namespace pm {
namespace {

constexpr int kThreads = 1024;
constexpr int kOutStride = 40;  // bytes of result staging per lane (16 x u16 + padding against bank conflicts)
constexpr int kHotSteps = 4;    // shared-memory steps a lane in a hot state takes per loop iteration

__device__ __forceinline__ void ldg_rec(const uint32_t* ptr, uint32_t (&w)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr));
}

struct Lane {
    uint64_t wlo, whi;   // the next stream bytes, byte 0 of wlo = the byte at position q
    int64_t q;           // next position to consume, relative to p.stream (negative = history)
    int64_t s0, s1;      // this segment reports [s0, s1)
};

// (re)fill the window so that its first byte is position q; bytes outside [lo, hi) read as 0 and are never consumed
__device__ __forceinline__ void refill(Lane& L, const uint8_t* __restrict__ stream, int64_t lo, int64_t hi) {
    const int64_t qa = (L.q >> 4) << 4;   // floor to 16 (arithmetic shift: q may be negative)
    uint64_t a = 0, b = 0;
    if (qa >= lo && qa + 16 <= hi) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(stream + qa));
        a = uint64_t(v.x) | (uint64_t(v.y) << 32);
        b = uint64_t(v.z) | (uint64_t(v.w) << 32);
    } else {
        for (int k = 0; k < 16; ++k) {
            const int64_t g = qa + k;
            const uint64_t c = (g >= lo && g < hi) ? uint64_t(stream[g]) : 0ull;
            if (k < 8) a |= c << (8 * k); else b |= c << (8 * (k - 8));
        }
    }
    const uint32_t sk = uint32_t(L.q - qa);   // bytes of the window that lie before q (only at the start of a segment)
    if (sk >= 8) { a = b >> (8 * (sk - 8)); b = 0; }
    else if (sk) { a = (a >> (8 * sk)) | (b << (64 - 8 * sk)); b >>= 8 * sk; }
    L.wlo = a; L.whi = b;
}
__device__ __forceinline__ uint32_t peek(const Lane& L) { return uint32_t(L.wlo) & 0xFFu; }
__device__ __forceinline__ void advance(Lane& L, const uint8_t* __restrict__ stream, int64_t lo, int64_t hi) {
    ++L.q;
    if ((L.q & 15) == 0) { if (L.q < L.s1) refill(L, stream, lo, hi); }
    else { L.wlo = (L.wlo >> 8) | (L.whi << 56); L.whi >>= 8; }
}

// result for position pos: staged in the lane's strip of shared memory; a full strip of 16 leaves as one 32-byte store
__device__ __forceinline__ void emit(const Lane& L, int64_t pos, uint32_t pid, uint8_t* s_strip, uint16_t* __restrict__ out, bool wide) {
    if (pos < L.s0) return;   // warm-up bytes are walked, not reported
    const uint32_t k = uint32_t(pos) & 15u;
    *reinterpret_cast<uint16_t*>(s_strip + 2 * k) = uint16_t(pid);
    if (k == 15u) {
        const uint2* s2 = reinterpret_cast<const uint2*>(s_strip);
        const uint2 a = s2[0], b = s2[1], c = s2[2], d = s2[3];
        uint16_t* dst = out + (pos - 15);
        if (wide) {
            asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         :: "l"(dst), "r"(a.x), "r"(a.y), "r"(b.x), "r"(b.y), "r"(c.x), "r"(c.y), "r"(d.x), "r"(d.y) : "memory");
        } else {
            __stcs(reinterpret_cast<uint4*>(dst), make_uint4(a.x, a.y, b.x, b.y));
            __stcs(reinterpret_cast<uint4*>(dst) + 1, make_uint4(c.x, c.y, d.x, d.y));
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1) deep_scan_kernel(const DeepParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint16_t* s_hot = reinterpret_cast<uint16_t*>(smem);                              // n_hot x 256
    uint16_t* s_long = s_hot + (size_t(p.n_hot) << 8);                                // n_hot
    uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_long + ((p.n_hot + 7) & ~7u));    // kThreads x kOutStride
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.hot_rows);
        uint4* dst = reinterpret_cast<uint4*>(s_hot);
        for (uint32_t i = threadIdx.x; i < (p.n_hot << 8) / 8; i += kThreads) dst[i] = __ldg(src + i);
        for (uint32_t i = threadIdx.x; i < p.n_hot; i += kThreads) s_long[i] = __ldg(p.hot_longest + i);
    }
    __syncthreads();
    uint8_t* s_strip = s_stage + threadIdx.x * kOutStride;
    const uint8_t* __restrict__ stream = p.stream;
    const int64_t lo = -int64_t(p.hist_valid), hi = int64_t(p.n);
    const uint32_t n_hot = p.n_hot;
    const bool wide = p.wide != 0;

    for (uint64_t seg = uint64_t(blockIdx.x) * kThreads + threadIdx.x; seg < p.n_seg; seg += uint64_t(gridDim.x) * kThreads) {
        Lane L;
        L.s0 = int64_t(seg * uint64_t(p.seg));
        L.s1 = min(hi, L.s0 + int64_t(p.seg));
        L.q = max(lo, L.s0 - int64_t(p.warm));
        refill(L, stream, lo, hi);
        uint32_t s = 0;          // current state
        bool pending = false;    // the result of position q-1 is the longest pid of (cold) state s: comes with its record
        for (;;) {
            if (s < n_hot) {
                if (L.q >= L.s1) break;
#pragma unroll 1
                for (int h = 0; h < kHotSteps && s < n_hot && L.q < L.s1; ++h) {
                    s = s_hot[(s << 8) | peek(L)];
                    if (s < n_hot) emit(L, L.q, s_long[s], s_strip, p.out, wide); else pending = true;
                    advance(L, stream, lo, hi);
                }
                continue;
            }
            uint32_t w[8];
            ldg_rec(p.recs + size_t(s) * 8, w);
            if (pending) { emit(L, L.q - 1, w[1], s_strip, p.out, wide); pending = false; }
            if (L.q >= L.s1) break;
            const uint32_t kind = (w[0] >> 24) & 3u, cnt = w[0] >> 26, fail = w[0] & 0xFFFFFFu;
            if (kind == 1u) {   // CHAIN: states s+1 .. s+cnt, one byte each
                uint64_t labels = uint64_t(w[2]) | (uint64_t(w[3]) << 32);
                uint64_t l0 = uint64_t(w[4]) | (uint64_t(w[5]) << 32), l1 = uint64_t(w[6]) | (uint64_t(w[7]) << 32);
                uint32_t j = 0;
                while (j < cnt && L.q < L.s1 && peek(L) == (uint32_t(labels) & 0xFFu)) {
                    emit(L, L.q, uint32_t(l0) & 0xFFFFu, s_strip, p.out, wide);
                    advance(L, stream, lo, hi);
                    labels >>= 8;
                    l0 = (l0 >> 16) | (l1 << 48); l1 >>= 16;
                    ++j;
                }
                if (j == 0 && L.q < L.s1) s = fail;   // the only child does not match: failure transition, byte not consumed
                else s += j;                           // inside (or at the end of) the run; a mismatch there is found at j == 0 next time
            } else if (kind == 0u) {   // BRANCH
                const uint32_t c = peek(L);
                uint32_t next = 0xFFFFFFFFu;
#pragma unroll
                for (int k = 0; k < 6; ++k)
                    if (uint32_t(k) < cnt && (w[2 + k] & 0xFFu) == c) next = w[2 + k] >> 8;
                if (next != 0xFFFFFFFFu) { s = next; pending = true; advance(L, stream, lo, hi); }
                else s = fail;
            } else {   // DENSE: a complete row, one more (dependent) fetch
                s = __ldg(p.dense_rows + ((size_t(w[2]) << 8) | peek(L)));
                if (s < n_hot) emit(L, L.q, s_long[s], s_strip, p.out, wide); else pending = true;
                advance(L, stream, lo, hi);
            }
        }
        // ragged end of the stream: the last strip is not full
        const uint32_t rest = uint32_t(L.s1 - L.s0) & 15u;
        if (rest) {
            const int64_t base = L.s1 - rest;
            for (uint32_t k = 0; k < rest; ++k) p.out[base + k] = *reinterpret_cast<const uint16_t*>(s_strip + 2 * k);
        }
    }
}

}  // namespace

size_t deep_smem_bytes(uint32_t n_hot) {
    return (size_t(n_hot) << 9) + size_t((n_hot + 7) & ~7u) * 2 + size_t(kThreads) * kOutStride;
}

cudaError_t deep_scan_launch(const DeepParams& p_in, int n_sms, cudaStream_t st, uint64_t* launches) {
    DeepParams p = p_in;
    if (p.n == 0) return cudaSuccess;
    const size_t smem = deep_smem_bytes(p.n_hot);
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    p.wide = ((reinterpret_cast<uintptr_t>(p.out)) & 31) == 0;
    // segments: about 4 KiB each (the warm-up is max_pat_len-1 bytes), cut so that the persistent grid's lanes get the
    // same number of them; multiples of 16 bytes (result strips)
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
