// dfa_scan.cu -- exact dictionary scan by per-thread forward walks of the flat Aho-Corasick DFA.
//
// Semantics identical to sfx_scan.cu: out[i] = pid of the longest pattern that is a suffix of
// stream[..i] (ac_read_char, Core/src/mpac.c:304-319).  The automaton is the reference's goto +
// failure + nearest-output-link machine (Core/src/mpac.c:147-210) completed on the host to a full
// DFA, so one table lookup per byte replaces the failure-link loop; worst-case work per byte is
// constant whatever the input (the backward scan's is O(match depth)).
//
// The stream is cut into per-thread segments of kDfaSeg bytes; a thread first walks the
// max_pat_len-1 bytes before its segment from the root without reporting (after that many bytes the
// state reports exactly what a continuous scan reports -- SURVEY Q8) and then reports its own bytes.
#include "dfa_scan.cuh"
#include "pm_dev.cuh"

namespace pm {
namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) dfa_scan_kernel(const DfaParams p) {
    const uint64_t seg = uint64_t(blockIdx.x) * kThreads + threadIdx.x;
    const uint64_t s0 = seg * uint64_t(kDfaSeg);
    if (s0 >= p.n) return;
    const uint64_t s1 = min(p.n, s0 + uint64_t(kDfaSeg));
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

}  // namespace

cudaError_t dfa_scan_launch(const DfaParams& p, cudaStream_t st, uint64_t* launches) {
    if (p.n == 0) return cudaSuccess;
    const uint64_t segs = (p.n + kDfaSeg - 1) / kDfaSeg;
    const uint32_t grid = uint32_t((segs + kThreads - 1) / kThreads);
    dfa_scan_kernel<<<grid, kThreads, 0, st>>>(p);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pm
