// kr_scan.cu -- the randomized Karp-Rabin variant for sm_100a ("mpbg / bgps / kmprt style").
//
// Reference ideas kept: fingerprints fp(s) = sum s[i] r^i mod p with p = 2^31-1
// (Core/src/Fingerprint.c:29-42, Core/src/mpbg.c:83); the sliding identities of
// Core/src/Fingerprint.h:66-89 (a window fingerprint from two prefix fingerprints and a power of
// r); filtering by doubling stage lengths before the full-length check (Core/src/bgps.c:215-249);
// patterns of <= 8 bytes matched exactly (Core/src/bgps.c:459-464, bgps.h:36).
// Reference behaviour NOT kept: the O(#patterns) loop per byte (Core/src/mpbg.c:132-145), the
// unseeded r (bgps.c:469-475) and the bugs that stop it reporting any pattern > 8 bytes.
//
// Per tile of 8192 positions (+352-byte halo) a CTA
//   1. stages the bytes with one bulk async copy,
//   2. builds the prefix fingerprints PHI(x) = sum_{t<=x} s[t] r^t mod p with a block-wide modular
//      prefix sum (per-thread serial part, warp-shuffle scan, cross-warp scan),
//   3. for every position forms the fingerprint of the last 8 bytes, (PHI(x)-PHI(x-8)) r^-(x-7),
//      tests it against a 64 KiB two-hash Bloom bitmap in shared memory, and only on a hit probes the
//      open-addressing table of 8-byte-suffix fingerprints in global memory and verifies the
//      candidates stage by stage (16, 32, ... bytes, then the full length), longest first.
// A pattern > 8 bytes is reported iff ALL its stage fingerprints agree; a false positive needs a
// simultaneous collision in every stage.
#include "kr_scan.cuh"
#include "pm_dev.cuh"

namespace pm {
namespace {

constexpr int kThreads = 512;                                  // two CTAs per SM: one computes while the other waits at a barrier
constexpr int kSpan = kHalo + kKrTile;                        // staged bytes
constexpr int kElems = (kSpan + kThreads - 1) / kThreads;     // elements per thread in the prefix sum (9)
constexpr int kPosPer = kKrTile / kThreads;                   // reported positions per thread (8)
constexpr int kBloomWords = (1 << 19) / 32;

constexpr int kOffBloom = 0;                                  // 65536
constexpr int kOffPhi = kOffBloom + kBloomWords * 4;          // (kSpan + 1) x u32
constexpr int kOffBytes = ((kOffPhi + (kSpan + 1) * 4 + 15) / 16) * 16;
constexpr int kOffWarp = kOffBytes + kElems * kThreads;       // padded so every thread may read its 9 bytes
constexpr int kOffBar = kOffWarp + 32 * 4;
constexpr int kSmem = kOffBar + 16;
static_assert(kSmem <= 227 * 1024, "shared memory budget");

struct KrParams {
    KrDevTables t;
    const uint8_t* stream;
    uint64_t n, hist_valid;
    uint16_t* out;
    const uint32_t* pat_len;  // by canonical index
    uint32_t n_tiles;
};

__device__ __forceinline__ uint32_t win_fp(const uint32_t* phi, const uint32_t* __restrict__ rinvpow, int x, int l) {
    // fingerprint of the l bytes ending at tile-relative index x (Fingerprint.h:66-73, calc_fp_suffix)
    return kr_mulmod(kr_submod(phi[x + 1], phi[x + 1 - l]), __ldg(rinvpow + (x + 1 - l)));
}

__device__ __noinline__ uint32_t kr_verify(const KrDevTables& t, const uint32_t* phi, int x, uint32_t f8, uint64_t avail) {
    uint32_t h = uint32_t(splitmix64_d(f8)) & t.bucket_mask;
    for (;;) {
        const uint32_t sf = __ldg(t.slot_fp + h);
        if (sf == 0xFFFFFFFFu) return 0;
        if (sf == f8) break;
        h = (h + 1) & t.bucket_mask;
    }
    const uint32_t b = __ldg(t.slot_begin + h), c = __ldg(t.slot_count + h);
    for (uint32_t k = b; k < b + c; ++k) {  // longest candidate first
        const uint32_t len = __ldg(t.cand_len + k);
        if (uint64_t(len) > avail) continue;
        const uint32_t* sfp = t.stage_fp + __ldg(t.cand_stage_off + k);
        bool ok = true;
        for (uint32_t l = 16; l <= len && ok; l <<= 1, ++sfp) ok = win_fp(phi, t.rinvpow, x, int(l)) == __ldg(sfp);
        if (ok && win_fp(phi, t.rinvpow, x, int(len)) == __ldg(sfp)) return __ldg(t.cand_pid + k);
    }
    return 0;
}

__global__ void __launch_bounds__(kThreads, 2) kr_scan_kernel(const KrParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* s_bloom = reinterpret_cast<uint32_t*>(smem + kOffBloom);
    uint32_t* s_phi = reinterpret_cast<uint32_t*>(smem + kOffPhi);
    uint8_t* s_bytes = smem + kOffBytes;
    uint32_t* s_warp = reinterpret_cast<uint32_t*>(smem + kOffWarp);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kOffBar);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool have_halo = p.hist_valid >= uint64_t(kHalo);

    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    for (int i = tid; i < kBloomWords; i += kThreads) s_bloom[i] = __ldg(p.t.bloom + i);
    for (int i = tid; i < (kElems * kThreads) / 4; i += kThreads) reinterpret_cast<uint32_t*>(s_bytes)[i] = 0;
    fence_proxy_async();
    __syncthreads();

    uint32_t it = 0;
    for (uint64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x, ++it) {
        const uint64_t s0 = t * uint64_t(kKrTile);
        const uint32_t len = uint32_t(min(uint64_t(kKrTile), p.n - s0));
        const bool halo = (t > 0) || have_halo;
        const uint32_t body = len & ~15u;
        if (tid == 0) {
            const uint32_t bytes = body + (halo ? kHalo : 0);
            mbar_arrive_expect_tx(bar, bytes);
            if (bytes) bulk_g2s(s_bytes + (halo ? 0 : kHalo), p.stream + s0 - (halo ? kHalo : 0), bytes, bar);
        }
        // exact results of this tile's positions -> "longest pattern of <= 8 bytes" (two dependent global loads):
        // issued now, consumed after the prefix sum, so their latency hides behind the tile load and the scan
        uint32_t short_pid[kPosPer];
#pragma unroll
        for (int k = 0; k < kPosPer; ++k) {
            const uint32_t q = uint32_t(k) * kThreads + tid;
            short_pid[k] = q < len ? uint32_t(p.out[s0 + q]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < kPosPer; ++k) short_pid[k] = __ldg(p.t.short_of + short_pid[k]);
        if (!halo) for (int i = tid; i < kHalo; i += kThreads) s_bytes[i] = 0;
        if (uint32_t(tid) < (len & 15u)) s_bytes[kHalo + body + tid] = p.stream[s0 + body + tid];
        mbar_wait(bar, it & 1);
        __syncthreads();

        // ---- block-wide modular prefix sum of s[x] * r^x ----
        // pass 1, interleaved mapping (coalesced reads of the power table): the terms go to s_phi[x + 1]
        for (int x = tid; x < kSpan; x += kThreads) s_phi[x + 1] = kr_mulmod(s_bytes[x], __ldg(p.t.rpow + x));
        __syncthreads();
        // pass 2, blocked mapping: every thread sums its kElems consecutive terms (stride-kElems reads, kElems odd:
        // conflict-free), then the thread totals are scanned across the block
        uint32_t loc[kElems];
        uint32_t acc = 0;
        const int x0 = tid * kElems;
#pragma unroll
        for (int k = 0; k < kElems; ++k) {
            const int x = x0 + k;
            const uint32_t term = x < kSpan ? s_phi[x + 1] : 0u;
            acc = kr_addmod(acc, term);
            loc[k] = acc;
        }
        uint32_t inc = acc;  // inclusive warp scan of the thread totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc = kr_addmod(inc, y);
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const uint32_t w = s_warp[lane];
            uint32_t z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, z, o);
                if (lane >= o) z = kr_addmod(z, y);
            }
            s_warp[lane] = kr_submod(z, w);  // exclusive offset of each warp
        }
        __syncthreads();
        const uint32_t offs = kr_addmod(s_warp[wid], kr_submod(inc, acc));
#pragma unroll
        for (int k = 0; k < kElems; ++k) {
            const int x = x0 + k;
            if (x < kSpan) s_phi[x + 1] = kr_addmod(offs, loc[k]);
        }
        if (tid == 0) s_phi[0] = 0;
        __syncthreads();

        // ---- per position: stage-8 fingerprint, Bloom test, verification ----
        // position k * kThreads + tid: consecutive lanes take consecutive positions, so the PHI reads are
        // conflict-free (a blocked mapping would be an 8- or 16-way bank conflict) and the result stores coalesce
#pragma unroll
        for (int k = 0; k < kPosPer; ++k) {
            const uint32_t q = uint32_t(k) * kThreads + tid;
            if (q < len) {
                const uint64_t i = s0 + q;
                uint32_t r = short_pid[k];
                const uint64_t avail = i + p.hist_valid + 1;  // bytes of the stream up to and incl. c[i]
                if (avail >= 9) {
                    const int x = kHalo + int(q);
                    const uint32_t f8 = win_fp(s_phi, p.t.rinvpow, x, 8);
                    const uint32_t bit = f8 & ((1u << 19) - 1);
                    const uint32_t bit2 = (f8 >> 12) & ((1u << 19) - 1);
                    if (((s_bloom[bit >> 5] >> (bit & 31)) & 1u) && ((s_bloom[bit2 >> 5] >> (bit2 & 31)) & 1u)) {
                        const uint32_t hit = kr_verify(p.t, s_phi, x, f8, avail);
                        if (hit) r = hit;
                    }
                }
                p.out[i] = uint16_t(r);
            }
        }
        __syncthreads();  // s_bytes / s_phi are rewritten by the next tile
    }
}

}  // namespace

cudaError_t kr_upload_tables(const Dict& d, KrDevTables* t, size_t* bytes) {
    const KrTables& k = d.kr;
    *t = KrDevTables();
    t->r = uint32_t(k.r);
    t->bucket_mask = (1u << k.bucket_bits) - 1;
    auto up = [&](const void* src, size_t n, void** dst) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, n ? n : 4);
        if (e != cudaSuccess) return e;
        *bytes += n;
        return n ? cudaMemcpy(*dst, src, n, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    const int span = kHalo + kKrTile;
    std::vector<uint32_t> rpow(span + 1), rinvpow(span + 1);
    const uint64_t rinv = kr_inv(k.r);
    uint64_t a = 1, b = 1;
    for (int i = 0; i <= span; ++i) { rpow[i] = uint32_t(a); rinvpow[i] = uint32_t(b); a = kr_mul(a, k.r); b = kr_mul(b, rinv); }
    std::vector<uint16_t> short_of(d.pats.size() + 1, 0);
    for (size_t i = 0; i < d.pats.size(); ++i) {
        uint32_t q = uint32_t(i + 1);
        while (q && d.pats[q - 1].len > 8) q = d.pats[q - 1].parent;
        short_of[i + 1] = uint16_t(q);
    }
    cudaError_t e;
#define UP(vec, field) \
    if ((e = up(vec.data(), vec.size() * sizeof(vec[0]), reinterpret_cast<void**>(&t->field))) != cudaSuccess) return e;
    UP(k.slot_fp, slot_fp) UP(k.slot_begin, slot_begin) UP(k.slot_count, slot_count) UP(k.cand_pid, cand_pid)
    UP(k.cand_len, cand_len) UP(k.cand_stage_off, cand_stage_off) UP(k.stage_fp, stage_fp) UP(k.bloom, bloom)
    UP(rpow, rpow) UP(rinvpow, rinvpow) UP(short_of, short_of)
#undef UP
    return cudaSuccess;
}

void kr_free_tables(KrDevTables* t) {
    void* ptrs[] = {t->slot_fp, t->slot_begin, t->slot_count, t->cand_pid, t->cand_len, t->cand_stage_off,
                    t->stage_fp, t->bloom, t->rpow, t->rinvpow, t->short_of};
    for (void* p : ptrs) if (p) cudaFree(p);
    *t = KrDevTables();
}

cudaError_t kr_scan_launch(const KrDevTables& t, const uint8_t* stream, uint64_t n, uint64_t hist_valid, uint16_t* out,
                           const PatTables& pt, int n_sms, cudaStream_t st, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    KrParams p{};
    p.t = t; p.stream = stream; p.n = n; p.hist_valid = hist_valid; p.out = out; p.pat_len = pt.len;
    p.n_tiles = uint32_t((n + kKrTile - 1) / kKrTile);
    cudaError_t e = cudaFuncSetAttribute(kr_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    const uint32_t grid = p.n_tiles < uint32_t(2 * n_sms) ? p.n_tiles : uint32_t(2 * n_sms);
    kr_scan_kernel<<<grid, kThreads, kSmem, st>>>(p);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pm
