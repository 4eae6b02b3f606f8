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
// Per tile of 512 positions (+32 bytes before it) a WARP
//   1. stages the bytes in its private part of shared memory (coalesced 16-byte loads),
//   2. builds the prefix fingerprints PHI(x) = sum_{t<x} s[t] r^t mod p with a warp-wide modular prefix sum
//      (17 serial terms per lane, warp-shuffle scan of the lane totals, offsets added back),
//   3. for every position forms the fingerprint of the last 8 bytes, (PHI(x+1)-PHI(x-7)) r^-(x-7),
//      tests it against a 64 KiB four-hash Bloom bitmap in shared memory, and only on a hit probes the
//      open-addressing table of 8-byte-suffix fingerprints in global memory and verifies the
//      candidates stage by stage (16, 32, ... bytes, then the full length), longest first; the longer window
//      fingerprints are extended from the 8-byte one, one stream byte and one multiplication at a time.
// A pattern > 8 bytes is reported iff ALL its stage fingerprints agree; a false positive needs a
// simultaneous collision in every stage.
#include "kr_scan.cuh"
#include "pm_dev.cuh"

namespace pm {
namespace {

// Every WARP owns a tile: no block-wide barrier anywhere after start-up, so the warps of an SM hide each other's
// latencies (the first version built the prefix sum per CTA over 8 KiB tiles and spent 63% of its issue slots
// waiting at the five barriers per tile).  The prefix sum only has to reach 8 bytes back (the stage every
// position is tested for); the longer stages of the rare candidates are extended byte by byte from the stream.
constexpr int kWarps = 32;
constexpr int kThreads = kWarps * 32;
constexpr int kWT = kKrTile;                                   // positions per warp tile (512)
constexpr int kLead = 32;                                     // staged bytes before the tile (>= 7, multiple of 16)
constexpr int kSpan = kLead + kWT;                            // staged bytes per tile (544)
constexpr int kPerLane = kSpan / 32;                          // elements per lane in the prefix sum (17: odd => conflict-free)
static_assert(kSpan % 32 == 0 && (kPerLane & 1) == 1 && kSpan % 16 == 0 && kLead >= 7, "tile geometry");
constexpr int kBloomWords = (1 << 19) / 32;

constexpr int kPowBytes = ((kSpan + 1) * 4 + 15) / 16 * 16;   // one power table
constexpr int kOffBloom = 0;                                  // 65536
constexpr int kOffRpow = kOffBloom + kBloomWords * 4;
constexpr int kOffRinv = kOffRpow + kPowBytes;
constexpr int kOffLong = kOffRinv + kPowBytes;                // one bit per pid: the pattern has more than 8 bytes (8 KiB for 65,536 pids)
constexpr int kLongWords = 65536 / 32;
constexpr int kOffWarp = kOffLong + kLongWords * 4;           // per warp: phi[(kSpan + 1)] u32, then the staged bytes
constexpr int kWarpBytes = kPowBytes + kSpan + 16;             // + the warp's mbarrier (bulk-copy staging)
constexpr int kSmem = kOffWarp + kWarps * kWarpBytes;
static_assert(kWarpBytes % 16 == 0 && kSmem <= 227 * 1024, "shared memory budget");

struct KrParams {
    KrDevTables t;
    const uint8_t* stream;
    uint64_t n, hist_valid;
    uint16_t* out;
    const uint32_t* pat_len;  // by canonical index
    uint32_t n_tiles;
    uint32_t bulk;            // stage full tiles with one bulk async copy (TMA engine) per tile
};

// A position whose 8-byte fingerprint passed the Bloom test: probe the table of 8-byte-suffix fingerprints and
// check the candidates, longest first.  The window fingerprint is extended backwards one byte at a time from the
// stream, fp_{m+1} = s[i-m] + r * fp_m (the sliding identity of Fingerprint.h:75-89 turned around), and compared
// at every doubling stage and at the full length.  `ci` points at c[i].
__device__ __noinline__ uint32_t kr_verify(const KrDevTables& t, const uint8_t* __restrict__ ci, uint32_t f8, uint64_t avail) {
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
        uint32_t fp = f8, m = 8;
        bool ok = true;
        for (uint32_t l = 16; l <= len && ok; l <<= 1, ++sfp) {
            for (; m < l; ++m) fp = kr_addmod(kr_mulmod(fp, t.r), uint32_t(__ldg(ci - m)));
            ok = fp == __ldg(sfp);
        }
        if (!ok) continue;
        for (; m < len; ++m) fp = kr_addmod(kr_mulmod(fp, t.r), uint32_t(__ldg(ci - m)));
        if (fp == __ldg(sfp)) return __ldg(t.cand_pid + k);
    }
    return 0;
}

// s[x] * r^x mod p for a byte s[x] < 256: the product is < 2^39, one fold and one conditional subtract
__device__ __forceinline__ uint32_t kr_mul_byte(uint32_t byte, uint32_t pw) {
    const uint64_t x = uint64_t(byte) * pw;
    const uint32_t v = uint32_t(x & 0x7FFFFFFFu) + uint32_t(x >> 31);
    return v >= 0x7FFFFFFFu ? v - 0x7FFFFFFFu : v;
}

__global__ void __launch_bounds__(kThreads, 1) kr_scan_kernel(const KrParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* s_bloom = reinterpret_cast<uint32_t*>(smem + kOffBloom);
    uint32_t* s_rpow = reinterpret_cast<uint32_t*>(smem + kOffRpow);
    uint32_t* s_rinv = reinterpret_cast<uint32_t*>(smem + kOffRinv);
    uint32_t* s_longbits = reinterpret_cast<uint32_t*>(smem + kOffLong);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t* s_phi = reinterpret_cast<uint32_t*>(smem + kOffWarp + wid * kWarpBytes);
    uint8_t* s_bytes = reinterpret_cast<uint8_t*>(s_phi) + kPowBytes;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bytes + kSpan);     // this warp's mbarrier
    uint32_t bar_parity = 0;
    if (lane == 0) { mbar_init(s_bar, 1); fence_mbar_init(); }

    for (int i = tid; i < kBloomWords; i += kThreads) s_bloom[i] = __ldg(p.t.bloom + i);
    for (int i = tid; i <= kSpan; i += kThreads) { s_rpow[i] = __ldg(p.t.rpow + i); s_rinv[i] = __ldg(p.t.rinvpow + i); }
    for (int i = tid; i < kLongWords; i += kThreads) s_longbits[i] = __ldg(p.t.long_bits + i);
    __syncthreads();  // the only CTA-wide barrier

    const uint64_t gw = uint64_t(blockIdx.x) * kWarps + wid, G = uint64_t(gridDim.x) * kWarps;
    for (uint64_t t = gw; t < p.n_tiles; t += G) {
        const uint64_t s0 = t * uint64_t(kWT);
        const uint32_t len = uint32_t(min(uint64_t(kWT), p.n - s0));
        // ---- stage [s0 - kLead, s0 + kWT): 16-byte pieces, zeros where the stream has no byte ----
        const int64_t lo = -int64_t(p.hist_valid), hi = int64_t(p.n);   // readable range, relative to p.stream
        // A tile that is readable as a whole (all but the first and the last of a scan) is staged by ONE bulk async copy:
        // lane 0 arms the warp's mbarrier with the byte count and hands the 544-byte copy to the TMA engine
        // (cp.async.bulk, SASS UBLKCP); the warp waits on the barrier's phase.  The fence orders the previous tile's
        // generic-proxy reads of the buffer before the async-proxy write.
        const bool whole = p.bulk && int64_t(s0) - kLead >= lo && int64_t(s0) + kWT <= hi;
        if (whole) {
            if (lane == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(s_bar, kSpan);
                bulk_g2s(s_bytes, p.stream + (int64_t(s0) - kLead), kSpan, s_bar);
            }
            mbar_wait(s_bar, bar_parity);
            bar_parity ^= 1u;
        }
#pragma unroll
        for (int k = 0; !whole && k < (kSpan / 16 + 31) / 32; ++k) {
            const int j = k * 32 + lane;
            if (j < kSpan / 16) {
                const int64_t g = int64_t(s0) - kLead + 16 * j;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (g >= lo && g + 16 <= hi) {
                    v = __ldg(reinterpret_cast<const uint4*>(p.stream + g));
                } else if (g + 16 > lo && g < hi) {
                    uint32_t w[4] = {0, 0, 0, 0};
                    for (int b = 0; b < 16; ++b)
                        if (g + b >= lo && g + b < hi) w[b >> 2] |= uint32_t(p.stream[g + b]) << (8 * (b & 3));
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
                reinterpret_cast<uint4*>(s_bytes)[j] = v;
            }
        }
        __syncwarp();

        // ---- warp-wide modular prefix sum of s[x] * r^x: PHI(x) = sum_{t < x} term(t), phi[0] = 0 ----
        // blocked mapping: lane l owns x in [17 l, 17 l + 17); the stride (17 words) is odd: conflict-free
        const int x0 = lane * kPerLane;
        uint32_t acc = 0;
#pragma unroll
        for (int k = 0; k < kPerLane; ++k) {
            acc = kr_addmod(acc, kr_mul_byte(s_bytes[x0 + k], s_rpow[x0 + k]));
            s_phi[x0 + k + 1] = acc;               // inclusive within the lane
        }
        uint32_t inc = acc;                         // inclusive warp scan of the lane totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc = kr_addmod(inc, y);
        }
        const uint32_t offs = kr_submod(inc, acc);  // exclusive offset of this lane
        if (lane == 0) s_phi[0] = 0;
#pragma unroll
        for (int k = 0; k < kPerLane; ++k) s_phi[x0 + k + 1] = kr_addmod(s_phi[x0 + k + 1], offs);
        __syncwarp();

        // ---- per position: stage-8 fingerprint, Bloom test, verification ----
        // position k * 32 + lane: consecutive lanes take consecutive positions (conflict-free PHI reads, coalesced
        // result traffic).  The exact result of the position is read first: it supplies, through short_of[], the
        // longest pattern of <= 8 bytes ending there (those are matched exactly).
#pragma unroll 1
        for (int kb = 0; kb < kWT / 32; kb += 4) {
            uint32_t sp[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t q = uint32_t(kb + k) * 32 + lane;
                sp[k] = q < len ? uint32_t(p.out[s0 + q]) : 0u;
            }
            // short_of[] is the identity for patterns of <= 8 bytes -- 99.9 % of the matches of binary traffic; only the
            // others pay for the gather from the 111 KB table (as a gather for every position it was 0.8 L1 sectors
            // per stream byte, most of this kernel's global-load traffic)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((s_longbits[sp[k] >> 5] >> (sp[k] & 31u)) & 1u) sp[k] = __ldg(p.t.short_of + sp[k]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t q = uint32_t(kb + k) * 32 + lane;
                if (q < len) {
                    const uint64_t i = s0 + q;
                    uint32_t r = sp[k];
                    const uint64_t avail = i + p.hist_valid + 1;  // bytes of the stream up to and incl. c[i]
                    if (avail >= 9) {
                        const int x = kLead + int(q);
                        const uint32_t f8 = kr_mulmod(kr_submod(s_phi[x + 1], s_phi[x - 7]), s_rinv[x - 7]);
                        bool pass = true;
#pragma unroll
                        for (int hk = 0; hk < kKrBloomHashes; ++hk) {
                            if (pass) {
                                const uint32_t bit = kr_bloom_bit(f8, hk);
                                pass = ((s_bloom[bit >> 5] >> (bit & 31)) & 1u) != 0;
                            }
                        }
                        if (pass) {
                            const uint32_t hit = kr_verify(p.t, p.stream + i, f8, avail);
                            if (hit) r = hit;
                        }
                    }
                    p.out[i] = uint16_t(r);
                }
            }
        }
        __syncwarp();  // s_bytes / s_phi are rewritten by the next tile
    }
}

// The reference's MPBG as shipped (Core/src/mpbg.c:132-145): its fingerprint stages never fire, so what it
// reports at a position is the longest pattern of <= 8 bytes ending there -- the patterns its exact KMP path handles
// (bgps.c:459-464).  That is short_of[] of the exact answer: out[i] = short_of[out[i]], eight positions per thread.
__global__ void __launch_bounds__(256) short_only_kernel(uint16_t* __restrict__ out, uint64_t n, const uint16_t* __restrict__ short_of) {
    const uint64_t n8 = n / 8;
    uint4* o8 = reinterpret_cast<uint4*>(out);
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 v = o8[i];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t lo = w[k] & 0xFFFFu, hi = w[k] >> 16;
            w[k] = (lo ? uint32_t(__ldg(short_of + lo)) : 0u) | ((hi ? uint32_t(__ldg(short_of + hi)) : 0u) << 16);
        }
        o8[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) { const uint64_t i = n8 * 8 + threadIdx.x; out[i] = __ldg(short_of + out[i]); }
}

}  // namespace

cudaError_t kr_short_only_launch(const KrDevTables& t, uint16_t* out, uint64_t n, int n_sms, cudaStream_t st, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    const uint64_t want = (n / 8 + 255) / 256 + 1;
    short_only_kernel<<<uint32_t(want < uint64_t(n_sms) * 8 ? want : uint64_t(n_sms) * 8), 256, 0, st>>>(out, n, t.short_of);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t kr_upload_tables(const Dict& d, const KrTables& k, KrDevTables* t, size_t* bytes) {
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
    std::vector<uint32_t> long_bits(65536 / 32, 0);
    for (size_t i = 0; i < d.pats.size() && i + 1 < 65536; ++i)
        if (d.pats[i].len > 8) long_bits[(i + 1) >> 5] |= 1u << ((i + 1) & 31);
    UP(rpow, rpow) UP(rinvpow, rinvpow) UP(short_of, short_of) UP(long_bits, long_bits)
#undef UP
    return cudaSuccess;
}

void kr_free_tables(KrDevTables* t) {
    void* ptrs[] = {t->slot_fp, t->slot_begin, t->slot_count, t->cand_pid, t->cand_len, t->cand_stage_off,
                    t->stage_fp, t->bloom, t->rpow, t->rinvpow, t->short_of, t->long_bits};
    for (void* p : ptrs) if (p) cudaFree(p);
    *t = KrDevTables();
}

cudaError_t kr_scan_launch(const KrDevTables& t, const uint8_t* stream, uint64_t n, uint64_t hist_valid, uint16_t* out,
                           const PatTables& pt, int n_sms, cudaStream_t st, uint64_t* launches, bool bulk) {
    if (n == 0) return cudaSuccess;
    KrParams p{};
    p.t = t; p.stream = stream; p.n = n; p.hist_valid = hist_valid; p.out = out; p.pat_len = pt.len;
    p.n_tiles = uint32_t((n + kKrTile - 1) / kKrTile);
    p.bulk = (bulk && (reinterpret_cast<uintptr_t>(stream) & 15) == 0) ? 1u : 0u;
    cudaError_t e = cudaFuncSetAttribute(kr_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    const uint32_t ctas = (p.n_tiles + kWarps - 1) / kWarps;
    const uint32_t grid = ctas < uint32_t(n_sms) ? ctas : uint32_t(n_sms);
    kr_scan_kernel<<<grid, kThreads, kSmem, st>>>(p);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace pm
