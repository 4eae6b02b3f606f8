// aux_kernels.cu -- kernels around the scan: seeded stream generators in HBM, the device-side
// reduction of a dense result (counts + order-independent digests), and compaction of a dense
// result into position-sorted (pos, pid) records with optional PatternsTree-ancestor expansion
// (Core/src/PatternsTree.c:485-494 semantics: all matches at a position = longest + its ancestors).
#include "aux_kernels.cuh"
#include "pm_dev.cuh"

namespace pm {
namespace {

// ------------------------------------------------------------------------------------------------
// Stream generators.  Definitions are SURVEY.md 8d / oracle/pm_oracle.c; every byte is a pure
// function of its absolute offset so that shards are generated independently.
// ------------------------------------------------------------------------------------------------
__global__ void gen_uniform_kernel(uint64_t off, uint64_t n, uint8_t* __restrict__ dst) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;  // 8-byte word index
    if (i * 8 >= n) return;
    reinterpret_cast<uint64_t*>(dst)[i] = splitmix64_d(0x5EED0001ull + (off >> 3) + i);
}

__global__ void gen_ab_kernel(uint64_t off, uint64_t n, uint8_t* __restrict__ dst) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i * 8 >= n) return;
    const uint64_t w = splitmix64_d(0xADE50004ull + (off >> 3) + i);
    uint64_t o = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) o |= uint64_t(((w >> (8 * k)) & 0xFF) < 192 ? 'a' : 'b') << (8 * k);
    reinterpret_cast<uint64_t*>(dst)[i] = o;
}

__global__ void gen_ascii_kernel(uint64_t off, uint64_t n, uint8_t* __restrict__ dst) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i * 8 >= n) return;
    const uint64_t w = splitmix64_d(0xA5C11006ull + (off >> 3) + i);
    uint64_t o = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) o |= uint64_t(0x20 + ((((w >> (8 * k)) & 0xFF) * 95) >> 8)) << (8 * k);
    reinterpret_cast<uint64_t*>(dst)[i] = o;
}

// one planted pattern per 4096-byte block, on top of the uniform background
__global__ void gen_plant_kernel(uint64_t off, uint64_t n, uint8_t* __restrict__ dst, PatTables t) {
    const uint64_t bi = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (bi * 4096 >= n || t.n_patterns == 0) return;
    const uint64_t b = (off >> 12) + bi;
    const uint64_t h = splitmix64_d(0xD1C70002ull + b);
    const uint32_t k = uint32_t(h % t.n_patterns);
    const uint32_t len = t.len[k];
    if (len > 4096) return;
    const uint32_t o = uint32_t((h >> 32) % (4096 - len + 1));
    const uint8_t* src = t.bytes + t.off[k];
    uint8_t* d = dst + bi * 4096 + o;
    for (uint32_t u = 0; u < len; ++u) d[u] = src[u];
}

// every 4096-byte block = concatenated random-length prefixes of random patterns
__global__ void gen_almost_kernel(uint64_t off, uint64_t n, uint8_t* __restrict__ dst, PatTables t) {
    const uint64_t bi = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (bi * 4096 >= n) return;
    uint8_t* d = dst + bi * 4096;
    if (t.n_patterns == 0) { for (int u = 0; u < 4096; ++u) d[u] = 0; return; }
    const uint64_t b = (off >> 12) + bi;
    uint32_t fill = 0;
    for (uint64_t piece = 0; fill < 4096; ++piece) {
        const uint64_t h = splitmix64_d(0xA1A50005ull + (b << 12) + piece);
        const uint32_t k = uint32_t(h % t.n_patterns);
        const uint32_t plen = 1 + uint32_t((h >> 32) % t.len[k]);
        const uint8_t* src = t.bytes + t.off[k];
        for (uint32_t u = 0; u < plen && fill < 4096; ++u, ++fill) d[fill] = src[u];
    }
}

// ------------------------------------------------------------------------------------------------
// Summary of a dense result
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// One persistent CTA per SM.  What a position costs is the gather of its pattern's digest key: 8 bytes from a 445 KB
// table (merged dictionary) that no L1 holds -- as plain global loads ~25 L1 wavefronts per warp and position: 58 ms for
// the 16 GiB result, 9 % of the HBM rate.  The positions of any traffic are dominated by the SHORTEST patterns (the one-
// and two-byte patterns answer 89 % of the matching positions of binary traffic), so their keys live in shared memory
// behind a 2-byte pid -> slot map: two shared-memory gathers (~3.5 + ~7 wavefronts) instead of the global one; everything
// else takes the general path below.
constexpr int kSumThreads = 1024;

__global__ void __launch_bounds__(kSumThreads, 1) summarize_kernel(const uint16_t* __restrict__ out, uint64_t n, uint64_t pos_base,
                                                                  PatTables t, unsigned long long* __restrict__ acc) {
    extern __shared__ __align__(16) uint8_t sum_smem[];
    uint64_t* s_own = reinterpret_cast<uint64_t*>(sum_smem);
    uint64_t* s_anc = s_own + t.n_hot;
    uint16_t* s_map = reinterpret_cast<uint16_t*>(s_anc + t.n_hot);
    for (uint32_t i = threadIdx.x; i < t.n_hot; i += kSumThreads) { s_own[i] = __ldg(t.hot_own + i); s_anc[i] = __ldg(t.hot_anc + i); }
    for (uint32_t i = threadIdx.x; i <= t.n_patterns; i += kSumThreads) s_map[i] = i ? __ldg(t.hot_map + i) : uint16_t(0x4000u);   // 0x4000: no pattern (slot 0 is read, its digest masked out)
    __syncthreads();
    // Accumulators: h0 = digest sum of the longest matches; hx = digest sum of the ANCESTORS only (h_all = h0 + hx);
    // positions; extra = ancestor count (matches = positions + extra).
    uint64_t h0 = 0, hx = 0;
    uint32_t positions = 0, extra = 0;
    // the general path: a pattern outside the shared-memory table (0.1 % of the matches of binary traffic)
    auto slow = [&](uint32_t pid, uint64_t pos) {
        h0 += splitmix64_d(pos ^ __ldg(t.pidhash + pid));
        const uint32_t nanc = __ldg(t.chain + pid);
        extra += nanc;
        if (nanc) {
            const uint32_t b = __ldg(t.anc_off + pid) + 1;
            for (uint32_t k = b; k < b + nanc; ++k) hx += splitmix64_d(pos ^ __ldg(t.pidhash + __ldg(t.anc_list + k)));
        }
    };
    // One position, straight-line for the common cases: s_map[0] is the "no pattern" code 0x4000, so a
    // position without a match runs the same instructions with its digest masked out -- the 29 % of such positions cost
    // less than the divergence of a branch around two 64-bit mixes did (18 of 32 lanes active, 72 instructions per position).
    auto one = [&](uint32_t pid, uint64_t pos) {
        const uint32_t m = s_map[pid];
        if (m == 0xFFFFu) { ++positions; slow(pid, pos); return; }
        const uint64_t own = splitmix64_d(pos ^ s_own[m & 0x3FFFu]);
        const bool hit = !(m & 0x4000u);
        h0 += hit ? own : 0ull;
        positions += hit ? 1u : 0u;
        if (m & 0x8000u) { hx += splitmix64_d(pos ^ s_anc[m & 0x3FFFu]); ++extra; }
    };
    // 8 positions (one 16-byte load) per thread and step; `out` is 16-byte aligned
    const uint64_t n8 = n / 8;
    const uint4* out8 = reinterpret_cast<const uint4*>(out);
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 v = __ldcs(out8 + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        if ((v.x | v.y | v.z | v.w) == 0) continue;
        const uint64_t pos0 = pos_base + i * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) one((w[k >> 1] >> (16 * (k & 1))) & 0xFFFF, pos0 + k);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) one(out[n8 * 8 + threadIdx.x], pos_base + n8 * 8 + threadIdx.x);
    __shared__ uint64_t sh[4][kSumThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t v[4] = {warp_sum(uint64_t(positions)), warp_sum(uint64_t(positions) + extra), warp_sum(h0), warp_sum(h0 + hx)};
    if (lane == 0) for (int k = 0; k < 4; ++k) sh[k][wid] = v[k];
    __syncthreads();
    if (threadIdx.x < 4) {
        uint64_t s = 0;
        for (int w = 0; w < kSumThreads / 32; ++w) s += sh[threadIdx.x][w];
        atomicAdd(&acc[threadIdx.x], (unsigned long long)s);
    }
}

// ------------------------------------------------------------------------------------------------
// pid -> caller's 64-bit pattern id (the plugin's pattern_id_t, Core/src/PatternsTree.h:104): ids[i] = table[out[i]].
// Used when the caller's result buffer is page-locked: the 8 bytes per position then leave the GPU by DMA instead of
// being produced by host threads.  Four positions per thread: one 8-byte load, two 16-byte stores.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_ids_kernel(const uint16_t* __restrict__ out, uint64_t n,
                                                        const unsigned long long* __restrict__ table,
                                                        unsigned long long* __restrict__ ids) {
    const uint64_t n4 = n / 4;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint2 v = __ldcs(reinterpret_cast<const uint2*>(out) + i);
        ulonglong2 a, b;
        a.x = __ldg(table + (v.x & 0xFFFFu)); a.y = __ldg(table + (v.x >> 16));
        b.x = __ldg(table + (v.y & 0xFFFFu)); b.y = __ldg(table + (v.y >> 16));
        __stcs(reinterpret_cast<ulonglong2*>(ids) + 2 * i, a);
        __stcs(reinterpret_cast<ulonglong2*>(ids) + 2 * i + 1, b);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) ids[n4 * 4 + threadIdx.x] = __ldg(table + out[n4 * 4 + threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// 32-bit results.  out32[i] = global pid of one part's 16-bit answer (local pid + base, 0 stays 0); unless `first`, the
// LONGER of that and what out32[i] already holds: two different patterns that end at the same position have different
// lengths, so the longest over all parts of a dictionary is the longest over the whole dictionary.  Four positions per thread.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_parts_kernel(uint32_t* __restrict__ out32, const uint16_t* __restrict__ part, uint64_t n,
                                                         uint32_t base, const uint32_t* __restrict__ glen, bool first) {
    auto one = [&](uint32_t t, uint32_t o) -> uint32_t {
        const uint32_t g = t ? t + base : 0u;
        if (first || o == 0) return g;
        if (g == 0) return o;
        return __ldg(glen + o) > __ldg(glen + g) ? o : g;
    };
    const uint64_t n4 = n / 4;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint2 v = __ldcs(reinterpret_cast<const uint2*>(part) + i);
        uint4 o = make_uint4(0, 0, 0, 0);
        if (!first) o = reinterpret_cast<const uint4*>(out32)[i];
        o.x = one(v.x & 0xFFFFu, o.x); o.y = one(v.x >> 16, o.y); o.z = one(v.y & 0xFFFFu, o.z); o.w = one(v.y >> 16, o.w);
        reinterpret_cast<uint4*>(out32)[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const uint64_t i = n4 * 4 + threadIdx.x;
        out32[i] = one(part[i], first ? 0u : out32[i]);
    }
}

// ------------------------------------------------------------------------------------------------
// Success classification of one dense result against another (Core/src/measure.c:174-190 on pids): per position
// equal -> success; the algorithm's answer is a PatternsTree ancestor of the real one -> partial success; the algorithm
// reported nothing -> false negative; anything else -> false positive.  acc[0..3] = success, partial, false_neg, false_pos.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) classify_kernel(const uint16_t* __restrict__ algo, const uint16_t* __restrict__ real,
                                                      uint64_t n, PatTables t, unsigned long long* __restrict__ acc) {
    uint32_t cnt[4] = {0, 0, 0, 0};
    auto one = [&](uint32_t a, uint32_t r) {
        if (a == r) { ++cnt[0]; return; }
        uint32_t c = r;
        while (c && c != a) c = __ldg(t.parent + c);   // is_pattern_suffix(algo, real), PatternsTree.c:485-494
        if (a && c == a) ++cnt[1];
        else if (!a) ++cnt[2];
        else ++cnt[3];
    };
    const uint64_t n8 = n / 8;
    const uint4* a8 = reinterpret_cast<const uint4*>(algo);
    const uint4* r8 = reinterpret_cast<const uint4*>(real);
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint4 a = __ldcs(a8 + i), r = __ldcs(r8 + i);
        if (a.x == r.x && a.y == r.y && a.z == r.z && a.w == r.w) { cnt[0] += 8; continue; }
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) one((aw[k >> 1] >> (16 * (k & 1))) & 0xFFFF, (rw[k >> 1] >> (16 * (k & 1))) & 0xFFFF);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) one(algo[n8 * 8 + threadIdx.x], real[n8 * 8 + threadIdx.x]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t v = cnt[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&acc[k], (unsigned long long)v);
    }
}

// ------------------------------------------------------------------------------------------------
// Compaction: dense -> position-sorted records, deterministic (count, scan, scatter)
// ------------------------------------------------------------------------------------------------
constexpr int kCompactChunk = 2048;  // positions per CTA
constexpr int kCompactThreads = 256;

// records position i contributes: none without a match or when the matched pattern is shorter than min_len
__device__ __forceinline__ uint32_t recs_at(uint32_t pid, bool expand, uint32_t min_len, const PatTables& t) {
    if (!pid || (min_len > 1 && t.len[pid - 1] < min_len)) return 0u;
    return expand ? t.anc_off[pid + 1] - t.anc_off[pid] : 1u;
}

__global__ void __launch_bounds__(kCompactThreads) compact_count_kernel(const uint16_t* __restrict__ out, uint64_t n,
                                                                         bool expand, uint32_t min_len, PatTables t,
                                                                         unsigned long long* __restrict__ block_counts) {
    const uint64_t base = uint64_t(blockIdx.x) * kCompactChunk;
    uint32_t c = 0;
    for (int k = threadIdx.x; k < kCompactChunk; k += kCompactThreads) {
        const uint64_t i = base + k;
        if (i < n) c += recs_at(out[i], expand, min_len, t);
    }
    __shared__ uint32_t sh[kCompactThreads / 32];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kCompactThreads / 32; ++w) s += sh[w];
        block_counts[blockIdx.x] = s;
    }
}

// exclusive scan of the per-CTA counts by ONE CTA (the array has n/2048 entries)
__global__ void __launch_bounds__(1024) compact_scan_kernel(unsigned long long* __restrict__ counts, uint64_t nb,
                                                            unsigned long long* __restrict__ total) {
    __shared__ unsigned long long sh[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t base = 0; base < nb; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        unsigned long long v = i < nb ? counts[i] : 0ull, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) sh[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long w = sh[lane], z = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, z, o);
                if (lane >= o) z += y;
            }
            sh[lane] = z - w;  // exclusive per-warp offsets
        }
        __syncthreads();
        const unsigned long long carry = carry_s;
        if (i < nb) counts[i] = carry + sh[wid] + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + sh[wid] + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(kCompactThreads) compact_write_kernel(const uint16_t* __restrict__ out, uint64_t n,
                                                                         uint64_t pos_base, bool expand, uint32_t min_len, PatTables t,
                                                                         const unsigned long long* __restrict__ block_offs,
                                                                         unsigned long long* __restrict__ recs, uint64_t cap) {
    // each thread owns 8 consecutive positions so that records stay position-sorted
    constexpr int kPer = kCompactChunk / kCompactThreads;
    const uint64_t base = uint64_t(blockIdx.x) * kCompactChunk + uint64_t(threadIdx.x) * kPer;
    uint32_t pid[kPer], c = 0;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        pid[k] = (base + k < n) ? out[base + k] : 0;
        if (!recs_at(pid[k], expand, min_len, t)) pid[k] = 0;
        c += recs_at(pid[k], expand, min_len, t);
    }
    // exclusive scan of c over the CTA
    __shared__ uint32_t sh[kCompactThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < wid; ++w) woff += sh[w];
    uint64_t dst = block_offs[blockIdx.x] + woff + x - c;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        uint32_t q = pid[k];
        if (!q) continue;
        const uint64_t pos = pos_base + base + k;
        if (!expand) {
            if (dst < cap) recs[dst] = (pos << 24) | q;
            ++dst;
        } else {
            for (uint32_t a = t.anc_off[q]; a < t.anc_off[q + 1]; ++a, ++dst)
                if (dst < cap) recs[dst] = (pos << 24) | t.anc_list[a];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Compaction of the scan kernel's sparse-mode bitmap (sfx_scan.cu, kFlags): bit i set = position i holds a match that
// qualifies.  Reads the bitmap (1 bit per position) and only the flagged entries of the dense result; deterministic and
// position-sorted like the dense compaction above (count, single-CTA scan, scatter).
// ------------------------------------------------------------------------------------------------
constexpr int kBmThreads = 256;
constexpr int kBmWords = 8;                                // words (of 32 positions) per thread
constexpr int kBmChunkWords = kBmThreads * kBmWords;       // per CTA: 65,536 positions

__global__ void __launch_bounds__(kBmThreads) bitmap_count_kernel(const uint32_t* __restrict__ flags, uint64_t n_words,
                                                                   unsigned long long* __restrict__ block_counts) {
    const uint64_t w0 = uint64_t(blockIdx.x) * kBmChunkWords + uint64_t(threadIdx.x) * kBmWords;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kBmWords; ++k) if (w0 + k < n_words) c += __popc(__ldg(flags + w0 + k));
    __shared__ uint32_t sh[kBmThreads / 32];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int w = 0; w < kBmThreads / 32; ++w) s += sh[w];
        block_counts[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(kBmThreads) bitmap_write_kernel(const uint32_t* __restrict__ flags, uint64_t n_words,
                                                                   const uint16_t* __restrict__ out, uint64_t pos_base,
                                                                   const unsigned long long* __restrict__ block_offs,
                                                                   unsigned long long* __restrict__ recs, uint64_t cap) {
    const uint64_t w0 = uint64_t(blockIdx.x) * kBmChunkWords + uint64_t(threadIdx.x) * kBmWords;
    uint32_t word[kBmWords], c = 0;
#pragma unroll
    for (int k = 0; k < kBmWords; ++k) {
        word[k] = (w0 + k < n_words) ? __ldg(flags + w0 + k) : 0u;
        c += __popc(word[k]);
    }
    __shared__ uint32_t sh[kBmThreads / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < wid; ++w) woff += sh[w];
    uint64_t dst = block_offs[blockIdx.x] + woff + x - c;
#pragma unroll
    for (int k = 0; k < kBmWords; ++k) {
        uint32_t m = word[k];
        while (m) {
            const uint32_t b = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t i = (w0 + k) * 32 + b;
            if (dst < cap) recs[dst] = ((pos_base + i) << 24) | out[i];
            ++dst;
        }
    }
}

}  // namespace

size_t bitmap_blocks(uint64_t n) { return size_t(((n + 31) / 32 + kBmChunkWords - 1) / kBmChunkWords); }

cudaError_t compact_bitmap_launch(const uint32_t* flags, const uint16_t* out, uint64_t n, uint64_t pos_base,
                                  unsigned long long* d_block_counts, unsigned long long* d_total, unsigned long long* recs,
                                  uint64_t cap, cudaStream_t st, uint64_t* launches) {
    const uint64_t n_words = (n + 31) / 32, nb = bitmap_blocks(n);
    if (nb == 0) return cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st);
    bitmap_count_kernel<<<uint32_t(nb), kBmThreads, 0, st>>>(flags, n_words, d_block_counts);
    compact_scan_kernel<<<1, 1024, 0, st>>>(d_block_counts, nb, d_total);
    bitmap_write_kernel<<<uint32_t(nb), kBmThreads, 0, st>>>(flags, n_words, out, pos_base, d_block_counts, recs, cap);
    *launches += 3;
    return cudaGetLastError();
}

cudaError_t generate_launch(int kind, uint64_t off, uint64_t n, uint8_t* dst, const PatTables& t, cudaStream_t st,
                            uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    if (n > (uint64_t(1) << 38)) return cudaErrorInvalidValue;   // 256 GiB per call: the grids below are sized in 32 bits
    const uint64_t words = (n + 7) / 8, blocks4k = (n + 4095) / 4096;
    switch (kind) {
        case 0:
        case 1:
            gen_uniform_kernel<<<uint32_t((words + 255) / 256), 256, 0, st>>>(off, n, dst);
            ++*launches;
            if (kind == 1) {
                gen_plant_kernel<<<uint32_t((blocks4k + 127) / 128), 128, 0, st>>>(off, n, dst, t);
                ++*launches;
            }
            break;
        case 2:
            gen_almost_kernel<<<uint32_t((blocks4k + 63) / 64), 64, 0, st>>>(off, n, dst, t);
            ++*launches;
            break;
        case 3:
            gen_ab_kernel<<<uint32_t((words + 255) / 256), 256, 0, st>>>(off, n, dst);
            ++*launches;
            break;
        case 4:
            gen_ascii_kernel<<<uint32_t((words + 255) / 256), 256, 0, st>>>(off, n, dst);
            ++*launches;
            break;
        default:
            return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t summarize_launch(const uint16_t* out, uint64_t n, uint64_t pos_base, const PatTables& t,
                             unsigned long long* d_acc4, int n_sms, cudaStream_t st, uint64_t* launches) {
    cudaError_t e = cudaMemsetAsync(d_acc4, 0, 4 * sizeof(unsigned long long), st);
    if (e != cudaSuccess || n == 0) return e;
    const uint64_t want = (n / 8 + kSumThreads - 1) / kSumThreads + 1;
    const uint32_t grid = uint32_t(want < uint64_t(n_sms) ? want : uint64_t(n_sms));
    const size_t smem = size_t(t.n_hot) * 16 + (size_t(t.n_patterns) + 1) * 2 + 16;
    e = cudaFuncSetAttribute(summarize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    summarize_kernel<<<grid, kSumThreads, smem, st>>>(out, n, pos_base, t, d_acc4);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t expand_ids_launch(const uint16_t* out, uint64_t n, const unsigned long long* table, unsigned long long* ids,
                              int n_sms, cudaStream_t st, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    const uint64_t want = (n / 4 + 255) / 256 + 1;
    expand_ids_kernel<<<uint32_t(want < uint64_t(n_sms) * 16 ? want : uint64_t(n_sms) * 16), 256, 0, st>>>(out, n, table, ids);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t merge_parts_launch(uint32_t* out32, const uint16_t* part, uint64_t n, uint32_t base, const uint32_t* glen, bool first,
                               int n_sms, cudaStream_t st, uint64_t* launches) {
    if (n == 0) return cudaSuccess;
    const uint64_t want = (n / 4 + 255) / 256 + 1;
    merge_parts_kernel<<<uint32_t(want < uint64_t(n_sms) * 16 ? want : uint64_t(n_sms) * 16), 256, 0, st>>>(out32, part, n, base, glen, first);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t classify_launch(const uint16_t* algo, const uint16_t* real, uint64_t n, const PatTables& t,
                            unsigned long long* d_acc4, int n_sms, cudaStream_t st, uint64_t* launches) {
    cudaError_t e = cudaMemsetAsync(d_acc4, 0, 4 * sizeof(unsigned long long), st);
    if (e != cudaSuccess || n == 0) return e;
    const uint64_t want = (n / 8 + 255) / 256 + 1;
    classify_kernel<<<uint32_t(want < uint64_t(n_sms) * 8 ? want : uint64_t(n_sms) * 8), 256, 0, st>>>(algo, real, n, t, d_acc4);
    ++*launches;
    return cudaGetLastError();
}

size_t compact_blocks(uint64_t n) { return size_t((n + kCompactChunk - 1) / kCompactChunk); }

cudaError_t compact_launch(const uint16_t* out, uint64_t n, uint64_t pos_base, bool expand, uint32_t min_len, const PatTables& t,
                           unsigned long long* d_block_counts, unsigned long long* d_total, unsigned long long* recs,
                           uint64_t cap, cudaStream_t st, uint64_t* launches) {
    const uint64_t nb = compact_blocks(n);
    if (nb == 0) return cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), st);
    compact_count_kernel<<<uint32_t(nb), kCompactThreads, 0, st>>>(out, n, expand, min_len, t, d_block_counts);
    compact_scan_kernel<<<1, 1024, 0, st>>>(d_block_counts, nb, d_total);
    compact_write_kernel<<<uint32_t(nb), kCompactThreads, 0, st>>>(out, n, pos_base, expand, min_len, t, d_block_counts, recs, cap);
    *launches += 3;
    return cudaGetLastError();
}

}  // namespace pm
