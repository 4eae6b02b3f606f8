// sfx_scan.cu -- exact dictionary scan for sm_100a, dense uint16 result.
//
// Semantics (bit-exact): out[i] = pid of the LONGEST dictionary pattern that is a suffix of
// stream[..i], 0 if none -- what ac_read_char returns per byte (Core/src/mpac.c:304-319) when driven
// by the loop at Core/src/measure.c:292-294.
//
// Method.  For position i the longest pattern ending at i is the deepest terminal on the path
// c[i], c[i-1], c[i-2], ... in the trie of REVERSED patterns.  The host compiler (dict.cpp) has
// flattened that trie so that every table entry is either the final answer or "continue at row r":
//   level 1+2 : root2[c[i] << 8 | c[i-1]]   65,536 x u16 = 128 KiB, resident in SHARED memory
//   level >=3 : rows[row << log2_ncp | cls(c[i-k])]   u32, global memory (hot part L2-resident)
// All positions are independent, so there is no per-thread warm-up: a CTA stages one 16 KiB tile
// plus a 352-byte left halo (>= max_pat_len-1) with ONE bulk async copy (TMA engine,
// SASS UBLKCP), double buffered behind an mbarrier; every thread resolves 16 consecutive
// positions (16 shared-memory gathers, ~10% of them followed by one L2 lookup), results are packed
// in shared memory and leave with one bulk async store per tile.
#include "pm_dev.cuh"
#include "sfx_scan.cuh"

namespace pm {

namespace {

constexpr int kThreads = kSfxThreads;
constexpr int kPos = kSfxPosPerThread;          // positions per thread per tile
constexpr int kTile = kSfxTile;                 // bytes per tile
constexpr int kTileBuf = kHalo + kTile;         // staged bytes per buffer
constexpr uint32_t kCont = 0x80000000u;

// shared memory carve-up (bytes)
constexpr int kOffRoot2 = 0;                           // 131072
constexpr int kOffTile0 = 131072;
constexpr int kOffTile1 = kOffTile0 + kTileBuf;
constexpr int kOffOut = kOffTile1 + kTileBuf;          // kTile * 2
constexpr int kOffCls = kOffOut + kTile * 2;           // 256
constexpr int kOffBar = kOffCls + 256;                 // 2 x 8
constexpr int kSmemBytes = kOffBar + 16;
static_assert(kOffTile1 % 16 == 0 && kOffOut % 16 == 0 && kOffBar % 8 == 0, "alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// byte / 16-bit window extraction from the 20-byte register window W = {prev, w.x, w.y, w.z, w.w}
template <int O>
__device__ __forceinline__ uint32_t win_u8(const uint32_t (&W)[5]) {
    return (W[O >> 2] >> (8 * (O & 3))) & 0xFFu;
}
template <int O>
__device__ __forceinline__ uint32_t win_u16(const uint32_t (&W)[5]) {
    if constexpr ((O & 3) < 3) return (W[O >> 2] >> (8 * (O & 3))) & 0xFFFFu;
    else return __funnelshift_r(W[O >> 2], W[(O >> 2) + 1], 24) & 0xFFFFu;
}

// Levels >= 4 (about 1e-3 of the positions on uniform bytes): follow the rows until a final entry.
// `pb` points at c[i] inside the staged tile; the halo guarantees pb[-k] is staged for every k the
// trie can ask for (k < max_pat_len <= kHalo + 1).
template <bool kIdentCls>
__device__ __noinline__ uint32_t sfx_walk_deep(uint32_t v, const uint8_t* pb, int k, const uint32_t* __restrict__ rows,
                                               uint32_t log2_ncp, const uint8_t* s_cls) {
    while (v & kCont) {
        uint32_t c = pb[-k];
        if constexpr (!kIdentCls) c = s_cls[c];
        v = __ldg(rows + ((size_t(v & ~kCont) << log2_ncp) | c));
        ++k;
    }
    return v;
}

template <int J, bool kIdentCls>
__device__ __forceinline__ void resolve_one(uint32_t (&e)[kPos], const uint32_t (&W)[5], const SfxParams& p,
                                            const uint8_t* base, const uint8_t* s_cls) {
    if (e[J] >= p.cont_base) {  // 2-byte suffix continues below the shared-memory table
        uint32_t c2 = win_u8<2 + J>(W);  // c[i-2]
        if constexpr (!kIdentCls) c2 = s_cls[c2];
        uint32_t row = e[J] - p.cont_base + p.row2_base;
        uint32_t v = __ldg(p.rows + ((size_t(row) << p.log2_ncp) | c2));
        if (v & kCont) v = sfx_walk_deep<kIdentCls>(v, base + J, 3, p.rows, p.log2_ncp, s_cls);
        e[J] = v;
    }
}

template <bool kIdentCls>
__global__ void __launch_bounds__(kThreads, 1) sfx_scan_kernel(const SfxParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint16_t* s_root2 = reinterpret_cast<uint16_t*>(smem + kOffRoot2);
    auto s_tile = [&](int b) -> uint8_t* { return smem + kOffTile0 + b * kTileBuf; };
    uint16_t* s_out = reinterpret_cast<uint16_t*>(smem + kOffOut);
    uint8_t* s_cls = smem + kOffCls;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const bool have_halo = p.hist_valid >= uint64_t(kHalo);

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    // zero the halo of buffer 0 (tile 0 without history reads it; the fix-up pass owns those positions)
    for (int i = tid; i < kHalo / 4; i += kThreads) reinterpret_cast<uint32_t*>(s_tile(0))[i] = 0;
    if (tid < 256) s_cls[tid] = p.cls[tid];
    fence_proxy_async();
    __syncthreads();

    auto issue_tile = [&](uint64_t t, int buf) {  // one elected thread
        const uint64_t s0 = t * uint64_t(kTile);
        const uint32_t len = uint32_t(min(uint64_t(kTile), p.n - s0));
        const uint32_t body = len & ~15u;
        const bool halo = (t > 0) || have_halo;
        const uint32_t bytes = body + (halo ? kHalo : 0);
        if (bytes) {
            mbar_arrive_expect_tx(&bars[buf], bytes);
            bulk_g2s(s_tile(buf) + (halo ? 0 : kHalo), p.stream + s0 - (halo ? kHalo : 0), bytes, &bars[buf]);
        } else {
            mbar_arrive_expect_tx(&bars[buf], 0);
        }
    };

    uint64_t t = blockIdx.x;
    if (t < p.n_tiles && tid == 0) issue_tile(t, 0);

    // root2 -> shared memory (once per CTA; 128 KiB of coalesced 16-byte loads, L2 hits after the first CTA)
    {
        const int4* src = reinterpret_cast<const int4*>(p.root2);
        int4* dst = reinterpret_cast<int4*>(s_root2);
        for (int i = tid; i < 131072 / 16; i += kThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    for (uint32_t it = 0; t < p.n_tiles; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint64_t s0 = t * uint64_t(kTile);
        const uint32_t len = uint32_t(min(uint64_t(kTile), p.n - s0));
        const uint64_t tn = t + gridDim.x;
        if (tid == 0 && tn < p.n_tiles) issue_tile(tn, buf ^ 1);  // buffer was released by the barriers of it-1
        if (len & 15u) {  // ragged end of the stream: the last <16 bytes come in with plain loads
            const uint32_t body = len & ~15u;
            if (uint32_t(tid) < (len & 15u)) s_tile(buf)[kHalo + body + tid] = p.stream[s0 + body + tid];
            __syncthreads();
        }
        mbar_wait(&bars[buf], (it >> 1) & 1);

        const uint8_t* base = s_tile(buf) + kHalo + tid * kPos;
        uint32_t e[kPos];
        const bool active = uint32_t(tid * kPos) < len;
        // the 20-byte register window {c[p0-4..p0-1], c[p0..p0+15]}; loaded by every thread so that the
        // shuffle stays warp-convergent (inactive threads of a ragged last tile read stale bytes, unused)
        uint32_t W[5];
        {
            const uint4 w = *reinterpret_cast<const uint4*>(base);
            W[1] = w.x; W[2] = w.y; W[3] = w.z; W[4] = w.w;
            W[0] = __shfl_up_sync(0xFFFFFFFFu, w.w, 1);
            if (lane == 0) W[0] = *reinterpret_cast<const uint32_t*>(base - 4);
        }
        if (active) {
            // phase A: levels 1+2, one shared-memory gather per position
#pragma unroll
            for (int j = 0; j < kPos; ++j) e[j] = 0;
            e[0] = s_root2[win_u16<3>(W)];   e[1] = s_root2[win_u16<4>(W)];
            e[2] = s_root2[win_u16<5>(W)];   e[3] = s_root2[win_u16<6>(W)];
            e[4] = s_root2[win_u16<7>(W)];   e[5] = s_root2[win_u16<8>(W)];
            e[6] = s_root2[win_u16<9>(W)];   e[7] = s_root2[win_u16<10>(W)];
            e[8] = s_root2[win_u16<11>(W)];  e[9] = s_root2[win_u16<12>(W)];
            e[10] = s_root2[win_u16<13>(W)]; e[11] = s_root2[win_u16<14>(W)];
            e[12] = s_root2[win_u16<15>(W)]; e[13] = s_root2[win_u16<16>(W)];
            e[14] = s_root2[win_u16<17>(W)]; e[15] = s_root2[win_u16<18>(W)];
            // phase B/C: level 3 from L2 (predicated, independent loads), deeper levels rarely
            resolve_one<0, kIdentCls>(e, W, p, base, s_cls);   resolve_one<1, kIdentCls>(e, W, p, base, s_cls);
            resolve_one<2, kIdentCls>(e, W, p, base, s_cls);   resolve_one<3, kIdentCls>(e, W, p, base, s_cls);
            resolve_one<4, kIdentCls>(e, W, p, base, s_cls);   resolve_one<5, kIdentCls>(e, W, p, base, s_cls);
            resolve_one<6, kIdentCls>(e, W, p, base, s_cls);   resolve_one<7, kIdentCls>(e, W, p, base, s_cls);
            resolve_one<8, kIdentCls>(e, W, p, base, s_cls);   resolve_one<9, kIdentCls>(e, W, p, base, s_cls);
            resolve_one<10, kIdentCls>(e, W, p, base, s_cls);  resolve_one<11, kIdentCls>(e, W, p, base, s_cls);
            resolve_one<12, kIdentCls>(e, W, p, base, s_cls);  resolve_one<13, kIdentCls>(e, W, p, base, s_cls);
            resolve_one<14, kIdentCls>(e, W, p, base, s_cls);  resolve_one<15, kIdentCls>(e, W, p, base, s_cls);
        }

        if (tid == 0) bulk_wait_read<0>();  // the previous tile's result has left shared memory
        __syncthreads();                    // s_out is free; every thread is done reading s_tile[buf]
        if (active) {
            uint4 lo, hi;
            lo.x = e[0] | (e[1] << 16);   lo.y = e[2] | (e[3] << 16);
            lo.z = e[4] | (e[5] << 16);   lo.w = e[6] | (e[7] << 16);
            hi.x = e[8] | (e[9] << 16);   hi.y = e[10] | (e[11] << 16);
            hi.z = e[12] | (e[13] << 16); hi.w = e[14] | (e[15] << 16);
            uint4* so = reinterpret_cast<uint4*>(s_out + tid * kPos);
            so[0] = lo;
            so[1] = hi;
        }
        fence_proxy_async();
        __syncthreads();
        const uint32_t out_bytes = len * 2;
        if (tid == 0) {
            if (out_bytes & ~15u) bulk_s2g(p.out + s0, s_out, out_bytes & ~15u);
            bulk_commit();
        }
        if (out_bytes & 15u) {  // ragged end: < 8 results with plain stores
            const uint32_t done = (out_bytes & ~15u) / 2;
            if (uint32_t(tid) < len - done) p.out[s0 + done + tid] = s_out[done + tid];
        }
    }
    if (tid == 0) bulk_wait<0>();
}

// Positions whose history is shorter than max_pat_len-1 (only at the very start of a stream):
// bounded walk straight from global memory.  avail(i) = i + hist_valid + 1 bytes exist up to c[i].
__global__ void sfx_fixup_kernel(const SfxParams p, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count || i >= p.n) return;
    const uint64_t avail = uint64_t(i) + p.hist_valid + 1;
    uint32_t v = p.root1[p.stream[i]];
    uint64_t k = 1;
    while (v & kCont) {
        const uint32_t row = v & ~kCont;
        if (k >= avail) { v = p.row_best[row]; break; }
        v = p.rows[(size_t(row) << p.log2_ncp) | p.cls[*(p.stream + i - k)]];
        ++k;
    }
    p.out[i] = uint16_t(v);
}

}  // namespace

size_t sfx_smem_bytes() { return kSmemBytes; }

cudaError_t sfx_scan_launch(const SfxParams& p_in, bool ident_cls, int n_sms, uint32_t max_pat_len, cudaStream_t st,
                            uint64_t* launches) {
    SfxParams p = p_in;
    if (p.n == 0) return cudaSuccess;
    p.n_tiles = (p.n + kTile - 1) / kTile;
    auto kern = ident_cls ? sfx_scan_kernel<true> : sfx_scan_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    const uint32_t grid = uint32_t(p.n_tiles < uint64_t(n_sms) ? p.n_tiles : uint64_t(n_sms));
    kern<<<grid, kThreads, kSmemBytes, st>>>(p);
    ++*launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (max_pat_len > 1 && p.hist_valid < max_pat_len - 1) {
        const uint32_t count = uint32_t(p.n < uint64_t(max_pat_len - 1 - p.hist_valid) ? p.n : uint64_t(max_pat_len - 1 - p.hist_valid));
        sfx_fixup_kernel<<<(count + 127) / 128, 128, 0, st>>>(p, count);
        ++*launches;
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace pm
