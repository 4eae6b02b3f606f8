// sfx_scan.cuh -- launch interface of the exact backward suffix-trie scan (sfx_scan.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

constexpr int kSfxThreads = 1024;
constexpr int kSfxTile = 512;    // positions per warp visit

struct SfxParams {
    const uint8_t* stream;    // device, 16-byte aligned; first reported byte
    uint64_t n;               // bytes to report on
    uint64_t hist_valid;      // readable bytes of the same stream directly before `stream`
    uint16_t* out;            // device, 16-byte aligned, n entries
    const uint16_t* root2;    // 65,536 entries
    const uint32_t* root1;    // 256 entries
    const uint32_t* rows;     // n_rows << log2_ncp entries
    const uint32_t* row_best; // n_rows entries
    const uint8_t* cls;       // 256 entries (device)
    const uint32_t* l3f;      // level-3 filter words, n_l3 entries (see dict.hpp); nullptr = not used
    uint32_t n_l3;
    cudaTextureObject_t rows_tex;  // rows as a linear texture (u32 texels), 0 = read rows with plain loads
    uint32_t l3_min_b;        // same for the second half of a visit
    uint32_t l3_min;          // a warp takes the filter path when >= l3_min of 32 sampled positions continue below root2
    const uint4* tail_rec;    // by pid: {text offset, length, next terminal length, best at tail start} (see dict.hpp)
    const uint8_t* pat_bytes; // pattern text (padded in front so that 8-byte windows never underrun)
    const uint32_t* pat_len;  // by canonical index (pid - 1)
    const uint16_t* parent;   // by pid: PatternsTree parent
    uint32_t cont_base, row2_base, log2_ncp;   // a root2 entry >= cont_base is the index of the row to continue at (row2_base == cont_base)
    uint64_t* queue;          // deferred deep walks: (position << 25) | row, or (position << 25) | (depth << 16) | pid;
                              // CTA b owns queue[b * q_per_cta .. (b+1) * q_per_cta)
    uint32_t* qcount;         // [2 * grid] per CTA: "continue at row" items (front of the strip), "tail" items (back)
    uint32_t q_per_cta;       // strip length; a CTA that fills its strip finishes further walks inline
    uint64_t n_tiles;         // full visits (n / kSfxTile); filled by the launcher
    // sparse mode (nullptr = off): bit i of this bitmap (bit i & 7 of byte i >> 3) is set iff the longest match at position i
    // is a pattern of at least min_len (>= 3) bytes; ceil(n / 32) words, 4-byte aligned.  Compacted by compact_bitmap_launch.
    uint8_t* flags;
    uint32_t min_len;
};

size_t sfx_smem_bytes();
// Launches the scan, the deep kernel and (for ragged ends) the edge kernel on `st`.
// ev[0..2], when non-null, are recorded on `st` before the main kernel, after it, and after the last kernel.
// number of CTAs the launcher will use for n bytes: sizes the queue
size_t sfx_scan_ctas(uint64_t n, int n_sms);
constexpr uint32_t kSfxMaxL3 = 16384;  // filter words that fit beside root2 in shared memory
cudaError_t sfx_scan_launch(const SfxParams& p, bool ident_cls, int n_sms, uint32_t max_pat_len, cudaStream_t st,
                            uint64_t* launches, cudaEvent_t* ev = nullptr);

// every position by the bounded walker straight from global memory: one launch, for small calls
cudaError_t sfx_walk_launch(const SfxParams& p, cudaStream_t st, uint64_t* launches);

}  // namespace pm
