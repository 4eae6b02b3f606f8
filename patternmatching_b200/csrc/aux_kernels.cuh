// aux_kernels.cuh -- launch interfaces of the generator / summary / compaction kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

// Device-resident pattern tables.  off/len/bytes are indexed by the canonical pattern index
// (pid - 1); parent/chain/pidhash by pid (entry 0 = "no pattern").
struct PatTables {
    uint32_t n_patterns;
    const uint32_t* off;
    const uint32_t* len;
    const uint8_t* bytes;
    const uint16_t* parent;   // PatternsTree parent pid (Core/src/PatternsTree.h:90-94), 0 = root
    const uint16_t* chain;    // number of ancestors
    const uint64_t* pidhash;  // splitmix64(((file+1) << 32) | line), see oracle/pm_oracle.h match_digest
    const uint32_t* anc_off;  // output links flattened to ranges: the patterns ending where pid ends are
    const uint16_t* anc_list; // anc_list[anc_off[pid] .. anc_off[pid+1]) (pid first, then its ancestors)
    // summary fast path (summarize_kernel keeps these in shared memory): the shortest patterns -- those that match at
    // most positions of any traffic -- with at most one ancestor.  hot_map[pid] = slot | 0x8000 when the pattern has an
    // ancestor, 0xFFFF = not in the table; hot_own[slot] = pidhash of the pattern, hot_anc[slot] = pidhash of its ancestor.
    const uint16_t* hot_map;  // n_patterns + 1 entries
    const uint64_t* hot_own;
    const uint64_t* hot_anc;
    uint32_t n_hot;
};
constexpr uint32_t kSummaryHotMax = 3072;   // slots of the summary fast path

cudaError_t generate_launch(int kind, uint64_t off, uint64_t n, uint8_t* dst, const PatTables& t, cudaStream_t st,
                            uint64_t* launches);
cudaError_t summarize_launch(const uint16_t* out, uint64_t n, uint64_t pos_base, const PatTables& t,
                             unsigned long long* d_acc4, int n_sms, cudaStream_t st, uint64_t* launches);
// ids[i] = table[out[i]] (pid -> the caller's 64-bit pattern id); out 8-byte aligned, ids 16-byte aligned
cudaError_t expand_ids_launch(const uint16_t* out, uint64_t n, const unsigned long long* table, unsigned long long* ids,
                              int n_sms, cudaStream_t st, uint64_t* launches);
// out32[i] = part[i] + base as a global pid (0 stays 0); unless `first`, the longer (glen, by global pid) of that and out32[i]
cudaError_t merge_parts_launch(uint32_t* out32, const uint16_t* part, uint64_t n, uint32_t base, const uint32_t* glen, bool first,
                               int n_sms, cudaStream_t st, uint64_t* launches);
// measure_success_rate (Core/src/measure.c:174-190) of one dense result against another; d_acc4 = success, partial, false_neg, false_pos
cudaError_t classify_launch(const uint16_t* algo, const uint16_t* real, uint64_t n, const PatTables& t,
                            unsigned long long* d_acc4, int n_sms, cudaStream_t st, uint64_t* launches);
size_t compact_blocks(uint64_t n);
// min_len > 1: only matches whose (longest) pattern has at least min_len bytes produce records
cudaError_t compact_launch(const uint16_t* out, uint64_t n, uint64_t pos_base, bool expand, uint32_t min_len, const PatTables& t,
                           unsigned long long* d_block_counts, unsigned long long* d_total, unsigned long long* recs,
                           uint64_t cap, cudaStream_t st, uint64_t* launches);

// sparse mode: compact the scan kernel's flag bitmap (bit i = position i qualifies) into position-sorted records
// (pos_base + i) << 24 | out[i]; d_block_counts needs bitmap_blocks(n) + 1 entries
size_t bitmap_blocks(uint64_t n);
cudaError_t compact_bitmap_launch(const uint32_t* flags, const uint16_t* out, uint64_t n, uint64_t pos_base,
                                  unsigned long long* d_block_counts, unsigned long long* d_total, unsigned long long* recs,
                                  uint64_t cap, cudaStream_t st, uint64_t* launches);

}  // namespace pm
