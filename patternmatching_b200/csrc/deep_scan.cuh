// deep_scan.cuh -- launch interface of the deep-match forward walker (deep_scan.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

struct DeepParams {
    const uint8_t* stream;       // device, 16-byte aligned
    uint64_t n;
    uint64_t hist_valid;
    uint16_t* out;               // device, 16-byte aligned
    const uint16_t* hot_rows;    // [hot state << 8 | byte] complete DFA rows of the root and the depth-1 states
    const uint16_t* hot_longest; // [state < n_small]: hot, DENSE and depth-2 states
    uint32_t n_hot, dense_end, n_small;   // ids < n_hot: shared-memory rows; < dense_end: DENSE rows; < n_small: longest id in shared memory
    const uint32_t* recs;        // 8 words per state (dict.hpp: DeepTables)
    const uint32_t* dense_rows;  // [(state - n_hot) << 8 | byte]: 256 entries per DENSE state
    uint32_t warm;               // max_pat_len - 1
    uint32_t seg;                // bytes reported per segment (filled by the launcher)
    uint64_t n_seg;              // (filled by the launcher)
};

size_t deep_smem_bytes(uint32_t n_hot, uint32_t n_small);
cudaError_t deep_scan_launch(const DeepParams& p, int n_sms, cudaStream_t st, uint64_t* launches);

}  // namespace pm
