// dfa_scan.cuh -- launch interface of the forward Aho-Corasick DFA walker (dfa_scan.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

constexpr int kDfaSeg = 4096;  // bytes reported per thread

struct DfaParams {
    const uint8_t* stream;   // device, 16-byte aligned
    uint64_t n;
    uint64_t hist_valid;
    uint16_t* out;           // device, 16-byte aligned
    const uint32_t* delta;   // [state << log2_ncp | cls]
    const uint16_t* longest; // [state]
    const uint8_t* cls;      // 256 entries
    uint32_t log2_ncp;
    uint32_t warm;           // max_pat_len - 1
};

cudaError_t dfa_scan_launch(const DfaParams& p, cudaStream_t st, uint64_t* launches);

}  // namespace pm
