// dfa_scan.cuh -- launch interface of the forward Aho-Corasick DFA walker (dfa_scan.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

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
    uint32_t hot_rows;       // leading states whose rows are kept in shared memory as u16 (hot variant)
    uint32_t hot_long;       // leading states whose longest-ids are kept in shared memory
    const uint32_t* fb_meta; // [state] Bloom of the goto children | failure state << 16 (see dict.hpp)
    uint32_t fb_count;       // states [hot_rows, hot_rows + fb_count) -- the first BFS level below the hot rows -- keep
                             // their fb_meta word in shared memory: no child on c => the step is the failure state's hot row
    uint32_t n_states;       // states of the automaton
    uint32_t no_fused;       // do not use the fused-entry kernel for automata that fit shared memory entirely (A/B switch)
    uint32_t seg;            // bytes reported per thread (filled by the launcher)
    uint32_t wide;           // stream and out are 32-byte aligned: 256-bit segment I/O (filled by the launcher)
};

// how many leading states go to shared memory (whole BFS levels whose targets fit u16)
void dfa_plan_hot(uint32_t n_states, uint32_t log2_ncp, const uint32_t* depth_count, uint32_t n_depths,
                  uint32_t* hot_rows, uint32_t* hot_long, uint32_t* fb_count);

cudaError_t dfa_scan_launch(const DfaParams& p, bool ident_cls, bool flat, int n_sms, cudaStream_t st, uint64_t* launches);

}  // namespace pm
