// kr_scan.cuh -- launch interface of the randomized Karp-Rabin variant (kr_scan.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "aux_kernels.cuh"
#include "dict.hpp"

namespace pm {

constexpr int kKrTile = 512;   // positions per warp tile

struct KrDevTables {
    uint32_t r = 0;
    uint32_t bucket_mask = 0;
    uint32_t *slot_fp = nullptr, *slot_begin = nullptr, *slot_count = nullptr;
    uint32_t *cand_pid = nullptr, *cand_len = nullptr, *cand_stage_off = nullptr, *stage_fp = nullptr;
    uint32_t* bloom = nullptr;     // 2^19 bits
    uint32_t* rpow = nullptr;      // r^k,  k < kHalo + kKrTile
    uint32_t* rinvpow = nullptr;   // r^-k
    uint16_t* short_of = nullptr;  // pid -> longest ancestor-or-self of <= 8 bytes (0 = none)
    uint32_t* long_bits = nullptr; // bit pid: the pattern has more than 8 bytes (short_of[pid] != pid); 2048 words
};

cudaError_t kr_upload_tables(const Dict& d, const KrTables& k, KrDevTables* t, size_t* bytes);
void kr_free_tables(KrDevTables* t);
// d_out holds the exact dense result on entry (used ONLY through short_of[], i.e. for the patterns
// of <= 8 bytes that the variant matches exactly) and the variant's result on exit.
cudaError_t kr_scan_launch(const KrDevTables& t, const uint8_t* stream, uint64_t n, uint64_t hist_valid, uint16_t* out,
                           const PatTables& pt, int n_sms, cudaStream_t st, uint64_t* launches, bool bulk = true);

// out[i] = short_of[out[i]] in place: the observable behaviour of the reference's MPBG (only patterns of <= 8 bytes are
// ever reported, Core/src/mpbg.c:132-145 + SURVEY Q5); `out` holds the exact dense result on entry, 16-byte aligned.
cudaError_t kr_short_only_launch(const KrDevTables& t, uint16_t* out, uint64_t n, int n_sms, cudaStream_t st, uint64_t* launches);

}  // namespace pm
