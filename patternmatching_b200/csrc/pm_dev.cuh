// pm_dev.cuh -- small device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pm {

constexpr int kHalo = 352;  // smallest history an engine keeps (>= 346 = snort's max_pat_len - 1), multiple of 16; longer patterns: more

__host__ __device__ __forceinline__ uint64_t splitmix64_d(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP): the staging of kr_scan_kernel's tiles ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst, src and bytes multiples of 16.  Completion is signalled on `bar`.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// ---- GF(2^31-1) arithmetic (Core/src/field.h:61-80 restated for a Mersenne prime) ----
__host__ __device__ __forceinline__ uint32_t kr_reduce(uint64_t x) {  // x < 2^62
    uint32_t lo = uint32_t(x & 0x7FFFFFFFu), hi = uint32_t(x >> 31);
    uint32_t s = lo + hi;  // < 2^32
    s = (s & 0x7FFFFFFFu) + (s >> 31);
    return s >= 0x7FFFFFFFu ? s - 0x7FFFFFFFu : s;
}
__host__ __device__ __forceinline__ uint32_t kr_mulmod(uint32_t a, uint32_t b) { return kr_reduce(uint64_t(a) * b); }
__host__ __device__ __forceinline__ uint32_t kr_addmod(uint32_t a, uint32_t b) {
    uint32_t s = a + b;
    return s >= 0x7FFFFFFFu ? s - 0x7FFFFFFFu : s;
}
__host__ __device__ __forceinline__ uint32_t kr_submod(uint32_t a, uint32_t b) { return a >= b ? a - b : a + 0x7FFFFFFFu - b; }

}  // namespace pm
