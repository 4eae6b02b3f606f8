// Microbenchmark: what a shared-memory gather costs the LSU data pipe (wavefronts per LDS instruction).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_gather lds_gather.cu ; run under
// ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,gpu__time_duration.sum
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <typename T, int kMode>   // kMode 0: lane-linear (conflict-free), 1: random per lane, 2: random, half the lanes predicated off
__global__ void __launch_bounds__(1024) gather(const uint32_t* __restrict__ seeds, uint32_t* __restrict__ sink) {
    extern __shared__ __align__(16) uint8_t smem[];
    T* tab = reinterpret_cast<T*>(smem);
    const int n = 131072 / sizeof(T);
    for (int i = threadIdx.x; i < n; i += blockDim.x) tab[i] = T(i * 2654435761u);
    __syncthreads();
    uint32_t x = seeds[blockIdx.x * blockDim.x + threadIdx.x], acc = 0;
    const int lane = threadIdx.x & 31;
#pragma unroll 8
    for (int it = 0; it < kIters; ++it) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx = kMode == 0 ? uint32_t((it * 32 + lane) & (n - 1)) : (x >> 8) & uint32_t(n - 1);
        if (kMode == 2) { if (x & 0x80) acc += tab[idx]; }
        else if (kMode == 3) acc += __shfl_up_sync(0xFFFFFFFFu, x, 1);
        else if (kMode == 4) acc += __shfl_sync(0xFFFFFFFFu, x, 31);
        else if (kMode == 5) acc += __ballot_sync(0xFFFFFFFFu, x & 0x80);
        else acc += tab[idx];
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    uint32_t *seeds, *sink;
    const int blocks = 148, threads = 1024;
    cudaMalloc(&seeds, blocks * threads * 4); cudaMalloc(&sink, blocks * threads * 4);
    uint32_t* h = new uint32_t[blocks * threads];
    for (int i = 0; i < blocks * threads; ++i) h[i] = i * 2246822519u + 12345u;
    cudaMemcpy(seeds, h, blocks * threads * 4, cudaMemcpyHostToDevice);
    auto run = [&](auto kern, const char* name) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        kern<<<blocks, threads, 131072>>>(seeds, sink);
        cudaEventRecord(a);
        kern<<<blocks, threads, 131072>>>(seeds, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        // cycles per warp-level LDS per SM, assuming 1.9 GHz
        const double lds = double(kIters) * 32;  // warp instructions per SM
        printf("%-28s %8.3f ms  %6.2f cycles per warp LDS per SM (at 1.9 GHz)  err=%d\n", name, ms, ms * 1e-3 * 1.9e9 / lds, int(cudaGetLastError()));
    };
    run(gather<uint16_t, 0>, "u16 linear");
    run(gather<uint32_t, 0>, "u32 linear");
    run(gather<uint16_t, 1>, "u16 random");
    run(gather<uint32_t, 1>, "u32 random");
    run(gather<uint16_t, 2>, "u16 random, half predicated");
    run(gather<uint32_t, 3>, "shfl_up");
    run(gather<uint32_t, 4>, "shfl_idx 31");
    run(gather<uint32_t, 5>, "ballot");
    return 0;
}
