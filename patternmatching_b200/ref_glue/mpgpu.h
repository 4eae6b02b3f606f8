/*
 * mpgpu.h -- reference-side binding of the B200 matcher: the two files a maintainer of
 * yehonatan145/PatternMatching drops into Core/src (next to mpac.h / mplmac.h / mpbg.h) to follow the
 * reference's own recipe for adding an algorithm (Core/src/README.md:119-130):
 *
 *   1. Core/src/mps.h:20-25    add MPS_GPU and MPS_GPU_KR to the enum, before MPS_SIZE
 *   2. this file + mpgpu.c     mps_gpu_register() / mps_gpu_kr_register() fill mps_table[MPS_GPU*]
 *   3. Core/src/mps.c:17-20    #include "mpgpu.h";  mps.c:120-124  call the two register functions
 *   4. Core/src/measure.c:292-294  use mps_read_block_of(algo) when it is non-NULL
 *
 * Compiled inside the reference tree (it includes the reference's mps.h); links against libpm_b200.so.
 * oracle/make_gpu_exe.py applies exactly these edits to a scratch copy and builds oracle/_ref/exe_gpu.
 */
#ifndef MPGPU_H
#define MPGPU_H

#include <stddef.h>
#include "mps.h" /* reference: MpsElem, mps_table[], pattern_id_t, MPS_GPU, MPS_GPU_KR */

void mps_gpu_register();    /* exact matcher (B200 dictionary scan)        -> mps_table[MPS_GPU]    */
void mps_gpu_kr_register(); /* randomized Karp-Rabin variant (mpbg style)  -> mps_table[MPS_GPU_KR] */

/* Batched form of MpsElem.read_char for the rows that have one (the GPU rows), NULL otherwise:
 * out[j] = what read_char(obj, buf[j]) would have returned, state carried across calls. */
typedef size_t (*mps_read_block_t)(void* obj, const char* buf, size_t n, pattern_id_t* out);
mps_read_block_t mps_read_block_of(int algo);

#endif
