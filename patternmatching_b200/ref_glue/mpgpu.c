/* mpgpu.c -- see mpgpu.h.  The reference-side half of the drop-in: registration only; every callback lives in
 * libpm_b200.so (include/pm_b200.h, patternmatching_b200/csrc/mps_gpu_shim.c). */
#include "mpgpu.h"
#include "pm_b200.h"

/* MpsElem (Core/src/mps.h:71-80) and pm_mps_elem (pm_b200.h) are the same eight pointers */
typedef char mpgpu_layout_check[(sizeof(MpsElem) == sizeof(pm_mps_elem)) ? 1 : -1];

void mps_gpu_register() { mps_gpu_register_into((pm_mps_elem*)&mps_table[MPS_GPU]); }
void mps_gpu_kr_register() { mps_gpu_kr_register_into((pm_mps_elem*)&mps_table[MPS_GPU_KR]); }

static size_t read_block(void* obj, const char* buf, size_t n, pattern_id_t* out) {
#ifdef MPGPU_REGISTER_BUFFERS
	/* measure.c hands over the same two arrays for every chunk.  When they were made static (a raised
	 * STREAM_BUFFER_SIZE does not fit the stack, Core/src/measure.c:243-245) they live as long as the program:
	 * page-lock them once, so that the stream leaves and the 8-byte ids arrive by DMA. */
	static const char* reg_buf;
	static pattern_id_t* reg_out;
	if (buf != reg_buf || out != reg_out) {
		pm_host_register((void*)buf, n);
		pm_host_register((void*)out, n * sizeof(pattern_id_t));
		reg_buf = buf;
		reg_out = out;
	}
#endif
	return gpu_read_block(obj, buf, n, (void**)out);
}

mps_read_block_t mps_read_block_of(int algo) {
	return (algo == MPS_GPU || algo == MPS_GPU_KR) ? read_block : NULL;
}
