/* mpgpu.c -- see mpgpu.h.  The reference-side half of the drop-in: registration only; every callback lives in
 * libpm_b200.so (include/pm_b200.h, patternmatching_b200/csrc/mps_gpu_shim.c). */
#include "mpgpu.h"
#include "pm_b200.h"

/* MpsElem (Core/src/mps.h:71-80) and pm_mps_elem (pm_b200.h) are the same eight pointers */
typedef char mpgpu_layout_check[(sizeof(MpsElem) == sizeof(pm_mps_elem)) ? 1 : -1];

void mps_gpu_register() { mps_gpu_register_into((pm_mps_elem*)&mps_table[MPS_GPU]); }
void mps_gpu_kr_register() { mps_gpu_kr_register_into((pm_mps_elem*)&mps_table[MPS_GPU_KR]); }

static size_t read_block(void* obj, const char* buf, size_t n, pattern_id_t* out) {
	return gpu_read_block(obj, buf, n, (void**)out);
}

mps_read_block_t mps_read_block_of(int algo) {
	return (algo == MPS_GPU || algo == MPS_GPU_KR) ? read_block : NULL;
}
