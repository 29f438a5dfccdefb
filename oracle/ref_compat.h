/*
 * ref_compat.h — force-included (nvcc -include) when compiling the UNMODIFIED reference
 * pointcloud_processing/src/kernels.cu in place for sm_100a with CUDA 12.9.
 *
 * The reference was written against CUDA 10.2 NPP, where nppiSumGetBufferHostSize_8u_C1R took an
 * `int*` (kernels.cu:514-516); NPP 12 takes `size_t*`. Rather than patching a copy of the
 * reference source, redirect that one call to an overload with the old signature.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref).
 */
#pragma once
#include <npp.h>
#include <nppi.h>
#include <cstddef>

static inline NppStatus dfb_ref_compat_sum_buffer_size(NppiSize roi, int* size_out) {
	size_t n = 0;
	const NppStatus st = nppiSumGetBufferHostSize_8u_C1R(roi, &n);
	*size_out = static_cast<int>(n);
	return st;
}
static inline NppStatus dfb_ref_compat_sum_buffer_size(NppiSize roi, size_t* size_out) {
	return nppiSumGetBufferHostSize_8u_C1R(roi, size_out);
}
#define nppiSumGetBufferHostSize_8u_C1R dfb_ref_compat_sum_buffer_size
