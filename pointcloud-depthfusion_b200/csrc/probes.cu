/*
 * probes.cu — device self tests and microbenchmarks exported through the C ABI:
 *   dfb_selftest_division : shared-reciprocal division (dfb_math.cuh) == __fdiv_rn, bit for bit
 *   dfb_probe_atomics     : L2 64-bit atomicMin throughput for contiguous / strided / random slots —
 *                           the secondary roofline of the splat
 */
#include "dfb/dfb_synth.h"
#include "dfb_internal.h"

namespace dfb {
namespace {

__device__ __forceinline__ float rnd_float(uint64_t h, float lo_exp, float hi_exp) {
	/* log-uniform magnitude in [2^lo_exp, 2^hi_exp), random mantissa */
	const float u = (float)(h & 0xFFFFFF) * (1.0f / 16777216.0f);
	const float e = lo_exp + (hi_exp - lo_exp) * (float)((h >> 24) & 0xFFFF) * (1.0f / 65536.0f);
	return exp2f(e) * (1.0f + u);
}

__global__ void __launch_bounds__(256) division_selftest_kernel(uint64_t n, uint64_t seed, unsigned long long* mismatches) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	unsigned long long bad = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint64_t h0 = dfs_mix64(seed ^ (i * 3 + 0));
		const uint64_t h1 = dfs_mix64(seed ^ (i * 3 + 1));
		const uint64_t h2 = dfs_mix64(seed ^ (i * 3 + 2));
		/* domain of the projection: Z in ~[0.05, 60] m mostly, numerators fx*X up to ~1e5; a slice of
		 * the draws uses the full guarded range to exercise the slow path */
		const bool wide = (h0 >> 60) == 0;
		float Z = wide ? rnd_float(h0, -45.f, 45.f) : rnd_float(h0, -4.5f, 6.f);
		float ax = wide ? rnd_float(h1, -70.f, 70.f) : rnd_float(h1, -12.f, 17.f);
		float ay = wide ? rnd_float(h2, -70.f, 70.f) : rnd_float(h2, -12.f, 17.f);
		if (h1 >> 63) ax = -ax;
		if (h2 >> 63) ay = -ay;
		if (((h0 >> 48) & 0xFFF) == 0) ax = 0.0f;
		float qx = 0.f, qy = 0.f;
		const bool ok = divide_pair(ax, ay, Z, qx, qy);
		const float rx = __fdiv_rn(ax, Z);
		const float ry = __fdiv_rn(ay, Z);
		if (!ok || __float_as_uint(qx) != __float_as_uint(rx) || __float_as_uint(qy) != __float_as_uint(ry)) ++bad;
	}
	if (bad) atomicAdd(mismatches, bad);
}

template <int kPattern>
__global__ void __launch_bounds__(256)
atomics_probe_kernel(unsigned long long* __restrict__ keys, uint64_t n_slots, uint64_t n_ops, uint32_t stride, uint32_t salt) {
	const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ops; i += total) {
		uint64_t slot;
		if (kPattern == 0) {
			slot = i % n_slots;
		} else if (kPattern == 1) {
			/* lane l of warp w -> (w + l*stride): each lane of a warp hits a different row */
			const uint64_t w = i >> 5, l = i & 31;
			slot = (w + l * (uint64_t)stride) % n_slots;
		} else {
			slot = dfs_mix64(i ^ salt) % n_slots;
		}
		const unsigned long long key = ((unsigned long long)(0x3F800000u + (uint32_t)(dfs_mix64(i + salt) & 0xFFFFF)) << 32) | (uint32_t)i;
		atomicMin(keys + slot, key);
	}
}

}  // namespace
}  // namespace dfb

extern "C" int64_t dfb_selftest_division(uint64_t n, uint64_t seed) {
	using namespace dfb;
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	unsigned long long* d_bad = nullptr;
	unsigned long long bad = 0;
	DFB_CUDA(cudaMalloc(&d_bad, 8));
	DFB_CUDA(cudaMemset(d_bad, 0, 8));
	division_selftest_kernel<<<148 * 8, 256>>>(n, seed, d_bad);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	DFB_CUDA(cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost));
	cudaFree(d_bad);
	return (int64_t)bad;
}

extern "C" float dfb_probe_atomics(uint64_t n_slots, uint64_t n_ops, int pattern, uint32_t stride, int iters) {
	using namespace dfb;
	if (require_device() != DFB_OK) return -1.0f;
	if (n_slots == 0 || iters <= 0 || pattern < 0 || pattern > 2) {
		set_error("dfb_probe_atomics: bad argument");
		return -1.0f;
	}
	unsigned long long* d_keys = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	float ms = -1.0f;
	if (cudaMalloc(&d_keys, n_slots * 8) != cudaSuccess) return -1.0f;
	cudaMemset(d_keys, 0xFF, n_slots * 8);
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int blocks = 148 * 8;
	for (int it = -2; it < iters; ++it) { /* two warm-up launches */
		if (it == 0) cudaEventRecord(e0);
		const uint32_t salt = (uint32_t)(it + 7);
		if (pattern == 0)
			atomics_probe_kernel<0><<<blocks, 256>>>(d_keys, n_slots, n_ops, stride, salt);
		else if (pattern == 1)
			atomics_probe_kernel<1><<<blocks, 256>>>(d_keys, n_slots, n_ops, stride, salt);
		else
			atomics_probe_kernel<2><<<blocks, 256>>>(d_keys, n_slots, n_ops, stride, salt);
		count_launch();
	}
	cudaEventRecord(e1);
	if (cudaEventSynchronize(e1) == cudaSuccess && cudaGetLastError() == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d_keys);
	return ms;
}
