/*
 * probes.cu — device self tests and microbenchmarks exported through the C ABI:
 *   dfb_selftest_division : shared-reciprocal division (dfb_math.cuh) == __fdiv_rn, bit for bit
 *   dfb_probe_atomics     : L2 64-bit atomicMin throughput for several address patterns — the secondary
 *                           roofline of the splat. The probe kernels spend ~6 integer
 *                           instructions per atomic so that the atomic path, not instruction issue, is
 *                           what is measured.
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

/*
 * patterns (n_slots is a power of two; slot = f(i) & mask):
 *   0 contiguous       lane i of the grid -> slot i (+ pass offset)
 *   1 lane-strided     lane l of warp w   -> w + l*stride        (each lane its own row)
 *   2 random           multiplicative hash of i
 *   3 warp-scattered   warps land on pseudo-random 32-key runs, lanes contiguous inside the run
 *   4 pair-duplicated  like 0 but lanes 2k and 2k+1 hit the same slot (in-warp collisions)
 */
template <int kPattern, bool kReturn>
__global__ void __launch_bounds__(256)
atomics_probe_kernel(unsigned long long* __restrict__ keys, uint32_t mask, uint32_t ops_per_thread, uint32_t stride,
                     uint32_t salt, unsigned long long* sink) {
	const uint32_t total = gridDim.x * blockDim.x;
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long acc = 0;
	for (uint32_t k = 0; k < ops_per_thread; ++k, i += total) {
		uint32_t slot;
		if (kPattern == 0) slot = i;
		else if (kPattern == 1) slot = (i >> 5) + (i & 31u) * stride;
		else if (kPattern == 2) slot = (i ^ salt) * 2654435761u;
		else if (kPattern == 3) slot = (((i >> 5) ^ salt) * 2654435761u & ~31u) | (i & 31u);
		else slot = i >> 1;
		slot &= mask;
		const unsigned long long key = ((unsigned long long)(0x3F800000u + ((i * 40503u + salt) & 0xFFFFFu)) << 32) | i;
		if (kReturn)
			acc += atomicMin(keys + slot, key);
		else
			atomicMin(keys + slot, key);
	}
	if (kReturn && acc == 0x1234567ull) *sink = acc;
}

template <int kPattern>
void launch_probe(bool ret, unsigned long long* d_keys, uint32_t mask, uint32_t opt, uint32_t stride, uint32_t salt,
                  unsigned long long* sink, int blocks) {
	if (ret)
		atomics_probe_kernel<kPattern, true><<<blocks, 256>>>(d_keys, mask, opt, stride, salt, sink);
	else
		atomics_probe_kernel<kPattern, false><<<blocks, 256>>>(d_keys, mask, opt, stride, salt, sink);
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

/* pattern: see atomics_probe_kernel; bit 8 set = use the returning form (ATOM instead of RED).
 * n_slots is rounded down to a power of two. Returns milliseconds for `iters` launches of n_ops atomics. */
extern "C" float dfb_probe_atomics(uint64_t n_slots, uint64_t n_ops, int pattern, uint32_t stride, int iters) {
	using namespace dfb;
	if (require_device() != DFB_OK) return -1.0f;
	const bool ret = (pattern & 0x100) != 0;
	pattern &= 0xFF;
	if (n_slots < 64 || iters <= 0 || pattern < 0 || pattern > 4) {
		set_error("dfb_probe_atomics: bad argument");
		return -1.0f;
	}
	uint64_t pow2 = 64;
	while (pow2 * 2 <= n_slots && pow2 < (1ull << 31)) pow2 *= 2;
	const int blocks = 148 * 8;
	const uint32_t opt = (uint32_t)((n_ops + (uint64_t)blocks * 256 - 1) / ((uint64_t)blocks * 256));
	unsigned long long *d_keys = nullptr, *d_sink = nullptr;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	float ms = -1.0f;
	if (cudaMalloc(&d_keys, pow2 * 8) != cudaSuccess || cudaMalloc(&d_sink, 8) != cudaSuccess) return -1.0f;
	cudaMemset(d_keys, 0xFF, pow2 * 8);
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int it = -2; it < iters; ++it) { /* two warm-up launches */
		if (it == 0) cudaEventRecord(e0);
		const uint32_t salt = (uint32_t)(it * 7919 + 17);
		const uint32_t mask = (uint32_t)(pow2 - 1);
		switch (pattern) {
			case 0: launch_probe<0>(ret, d_keys, mask, opt, stride, salt, d_sink, blocks); break;
			case 1: launch_probe<1>(ret, d_keys, mask, opt, stride, salt, d_sink, blocks); break;
			case 2: launch_probe<2>(ret, d_keys, mask, opt, stride, salt, d_sink, blocks); break;
			case 3: launch_probe<3>(ret, d_keys, mask, opt, stride, salt, d_sink, blocks); break;
			default: launch_probe<4>(ret, d_keys, mask, opt, stride, salt, d_sink, blocks); break;
		}
		count_launch();
	}
	cudaEventRecord(e1);
	if (cudaEventSynchronize(e1) == cudaSuccess && cudaGetLastError() == cudaSuccess) {
		cudaEventElapsedTime(&ms, e0, e1);
		/* report per requested op count: the kernel rounds ops up to a whole number per thread */
		ms *= (float)((double)n_ops / ((double)opt * blocks * 256));
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d_keys);
	cudaFree(d_sink);
	return ms;
}

extern "C" int dfb_selftest_exclusive_scan(uint32_t* d_data, uint32_t n, uint32_t* total, void* stream) {
	using namespace dfb;
	DFB_REQUIRE(total && (d_data || n == 0), "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	uint32_t* d_scratch = nullptr;
	DFB_CUDA(cudaMalloc(&d_scratch, (exclusive_scan_scratch_words(n) + 1) * sizeof(uint32_t)));
	uint32_t* d_total = d_scratch + exclusive_scan_scratch_words(n);
	rc = launch_exclusive_scan(d_data, n, d_total, nullptr, 0, d_scratch, s);
	cudaError_t e = cudaSuccess;
	if (rc == DFB_OK) e = cudaMemcpyAsync(total, d_total, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
	if (rc == DFB_OK && e == cudaSuccess) e = cudaStreamSynchronize(s);
	cudaFree(d_scratch);
	if (rc != DFB_OK) return rc;
	DFB_CUDA(e);
	return DFB_OK;
}
