/*
 * post_ops.cu — the step right after the splat in FusionNode::processSyncedFrames
 * (fusion_node.cpp:789): Frameset::filterColor (frameset.cpp:194-208) =
 *   Kernels::filter_color_gauss_npp  (kernels.cu:632-653)  nppiFilterGauss_8u_C3R 3x3
 *   Kernels::filter_color_median_npp (kernels.cu:594-630)  nppiFilterMedian_8u_C3R 3x3
 * Both read a copy of the image (the reference does a D2D copy into the ColorFrame scratch buffer)
 * and write the interior (W-2)x(H-2); the one-pixel border keeps the splat output.
 * NPP is closed source: the reference pins nothing here; we pin against NPP 12.4 on the B200
 * (tests/test_ref_kernels_gpu.py).
 */
#include <cmath>

#include "dfb_internal.h"

namespace dfb {
namespace {

constexpr int kThreads = 256;

struct GaussTaps {
	float w[3];
};

__global__ void __launch_bounds__(kThreads)
gauss3x3_kernel(const GaussTaps G, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int width, int height) {
	const int row_bytes = width * 3;
	const int inner = (width - 2) * 3;
	const long long n = (long long)inner * (height - 2);
	const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
	if (t >= n) return;
	const int y = (int)(t / inner) + 1;
	const int xb = (int)(t % inner) + 3; /* byte offset inside the row (pixel 1 .. W-2, any channel) */
	float acc = 0.0f;
#pragma unroll
	for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
		for (int dx = -1; dx <= 1; ++dx)
			acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(G.w[dy + 1], G.w[dx + 1]),
			                               (float)src[(size_t)(y + dy) * row_bytes + xb + dx * 3]));
	int v = __float2int_rz(acc); /* NPP truncates (pinned against NPP 12.4 on a B200) */
	v = v > 255 ? 255 : v;
	dst[(size_t)y * row_bytes + xb] = (uint8_t)v;
}

__device__ __forceinline__ void sort2(int& a, int& b) {
	const int lo = min(a, b), hi = max(a, b);
	a = lo;
	b = hi;
}

__global__ void __launch_bounds__(kThreads)
median3x3_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int width, int height) {
	const int row_bytes = width * 3;
	const int inner = (width - 2) * 3;
	const long long n = (long long)inner * (height - 2);
	const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
	if (t >= n) return;
	const int y = (int)(t / inner) + 1;
	const int xb = (int)(t % inner) + 3;
	int p[9];
#pragma unroll
	for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
		for (int dx = -1; dx <= 1; ++dx) p[(dy + 1) * 3 + dx + 1] = src[(size_t)(y + dy) * row_bytes + xb + dx * 3];
	/* 19-exchange median-of-9 network */
	sort2(p[1], p[2]); sort2(p[4], p[5]); sort2(p[7], p[8]);
	sort2(p[0], p[1]); sort2(p[3], p[4]); sort2(p[6], p[7]);
	sort2(p[1], p[2]); sort2(p[4], p[5]); sort2(p[7], p[8]);
	sort2(p[0], p[3]); sort2(p[5], p[8]); sort2(p[4], p[7]);
	sort2(p[3], p[6]); sort2(p[1], p[4]); sort2(p[2], p[5]);
	sort2(p[4], p[7]); sort2(p[4], p[2]); sort2(p[6], p[4]);
	sort2(p[4], p[2]);
	/* reference quirk (kernels.cu:625 passes the un-offset destination): the result of the window centred on
	 * (y, x) is stored at (y-1, x-1); the last two rows / columns keep the splat output */
	dst[(size_t)(y - 1) * row_bytes + xb - 3] = (uint8_t)p[4];
}

}  // namespace
}  // namespace dfb

using namespace dfb;

extern "C" int dfb_op_filter_color(uint8_t* d_color, uint8_t* d_scratch, int width, int height, int median, void* stream) {
	DFB_REQUIRE(d_color && d_scratch, "NULL argument");
	DFB_REQUIRE(width > 0 && height > 0, "bad image size");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	if (width < 3 || height < 3) return DFB_OK;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const size_t bytes = (size_t)width * height * 3;
	DFB_CUDA(cudaMemcpyAsync(d_scratch, d_color, bytes, cudaMemcpyDeviceToDevice, s)); /* kernels.cu:616-620,644-647 */
	const long long n = (long long)(width - 2) * 3 * (height - 2);
	const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
	if (median) {
		median3x3_kernel<<<blocks, kThreads, 0, s>>>(d_scratch, d_color, width, height);
	} else {
		/* sigma = 0.4 + (3/2)*0.6 = 1.0 (comment at kernels.cu:649); separable taps exp(-d^2/2)/sum */
		const float e = expf(-0.5f);
		const float sum = 1.0f + 2.0f * e;
		GaussTaps G;
		G.w[0] = e / sum;
		G.w[1] = 1.0f / sum;
		G.w[2] = e / sum;
		gauss3x3_kernel<<<blocks, kThreads, 0, s>>>(G, d_scratch, d_color, width, height);
	}
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

