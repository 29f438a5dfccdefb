/*
 * align.cu — depth -> colour alignment for unaligned input (Frameset::alignDepthToColor, frameset.cpp:180-188 ->
 * Kernels::align_depth_to_color, kernels.cu:714-747). fusion_node uses it when `align_frames` is set
 * (fusion_node.cpp:443-457,719-733); the default deployment aligns in the camera driver.
 *
 * Reference: three kernels and a pixel map in HBM (kernel_map_depth_to_color -> int2[2*P], kernel_align_depth_to_color,
 * kernel_replace_to_zero) plus a memset. Here: memset + ONE kernel that maps both corners of a depth pixel in registers
 * and scatters it, + the 0xFFFF -> 0 sweep. Arithmetic: the exact operation sequence nvcc emitted for the reference
 * (read from its SASS; every FMA below is one the reference executes), so the result is bit-identical to the
 * reference kernels on the same inputs (tests/test_ref_kernels_gpu.py) — including its quirk of min-ing d<<16|d into the
 * 32-bit word that holds the target pixel AND its horizontal neighbour (kernels.cu:293-299).
 * The scatter is a commutative 32-bit atomicMin: deterministic. The F-theta colour model (kernels.cu:107-112) calls CUDA's own
 * sqrtf / tanf / atanf like the reference does: pinned against the compiled reference kernels on the GPU
 * (tests/test_ref_kernels_gpu.py), not against the CPU oracle, whose libm rounds those functions differently.
 */
#include "dfb_internal.h"

namespace dfb {
namespace {

struct AlignParams {
	dfb_intrinsics depth, color;
	dfb_extrinsics ex;
	float depth_scale;
};

/* kernel_transfer_pixels, kernels.cu:137-159 (deproject_pixel_to_point :46-68, transform_point_to_point :128-136,
 * project_point_to_pixel :92-116) */
__device__ __forceinline__ int2 map_corner(const AlignParams& A, float depth_val, int dx, int dy, float shift) {
	float x = __fdiv_rn(__fadd_rn(__fadd_rn((float)dx, shift), -A.depth.ppx), A.depth.fx);
	float y = __fdiv_rn(__fadd_rn(__fadd_rn((float)dy, shift), -A.depth.ppy), A.depth.fy);
	if (A.depth.model == DFB_DISTORTION_INVERSE_BROWN_CONRADY) {
		const float* c = A.depth.coeffs;
		const float r2 = __fmaf_rn(x, x, __fmul_rn(y, y));
		float f = __fmaf_rn(r2, c[0], 1.0f);
		f = __fmaf_rn(r2, __fmul_rn(r2, c[1]), f);
		f = __fmaf_rn(r2, __fmul_rn(r2, __fmul_rn(r2, c[4])), f);
		const float a = __fmul_rn(x, __fadd_rn(c[3], c[3])), b = __fmul_rn(x, __fadd_rn(c[2], c[2]));
		const float gx = __fmaf_rn(x, __fadd_rn(x, x), r2), gy = __fmaf_rn(y, __fadd_rn(y, y), r2);
		const float ux = __fmaf_rn(c[3], gx, __fmaf_rn(y, b, __fmul_rn(x, f)));
		const float uy = __fmaf_rn(c[2], gy, __fmaf_rn(y, f, __fmul_rn(y, a)));
		x = ux;
		y = uy;
	}
	const float p0 = __fmul_rn(depth_val, x), p1 = __fmul_rn(depth_val, y), p2 = depth_val;
	const float* R = A.ex.rotation;
	const float* t = A.ex.translation;
	const float q0 = __fadd_rn(__fmaf_rn(p2, R[6], __fmaf_rn(p0, R[0], __fmul_rn(p1, R[3]))), t[0]);
	const float q1 = __fadd_rn(__fmaf_rn(p2, R[7], __fmaf_rn(p0, R[1], __fmul_rn(p1, R[4]))), t[1]);
	const float q2 = __fadd_rn(__fmaf_rn(p2, R[8], __fmaf_rn(p0, R[2], __fmul_rn(p1, R[5]))), t[2]);
	float px = __fdiv_rn(q0, q2), py = __fdiv_rn(q1, q2);
	if (A.color.model == DFB_DISTORTION_MODIFIED_BROWN_CONRADY) {
		const float* c = A.color.coeffs;
		const float r2 = __fmaf_rn(px, px, __fmul_rn(py, py));
		float f = __fmaf_rn(r2, c[0], 1.0f);
		f = __fmaf_rn(r2, __fmul_rn(r2, c[1]), f);
		f = __fmaf_rn(r2, __fmul_rn(r2, __fmul_rn(r2, c[4])), f);
		const float xf = __fmul_rn(px, f), yf = __fmul_rn(py, f);
		const float ddx = __fmaf_rn(c[3], __fmaf_rn(xf, __fadd_rn(xf, xf), r2), __fmaf_rn(yf, __fmul_rn(xf, __fadd_rn(c[2], c[2])), xf));
		const float ddy = __fmaf_rn(c[2], __fmaf_rn(yf, __fadd_rn(yf, yf), r2), __fmaf_rn(yf, __fmul_rn(xf, __fadd_rn(c[3], c[3])), yf));
		px = ddx;
		py = ddy;
	}
	if (A.color.model == DFB_DISTORTION_FTHETA) { /* kernels.cu:107-112; sqrtf / tanf / atanf are CUDA's own, as in the reference */
		const float c0 = A.color.coeffs[0];
		const float r = sqrtf(__fmaf_rn(px, px, __fmul_rn(py, py)));
		const float rd = __fmul_rn(__fdiv_rn(1.0f, c0), atanf(__fmul_rn(__fmul_rn(2.0f, r), tanf(__fmul_rn(c0, 0.5f)))));
		const float q = __fdiv_rn(rd, r);
		px = __fmul_rn(px, q);
		py = __fmul_rn(py, q);
	}
	int2 m;
	m.x = __float2int_rz(__fadd_rn(__fmaf_rn(A.color.fx, px, A.color.ppx), 0.5f));
	m.y = __float2int_rz(__fadd_rn(__fmaf_rn(A.color.fy, py, A.color.ppy), 0.5f));
	return m;
}

__global__ void __launch_bounds__(256) align_scatter_kernel(const __grid_constant__ AlignParams A, const uint16_t* __restrict__ depth,
                                                            uint32_t* __restrict__ aligned_words) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int n = A.depth.width * A.depth.height;
	if (i >= n) return;
	const int dy = i / A.depth.width, dx = i - dy * A.depth.width;
	const uint32_t d = depth[i];
	const float depth_val = __fmul_rn((float)d, A.depth_scale);
	if (depth_val == 0.0f) return; /* mapped = {-1,-1}: rejected by the p0 test (kernels.cu:143-146,289) */
	const int2 p0 = map_corner(A, depth_val, dx, dy, -0.5f);
	const int2 p1 = map_corner(A, depth_val, dx, dy, 0.5f);
	if (p0.x < 0 || p0.y < 0 || p1.x >= A.color.width || p1.y >= A.color.height) return;
	const uint32_t v = (d << 16) | d;
	for (int y = p0.y; y <= p1.y; ++y)
		for (int x = p0.x; x <= p1.x; ++x) atomicMin(aligned_words + ((size_t)y * A.color.width + x) / 2, v);
}

__global__ void __launch_bounds__(256) align_finish_kernel(uint32_t* __restrict__ words, size_t n_words) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_words) return;
	uint32_t w = words[i];
	if ((w & 0xFFFFu) == 0xFFFFu) w &= 0xFFFF0000u; /* kernel_replace_to_zero, kernels.cu:305-311 */
	if ((w >> 16) == 0xFFFFu) w &= 0x0000FFFFu;
	words[i] = w;
}

}  // namespace
}  // namespace dfb

using namespace dfb;

extern "C" int dfb_op_align_depth_to_color(uint16_t* d_aligned, const uint16_t* d_depth, const dfb_intrinsics* depth_intr,
                                           const dfb_intrinsics* color_intr, const dfb_extrinsics* depth_to_color, float depth_scale,
                                           void* stream) {
	DFB_REQUIRE(d_aligned && d_depth && depth_intr && color_intr && depth_to_color, "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	DFB_REQUIRE(depth_intr->width > 0 && depth_intr->height > 0 && color_intr->width > 0 && color_intr->height > 0, "empty image");
	const size_t n_color = (size_t)color_intr->width * color_intr->height;
	DFB_REQUIRE((n_color & 1u) == 0 && (reinterpret_cast<uintptr_t>(d_aligned) & 3u) == 0,
	            "the aligned image is updated in 32-bit pixel pairs: even pixel count and 4-byte alignment required");
	/* the reference asserts these models away in deprojection (kernels.cu:48-49) */
	DFB_REQUIRE(depth_intr->model != DFB_DISTORTION_MODIFIED_BROWN_CONRADY && depth_intr->model != DFB_DISTORTION_FTHETA,
	            "depth intrinsics: distortion model cannot be deprojected");
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	AlignParams A;
	A.depth = *depth_intr;
	A.color = *color_intr;
	A.ex = *depth_to_color;
	A.depth_scale = depth_scale;
	DFB_CUDA(cudaMemsetAsync(d_aligned, 0xff, n_color * 2, s)); /* kernels.cu:725 */
	const int n_depth = depth_intr->width * depth_intr->height;
	align_scatter_kernel<<<(n_depth + 255) / 256, 256, 0, s>>>(A, d_depth, reinterpret_cast<uint32_t*>(d_aligned));
	count_launch();
	const size_t n_words = n_color / 2;
	align_finish_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint32_t*>(d_aligned), n_words);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}
