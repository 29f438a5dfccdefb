/*
 * fuse_direct.cu — first-generation fused kernels of the hot path (sm_100a):
 *
 *   splat_direct_kernel : filter -> deproject -> [T_cam] -> T_fused -> project -> 64-bit atomicMin.
 *       One pass over the input pixels; the intermediate float4 cloud of the reference
 *       (kernel_deproject_depth -> kernel_transform x2 -> append -> kernel_project_pointcloud,
 *       kernels.cu:161-274) is never written to memory. Each thread owns 4 consecutive pixels
 *       (one 8-byte depth load + three 4-byte colour loads, coalesced).
 *   resolve_kernel      : key -> rgb8 (background 0,0,0) and re-arm the key to all-ones, replacing the
 *       reference's two memsets + racy colour store (kernels.cu:582-587, 264-273).
 *
 * The tiled TMA variant (fuse_tiled.cu) shares dfb_math.cuh and must produce identical keys.
 */
#include "dfb_internal.h"

namespace dfb {

namespace {

constexpr int kThreads = 256;
constexpr int kPixPerThread = 4;

__device__ __forceinline__ void splat_one(const FuseParams& P, const CamParams& C, int cam, uint32_t d, uint32_t rgb,
                                          float xr, float yr, unsigned long long* __restrict__ keys) {
	/* kernels.cu:202 z = depth_scale * (float)d ; :75-77 X = z*x, Y = z*y, Z = z */
	const float z = __fmul_rn(C.scale, (float)d);
	float X = __fmul_rn(z, xr);
	float Y = __fmul_rn(z, yr);
	float Z = z;
	if (C.has_transform) transform_point(X, Y, Z, C.to_ref); /* fusion_node.cpp:757 */
	transform_point(X, Y, Z, P.fused);                        /* fusion_node.cpp:779 */
	int idx;
	if (!project_point(X, Y, Z, P.proj, idx)) return;
	atomicMin(keys + idx, make_key(Z, ((uint32_t)cam << 24) | rgb));
}

template <bool kVec>
__global__ void __launch_bounds__(kThreads)
splat_direct_kernel(const __grid_constant__ FuseParams P, const uint16_t* __restrict__ depth,
                    const uint8_t* __restrict__ color, unsigned long long* __restrict__ keys) {
	const int cam = blockIdx.y;
	const int fs = blockIdx.z;
	const CamParams& C = P.cam[cam];
	const int npix = C.width * C.height;
	const int i0 = (blockIdx.x * kThreads + threadIdx.x) * kPixPerThread;
	if (i0 >= npix) return;
	const size_t img = (size_t)fs * P.n_cameras + cam;
	const uint16_t* dp = depth + img * (size_t)npix + i0;
	const uint8_t* cp = color + (img * (size_t)npix + i0) * 3;
	unsigned long long* k = keys + (size_t)fs * P.n_out_pixels;

	uint32_t d[kPixPerThread];
	uint32_t rgb[kPixPerThread]; /* R<<16 | G<<8 | B */
	const int n = min(kPixPerThread, npix - i0);
	if (kVec) { /* npix % 4 == 0 and 16-byte aligned pools: n == 4 */
		const uint2 dd = __ldg(reinterpret_cast<const uint2*>(dp));
		const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(cp));
		const uint32_t w1 = __ldg(reinterpret_cast<const uint32_t*>(cp) + 1);
		const uint32_t w2 = __ldg(reinterpret_cast<const uint32_t*>(cp) + 2);
		d[0] = dd.x & 0xFFFFu;
		d[1] = dd.x >> 16;
		d[2] = dd.y & 0xFFFFu;
		d[3] = dd.y >> 16;
		/* bytes in memory: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3 */
		rgb[0] = __byte_perm(w0, 0, 0x4012);
		rgb[1] = __byte_perm(__byte_perm(w0, w1, 0x3345), 0, 0x4210);
		rgb[2] = __byte_perm(__byte_perm(w1, w2, 0x2234), 0, 0x4210);
		rgb[3] = __byte_perm(w2, 0, 0x4123);
	} else {
#pragma unroll
		for (int j = 0; j < kPixPerThread; ++j) {
			d[j] = 0;
			rgb[j] = 0;
			if (j < n) {
				d[j] = dp[j];
				rgb[j] = ((uint32_t)cp[3 * j] << 16) | ((uint32_t)cp[3 * j + 1] << 8) | (uint32_t)cp[3 * j + 2];
			}
		}
	}

	int v = i0 / C.width;
	int u = i0 - v * C.width;
	float yr = deproject_ray(v, C.ppy, C.fy);
#pragma unroll
	for (int j = 0; j < kPixPerThread; ++j) {
		if (j < n) {
			/* DepthFrame::filter folded in: keep iff lt <= d <= gt, d > 0, inside the ROI
			 * (kernels.cu:357-368, 379-394) */
			const bool keep = d[j] >= C.lt && d[j] <= C.gt && d[j] > 0u &&
			                  (unsigned)(u - C.roi_x0) < (unsigned)C.roi_w && (unsigned)(v - C.roi_y0) < (unsigned)C.roi_h;
			if (keep) splat_one(P, C, cam, d[j], rgb[j], deproject_ray(u, C.ppx, C.fx), yr, k);
		}
		if (++u == C.width) {
			u = 0;
			++v;
			yr = deproject_ray(v, C.ppy, C.fy);
		}
	}
}

__global__ void __launch_bounds__(kThreads)
resolve_kernel(unsigned long long* __restrict__ keys, uint8_t* __restrict__ out, size_t n_pixels) {
	const size_t i0 = ((size_t)blockIdx.x * kThreads + threadIdx.x) * 4;
	if (i0 >= n_pixels) return;
	if (i0 + 4 <= n_pixels) {
		ulonglong2* kp = reinterpret_cast<ulonglong2*>(keys + i0);
		const ulonglong2 a = kp[0];
		const ulonglong2 b = kp[1];
		const ulonglong2 empty = make_ulonglong2(KEY_EMPTY, KEY_EMPTY);
		kp[0] = empty;
		kp[1] = empty;
		/* low 24 bits of a key are R<<16|G<<8|B; empty pixel -> background 0 (kernels.cu:582) */
		const uint32_t c0 = (a.x == KEY_EMPTY) ? 0u : (uint32_t)a.x;
		const uint32_t c1 = (a.y == KEY_EMPTY) ? 0u : (uint32_t)a.y;
		const uint32_t c2 = (b.x == KEY_EMPTY) ? 0u : (uint32_t)b.x;
		const uint32_t c3 = (b.y == KEY_EMPTY) ? 0u : (uint32_t)b.y;
		/* output bytes: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3 ; cX bytes: [B, G, R, cam] */
		const uint32_t w0 = __byte_perm(c0, c1, 0x6012);
		const uint32_t w1 = __byte_perm(c1, c2, 0x5601);
		const uint32_t w2 = __byte_perm(c2, c3, 0x4560);
		uint32_t* op = reinterpret_cast<uint32_t*>(out + i0 * 3);
		op[0] = w0;
		op[1] = w1;
		op[2] = w2;
	} else {
		for (size_t i = i0; i < n_pixels; ++i) {
			const unsigned long long key = keys[i];
			keys[i] = KEY_EMPTY;
			const uint32_t c = (key == KEY_EMPTY) ? 0u : (uint32_t)key;
			out[i * 3 + 0] = (uint8_t)(c >> 16);
			out[i * 3 + 1] = (uint8_t)(c >> 8);
			out[i * 3 + 2] = (uint8_t)c;
		}
	}
}

__global__ void __launch_bounds__(kThreads) keys_clear_kernel(unsigned long long* __restrict__ keys, size_t n) {
	const size_t stride = (size_t)gridDim.x * kThreads;
	for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) keys[i] = KEY_EMPTY;
}

}  // namespace

int launch_splat_direct(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color,
                        unsigned long long* d_keys, int group, cudaStream_t s) {
	const int npix = P.cam[0].width * P.cam[0].height;
	const dim3 grid((npix + kThreads * kPixPerThread - 1) / (kThreads * kPixPerThread), P.n_cameras, group);
	const bool vec = (npix % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_depth) & 7u) == 0) &&
	                 ((reinterpret_cast<uintptr_t>(d_color) & 3u) == 0);
	if (vec)
		splat_direct_kernel<true><<<grid, kThreads, 0, s>>>(P, d_depth, d_color, d_keys);
	else
		splat_direct_kernel<false><<<grid, kThreads, 0, s>>>(P, d_depth, d_color, d_keys);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

int launch_resolve(const FuseParams& P, unsigned long long* d_keys, uint8_t* d_out, int group, cudaStream_t s) {
	const size_t n = (size_t)P.n_out_pixels * group;
	const unsigned blocks = (unsigned)((n + kThreads * 4 - 1) / (kThreads * 4));
	resolve_kernel<<<blocks, kThreads, 0, s>>>(d_keys, d_out, n);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

int launch_keys_clear(unsigned long long* d_keys, size_t n, cudaStream_t s) {
	if (n == 0) return DFB_OK;
	size_t nb = (n + kThreads - 1) / kThreads;
	if (nb > 148 * 8) nb = 148 * 8;
	const unsigned blocks = (unsigned)nb;
	keys_clear_kernel<<<blocks, kThreads, 0, s>>>(d_keys, n);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

}  // namespace dfb
