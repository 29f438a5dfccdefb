/*
 * cloud_ops.cu — eager operators, one per reference Kernels::* wrapper, for callers that need the
 * intermediate products the fused path never materialises (drop-in Frameset/Pointcloud classes, the
 * registration feed). All arithmetic comes from dfb_math.cuh, so a cloud built here and splatted
 * with dfb_op_project yields exactly the keys the fused kernel produces.
 *
 *   dfb_op_filter_depth    DepthFrame::filter              depth_frame.cpp:153-182, kernels.cu:353-395,501-547
 *   dfb_op_deproject       Kernels::deproject_depth[_color] kernels.cu:183-221,549-568
 *   dfb_op_transform       Kernels::transform              kernels.cu:161-181,570-575
 *   dfb_op_project         Kernels::project_pointcloud     kernels.cu:239-274,577-592
 *   dfb_registration_cloud registration_node.cpp:313-333 (filter + deproject fused, ordered)
 *
 * Ordered stream compaction (replaces compactor.cuh + thrust::exclusive_scan): block counts ->
 * one-block exclusive scan -> scatter; output order = ascending pixel index, as the reference.
 */
#include <new>

#include "dfb_internal.h"

namespace dfb {
namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;                      /* pixels per thread in the compaction kernels */
constexpr int kTile = kThreads * kItems;       /* 2048 pixels per block */

struct FilterParams {
	int width, height;
	int roi_x0, roi_y0, roi_w, roi_h;
	uint32_t lt, gt;
};

__device__ __forceinline__ bool in_roi(const FilterParams& F, int i) {
	const int v = i / F.width;
	const int u = i - v * F.width;
	return (unsigned)(u - F.roi_x0) < (unsigned)F.roi_w && (unsigned)(v - F.roi_y0) < (unsigned)F.roi_h;
}

/* block-wide exclusive scan of one value per thread; returns the exclusive prefix, total in *total */
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
	__shared__ uint32_t warp_sums[kThreads / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) warp_sums[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = (lane < kThreads / 32) ? warp_sums[lane] : 0u;
#pragma unroll
		for (int o = 1; o < kThreads / 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
			if (lane >= o) w += t;
		}
		if (lane < kThreads / 32) warp_sums[lane] = w; /* inclusive */
	}
	__syncthreads();
	const uint32_t warp_off = (warp == 0) ? 0u : warp_sums[warp - 1];
	*total = warp_sums[kThreads / 32 - 1];
	const uint32_t r = warp_off + incl - v;
	__syncthreads();
	return r;
}

/* pass A of DepthFrame::filter: threshold in place (whole image), mask inside the ROI, count */
__global__ void __launch_bounds__(kThreads)
filter_mask_count_kernel(const FilterParams F, uint16_t* __restrict__ depth, uint8_t* __restrict__ mask,
                         uint32_t* __restrict__ block_counts, uint32_t* __restrict__ acc, uint32_t* __restrict__ host_total) {
	const int npix = F.width * F.height;
	const int base = blockIdx.x * kTile + threadIdx.x * kItems;
	uint32_t cnt = 0;
	static_assert(kItems == 8, "the vector form below moves 8 pixels as one uint4 of depth and one uint2 of mask");
	const bool vec = base + kItems <= npix && ((reinterpret_cast<uintptr_t>(depth) & 15u) | (reinterpret_cast<uintptr_t>(mask) & 7u)) == 0;
	if (vec) {
		/* 8 pixels of one thread as ONE 16-byte load (scalar 2-byte loads at a 16-byte thread stride took 7.9 us for a
		 * 720p frame, 2.4 x the time of the fused splat's share of it) */
		const int v = base / F.width, u0 = base - v * F.width;
		const bool one_row = u0 + kItems <= F.width;
		const bool row_in = (unsigned)(v - F.roi_y0) < (unsigned)F.roi_h;
		const bool all_in = one_row && row_in && u0 >= F.roi_x0 && u0 + kItems <= F.roi_x0 + F.roi_w;
		uint4 q = *reinterpret_cast<const uint4*>(depth + base);
		uint32_t w[4] = {q.x, q.y, q.z, q.w};
		uint2 old = make_uint2(0u, 0u);
		if (!all_in) old = *reinterpret_cast<const uint2*>(mask + base); /* stale mask outside the ROI is kept and counted */
		const uint32_t om[2] = {old.x, old.y};
		uint32_t nm[2] = {0u, 0u};
		bool changed = false;
#pragma unroll
		for (int j = 0; j < kItems; ++j) {
			uint32_t d = (w[j >> 1] >> (16 * (j & 1))) & 0xFFFFu;
			if (d < F.lt || d > F.gt) { /* nppiThreshold_LTValGTVal, kernels.cu:366-368 */
				changed |= d != 0u;
				d = 0;
				w[j >> 1] &= (j & 1) ? 0x0000FFFFu : 0xFFFF0000u;
			}
			const bool inside = all_in || (one_row ? (row_in && (unsigned)(u0 + j - F.roi_x0) < (unsigned)F.roi_w) : in_roi(F, base + j));
			const uint32_t m = inside ? (d > 0 ? 255u : 0u) : ((om[j >> 2] >> (8 * (j & 3))) & 0xFFu);
			nm[j >> 2] |= m << (8 * (j & 3));
			cnt += (m > 0);
		}
		if (changed) *reinterpret_cast<uint4*>(depth + base) = make_uint4(w[0], w[1], w[2], w[3]);
		*reinterpret_cast<uint2*>(mask + base) = make_uint2(nm[0], nm[1]);
	} else {
#pragma unroll
		for (int j = 0; j < kItems; ++j) {
			const int i = base + j;
			if (i < npix) {
				uint32_t d = depth[i];
				if (d < F.lt || d > F.gt) { /* nppiThreshold_LTValGTVal, kernels.cu:366-368 */
					d = 0;
					depth[i] = 0;
				}
				uint8_t m;
				if (in_roi(F, i)) { /* nppiCompareC > 0 inside the ROI only, kernels.cu:392-394 */
					m = d > 0 ? 255 : 0;
					mask[i] = m;
				} else {
					m = mask[i]; /* stale mask outside the ROI is kept and counted */
				}
				cnt += (m > 0);
			}
		}
	}
	uint32_t total;
	block_exclusive_scan(cnt, &total);
	if (threadIdx.x == 0) {
		block_counts[blockIdx.x] = total;
		/* mask_count for the host (filter_mask_count_npp, kernels.cu:521-532) without a second pass or a copy: blocks add
		 * to an accumulator, the last one to finish (self-resetting ticket) moves the total into pinned host memory and
		 * clears the accumulator for the next call */
		atomicAdd(acc, total);
		__threadfence();
		if (atomicInc(acc + 1, gridDim.x - 1) == gridDim.x - 1) {
			*host_total = atomicExch(acc, 0u);
			__threadfence_system();
		}
	}
}

/* block counts of an existing mask (the index list is built on demand, after a filter that skipped it) */
__global__ void __launch_bounds__(kThreads)
mask_count_kernel(int npix, const uint8_t* __restrict__ mask, uint32_t* __restrict__ block_counts) {
	const int base = blockIdx.x * kTile + threadIdx.x * kItems;
	uint32_t cnt = 0;
#pragma unroll
	for (int j = 0; j < kItems; ++j) {
		const int i = base + j;
		cnt += (i < npix && mask[i] > 0);
	}
	uint32_t total;
	block_exclusive_scan(cnt, &total);
	if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

/* pass A of the registration cloud: predicate straight from the unmodified depth (fresh DepthFrame) */
__global__ void __launch_bounds__(kThreads)
valid_count_kernel(const FilterParams F, const uint16_t* __restrict__ depth, uint32_t* __restrict__ block_counts) {
	const int npix = F.width * F.height;
	const int base = blockIdx.x * kTile + threadIdx.x * kItems;
	uint32_t cnt = 0;
#pragma unroll
	for (int j = 0; j < kItems; ++j) {
		const int i = base + j;
		if (i < npix) {
			const uint32_t d = depth[i];
			cnt += (d >= F.lt && d <= F.gt && d > 0 && in_roi(F, i));
		}
	}
	uint32_t total;
	block_exclusive_scan(cnt, &total);
	if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

/* counts[0..n) -> exclusive offsets in place, counts[n] = total; one block */
__global__ void __launch_bounds__(kThreads) scan_block_counts_kernel(uint32_t* __restrict__ counts, int n) {
	uint32_t carry = 0;
	for (int base = 0; base < n; base += kThreads) {
		const int i = base + threadIdx.x;
		const uint32_t v = (i < n) ? counts[i] : 0u;
		uint32_t total;
		const uint32_t ex = block_exclusive_scan(v, &total);
		if (i < n) counts[i] = carry + ex;
		carry += total;
	}
	if (threadIdx.x == 0) counts[n] = carry;
}

/* pass B: ascending pixel indices of mask > 0 (compactor::kernel_compact semantics, compactor.cuh:28-68) */
__global__ void __launch_bounds__(kThreads)
compact_indices_kernel(int npix, const uint8_t* __restrict__ mask, const uint32_t* __restrict__ block_offsets,
                       uint32_t* __restrict__ indices) {
	const int base = blockIdx.x * kTile + threadIdx.x * kItems;
	uint32_t flags = 0, cnt = 0;
#pragma unroll
	for (int j = 0; j < kItems; ++j) {
		const int i = base + j;
		if (i < npix && mask[i] > 0) {
			flags |= 1u << j;
			++cnt;
		}
	}
	uint32_t total;
	uint32_t o = block_offsets[blockIdx.x] + block_exclusive_scan(cnt, &total);
#pragma unroll
	for (int j = 0; j < kItems; ++j)
		if (flags & (1u << j)) indices[o++] = (uint32_t)(base + j);
}

struct DeprojParams {
	int width;
	float ppx, ppy, fx, fy, scale;
};

/* pass B of the registration cloud: write {X,Y,Z,1.0f} directly in ascending pixel order */
__global__ void __launch_bounds__(kThreads)
compact_cloud_kernel(const FilterParams F, const DeprojParams D, const uint16_t* __restrict__ depth,
                     const uint32_t* __restrict__ block_offsets, float4* __restrict__ points) {
	const int npix = F.width * F.height;
	const int base = blockIdx.x * kTile + threadIdx.x * kItems;
	uint32_t dv[kItems];
	uint32_t flags = 0, cnt = 0;
#pragma unroll
	for (int j = 0; j < kItems; ++j) {
		const int i = base + j;
		dv[j] = 0;
		if (i < npix) {
			const uint32_t d = depth[i];
			if (d >= F.lt && d <= F.gt && d > 0 && in_roi(F, i)) {
				dv[j] = d;
				flags |= 1u << j;
				++cnt;
			}
		}
	}
	uint32_t total;
	uint32_t o = block_offsets[blockIdx.x] + block_exclusive_scan(cnt, &total);
#pragma unroll
	for (int j = 0; j < kItems; ++j)
		if (flags & (1u << j)) {
			const int i = base + j;
			const int v = i / D.width;
			const int u = i - v * D.width;
			const float z = __fmul_rn(D.scale, (float)dv[j]);
			points[o++] = make_float4(__fmul_rn(z, deproject_ray(u, D.ppx, D.fx)),
			                          __fmul_rn(z, deproject_ray(v, D.ppy, D.fy)), z, 1.0f);
		}
}

/* kernel_deproject_depth (both overloads), kernels.cu:183-221 */
template <bool kColor>
__global__ void __launch_bounds__(kThreads)
deproject_kernel(const DeprojParams D, const uint16_t* __restrict__ depth, const uint8_t* __restrict__ color,
                 const uint32_t* __restrict__ indices, uint32_t count, float4* __restrict__ points) {
	const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= count) return;
	const uint32_t img = indices[i];
	const int v = (int)img / D.width;
	const int u = (int)img - v * D.width;
	const float z = __fmul_rn(D.scale, (float)depth[img]);
	float w = 1.0f; /* kernels.cu:89 */
	if (kColor) {
		/* bytes {B,G,R,0} reinterpreted as float, kernels.cu:198 */
		const uint32_t r = color[(size_t)img * 3], g = color[(size_t)img * 3 + 1], b = color[(size_t)img * 3 + 2];
		w = __uint_as_float(b | (g << 8) | (r << 16));
	}
	points[i] = make_float4(__fmul_rn(z, deproject_ray(u, D.ppx, D.fx)), __fmul_rn(z, deproject_ray(v, D.ppy, D.fy)), z, w);
}

__global__ void __launch_bounds__(kThreads) transform_kernel(const Mat34 T, float4* __restrict__ points, uint32_t count) {
	const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= count) return;
	float4 p = points[i];
	transform_point(p.x, p.y, p.z, T);
	points[i] = p; /* w untouched, kernels.cu:177-180 */
}

struct Segments {
	uint32_t end[DFB_MAX_CAMERAS]; /* exclusive end of each segment */
	int n;
};

__global__ void __launch_bounds__(kThreads)
project_cloud_kernel(const ProjParams P, const Segments S, const float4* __restrict__ points, uint32_t count,
                     unsigned long long* __restrict__ keys) {
	const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
	if (i >= count) return;
	const float4 p = points[i];
	int idx;
	if (!project_point(p.x, p.y, p.z, P, idx)) return;
	uint32_t cam = 0;
#pragma unroll
	for (int s = 0; s < DFB_MAX_CAMERAS - 1; ++s)
		if (s < S.n - 1 && i >= S.end[s]) cam = s + 1;
	/* w holds bytes {B,G,R,0} (kernels.cu:269-273): as a little-endian word that is R<<16|G<<8|B */
	const uint32_t rgb = __float_as_uint(p.w) & 0x00FFFFFFu;
	atomicMin(keys + idx, make_key(p.z, (cam << 24) | rgb));
}

int ensure_blocks(dfb_workspace* ws, size_t nblocks) {
	if (ws->capacity >= nblocks) return DFB_OK;
	if (ws->d_block_counts) DFB_CUDA(cudaFree(ws->d_block_counts));
	ws->d_block_counts = nullptr;
	ws->capacity = 0;
	DFB_CUDA(cudaMalloc(&ws->d_block_counts, (nblocks + 1) * sizeof(uint32_t)));
	ws->capacity = nblocks;
	return DFB_OK;
}

int make_filter_params(int width, int height, float min_depth, float max_depth, float depth_scale, const int32_t roi[4],
                       FilterParams& F) {
	dfb_camera_desc cd{};
	cd.intr.width = width;
	cd.intr.height = height;
	cd.intr.fx = cd.intr.fy = 32768.0f; /* unused by the filter; keeps the plan validation happy */
	cd.depth_scale = depth_scale;
	cd.min_depth = min_depth;
	cd.max_depth = max_depth;
	for (int i = 0; i < 4; ++i) cd.roi[i] = roi ? roi[i] : -1;
	CamParams C;
	int rc = make_cam_params(cd, C);
	if (rc != DFB_OK) return rc;
	F.width = C.width;
	F.height = C.height;
	F.roi_x0 = C.roi_x0;
	F.roi_y0 = C.roi_y0;
	F.roi_w = C.roi_w;
	F.roi_h = C.roi_h;
	F.lt = C.lt;
	F.gt = C.gt;
	return DFB_OK;
}

}  // namespace
}  // namespace dfb

using namespace dfb;

extern "C" {

int dfb_workspace_create(dfb_workspace** out) {
	DFB_REQUIRE(out != nullptr, "NULL argument");
	*out = nullptr;
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	dfb_workspace* ws = new (std::nothrow) dfb_workspace();
	if (!ws) return DFB_ERR_NOMEM;
	cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&ws->h_total), sizeof(uint32_t), cudaHostAllocMapped); /* kernels write the count here */
	if (e != cudaSuccess) {
		delete ws;
		return cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__);
	}
	*out = ws;
	return DFB_OK;
}

int dfb_workspace_destroy(dfb_workspace* ws) {
	if (!ws) return DFB_OK;
	cudaFree(ws->d_block_counts);
	cudaFree(ws->d_count_acc);
	cudaFree(ws->d_voxel_tmp);
	cudaFreeHost(ws->h_total);
	if (ws->count_ready) cudaEventDestroy(ws->count_ready);
	delete ws;
	return DFB_OK;
}

/* launches the filter; the kernel itself delivers the count to pinned host memory, dfb_op_filter_depth_count waits
 * for it. d_indices == NULL skips the index list (two of the three kernels): dfb_op_filter_depth_indices builds it
 * later from the mask if somebody asks. */
int dfb_op_filter_depth_begin(dfb_workspace* ws, uint16_t* d_depth, uint8_t* d_mask, uint32_t* d_indices, int width, int height,
                              float min_depth, float max_depth, float depth_scale, const int32_t roi[4], void* stream) {
	DFB_REQUIRE(ws && d_depth && d_mask, "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	FilterParams F;
	rc = make_filter_params(width, height, min_depth, max_depth, depth_scale, roi, F);
	if (rc != DFB_OK) return rc;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const int npix = width * height;
	const int nblocks = (npix + kTile - 1) / kTile;
	rc = ensure_blocks(ws, nblocks);
	if (rc != DFB_OK) return rc;
	if (!ws->d_count_acc) {
		DFB_CUDA(cudaMalloc(&ws->d_count_acc, 2 * sizeof(uint32_t)));
		DFB_CUDA(cudaMemsetAsync(ws->d_count_acc, 0, 2 * sizeof(uint32_t), s));
	}
	filter_mask_count_kernel<<<nblocks, kThreads, 0, s>>>(F, d_depth, d_mask, ws->d_block_counts, ws->d_count_acc, ws->h_total);
	count_launch();
	if (d_indices) {
		scan_block_counts_kernel<<<1, kThreads, 0, s>>>(ws->d_block_counts, nblocks);
		compact_indices_kernel<<<nblocks, kThreads, 0, s>>>(npix, d_mask, ws->d_block_counts, d_indices);
		count_launch(2);
	}
	DFB_CUDA(cudaGetLastError());
	if (!ws->count_ready) DFB_CUDA(cudaEventCreateWithFlags(&ws->count_ready, cudaEventDisableTiming));
	DFB_CUDA(cudaEventRecord(ws->count_ready, s));
	return DFB_OK;
}

/* the ascending index list of mask > 0 (create_img_indices, kernels.cu:542-547) for a mask written earlier */
int dfb_op_filter_depth_indices(dfb_workspace* ws, const uint8_t* d_mask, uint32_t* d_indices, int width, int height, void* stream) {
	DFB_REQUIRE(ws && d_mask && d_indices, "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const int npix = width * height;
	const int nblocks = (npix + kTile - 1) / kTile;
	rc = ensure_blocks(ws, nblocks);
	if (rc != DFB_OK) return rc;
	mask_count_kernel<<<nblocks, kThreads, 0, s>>>(npix, d_mask, ws->d_block_counts);
	scan_block_counts_kernel<<<1, kThreads, 0, s>>>(ws->d_block_counts, nblocks);
	compact_indices_kernel<<<nblocks, kThreads, 0, s>>>(npix, d_mask, ws->d_block_counts, d_indices);
	count_launch(3);
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

int dfb_op_filter_depth_count(dfb_workspace* ws, uint32_t* mask_count) {
	DFB_REQUIRE(ws && mask_count, "NULL argument");
	DFB_REQUIRE(ws->count_ready != nullptr, "dfb_op_filter_depth_count without dfb_op_filter_depth_begin");
	DFB_CUDA(cudaEventSynchronize(ws->count_ready));
	*mask_count = *ws->h_total;
	return DFB_OK;
}

int dfb_op_filter_depth(dfb_workspace* ws, uint16_t* d_depth, uint8_t* d_mask, uint32_t* d_indices, int width,
                        int height, float min_depth, float max_depth, float depth_scale, const int32_t roi[4],
                        uint32_t* mask_count, void* stream) {
	DFB_REQUIRE(mask_count && d_indices, "NULL argument");
	int rc = dfb_op_filter_depth_begin(ws, d_depth, d_mask, d_indices, width, height, min_depth, max_depth, depth_scale, roi, stream);
	if (rc != DFB_OK) return rc;
	DFB_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); /* the reference synchronises here too, kernels.cu:529 */
	return dfb_op_filter_depth_count(ws, mask_count);
}

int dfb_op_deproject(float* d_points, const uint16_t* d_depth, const uint8_t* d_color, const dfb_intrinsics* intr,
                     const uint32_t* d_indices, uint32_t count, float depth_scale, void* stream) {
	DFB_REQUIRE(d_points && d_depth && intr && d_indices, "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	if (count == 0) return DFB_OK;
	const DeprojParams D{intr->width, intr->ppx, intr->ppy, intr->fx, intr->fy, depth_scale};
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const unsigned blocks = (count + kThreads - 1) / kThreads;
	if (d_color)
		deproject_kernel<true><<<blocks, kThreads, 0, s>>>(D, d_depth, d_color, d_indices, count, reinterpret_cast<float4*>(d_points));
	else
		deproject_kernel<false><<<blocks, kThreads, 0, s>>>(D, d_depth, nullptr, d_indices, count, reinterpret_cast<float4*>(d_points));
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

int dfb_op_transform(float* d_points, uint32_t count, const float T[16], void* stream) {
	DFB_REQUIRE(d_points && T, "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	if (count == 0) return DFB_OK;
	Mat34 M;
	for (int i = 0; i < 12; ++i) M.m[i] = T[i];
	transform_kernel<<<(count + kThreads - 1) / kThreads, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
	    M, reinterpret_cast<float4*>(d_points), count);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

int dfb_op_keys_clear(uint64_t* d_keys, size_t n_pixels, void* stream) {
	DFB_REQUIRE(d_keys != nullptr, "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	return launch_keys_clear(reinterpret_cast<unsigned long long*>(d_keys), n_pixels, static_cast<cudaStream_t>(stream));
}

int dfb_op_project(uint8_t* d_color_out, uint64_t* d_keys, const float* d_points, const uint32_t* seg_counts,
                   int n_segments, const dfb_intrinsics* out_intr, int mirror, void* stream) {
	DFB_REQUIRE(d_color_out && d_keys && out_intr && seg_counts, "NULL argument");
	DFB_REQUIRE(n_segments >= 1 && n_segments <= DFB_MAX_CAMERAS, "n_segments must be in [1, DFB_MAX_CAMERAS]");
	DFB_REQUIRE(out_intr->width > 0 && out_intr->height > 0 && out_intr->width <= 32768 && out_intr->height <= 32768,
	            "image resolution must be in [1, 32768]");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	Segments S{};
	S.n = n_segments;
	uint32_t total = 0;
	for (int i = 0; i < n_segments; ++i) {
		total += seg_counts[i];
		S.end[i] = total;
	}
	FuseParams P{};
	P.n_out_pixels = out_intr->width * out_intr->height;
	P.proj.ppx = out_intr->ppx;
	P.proj.ppy = out_intr->ppy;
	P.proj.fx = out_intr->fx;
	P.proj.fy = out_intr->fy;
	P.proj.wf = (float)out_intr->width;
	P.proj.hf = (float)out_intr->height;
	P.proj.width = out_intr->width;
	P.proj.height = out_intr->height;
	P.proj.mirror = mirror ? 1 : 0;
	if (total > 0) {
		DFB_REQUIRE(d_points != nullptr, "NULL points");
		project_cloud_kernel<<<(total + kThreads - 1) / kThreads, kThreads, 0, s>>>(
		    P.proj, S, reinterpret_cast<const float4*>(d_points), total, reinterpret_cast<unsigned long long*>(d_keys));
		count_launch();
		DFB_CUDA(cudaGetLastError());
	}
	return launch_resolve(P, reinterpret_cast<unsigned long long*>(d_keys), d_color_out, 1, s);
}

int dfb_registration_cloud(dfb_workspace* ws, const dfb_camera_desc* cam, const uint16_t* d_depth, float* d_points,
                           uint32_t* count, void* stream) {
	DFB_REQUIRE(ws && cam && d_depth && d_points && count, "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	FilterParams F;
	rc = make_filter_params(cam->intr.width, cam->intr.height, cam->min_depth, cam->max_depth, cam->depth_scale, cam->roi, F);
	if (rc != DFB_OK) return rc;
	const DeprojParams D{cam->intr.width, cam->intr.ppx, cam->intr.ppy, cam->intr.fx, cam->intr.fy, cam->depth_scale};
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const int npix = F.width * F.height;
	const int nblocks = (npix + kTile - 1) / kTile;
	rc = ensure_blocks(ws, nblocks);
	if (rc != DFB_OK) return rc;
	valid_count_kernel<<<nblocks, kThreads, 0, s>>>(F, d_depth, ws->d_block_counts);
	scan_block_counts_kernel<<<1, kThreads, 0, s>>>(ws->d_block_counts, nblocks);
	compact_cloud_kernel<<<nblocks, kThreads, 0, s>>>(F, D, d_depth, ws->d_block_counts, reinterpret_cast<float4*>(d_points));
	count_launch(3);
	DFB_CUDA(cudaGetLastError());
	DFB_CUDA(cudaMemcpyAsync(ws->h_total, ws->d_block_counts + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	DFB_CUDA(cudaStreamSynchronize(s));
	*count = *ws->h_total;
	return DFB_OK;
}

}  // extern "C"
