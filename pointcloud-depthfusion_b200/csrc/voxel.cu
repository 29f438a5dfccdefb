/*
 * voxel.cu — GPU voxel grid for the registration feed, replacing Registration::voxelFilter
 * (registration_node/src/registration.cpp:144-150: pcl::ApproximateVoxelGrid<PointXYZ>, leaf = voxelsize).
 *
 * PCL's approximate grid is a sequential hash-history filter whose output depends on the visiting order and on
 * hash collisions (it may emit several centroids per voxel); what GICP needs from it is "about one point per
 * occupied voxel". This operator computes exactly that, deterministically: ONE centroid per occupied voxel,
 * voxel index (ix,iy,iz) = floor(p * (1/leaf)) in fp32 like PCL, centroid = double-precision mean of the voxel's
 * points summed in ascending source order, output sorted by (iz,iy,ix). The oracle's dfo_voxel_exact is the same
 * definition; dfo_voxel_approximate (PCL's algorithm restated) covers the same voxel set (tests/test_ops_gpu.py).
 *
 *   1. voxel_box_kernel    bounding box of the occupied voxel indices (one host read: it sizes the sort)
 *   2. voxel_keys_kernel   compact key = the rank of (iz,iy,ix) inside that box, payload = source index
 *   3. LSD radix sort of (key, index) pairs, 8 bits per pass, only as many passes as the box needs bits (3 for a
 *      room at 5 cm): per pass a stable block-local rank + digit histogram, one scan, one scatter. Stable, so equal
 *      keys stay in ascending source order. 12-14 launches for 2^20 points (the bitonic network this replaces: 65).
 *   4. run starts -> ordered compaction (count per block, scan of the block counts, emit) -> one thread per voxel
 *      walks its run and writes the centroid.
 * HBM-bound integer / index work: no tensor cores, coalesced 8-byte and 4-byte streams.
 */
#include "dfb_internal.h"

namespace dfb {
namespace {

constexpr int kRadixThreads = 1024; /* keys per block of a radix pass */
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr int kBias = 1 << 20; /* voxel indices must lie in [-2^20, 2^20) */

struct VoxelBox { /* device-side bounding box of the voxel indices + range flag */
	int lo[3], hi[3];
	uint32_t overflow;
};

__device__ __forceinline__ void voxel_index(const float4 p, float inv, int& ix, int& iy, int& iz) {
	/* approximate_voxel_grid.hpp: ijk = floor(p * inverse_leaf_size), in float */
	ix = (int)floorf(__fmul_rn(p.x, inv));
	iy = (int)floorf(__fmul_rn(p.y, inv));
	iz = (int)floorf(__fmul_rn(p.z, inv));
}

/* pass 0: bounding box of the occupied voxels (one atomic pair per warp and axis) */
__global__ void voxel_box_kernel(const float4* __restrict__ points, uint32_t n, float inv, VoxelBox* __restrict__ box) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	int v[3] = {0, 0, 0};
	const bool live = i < n;
	if (live) {
		const float4 p = __ldg(points + i);
		voxel_index(p, inv, v[0], v[1], v[2]);
		const float lim = (float)kBias;
		const float fx = __fmul_rn(p.x, inv), fy = __fmul_rn(p.y, inv), fz = __fmul_rn(p.z, inv);
		/* written so that NaN fails: a non-finite coordinate is out of range too */
		if (!(fx >= -lim && fx < lim && fy >= -lim && fy < lim && fz >= -lim && fz < lim)) atomicOr(&box->overflow, 1u);
	}
	const uint32_t mask = __ballot_sync(0xFFFFFFFFu, live);
	if (!live) return;
#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const int lo = __reduce_min_sync(mask, v[a]), hi = __reduce_max_sync(mask, v[a]);
		if ((threadIdx.x & 31) == __ffs(mask) - 1) {
			atomicMin(&box->lo[a], lo);
			atomicMax(&box->hi[a], hi);
		}
	}
}

/* compact key = ((iz - z0) * ny + (iy - y0)) * nx + (ix - x0): the order of (iz, iy, ix), in as few bits as the
 * occupied box needs (a 6 m room at 5 cm is 21 bits = three 8-bit passes instead of eight) */
__global__ void voxel_keys_kernel(const float4* __restrict__ points, uint32_t n, float inv, int x0, int y0, int z0,
                                  unsigned long long nx, unsigned long long ny, unsigned long long* __restrict__ keys,
                                  uint32_t* __restrict__ idx) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int ix, iy, iz;
	voxel_index(__ldg(points + i), inv, ix, iy, iz);
	keys[i] = ((unsigned long long)(iz - z0) * ny + (unsigned long long)(iy - y0)) * nx + (unsigned long long)(ix - x0);
	idx[i] = i;
}

/* one LSD radix pass, step 1: digit of every key, its STABLE rank among the equal digits of its block (warp match,
 * one counter row per warp, prefix over the warps in order) and the block's digit histogram, digit-major:
 * hist[digit * n_blocks + block] */
__global__ void __launch_bounds__(kRadixThreads)
radix_rank_kernel(const unsigned long long* __restrict__ keys, uint32_t n, int shift, uint16_t* __restrict__ local_rank,
                  uint32_t* __restrict__ hist, uint32_t n_blocks) {
	__shared__ uint16_t wcount[kRadixThreads / 32][kRadix];
	const uint32_t i = blockIdx.x * kRadixThreads + threadIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int d = lane; d < kRadix; d += 32) wcount[warp][d] = 0;
	const bool live = i < n;
	const uint32_t d = live ? (uint32_t)(keys[i] >> shift) & (kRadix - 1) : (uint32_t)kRadix; /* dead lanes match nobody live */
	const uint32_t same = __match_any_sync(0xFFFFFFFFu, d);
	const uint32_t before = __popc(same & ((1u << lane) - 1u));
	__syncwarp();
	if (live && before == 0) wcount[warp][d] = (uint16_t)__popc(same);
	__syncthreads();
	if (threadIdx.x < kRadix) { /* exclusive prefix over the warps, per digit */
		uint32_t run = 0;
#pragma unroll 8
		for (int w = 0; w < kRadixThreads / 32; ++w) {
			const uint32_t c = wcount[w][threadIdx.x];
			wcount[w][threadIdx.x] = (uint16_t)run;
			run += c;
		}
		hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = run;
	}
	__syncthreads();
	if (live) local_rank[i] = (uint16_t)(wcount[warp][d] + before);
}

/* in-place exclusive scan of data[0..n) by one block */
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ data, uint32_t n) {
	__shared__ uint32_t warp_sums[32];
	__shared__ uint32_t carry_s;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	/* four consecutive words per thread: 4096 per round */
	for (uint32_t base = 0; base < n; base += 4096) {
		const uint32_t i = base + threadIdx.x * 4;
		uint32_t v[4];
#pragma unroll
		for (int e = 0; e < 4; ++e) v[e] = i + e < n ? data[i + e] : 0u;
		const uint32_t mine = v[0] + v[1] + v[2] + v[3];
		uint32_t x = mine;
#pragma unroll
		for (int dd = 1; dd < 32; dd <<= 1) {
			const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, dd);
			if (lane >= dd) x += y;
		}
		if (lane == 31) warp_sums[warp] = x;
		__syncthreads();
		if (warp == 0) {
			uint32_t s = warp_sums[lane];
#pragma unroll
			for (int dd = 1; dd < 32; dd <<= 1) {
				const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, dd);
				if (lane >= dd) s += y;
			}
			warp_sums[lane] = s; /* inclusive */
		}
		__syncthreads();
		uint32_t excl = carry_s + (warp ? warp_sums[warp - 1] : 0u) + x - mine;
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			if (i + e < n) data[i + e] = excl;
			excl += v[e];
		}
		__syncthreads();
		if (threadIdx.x == 0) carry_s += warp_sums[31];
		__syncthreads();
	}
}

/* step 3: every pair to its place */
__global__ void radix_scatter_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ idx,
                                     const uint16_t* __restrict__ local_rank, const uint32_t* __restrict__ hist, uint32_t n,
                                     uint32_t n_blocks, int shift, unsigned long long* __restrict__ keys_out,
                                     uint32_t* __restrict__ idx_out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const unsigned long long k = keys[i];
	const uint32_t d = (uint32_t)(k >> shift) & (kRadix - 1);
	const uint32_t o = hist[(size_t)d * n_blocks + i / kRadixThreads] + local_rank[i];
	keys_out[o] = k;
	idx_out[o] = idx[i];
}

__device__ __forceinline__ uint32_t scan_block_exclusive(uint32_t v, uint32_t* total) {
	__shared__ uint32_t warp_sums[kScanThreads / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= d) incl += o;
	}
	if (lane == 31) warp_sums[warp] = incl;
	__syncthreads();
	uint32_t before = 0, all = 0;
#pragma unroll
	for (int w = 0; w < kScanThreads / 32; ++w) {
		const uint32_t s = warp_sums[w];
		if (w < warp) before += s;
		all += s;
	}
	__syncthreads();
	*total = all;
	return before + incl - v;
}

__device__ __forceinline__ bool run_start(const unsigned long long* keys, uint32_t i, uint32_t n) {
	return i < n && (i == 0 || keys[i] != keys[i - 1]);
}

__global__ void __launch_bounds__(kScanThreads) voxel_count_kernel(const unsigned long long* __restrict__ keys, uint32_t n,
                                                                     uint32_t* __restrict__ block_counts) {
	const uint32_t i0 = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
	uint32_t c = 0;
#pragma unroll
	for (int e = 0; e < kScanItems; ++e) c += run_start(keys, i0 + e, n) ? 1u : 0u;
	uint32_t total;
	scan_block_exclusive(c, &total);
	if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

/* exclusive scan of the block counts in place (one block); block_counts[n_blocks] = total */
__global__ void __launch_bounds__(kScanThreads) voxel_scan_kernel(uint32_t* __restrict__ block_counts, uint32_t n_blocks) {
	uint32_t carry = 0;
	for (uint32_t b0 = 0; b0 < n_blocks; b0 += kScanThreads) {
		const uint32_t b = b0 + threadIdx.x;
		const uint32_t v = b < n_blocks ? block_counts[b] : 0u;
		uint32_t total;
		const uint32_t ex = scan_block_exclusive(v, &total);
		if (b < n_blocks) block_counts[b] = carry + ex;
		carry += total;
		__syncthreads();
	}
	if (threadIdx.x == 0) block_counts[n_blocks] = carry;
}

__global__ void __launch_bounds__(kScanThreads)
voxel_emit_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ idx, const float4* __restrict__ points,
                  uint32_t n, const uint32_t* __restrict__ block_offsets, float4* __restrict__ out) {
	const uint32_t i0 = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
	bool flag[kScanItems];
	uint32_t c = 0;
#pragma unroll
	for (int e = 0; e < kScanItems; ++e) {
		flag[e] = run_start(keys, i0 + e, n);
		c += flag[e] ? 1u : 0u;
	}
	uint32_t total;
	uint32_t o = block_offsets[blockIdx.x] + scan_block_exclusive(c, &total);
#pragma unroll
	for (int e = 0; e < kScanItems; ++e) {
		if (!flag[e]) continue;
		const uint32_t first = i0 + e;
		const unsigned long long key = keys[first];
		double sx = 0.0, sy = 0.0, sz = 0.0;
		uint32_t j = first;
		for (; j < n && keys[j] == key; ++j) { /* ascending source index inside the voxel (the sort's tie-break) */
			const float4 p = __ldg(points + idx[j]);
			sx += (double)p.x;
			sy += (double)p.y;
			sz += (double)p.z;
		}
		const double cnt = (double)(j - first);
		out[o++] = make_float4((float)(sx / cnt), (float)(sy / cnt), (float)(sz / cnt), 1.0f);
	}
}

}  // namespace
}  // namespace dfb

using namespace dfb;

extern "C" int dfb_voxel_downsample(dfb_workspace* ws, const float* d_points, uint32_t n, float leaf, float* d_out,
                                    uint32_t* n_out, void* stream) {
	DFB_REQUIRE(ws && n_out, "NULL argument");
	DFB_REQUIRE(leaf > 0.0f && leaf == leaf, "leaf size must be positive");
	DFB_REQUIRE(n <= (1u << 27), "too many points for one voxel grid call");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	*n_out = 0;
	if (n == 0) return DFB_OK;
	DFB_REQUIRE(d_points && d_out, "NULL argument");
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const uint32_t n_blocks = (n + kScanTile - 1) / kScanTile;
	const uint32_t r_blocks = (n + kRadixThreads - 1) / kRadixThreads;
	/* workspace: keys u64[2][n] | idx u32[2][n] | block offsets u32[n_blocks + 2] | digit histograms u32[256 * r_blocks]
	 * | local ranks u16[n] | box */
	const size_t n_even = ((size_t)n + 1) & ~(size_t)1;
	const size_t need = n_even * 24 + ((size_t)n_blocks + 2 + (size_t)kRadix * r_blocks) * 4 + n_even * 2 + 64;
	if (ws->voxel_tmp_bytes < need) {
		DFB_CUDA(cudaStreamSynchronize(s));
		cudaFree(ws->d_voxel_tmp);
		ws->d_voxel_tmp = nullptr;
		ws->voxel_tmp_bytes = 0;
		DFB_CUDA(cudaMalloc(&ws->d_voxel_tmp, need));
		ws->voxel_tmp_bytes = need;
	}
	unsigned long long* keys_a = static_cast<unsigned long long*>(ws->d_voxel_tmp);
	unsigned long long* keys_b = keys_a + n_even;
	uint32_t* idx_a = reinterpret_cast<uint32_t*>(keys_b + n_even);
	uint32_t* idx_b = idx_a + n_even;
	uint32_t* offsets = idx_b + n_even;
	uint32_t* hist = offsets + n_blocks + 2;
	uint16_t* local_rank = reinterpret_cast<uint16_t*>(hist + (size_t)kRadix * r_blocks);
	VoxelBox* box = reinterpret_cast<VoxelBox*>(reinterpret_cast<char*>(local_rank + n_even) + ((8 - (n_even * 2) % 8) % 8));
	const float4* pts = reinterpret_cast<const float4*>(d_points);
	const float inv = 1.0f / leaf; /* inverse_leaf_size_, float like PCL */

	VoxelBox hb;
	for (int a = 0; a < 3; ++a) {
		hb.lo[a] = INT32_MAX;
		hb.hi[a] = INT32_MIN;
	}
	hb.overflow = 0;
	DFB_CUDA(cudaMemcpyAsync(box, &hb, sizeof(hb), cudaMemcpyHostToDevice, s));
	voxel_box_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts, n, inv, box);
	count_launch();
	DFB_CUDA(cudaMemcpyAsync(&hb, box, sizeof(hb), cudaMemcpyDeviceToHost, s));
	DFB_CUDA(cudaStreamSynchronize(s));
	DFB_REQUIRE(hb.overflow == 0, "voxel index out of range (|p / leaf| >= 2^20, or a non-finite coordinate)");
	const unsigned long long nx = (unsigned long long)((long long)hb.hi[0] - hb.lo[0] + 1), ny = (unsigned long long)((long long)hb.hi[1] - hb.lo[1] + 1),
	                         nz = (unsigned long long)((long long)hb.hi[2] - hb.lo[2] + 1);
	int bits = 0; /* every axis spans <= 2^21 cells: the product stays below 2^63 */
	while (bits < 63 && (1ull << bits) < nx * ny * nz) ++bits;
	voxel_keys_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts, n, inv, hb.lo[0], hb.lo[1], hb.lo[2], nx, ny, keys_a, idx_a);
	count_launch();
	unsigned long long* keys = keys_a;
	uint32_t* idx = idx_a;
	for (int shift = 0; shift < bits; shift += kRadixBits) {
		unsigned long long* keys_o = keys == keys_a ? keys_b : keys_a;
		uint32_t* idx_o = idx == idx_a ? idx_b : idx_a;
		radix_rank_kernel<<<r_blocks, kRadixThreads, 0, s>>>(keys, n, shift, local_rank, hist, r_blocks);
		radix_scan_kernel<<<1, 1024, 0, s>>>(hist, (uint32_t)kRadix * r_blocks);
		radix_scatter_kernel<<<(n + 255) / 256, 256, 0, s>>>(keys, idx, local_rank, hist, n, r_blocks, shift, keys_o, idx_o);
		count_launch();
		count_launch();
		count_launch();
		keys = keys_o;
		idx = idx_o;
	}
	voxel_count_kernel<<<n_blocks, kScanThreads, 0, s>>>(keys, n, offsets);
	voxel_scan_kernel<<<1, kScanThreads, 0, s>>>(offsets, n_blocks);
	voxel_emit_kernel<<<n_blocks, kScanThreads, 0, s>>>(keys, idx, pts, n, offsets, reinterpret_cast<float4*>(d_out));
	count_launch();
	count_launch();
	count_launch();
	DFB_CUDA(cudaGetLastError());
	uint32_t h[1] = {0};
	DFB_CUDA(cudaMemcpyAsync(&h[0], offsets + n_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	DFB_CUDA(cudaStreamSynchronize(s));
	*n_out = h[0];
	return DFB_OK;
}
