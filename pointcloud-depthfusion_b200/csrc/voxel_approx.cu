/*
 * voxel_approx.cu — pcl::ApproximateVoxelGrid<PointXYZ> (PCL 1.12, approximate_voxel_grid.hpp: the filter behind
 * Registration::voxelFilter, registration_node/src/registration.cpp:144-150) with ITS semantics, in parallel.
 *
 * PCL walks the cloud once with a history table of 512 entries: a point goes to entry
 *     hash = (ix*7171 + iy*3079 + iz*4231) & 511,      (ix,iy,iz) = floor(p * (1/leaf)) in float;
 * if the entry holds another voxel, its centroid (float sum in visiting order / count) is emitted and the entry
 * restarts; at the end the non-empty entries are emitted in table order. The output therefore depends on the
 * visiting order and may hold several centroids of one voxel. It is "sequential" only per table entry: the 512
 * entries are independent state machines, and an entry's life is a sequence of RUNS (maximal stretches of its points
 * with one voxel). Every run is one output point; a run that is ended by point j is emitted when PCL reaches j, so
 * the output order is: ended runs by the index of the point that ended them, then the open runs by table entry.
 *
 *   1. rank_kernel     voxel + hash per point; stable rank of the point among the points of its entry inside its block
 *                      (warp match + one shared counter per entry, warps in order) and the block's entry histogram
 *   2. scan (scan.cu)  exclusive scan of the histograms, entry-major: where each (entry, block) starts
 *   3. scatter_kernel  order[] = point indices grouped by entry, input order kept inside an entry
 *   4. heads_kernel    run starts; ender[j] = 1 when point j ends a run
 *   5. scan            exclusive scan of ender[] = output slot of the run that j ends; ranks of the non-empty entries
 *   6. emit_kernel     one thread per run: float sums in visiting order (bit-identical to PCL's accumulation), / count
 * No sort, 9 launches whatever the cloud size. Integer / index work bound by HBM and launch latency; no tensor cores.
 */
#include "dfb_internal.h"

#include <algorithm>

namespace dfb {
namespace {

constexpr int kHist = 512;       /* PCL: histsize_ */
constexpr int kRankThreads = 1024;

struct VoxelOf {
	int ix, iy, iz;
};

__device__ __forceinline__ VoxelOf voxel_of(const float4 p, float inv) {
	VoxelOf v;
	v.ix = (int)floorf(__fmul_rn(p.x, inv));
	v.iy = (int)floorf(__fmul_rn(p.y, inv));
	v.iz = (int)floorf(__fmul_rn(p.z, inv));
	return v;
}

__device__ __forceinline__ uint32_t hash_of(const VoxelOf& v) {
	return (uint32_t)(v.ix * 7171 + v.iy * 3079 + v.iz * 4231) & (kHist - 1);
}

/* one block = 1024 consecutive points. hist is entry-major: hist[entry * n_blocks + block] */
__global__ void __launch_bounds__(kRankThreads)
rank_kernel(const float4* __restrict__ points, uint32_t n, float inv, uint16_t* __restrict__ entry_of,
            uint16_t* __restrict__ local_rank, uint32_t* __restrict__ hist, uint32_t n_blocks) {
	__shared__ uint32_t count[kHist];
	const uint32_t i = blockIdx.x * kRankThreads + threadIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int b = threadIdx.x; b < kHist; b += kRankThreads) count[b] = 0;
	const bool live = i < n;
	uint32_t e = kHist; /* dead lanes: an entry nobody else has */
	if (live) e = hash_of(voxel_of(__ldg(points + i), inv));
	const uint32_t same = __match_any_sync(0xFFFFFFFFu, e);
	const uint32_t before = __popc(same & ((1u << lane) - 1u));
	const int leader = __ffs(same) - 1;
	uint32_t base = 0;
	__syncthreads();
	for (int w = 0; w < kRankThreads / 32; ++w) { /* warps in order: the rank is STABLE */
		if (warp == w && live && lane == leader) {
			base = count[e];
			count[e] = base + __popc(same);
		}
		__syncthreads();
	}
	base = __shfl_sync(0xFFFFFFFFu, base, leader);
	if (live) {
		entry_of[i] = (uint16_t)e;
		local_rank[i] = (uint16_t)(base + before);
	}
	for (int b = threadIdx.x; b < kHist; b += kRankThreads) hist[(size_t)b * n_blocks + blockIdx.x] = count[b];
}

__global__ void scatter_kernel(const uint16_t* __restrict__ entry_of, const uint16_t* __restrict__ local_rank,
                               const uint32_t* __restrict__ hist, uint32_t n, uint32_t n_blocks, uint32_t* __restrict__ order) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	order[hist[(size_t)entry_of[i] * n_blocks + i / kRankThreads] + local_rank[i]] = i;
}

/* position p of the grouped order: does a run start here? The point that starts a run — unless it is the first of its
 * entry — ENDS the previous run of that entry: PCL emits that run when it reaches this point */
__global__ void heads_kernel(const float4* __restrict__ points, const uint32_t* __restrict__ order, const uint16_t* __restrict__ entry_of,
                             const uint32_t* __restrict__ entry_begin, uint32_t n, float inv, uint8_t* __restrict__ head,
                             uint32_t* __restrict__ ender) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n) return;
	const uint32_t i = order[p];
	const bool first = p == entry_begin[entry_of[i]];
	bool h = first;
	if (!first) {
		const VoxelOf a = voxel_of(__ldg(points + order[p - 1]), inv), b = voxel_of(__ldg(points + i), inv);
		h = a.ix != b.ix || a.iy != b.iy || a.iz != b.iz;
	}
	head[p] = h ? 1 : 0;
	ender[i] = (h && !first) ? 1u : 0u;
}

/* ranks of the non-empty table entries (the open runs are emitted in table order); one block of kHist threads */
__global__ void __launch_bounds__(kHist)
entry_ranks_kernel(const uint32_t* __restrict__ entry_begin, uint32_t* __restrict__ entry_rank, const uint32_t* __restrict__ n_ended,
                   uint32_t* __restrict__ n_out) {
	__shared__ uint32_t sums[kHist / 32];
	const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
	const uint32_t v = entry_begin[b + 1] > entry_begin[b] ? 1u : 0u;
	uint32_t x = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
		if (lane >= d) x += y;
	}
	if (lane == 31) sums[warp] = x;
	__syncthreads();
	uint32_t carry = 0;
	for (int w = 0; w < warp; ++w) carry += sums[w];
	entry_rank[b] = carry + x - v;
	if (b == kHist - 1) *n_out = *n_ended + carry + x;
}

__global__ void emit_kernel(const float4* __restrict__ points, const uint32_t* __restrict__ order, const uint16_t* __restrict__ entry_of,
                            const uint32_t* __restrict__ entry_begin, const uint8_t* __restrict__ head,
                            const uint32_t* __restrict__ ender_slot, const uint32_t* __restrict__ entry_rank,
                            const uint32_t* __restrict__ n_ended, uint32_t n, float4* __restrict__ out) {
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= n || !head[p]) return;
	const uint32_t e = entry_of[order[p]], end = entry_begin[e + 1];
	float sx = 0.f, sy = 0.f, sz = 0.f; /* hhe->centroid += scratch, float, in visiting order */
	uint32_t q = p, count = 0;
	do {
		const float4 v = __ldg(points + order[q]);
		sx = __fadd_rn(sx, v.x);
		sy = __fadd_rn(sy, v.y);
		sz = __fadd_rn(sz, v.z);
		++count;
		++q;
	} while (q < end && !head[q]);
	const uint32_t slot = q < end ? ender_slot[order[q]] : *n_ended + entry_rank[e];
	const float c = (float)count;
	out[slot] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), 1.0f); /* centroid /= count */
}

}  // namespace
}  // namespace dfb

using namespace dfb;

extern "C" int dfb_voxel_downsample_approximate(dfb_workspace* ws, const float* d_points, uint32_t n, float leaf, float* d_out,
                                                uint32_t* n_out, void* stream) {
	DFB_REQUIRE(ws && n_out, "NULL argument");
	DFB_REQUIRE(leaf > 0.0f && leaf == leaf, "leaf size must be positive");
	DFB_REQUIRE(n <= (1u << 26), "too many points for one voxel grid call");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	*n_out = 0;
	if (n == 0) return DFB_OK;
	DFB_REQUIRE(d_points && d_out, "NULL argument");
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const uint32_t n_blocks = (n + kRankThreads - 1) / kRankThreads;
	/* workspace: hist u32[512 * n_blocks] | order u32[n] | ender u32[n] | entry_begin u32[513] | entry_rank u32[512] |
	 * totals u32[2] | entry_of u16[n] | local_rank u16[n] | head u8[n] */
	const size_t scan_words = exclusive_scan_scratch_words(std::max<uint32_t>(kHist * n_blocks, n));
	const size_t words = (size_t)kHist * n_blocks + 2 * (size_t)n + kHist + 1 + kHist + 4 + scan_words;
	const size_t need = words * 4 + (size_t)n * 5 + 64;
	if (ws->voxel_tmp_bytes < need) {
		DFB_CUDA(cudaStreamSynchronize(s));
		cudaFree(ws->d_voxel_tmp);
		ws->d_voxel_tmp = nullptr;
		ws->voxel_tmp_bytes = 0;
		DFB_CUDA(cudaMalloc(&ws->d_voxel_tmp, need));
		ws->voxel_tmp_bytes = need;
	}
	uint32_t* hist = static_cast<uint32_t*>(ws->d_voxel_tmp);
	uint32_t* order = hist + (size_t)kHist * n_blocks;
	uint32_t* ender = order + n;
	uint32_t* entry_begin = ender + n;
	uint32_t* entry_rank = entry_begin + kHist + 1;
	uint32_t* totals = entry_rank + kHist; /* [0] ended runs, [1] output points, [2] all points (histogram scan) */
	uint32_t* scan_scratch = totals + 4;
	uint16_t* entry_of = reinterpret_cast<uint16_t*>(scan_scratch + scan_words);
	uint16_t* local_rank = entry_of + n;
	uint8_t* head = reinterpret_cast<uint8_t*>(local_rank + n);
	const float4* pts = reinterpret_cast<const float4*>(d_points);
	const float inv = 1.0f / leaf; /* inverse_leaf_size_, float like PCL */
	const unsigned grid = (n + 255) / 256;

	rank_kernel<<<n_blocks, kRankThreads, 0, s>>>(pts, n, inv, entry_of, local_rank, hist, n_blocks);
	rc = launch_exclusive_scan(hist, kHist * n_blocks, totals + 2, entry_begin, n_blocks, scan_scratch, s);
	if (rc != DFB_OK) return rc;
	scatter_kernel<<<grid, 256, 0, s>>>(entry_of, local_rank, hist, n, n_blocks, order);
	heads_kernel<<<grid, 256, 0, s>>>(pts, order, entry_of, entry_begin, n, inv, head, ender);
	rc = launch_exclusive_scan(ender, n, totals, nullptr, 0, scan_scratch, s);
	if (rc != DFB_OK) return rc;
	entry_ranks_kernel<<<1, kHist, 0, s>>>(entry_begin, entry_rank, totals, totals + 1);
	emit_kernel<<<grid, 256, 0, s>>>(pts, order, entry_of, entry_begin, head, ender, entry_rank, totals, n,
	                                 reinterpret_cast<float4*>(d_out));
	count_launch(5);
	DFB_CUDA(cudaGetLastError());
	DFB_CUDA(cudaMemcpyAsync(ws->h_total, totals + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
	DFB_CUDA(cudaStreamSynchronize(s));
	*n_out = *ws->h_total;
	return DFB_OK;
}
