/*
 * scan.cu — in-place exclusive prefix sum of u32 words, the building block of the registration feed's counting /
 * radix sorts (voxel.cu, voxel_approx.cu). Two launches whatever the length:
 *   scan_tiles_kernel  every block scans its own 4096-word tile in place and leaves the tile's total; the block that
 *                      finishes LAST (a ticket) turns the tile totals into tile offsets — no second grid, no host
 *   scan_add_kernel    adds its tile's offset to every word
 * (a single block walking the whole array, the first version, took 49 us for the 1.4 MB histogram of one radix pass
 * and 270 us for a 366 k-word flag array: launch latency is ~3 us.) HBM-bound integer work, no tensor cores.
 */
#include "dfb_internal.h"

namespace dfb {
namespace {

constexpr int kThreads = 1024;
constexpr int kPer = 4;
constexpr int kTile = kThreads * kPer;

/* inclusive scan of one value per thread over the block; returns the exclusive prefix, *total = block sum */
__device__ __forceinline__ uint32_t block_exclusive(uint32_t mine, uint32_t* warp_sums, uint32_t* total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t x = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
		if (lane >= d) x += y;
	}
	if (lane == 31) warp_sums[warp] = x;
	__syncthreads();
	if (warp == 0) {
		uint32_t sum = warp_sums[lane];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, sum, d);
			if (lane >= d) sum += y;
		}
		warp_sums[lane] = sum; /* inclusive over the warps */
	}
	__syncthreads();
	const uint32_t excl = (warp ? warp_sums[warp - 1] : 0u) + x - mine;
	*total = warp_sums[kThreads / 32 - 1];
	__syncthreads(); /* warp_sums may be reused */
	return excl;
}

__global__ void __launch_bounds__(kThreads)
scan_tiles_kernel(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ tile_sums, uint32_t* __restrict__ ticket,
                  uint32_t* __restrict__ total_out) {
	__shared__ uint32_t warp_sums[kThreads / 32];
	__shared__ bool last;
	const uint32_t i = blockIdx.x * kTile + threadIdx.x * kPer;
	uint32_t v[kPer];
	const bool vec = (reinterpret_cast<uintptr_t>(data) & 15u) == 0 && i + kPer <= n; /* whole aligned 16-byte piece */
	if (vec) {
		const uint4 q = *reinterpret_cast<const uint4*>(data + i);
		v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
	} else {
#pragma unroll
		for (int e = 0; e < kPer; ++e) v[e] = i + e < n ? data[i + e] : 0u;
	}
	uint32_t total;
	uint32_t excl = block_exclusive(v[0] + v[1] + v[2] + v[3], warp_sums, &total);
	if (vec) {
		*reinterpret_cast<uint4*>(data + i) = make_uint4(excl, excl + v[0], excl + v[0] + v[1], excl + v[0] + v[1] + v[2]);
	} else {
#pragma unroll
		for (int e = 0; e < kPer; ++e) {
			if (i + e < n) data[i + e] = excl;
			excl += v[e];
		}
	}
	if (threadIdx.x == 0) {
		tile_sums[blockIdx.x] = total;
		__threadfence(); /* the total is visible before the ticket is */
		last = atomicAdd(ticket, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	if (!last) return;
	__threadfence();
	/* last block: exclusive scan of the tile totals in place, 1024 per round */
	uint32_t carry = 0;
	for (uint32_t b0 = 0; b0 < gridDim.x; b0 += kThreads) {
		const uint32_t b = b0 + threadIdx.x;
		const uint32_t t = b < gridDim.x ? __ldcg(tile_sums + b) : 0u;
		uint32_t round_total;
		const uint32_t ex = block_exclusive(t, warp_sums, &round_total);
		if (b < gridDim.x) tile_sums[b] = carry + ex;
		carry += round_total;
	}
	if (threadIdx.x == 0) {
		if (total_out) *total_out = carry;
		*ticket = 0; /* armed for the next scan on this stream */
	}
}

/* data[i] += offset of its tile; with entry_begin: entry_begin[k] = data[k * stride] (where each bucket of a bucket-major
 * histogram starts), entry_begin[n / stride] = total */
__global__ void __launch_bounds__(256)
scan_add_kernel(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ tile_offsets, const uint32_t* __restrict__ total,
                uint32_t* __restrict__ entry_begin, uint32_t stride) {
	const uint32_t i = (blockIdx.x * 256u + threadIdx.x) * kPer;
	if (i >= n) return;
	const uint32_t off = tile_offsets[i / kTile]; /* kPer divides kTile: the four words share a tile */
#pragma unroll
	for (int e = 0; e < kPer; ++e) {
		if (i + e >= n) break;
		const uint32_t v = data[i + e] + off;
		data[i + e] = v;
		if (entry_begin && (i + e) % stride == 0) entry_begin[(i + e) / stride] = v;
	}
	if (entry_begin && i == 0) entry_begin[n / stride] = *total;
}

}  // namespace

size_t exclusive_scan_scratch_words(uint32_t n) { return (size_t)(n + kTile - 1) / kTile + 2; }

/* d_scratch: exclusive_scan_scratch_words(n) words; d_total (device, optional) receives the sum of all words.
 * With d_entry_begin (needs d_total): see scan_add_kernel. Two launches + one 4-byte memset, all on s. */
int launch_exclusive_scan(uint32_t* d_data, uint32_t n, uint32_t* d_total, uint32_t* d_entry_begin, uint32_t stride,
                          uint32_t* d_scratch, cudaStream_t s) {
	if (n == 0) {
		if (d_total) DFB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), s));
		return DFB_OK;
	}
	DFB_REQUIRE(!d_entry_begin || (d_total && stride > 0 && n % stride == 0), "bad bucket layout for the scan");
	const uint32_t tiles = (n + kTile - 1) / kTile;
	uint32_t* ticket = d_scratch + tiles;
	DFB_CUDA(cudaMemsetAsync(ticket, 0, sizeof(uint32_t), s));
	scan_tiles_kernel<<<tiles, kThreads, 0, s>>>(d_data, n, d_scratch, ticket, d_total);
	scan_add_kernel<<<(n + 256 * kPer - 1) / (256 * kPer), 256, 0, s>>>(d_data, n, d_scratch, d_total, d_entry_begin, stride);
	count_launch(2);
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

}  // namespace dfb
