/*
 * fuse_tiled.cu — splat kernel with persistent CTAs fed by TMA tiles (one launch per camera and group).
 *
 * Why (ncu on a B200, profiles/): with direct loads (fuse_column.cu) every warp-level LDG.128 of the column
 * traversal touches 32 different cache lines; the L1TEX wavefront pipe becomes the top unit (~60 %) and warps
 * sit in long-scoreboard stalls at the head of every block. Here
 *   - each CTA is persistent and walks tiles of 64 x 32 source pixels (static stride over the tile list);
 *   - a tile is brought into shared memory by FOUR bulk tensor copies (cp.async.bulk.tensor.3d, SASS
 *     UTMALDG; tma_utils.cuh), completion tracked by one mbarrier per stage, kStages tiles in flight per
 *     CTA, L2 evict-first policy (the inputs are read once; the key ring must stay in L2); out-of-range
 *     rows / columns of edge tiles are zero-filled by the TMA unit = invalid depth;
 *   - a thread then reads ITS row segment (lane = source row, warp = 16-pixel column chunk) with five
 *     conflict-free LDS.128, frees the stage, and runs the shared per-pair body (splat_core.cuh) while the
 *     next tiles stream in.
 * Measured: same time per launch as fuse_column.cu — at ~20 us per launch both are dominated by launch ramp,
 * tail and the gaps between the three kernels of a group. fuse_stream.cu is the answer to that; this kernel
 * stays as the per-group building block (and the cross-check of the tile addressing).
 */
#include "splat_core.cuh"
#include "tma_utils.cuh"

namespace dfb {
namespace {

constexpr int kWarps = kTileW / kSeg; /* 4 */
constexpr int kThreads = kWarps * 32;
constexpr int kStages = 4;
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024; /* + slack to align the base to 1024 */

template <int kHasT, int kRoiAll>
__global__ void __launch_bounds__(kThreads)
splat_tiled_kernel(const __grid_constant__ SplatShared G, const __grid_constant__ SplatCam K, const __grid_constant__ FuseParams P,
                   const __grid_constant__ CUtensorMap map_depth, const __grid_constant__ CUtensorMap map_color,
                   const float* __restrict__ ray_tables, const uint16_t* __restrict__ depth,
                   const uint8_t* __restrict__ color, unsigned long long* __restrict__ keys_base, int n_tiles, int tiles_x,
                   int tiles_per_image) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ __align__(8) unsigned long long bars[kStages];
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u; /* SWIZZLE_128B wants 1024-byte aligned tiles */
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	uint64_t policy = 0;
	if (tid == 0) {
		policy = l2_evict_first_policy();
		for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
		mbar_fence_init();
	}
	__syncthreads();

	const int first = blockIdx.x, step = gridDim.x;
	auto issue = [&](int t, int s) { /* thread 0 only */
		const int fs = t / tiles_per_image;
		const int r = t - fs * tiles_per_image;
		const int ty = r / tiles_x;
		tma_issue_tile(base + (uint32_t)s * kStageBytes, smem_u32(&bars[s]), &map_depth, &map_color, (r - ty * tiles_x) * kTileW,
		               ty * kTileH, fs * G.n_cameras + (int)K.cam, policy);
	};
	if (tid == 0)
		for (int s = 0; s < kStages; ++s)
			if (first + s * step < n_tiles) issue(first + s * step, s);

	const TileReadOffsets off = tile_read_offsets((uint32_t)lane, (uint32_t)warp);
	int it = 0;
	for (int t = first; t < n_tiles; t += step, ++it) {
		const int s = it % kStages;
		const uint32_t stage = base + (uint32_t)s * kStageBytes;
		mbar_wait(smem_u32(&bars[s]), (uint32_t)(it / kStages) & 1u);
		const uint4 d0 = lds128(stage + off.d[0]), d1 = lds128(stage + off.d[1]);
		const uint4 c0 = lds128(stage + off.c[0]), c1 = lds128(stage + off.c[1]), c2 = lds128(stage + off.c[2]);
		__syncthreads(); /* every thread holds its segment in registers: the stage can be refilled */
		if (tid == 0 && t + kStages * step < n_tiles) issue(t + kStages * step, s);
		const int fs = t / tiles_per_image;
		const int r = t - fs * tiles_per_image;
		const int ty = r / tiles_x;
		const int v = ty * kTileH + lane;
		const int u0 = (r - ty * tiles_x) * kTileW + warp * kSeg;
		const int vt = min(v, G.height - 1); /* rows below the image are zero depth; keep the table read in range */
		const float4* xp = reinterpret_cast<const float4*>(ray_tables + K.x_off + u0);
		const float4 x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2), x3 = __ldg(xp + 3);
		const float yr = __ldg(ray_tables + K.y_off + vt);
		const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
		const uint32_t cw[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
		const float2 xr[8] = {{x0.x, x0.y}, {x0.z, x0.w}, {x1.x, x1.y}, {x1.z, x1.w}, {x2.x, x2.y}, {x2.z, x2.w}, {x3.x, x3.y}, {x3.z, x3.w}};
		unsigned long long* keys = keys_base + (size_t)fs * G.n_out_pixels;
		const bool row_in_roi = (unsigned)(v - K.roi_y0) < (unsigned)K.roi_h;
		uint32_t slow = splat_segment<kHasT, kRoiAll>(std::make_integer_sequence<int, kSeg / 2>{}, G, K, dw, cw, xr, yr, u0,
		                                              row_in_roi, KeyTarget{reinterpret_cast<char*>(keys), G.px_origin, 0u});
		if (slow) { /* rare: operands outside the guarded fast-division range (valid pixels only, so in range) */
			const size_t img = ((size_t)fs * G.n_cameras + K.cam) * ((size_t)G.width * G.height);
			while (slow) {
				const int k = __ffs(slow) - 1;
				slow &= slow - 1;
				splat_pixel_slow(P, (int)K.cam, depth + img, color + img * 3, u0 + 2 * k, v, keys);
				splat_pixel_slow(P, (int)K.cam, depth + img, color + img * 3, u0 + 2 * k + 1, v, keys);
			}
		}
	}
}

template <int kHasT, int kRoiAll>
int launch_one(const SplatShared& G, const SplatCam& K, const FuseParams& P, const CUtensorMap& md, const CUtensorMap& mc, const float* rays,
               const uint16_t* d_depth, const uint8_t* d_color, unsigned long long* d_keys, int n_tiles, int tiles_x,
               int tiles_per_image, int grid, cudaStream_t s) {
	auto kern = splat_tiled_kernel<kHasT, kRoiAll>;
	static bool configured = false;
	if (!configured) {
		DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
		configured = true;
	}
	kern<<<grid, kThreads, kSmemBytes, s>>>(G, K, P, md, mc, rays, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image);
	count_launch();
	return DFB_OK;
}

}  // namespace

bool splat_tiled_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color) {
	return pools_fit_tensor_maps(P.cam[0].width, d_depth, d_color) && ((size_t)P.n_out_pixels * 8 < (1ull << 32));
}

int launch_splat_tiled(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color,
                       unsigned long long* d_keys, int group, int ctas_per_sm, cudaStream_t s) {
	const int w = P.cam[0].width, h = P.cam[0].height;
	alignas(64) CUtensorMap md, mc;
	int rc = make_pool_tensor_maps(d_depth, d_color, w, h, (uint64_t)group * P.n_cameras, &md, &mc);
	if (rc != DFB_OK) return rc;
	const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;
	const int tiles_per_image = tiles_x * tiles_y;
	const int n_tiles = tiles_per_image * group;
	int sms = 148, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (ctas_per_sm <= 0) ctas_per_sm = 5;
	int grid = sms * ctas_per_sm;
	if (grid > n_tiles) grid = n_tiles;
	for (int c = 0; c < P.n_cameras; ++c) {
		const CamParams& C = P.cam[c];
		const SplatShared G = make_splat_shared(P);
		const SplatCam K = make_splat_cam(P, c, rays);
		if (C.has_transform)
			rc = C.roi_all ? launch_one<1, 1>(G, K, P, md, mc, rays.x, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image, grid, s)
			               : launch_one<1, 0>(G, K, P, md, mc, rays.x, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image, grid, s);
		else
			rc = C.roi_all ? launch_one<0, 1>(G, K, P, md, mc, rays.x, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image, grid, s)
			               : launch_one<0, 0>(G, K, P, md, mc, rays.x, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image, grid, s);
		if (rc != DFB_OK) return rc;
	}
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

}  // namespace dfb
