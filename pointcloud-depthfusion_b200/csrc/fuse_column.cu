/*
 * fuse_column.cu — column-traversal splat kernel with direct global loads (no shared memory).
 *
 * Measured on a B200 (tools/gpu_explore.py): 64-bit atomicMin into L2-resident keys sustains ~760 Gop/s when
 * the 32 lanes of a warp hit 32 consecutive keys, ~190 Gop/s when every lane hits its own sector. With the
 * reference's default `vertical_image` (prerotate 90 deg about Z, fusion_node.cpp:774-777) a source ROW
 * becomes a destination COLUMN, so here the lanes of a warp run down a source COLUMN (lane = source row v)
 * and land on consecutive pixels of one destination row. Each thread owns a 16-pixel segment of its row:
 * 2 x 16 B of depth + 3 x 16 B of colour = whole 32-byte sectors per lane.
 *
 * This variant keeps the loads in the LSU path: every warp-level LDG.128 touches 32 different lines, which
 * ncu shows as the L1TEX wavefront limiter (profiles/). fuse_tiled.cu removes that with TMA tiles; this
 * kernel remains as the fallback for pools the tensor maps cannot describe and as a cross-check.
 * The per-pair arithmetic lives in splat_core.cuh.
 */
#include "splat_core.cuh"

namespace dfb {
namespace {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;

/* grid: x = column strips of kWarps*16 pixels, y = row blocks of 32, z = frameset within the group */
template <int kHasT, int kRoiAll>
__global__ void __launch_bounds__(kThreads)
splat_column_kernel(const __grid_constant__ SplatShared G, const __grid_constant__ SplatCam K, const __grid_constant__ FuseParams P,
                    const float* __restrict__ ray_tables, const uint16_t* __restrict__ depth,
                    const uint8_t* __restrict__ color, unsigned long long* __restrict__ keys_base) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int v = blockIdx.y * 32 + lane;
	const int u0 = (blockIdx.x * kWarps + warp) * kSeg;
	if (v >= G.height || u0 >= G.width) return; /* W % 16 == 0: a segment is never partial */
	const size_t img = ((size_t)blockIdx.z * G.n_cameras + K.cam) * ((size_t)G.width * G.height);
	const size_t pix = img + (size_t)v * G.width + u0;
	const uint4* dp = reinterpret_cast<const uint4*>(depth + pix);
	const uint4* cp = reinterpret_cast<const uint4*>(color + pix * 3);
	const uint4 d0 = ldg_stream(dp), d1 = ldg_stream(dp + 1);
	const uint4 c0 = ldg_stream(cp), c1 = ldg_stream(cp + 1), c2 = ldg_stream(cp + 2);
	const float4* xp = reinterpret_cast<const float4*>(ray_tables + K.x_off + u0);
	const float4 x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2), x3 = __ldg(xp + 3);
	const float yr = __ldg(ray_tables + K.y_off + v);
	const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
	const uint32_t cw[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
	const float2 xr[8] = {{x0.x, x0.y}, {x0.z, x0.w}, {x1.x, x1.y}, {x1.z, x1.w}, {x2.x, x2.y}, {x2.z, x2.w}, {x3.x, x3.y}, {x3.z, x3.w}};
	unsigned long long* keys = keys_base + (size_t)blockIdx.z * G.n_out_pixels;
	const bool row_in_roi = (unsigned)(v - K.roi_y0) < (unsigned)K.roi_h;
	uint32_t slow = splat_segment<kHasT, kRoiAll>(std::make_integer_sequence<int, kSeg / 2>{}, G, K, dw, cw, xr, yr, u0, row_in_roi,
	                                              KeyTarget{reinterpret_cast<char*>(keys), G.px_origin, 0u});
	while (slow) { /* rare: operands outside the guarded fast-division range */
		const int k = __ffs(slow) - 1;
		slow &= slow - 1;
		splat_pixel_slow(P, (int)K.cam, depth + img, color + img * 3, u0 + 2 * k, v, keys);
		splat_pixel_slow(P, (int)K.cam, depth + img, color + img * 3, u0 + 2 * k + 1, v, keys);
	}
}

__global__ void __launch_bounds__(256)
ray_tables_kernel(const __grid_constant__ FuseParams P, float* __restrict__ xtab, float* __restrict__ ytab, int x_pitch, int y_pitch) {
	const int cam = blockIdx.y;
	const CamParams& C = P.cam[cam];
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < x_pitch) xtab[(size_t)cam * x_pitch + i] = deproject_ray(i, C.ppx, C.fx);
	if (i < y_pitch) ytab[(size_t)cam * y_pitch + i] = deproject_ray(i, C.ppy, C.fy);
}

/* the same rays regrouped per 64 x kStreamItemRows input item of the stream kernel: [camera][tile][64 column rays, the
 * item's row rays], so that an item's rays arrive with ONE bulk copy */
__global__ void __launch_bounds__(64 + kStreamItemRows)
tile_rays_kernel(const __grid_constant__ FuseParams P, float* __restrict__ tiles, int tiles_x, int tiles_per_image) {
	const int cam = blockIdx.y, tile = blockIdx.x, i = threadIdx.x;
	const CamParams& C = P.cam[cam];
	const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
	const float v = i < 64 ? deproject_ray(tx * 64 + i, C.ppx, C.fx) : deproject_ray(ty * kStreamItemRows + (i - 64), C.ppy, C.fy);
	tiles[((size_t)cam * tiles_per_image + tile) * (64 + kStreamItemRows) + i] = v;
}

}  // namespace

bool splat_column_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color) {
	const int w = P.cam[0].width;
	return (w % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_depth) & 15u) == 0) &&
	       ((reinterpret_cast<uintptr_t>(d_color) & 15u) == 0) && ((size_t)P.n_out_pixels * 8 < (1ull << 32));
}

int launch_splat_column(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color,
                        unsigned long long* d_keys, int group, cudaStream_t s) {
	const int w = P.cam[0].width, h = P.cam[0].height;
	const dim3 grid((w + kWarps * kSeg - 1) / (kWarps * kSeg), (h + 31) / 32, group);
	for (int c = 0; c < P.n_cameras; ++c) {
		const CamParams& C = P.cam[c];
		const SplatShared G = make_splat_shared(P);
		const SplatCam K = make_splat_cam(P, c, rays);
		if (C.has_transform) {
			if (C.roi_all)
				splat_column_kernel<1, 1><<<grid, kThreads, 0, s>>>(G, K, P, rays.x, d_depth, d_color, d_keys);
			else
				splat_column_kernel<1, 0><<<grid, kThreads, 0, s>>>(G, K, P, rays.x, d_depth, d_color, d_keys);
		} else {
			if (C.roi_all)
				splat_column_kernel<0, 1><<<grid, kThreads, 0, s>>>(G, K, P, rays.x, d_depth, d_color, d_keys);
			else
				splat_column_kernel<0, 0><<<grid, kThreads, 0, s>>>(G, K, P, rays.x, d_depth, d_color, d_keys);
		}
		count_launch();
	}
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

int launch_ray_tables(const FuseParams& P, float* d_xtab, float* d_ytab, int x_pitch, int y_pitch, cudaStream_t s) {
	const int n = x_pitch > y_pitch ? x_pitch : y_pitch;
	ray_tables_kernel<<<dim3((n + 255) / 256, P.n_cameras), 256, 0, s>>>(P, d_xtab, d_ytab, x_pitch, y_pitch);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

int launch_tile_rays(const FuseParams& P, float* d_tiles, int tiles_x, int tiles_per_image, cudaStream_t s) {
	tile_rays_kernel<<<dim3(tiles_per_image, P.n_cameras), 64 + kStreamItemRows, 0, s>>>(P, d_tiles, tiles_x, tiles_per_image);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

}  // namespace dfb
