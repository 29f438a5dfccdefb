/*
 * fuse_tiled.cu — third-generation splat kernel: persistent CTAs fed by TMA tiles.
 *
 * Why (ncu on a B200, profiles/): with direct loads (fuse_column.cu) every warp-level LDG.128 of the column
 * traversal touches 32 different cache lines; the L1TEX wavefront pipe becomes the top unit (~60 %), warps sit
 * in long-scoreboard stalls at the head of every block and the issue slots are only ~46 % used. Here
 *   - each CTA is persistent and walks tiles of 64 x 32 source pixels (static stride over the tile list);
 *   - a tile is brought into shared memory by FOUR bulk tensor copies (cp.async.bulk.tensor.3d, SASS
 *     UTMALDG): one 64x32 uint16 depth box (128-byte rows, SWIZZLE_128B) and three 64-byte x 32-row colour
 *     boxes (SWIZZLE_64B), completion tracked by one mbarrier per stage, kStages tiles in flight per CTA,
 *     L2 evict-first policy (the inputs are read once; the key ring must stay in L2);
 *     out-of-range rows / columns of edge tiles are zero-filled by the TMA unit = invalid depth;
 *   - a thread then reads ITS row segment (lane = source row, warp = 16-pixel column chunk) with five
 *     conflict-free LDS.128 (the swizzles spread the 8 lanes of a quarter-warp over all 32 banks), frees the
 *     stage, and runs the shared per-pair body (splat_core.cuh) while the next tiles stream in.
 * The arithmetic, and therefore every key, is identical to fuse_direct.cu / fuse_column.cu.
 */
#include <cuda.h>
#include <cudaTypedefs.h>

#include "splat_core.cuh"

namespace dfb {
namespace {

constexpr int kTileW = 64;
constexpr int kTileH = 32;
constexpr int kWarps = kTileW / kSeg; /* 4 */
constexpr int kThreads = kWarps * 32;
constexpr int kStages = 4;
constexpr uint32_t kDepthBytes = kTileW * 2 * kTileH;               /* 4096, 128-byte rows */
constexpr uint32_t kColorBoxBytes = 64 * kTileH;                     /* 2048, 64-byte rows */
constexpr uint32_t kStageBytes = kDepthBytes + 3 * kColorBoxBytes;   /* 10240 */
constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024;        /* + slack to align the base to 1024 */

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

/* bounded wait: a lost TMA completion must become an error, never a hung GPU */
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done = 0;
	for (uint32_t spins = 0; !done; ++spins) {
		asm volatile(
		    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(done)
		    : "r"(bar), "r"(parity)
		    : "memory");
		if (!done && spins > (1u << 24)) __trap();
	}
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar,
                                            uint64_t policy) {
	asm volatile(
	    "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
	    " [%0], [%1, {%2, %3, %4}], [%5], %6;"
	    :
	    : "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar), "l"(policy)
	    : "memory");
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
	uint4 r;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
	return r;
}

struct TileCoord {
	int fs, u0, v0;
};

__device__ __forceinline__ TileCoord tile_coord(int t, int tiles_x, int tiles_per_image) {
	TileCoord c;
	c.fs = t / tiles_per_image;
	const int r = t - c.fs * tiles_per_image;
	const int ty = r / tiles_x;
	c.v0 = ty * kTileH;
	c.u0 = (r - ty * tiles_x) * kTileW;
	return c;
}

template <int kHasT, int kRoiAll>
__global__ void __launch_bounds__(kThreads)
splat_tiled_kernel(const __grid_constant__ SplatParams K, const __grid_constant__ FuseParams P,
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
		asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
		for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	}
	__syncthreads();

	const int first = blockIdx.x, step = gridDim.x;
	auto issue = [&](int t, int s) { /* thread 0 only */
		const TileCoord c = tile_coord(t, tiles_x, tiles_per_image);
		const uint32_t bar = smem_u32(&bars[s]);
		const uint32_t dst = base + (uint32_t)s * kStageBytes;
		const int z = c.fs * K.n_cameras + (int)K.cam;
		mbar_expect_tx(bar, kStageBytes);
		tma_load_3d(dst, &map_depth, c.u0, c.v0, z, bar, policy);
		tma_load_3d(dst + kDepthBytes, &map_color, c.u0 * 3, c.v0, z, bar, policy);
		tma_load_3d(dst + kDepthBytes + kColorBoxBytes, &map_color, c.u0 * 3 + 64, c.v0, z, bar, policy);
		tma_load_3d(dst + kDepthBytes + 2 * kColorBoxBytes, &map_color, c.u0 * 3 + 128, c.v0, z, bar, policy);
	};
	if (tid == 0) {
		for (int s = 0; s < kStages; ++s) {
			const int t = first + s * step;
			if (t < n_tiles) issue(t, s);
		}
	}

	/* per-thread constant part of the shared-memory addresses */
	const uint32_t r = (uint32_t)lane;
	const uint32_t d_off0 = r * 128u + (((2u * warp + 0u) ^ (r & 7u)) << 4);
	const uint32_t d_off1 = r * 128u + (((2u * warp + 1u) ^ (r & 7u)) << 4);
	uint32_t c_off[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const uint32_t p = 48u * warp + 16u * k; /* byte position inside the 192-byte colour row */
		const uint32_t box = p >> 6, chunk = (p & 63u) >> 4;
		c_off[k] = kDepthBytes + box * kColorBoxBytes + r * 64u + ((chunk ^ ((r >> 1) & 3u)) << 4);
	}

	int it = 0;
	for (int t = first; t < n_tiles; t += step, ++it) {
		const int s = it % kStages;
		const uint32_t stage = base + (uint32_t)s * kStageBytes;
		mbar_wait(smem_u32(&bars[s]), (uint32_t)(it / kStages) & 1u);
		const uint4 d0 = lds128(stage + d_off0), d1 = lds128(stage + d_off1);
		const uint4 c0 = lds128(stage + c_off[0]), c1 = lds128(stage + c_off[1]), c2 = lds128(stage + c_off[2]);
		__syncthreads(); /* every thread holds its segment in registers: the stage can be refilled */
		if (tid == 0) {
			const int tn = t + kStages * step;
			if (tn < n_tiles) issue(tn, s);
		}
		const TileCoord tc = tile_coord(t, tiles_x, tiles_per_image);
		const int v = tc.v0 + lane;
		const int u0 = tc.u0 + warp * kSeg;
		const int vt = min(v, K.height - 1); /* rows below the image are zero depth; keep the table read in range */
		const float4* xp = reinterpret_cast<const float4*>(ray_tables + K.x_off + u0);
		const float4 x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2), x3 = __ldg(xp + 3);
		const float yr = __ldg(ray_tables + K.y_off + vt);
		const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
		const uint32_t cw[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
		const float2 xr[8] = {{x0.x, x0.y}, {x0.z, x0.w}, {x1.x, x1.y}, {x1.z, x1.w}, {x2.x, x2.y}, {x2.z, x2.w}, {x3.x, x3.y}, {x3.z, x3.w}};
		unsigned long long* keys = keys_base + (size_t)tc.fs * K.n_out_pixels;
		const bool row_in_roi = (unsigned)(v - K.roi_y0) < (unsigned)K.roi_h;
		uint32_t slow = splat_segment<kHasT, kRoiAll>(std::make_integer_sequence<int, kSeg / 2>{}, K, dw, cw, xr, yr, u0,
		                                              row_in_roi, reinterpret_cast<char*>(keys));
		if (slow) { /* rare: operands outside the guarded fast-division range (valid pixels only, so in range) */
			const size_t img = ((size_t)tc.fs * K.n_cameras + K.cam) * ((size_t)K.width * K.height);
			while (slow) {
				const int k = __ffs(slow) - 1;
				slow &= slow - 1;
				splat_pixel_slow(P, (int)K.cam, depth + img, color + img * 3, u0 + 2 * k, v, keys);
				splat_pixel_slow(P, (int)K.cam, depth + img, color + img * 3, u0 + 2 * k + 1, v, keys);
			}
		}
	}
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
	static PFN_cuTensorMapEncodeTiled_v12000 fn = []() -> PFN_cuTensorMapEncodeTiled_v12000 {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
		    q != cudaDriverEntryPointSuccess)
			return nullptr;
		return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
	}();
	return fn;
}

template <int kHasT, int kRoiAll>
int launch_one(const SplatParams& K, const FuseParams& P, const CUtensorMap& md, const CUtensorMap& mc, const float* rays,
               const uint16_t* d_depth, const uint8_t* d_color, unsigned long long* d_keys, int n_tiles, int tiles_x,
               int tiles_per_image, int grid, cudaStream_t s) {
	auto kern = splat_tiled_kernel<kHasT, kRoiAll>;
	static bool configured = false;
	if (!configured) {
		DFB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
		configured = true;
	}
	kern<<<grid, kThreads, kSmemBytes, s>>>(K, P, md, mc, rays, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image);
	count_launch();
	return DFB_OK;
}

}  // namespace

bool splat_tiled_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color) {
	const int w = P.cam[0].width;
	return (w % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_depth) & 15u) == 0) &&
	       ((reinterpret_cast<uintptr_t>(d_color) & 15u) == 0) && ((size_t)P.n_out_pixels * 8 < (1ull << 32)) &&
	       encode_fn() != nullptr;
}

int launch_splat_tiled(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color,
                       unsigned long long* d_keys, int group, int ctas_per_sm, cudaStream_t s) {
	const int w = P.cam[0].width, h = P.cam[0].height;
	const uint64_t n_img = (uint64_t)group * P.n_cameras;
	alignas(64) CUtensorMap md, mc;
	{
		const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, n_img};
		const cuuint64_t strides[2] = {(cuuint64_t)w * 2, (cuuint64_t)w * h * 2};
		const cuuint32_t box[3] = {kTileW, kTileH, 1};
		const cuuint32_t estr[3] = {1, 1, 1};
		CUresult r = encode_fn()(&md, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(d_depth), dims, strides, box, estr,
		                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
		                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) {
			set_error("cuTensorMapEncodeTiled(depth) failed: " + std::to_string((int)r));
			return DFB_ERR_CUDA;
		}
	}
	{
		const cuuint64_t dims[3] = {(cuuint64_t)w * 3, (cuuint64_t)h, n_img};
		const cuuint64_t strides[2] = {(cuuint64_t)w * 3, (cuuint64_t)w * h * 3};
		const cuuint32_t box[3] = {64, kTileH, 1};
		const cuuint32_t estr[3] = {1, 1, 1};
		CUresult r = encode_fn()(&mc, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_color), dims, strides, box, estr,
		                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
		                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) {
			set_error("cuTensorMapEncodeTiled(colour) failed: " + std::to_string((int)r));
			return DFB_ERR_CUDA;
		}
	}
	const int tiles_x = (w + kTileW - 1) / kTileW, tiles_y = (h + kTileH - 1) / kTileH;
	const int tiles_per_image = tiles_x * tiles_y;
	const int n_tiles = tiles_per_image * group;
	int sms = 148;
	{
		int dev = 0;
		cudaGetDevice(&dev);
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	}
	if (ctas_per_sm <= 0) ctas_per_sm = 5;
	int grid = sms * ctas_per_sm;
	if (grid > n_tiles) grid = n_tiles;
	for (int c = 0; c < P.n_cameras; ++c) {
		const CamParams& C = P.cam[c];
		const SplatParams K = make_splat_params(P, c, rays);
		int rc;
		if (C.has_transform)
			rc = C.roi_all ? launch_one<1, 1>(K, P, md, mc, rays.x, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image, grid, s)
			               : launch_one<1, 0>(K, P, md, mc, rays.x, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image, grid, s);
		else
			rc = C.roi_all ? launch_one<0, 1>(K, P, md, mc, rays.x, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image, grid, s)
			               : launch_one<0, 0>(K, P, md, mc, rays.x, d_depth, d_color, d_keys, n_tiles, tiles_x, tiles_per_image, grid, s);
		if (rc != DFB_OK) return rc;
	}
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

}  // namespace dfb
