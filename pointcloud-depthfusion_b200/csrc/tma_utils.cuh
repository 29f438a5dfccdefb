/*
 * tma_utils.cuh — thin inline-PTX wrappers for the Blackwell/Hopper async-copy machinery used by the tiled
 * kernels (mbarrier + cp.async.bulk.tensor, SASS SYNCS.* / UTMALDG), the input-tile geometry they share and
 * the host-side tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point,
 * so libdfb_b200.so does not link libcuda).
 */
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include <string>

#include "dfb_internal.h"

namespace dfb {
namespace {

/* one input tile = 64 x 32 source pixels: a 64x32 uint16 depth box (128-byte rows, SWIZZLE_128B) followed by
 * three 64-byte x 32-row colour boxes (SWIZZLE_64B) */
constexpr int kTileW = 64;
constexpr int kTileH = 32;
constexpr uint32_t kDepthBytes = kTileW * 2 * kTileH;             /* 4096 */
constexpr uint32_t kColorBoxBytes = 64 * kTileH;                   /* 2048 */
constexpr uint32_t kStageBytes = kDepthBytes + 3 * kColorBoxBytes; /* 10240 */

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

/* non-blocking: has the phase with this parity completed? */
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
	uint32_t done;
	asm volatile(
	    "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(done)
	    : "r"(bar), "r"(parity)
	    : "memory");
	return done != 0;
}

/* potentially blocking: the warp is suspended by the hardware until the phase completes or ~`ns` nanoseconds have
 * passed (no instructions are issued meanwhile, unlike a test_wait / nanosleep spin loop) */
__device__ __forceinline__ bool mbar_try_wait_ns(uint32_t bar, uint32_t parity, uint32_t ns) {
	uint32_t done;
	asm volatile(
	    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	    : "=r"(done)
	    : "r"(bar), "r"(parity), "r"(ns)
	    : "memory");
	return done != 0;
}

/* bounded wait: a lost TMA completion must become an error, never a hung GPU (~2^21 x 2 us) */
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	/* the suspended wait wakes up on EVERY mbarrier event of the CTA (each TMA completion counts), so an extra
	 * plain sleep keeps a waiting warp from spinning through issue slots */
	for (uint32_t spins = 0; !mbar_try_wait_ns(bar, parity, 2000u); ++spins) {
		__nanosleep(100);
		if (spins > (1u << 23)) __trap();
	}
}

__device__ __forceinline__ uint64_t l2_evict_first_policy() {
	uint64_t policy;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
	return policy;
}

__device__ __forceinline__ uint64_t l2_evict_last_policy() {
	uint64_t policy;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
	return policy;
}

/* 1-D bulk copy global -> shared with an L2 cache policy */
__device__ __forceinline__ void bulk_load_1d_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
	             :
	             : "r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
	             : "memory");
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

/* plain 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16), same completion mechanism */
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :
	             : "r"(dst), "l"(src), "r"(bytes), "r"(bar)
	             : "memory");
}

/* the four bulk copies of one input tile (thread 0 of a CTA) */
__device__ __forceinline__ void tma_issue_tile(uint32_t dst, uint32_t bar, const CUtensorMap* map_depth,
                                               const CUtensorMap* map_color, int u0, int v0, int image, uint64_t policy) {
	mbar_expect_tx(bar, kStageBytes);
	tma_load_3d(dst, map_depth, u0, v0, image, bar, policy);
	tma_load_3d(dst + kDepthBytes, map_color, u0 * 3, v0, image, bar, policy);
	tma_load_3d(dst + kDepthBytes + kColorBoxBytes, map_color, u0 * 3 + 64, v0, image, bar, policy);
	tma_load_3d(dst + kDepthBytes + 2 * kColorBoxBytes, map_color, u0 * 3 + 128, v0, image, bar, policy);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
	uint4 r;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
	return r;
}

/* byte offsets, inside a stage, of the five 16-byte pieces a thread (lane = row, warp = 16-pixel column
 * chunk) reads: 2 depth + 3 colour, with the TMA swizzles undone */
struct TileReadOffsets {
	uint32_t d[2], c[3];
};

__device__ __forceinline__ TileReadOffsets tile_read_offsets(uint32_t lane, uint32_t warp) {
	TileReadOffsets o;
	const uint32_t r = lane;
	o.d[0] = r * 128u + (((2u * warp + 0u) ^ (r & 7u)) << 4);
	o.d[1] = r * 128u + (((2u * warp + 1u) ^ (r & 7u)) << 4);
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const uint32_t p = 48u * warp + 16u * k; /* byte position inside the 192-byte colour row */
		const uint32_t box = p >> 6, chunk = (p & 63u) >> 4;
		o.c[k] = kDepthBytes + box * kColorBoxBytes + r * 64u + ((chunk ^ ((r >> 1) & 3u)) << 4);
	}
	return o;
}

inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
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

/* tensor maps over pools depth[n_img][h][w] (uint16) and color[n_img][h][w*3] (uint8) */
inline int make_pool_tensor_maps(const uint16_t* d_depth, const uint8_t* d_color, int w, int h, uint64_t n_img,
                                 CUtensorMap* map_depth, CUtensorMap* map_color, int box_rows = kTileH,
                                 size_t depth_img_stride = 0, size_t color_img_stride = 0) {
	if (!depth_img_stride) depth_img_stride = (size_t)w * h * 2;
	if (!color_img_stride) color_img_stride = (size_t)w * h * 3;
	auto enc = tensor_map_encoder();
	if (!enc) {
		set_error("cuTensorMapEncodeTiled is not available from this driver");
		return DFB_ERR_UNSUPPORTED;
	}
	const cuuint32_t estr[3] = {1, 1, 1};
	{
		const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, n_img};
		const cuuint64_t strides[2] = {(cuuint64_t)w * 2, (cuuint64_t)depth_img_stride};
		const cuuint32_t box[3] = {kTileW, (cuuint32_t)box_rows, 1};
		const CUresult r = enc(map_depth, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(d_depth), dims, strides, box,
		                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
		                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) {
			set_error("cuTensorMapEncodeTiled(depth) failed: " + std::to_string((int)r));
			return DFB_ERR_CUDA;
		}
	}
	{
		const cuuint64_t dims[3] = {(cuuint64_t)w * 3, (cuuint64_t)h, n_img};
		const cuuint64_t strides[2] = {(cuuint64_t)w * 3, (cuuint64_t)color_img_stride};
		const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
		const CUresult r = enc(map_color, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_color), dims, strides, box,
		                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
		                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS) {
			set_error("cuTensorMapEncodeTiled(colour) failed: " + std::to_string((int)r));
			return DFB_ERR_CUDA;
		}
	}
	return DFB_OK;
}

/* can the pools be described by the maps above? */
inline bool pools_fit_tensor_maps(int w, const uint16_t* d_depth, const uint8_t* d_color) {
	return (w % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_depth) & 15u) == 0) &&
	       ((reinterpret_cast<uintptr_t>(d_color) & 15u) == 0) && tensor_map_encoder() != nullptr;
}

}  // namespace
}  // namespace dfb
