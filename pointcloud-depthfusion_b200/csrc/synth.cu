/*
 * synth.cu — device-side generator of the synthetic framesets (include/dfb/dfb_synth.h), so that
 * bench.py can stage hundreds of distinct framesets in HBM in milliseconds. Bit-identical to the CPU
 * generator by construction (see the header); tests/test_synth_gpu.py checks it.
 * Stands in for camera_node (RealSense capture, camera_node/src/realsense.cpp) — out of scope.
 */
#include <vector>

#include "dfb_internal.h"

namespace dfb {
namespace {

__global__ void __launch_bounds__(256)
synth_kernel(const dfs_scene* __restrict__ scenes, const dfs_camera* __restrict__ cams, int n_cameras,
             uint16_t* __restrict__ depth, uint8_t* __restrict__ color) {
	const int cam = blockIdx.y;
	const int fs = blockIdx.z;
	const dfs_camera c = cams[cam];
	const int npix = c.width * c.height;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= npix) return;
	const int v = i / c.width;
	const int u = i - v * c.width;
	const dfs_scene sc = scenes[fs];
	uint16_t d;
	uint8_t rgb[3];
	dfs_pixel(&sc, &c, (uint32_t)cam, u, v, &d, rgb);
	const size_t o = ((size_t)fs * n_cameras + cam) * (size_t)npix + i;
	depth[o] = d;
	color[o * 3 + 0] = rgb[0];
	color[o * 3 + 1] = rgb[1];
	color[o * 3 + 2] = rgb[2];
}

}  // namespace
}  // namespace dfb

extern "C" int dfb_synth_generate(const dfs_scene* scenes, int n_framesets, const dfs_camera* cams, int n_cameras,
                                  uint16_t* d_depth, uint8_t* d_color, void* stream) {
	using namespace dfb;
	DFB_REQUIRE(scenes && cams && d_depth && d_color, "NULL argument");
	DFB_REQUIRE(n_framesets > 0 && n_framesets <= 65535 && n_cameras > 0 && n_cameras <= DFB_MAX_CAMERAS,
	            "n_framesets must be in [1, 65535] and n_cameras in [1, DFB_MAX_CAMERAS]");
	for (int c = 1; c < n_cameras; ++c)
		DFB_REQUIRE(cams[c].width == cams[0].width && cams[c].height == cams[0].height,
		            "all cameras must share one resolution");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	dfs_scene* d_scenes = nullptr;
	dfs_camera* d_cams = nullptr;
	DFB_CUDA(cudaMalloc(&d_scenes, sizeof(dfs_scene) * n_framesets));
	DFB_CUDA(cudaMalloc(&d_cams, sizeof(dfs_camera) * n_cameras));
	DFB_CUDA(cudaMemcpyAsync(d_scenes, scenes, sizeof(dfs_scene) * n_framesets, cudaMemcpyHostToDevice, s));
	DFB_CUDA(cudaMemcpyAsync(d_cams, cams, sizeof(dfs_camera) * n_cameras, cudaMemcpyHostToDevice, s));
	const int npix = cams[0].width * cams[0].height;
	const dim3 grid((npix + 255) / 256, n_cameras, n_framesets);
	synth_kernel<<<grid, 256, 0, s>>>(d_scenes, d_cams, n_cameras, d_depth, d_color);
	count_launch();
	cudaError_t e = cudaGetLastError();
	/* the scene tables are tiny; free them once the kernel is done */
	cudaError_t e2 = cudaStreamSynchronize(s);
	cudaFree(d_scenes);
	cudaFree(d_cams);
	DFB_CUDA(e);
	DFB_CUDA(e2);
	return DFB_OK;
}
