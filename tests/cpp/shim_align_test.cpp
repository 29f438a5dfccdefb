/*
 * shim_align_test.cpp — C++ consumer of the drop-in headers for the align_frames path of fusion_node
 * (fusion_node/src/fusion_node.cpp:443-447,545-551: allocateDepthFrame + allocateDepthFrameUnaligned +
 * allocateColorFrame, then setDepthFrameUnaligned + alignDepthToColor per frame). tests/test_shim_gpu.py compares the
 * aligned depth image with the CPU oracle.
 *
 *   shim_align_test <input.bin> <output.bin>
 * input : Intrinsics depth, Intrinsics colour, Extrinsics depth->colour, float depth_scale, then the depth image
 * output: the aligned uint16 image (colour resolution)
 */
#include <cstdint>
#include <cstdio>
#include <vector>

#include "pointcloud_processing/frameset.h"
#include "pointcloud_processing/kernels.cuh"

int main(int argc, char** argv) {
	if (argc != 3) return 2;
	FILE* in = std::fopen(argv[1], "rb");
	if (!in) return 2;
	Intrinsics depth_intr, color_intr;
	Extrinsics extr;
	float depth_scale;
	if (std::fread(&depth_intr, sizeof(depth_intr), 1, in) != 1 || std::fread(&color_intr, sizeof(color_intr), 1, in) != 1 ||
	    std::fread(&extr, sizeof(extr), 1, in) != 1 || std::fread(&depth_scale, sizeof(depth_scale), 1, in) != 1)
		return 2;
	std::vector<uint16_t> depth(static_cast<size_t>(depth_intr.width) * depth_intr.height);
	if (std::fread(depth.data(), 2, depth.size(), in) != depth.size()) return 2;
	std::fclose(in);

	Kernels::init_cuda(false);
	cudaStream_t stream;
	cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
	Frameset frameset;
	frameset.setStream(&stream);
	frameset.allocateDepthFrame(color_intr); /* the aligned depth frame has the colour camera's geometry */
	frameset.allocateDepthFrameUnaligned(depth_intr, extr);
	frameset.allocateColorFrame(color_intr);
	std::vector<uint16_t> aligned(static_cast<size_t>(color_intr.width) * color_intr.height);
	for (int frame = 0; frame < 2; ++frame) { /* twice: nothing may leak from one frame into the next */
		frameset.setDepthFrameUnaligned(depth.data());
		frameset.alignDepthToColor(depth_scale);
		frameset.copyDepthToHost(aligned.data());
	}
	FILE* out = std::fopen(argv[2], "wb");
	if (!out) return 2;
	std::fwrite(aligned.data(), 2, aligned.size(), out);
	std::fclose(out);
	cudaStreamDestroy(stream);
	return 0;
}
