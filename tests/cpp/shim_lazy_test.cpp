/*
 * shim_lazy_test.cpp — the lazy Pointcloud of the drop-in headers must be indistinguishable from the reference's eager
 * one (pointcloud.cpp:21-169). Same input file as shim_fusion_test; writes five fused images + one cloud:
 *   0  processSyncedFrames as fusion_node runs it            -> ONE fused launch in project()
 *   1  the same with fused_cloud.get() before project()       -> materialised: every recorded operator runs eagerly
 *   2  the left depth frame is overwritten AFTER deproject()  -> the cloud must have captured the OLD frame (== 0)
 *   3  a third transform on the fused cloud (lazy as long as it can be)
 *   4  the same sequence, materialised from the first call on (== 3)
 *   then uint32 n + float4[n]: right cloud after deproject + transform(T_R), read back with copyToHost
 */
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pointcloud_processing/frameset.h"
#include "pointcloud_processing/kernels.cuh"
#include "pointcloud_processing/pointcloud.h"

struct Header {
	int32_t width, height;
	float fx, fy, ppx, ppy;
	float depth_scale, min_depth, max_depth;
	int32_t mirror, post_filter;
	int32_t out_width, out_height;
	float out_fx, out_fy, out_ppx, out_ppy;
	double right_transform[16];
	double fused_transform[16];
};

template <typename T>
static std::vector<T> read_vec(FILE* f, size_t n) {
	std::vector<T> v(n);
	if (std::fread(v.data(), sizeof(T), n, f) != n) std::exit(2);
	return v;
}

int main(int argc, char** argv) {
	if (argc != 3) return 2;
	FILE* in = std::fopen(argv[1], "rb");
	if (!in) return 2;
	Header h;
	if (std::fread(&h, sizeof(h), 1, in) != 1) return 2;
	const size_t pix = static_cast<size_t>(h.width) * h.height;
	auto depth_left = read_vec<uint16_t>(in, pix);
	auto color_left = read_vec<uint8_t>(in, pix * 3);
	auto depth_right = read_vec<uint16_t>(in, pix);
	auto color_right = read_vec<uint8_t>(in, pix * 3);
	std::fclose(in);

	Kernels::init_cuda(false);
	cudaStream_t stream_left, stream_right;
	cudaStreamCreateWithFlags(&stream_left, cudaStreamNonBlocking);
	cudaStreamCreateWithFlags(&stream_right, cudaStreamNonBlocking);
	Intrinsics intr{};
	intr.width = h.width;
	intr.height = h.height;
	intr.fx = h.fx;
	intr.fy = h.fy;
	intr.ppx = h.ppx;
	intr.ppy = h.ppy;
	Intrinsics fused_intr = intr;
	fused_intr.width = h.out_width;
	fused_intr.height = h.out_height;
	fused_intr.fx = h.out_fx;
	fused_intr.fy = h.out_fy;
	fused_intr.ppx = h.out_ppx;
	fused_intr.ppy = h.out_ppy;

	const size_t out_bytes = static_cast<size_t>(h.out_width) * h.out_height * 3;
	std::vector<std::vector<uint8_t>> images(5, std::vector<uint8_t>(out_bytes));
	std::vector<float> cloud_host;
	uint32_t cloud_n = 0;
	{
		Frameset frameset_left, frameset_right, fused_frameset;
		frameset_left.setStream(&stream_left);
		frameset_right.setStream(&stream_right);
		frameset_left.allocateDepthFrame(intr);
		frameset_left.allocateColorFrame(intr);
		frameset_right.allocateDepthFrame(intr);
		frameset_right.allocateColorFrame(intr);
		fused_frameset.setStream(&stream_left);
		fused_frameset.allocateColorFrame(fused_intr);
		/* a small extra rigid motion for the three-transform cases */
		const double extra[16] = {1, 0, 0, 0.01, 0, 1, 0, -0.02, 0, 0, 1, 0.005, 0, 0, 0, 1};
		std::vector<uint16_t> zeros(pix, 0);

		for (int mode = 0; mode < 5; ++mode) {
			frameset_left.setColorFrame(color_left.data());
			frameset_left.setDepthFrame(depth_left.data());
			frameset_right.setColorFrame(color_right.data());
			frameset_right.setDepthFrame(depth_right.data());
			const std::array<int, 4> roi_left = {0, 0, h.width, h.height};
			frameset_left.filterDepth(h.min_depth, h.max_depth, h.depth_scale, roi_left);
			frameset_right.filterDepth(h.min_depth, h.max_depth, h.depth_scale);

			Pointcloud fused_cloud;
			fused_cloud.setStream(frameset_left.getStream());
			fused_cloud.allocate(frameset_left.getMaskCount() + frameset_right.getMaskCount());
			if (mode == 4) (void)fused_cloud.get(); /* eager from the start */
			fused_cloud.deproject(frameset_left, h.depth_scale);
			Pointcloud cloud_right;
			cloud_right.setStream(frameset_right.getStream());
			cloud_right.deproject(frameset_right, h.depth_scale);
			cloud_right.transform(h.right_transform);
			cudaStreamSynchronize(*frameset_right.getStream());
			if (mode == 0) { /* copyToHost materialises a lazy cloud; the appended copy must not care */
				cloud_n = cloud_right.getPointCount();
				Pointcloud probe;
				probe.setStream(frameset_right.getStream());
				probe.deproject(frameset_right, h.depth_scale);
				probe.transform(h.right_transform);
				cloud_host.resize(static_cast<size_t>(cloud_n) * 4 + 4);
				probe.copyToHost(cloud_host.data());
			}
			fused_cloud.append(cloud_right);
			fused_cloud.transform(h.fused_transform);
			if (mode == 3 || mode == 4) fused_cloud.transform(extra);
			if (mode == 1) (void)fused_cloud.get();
			if (mode == 2) {
				frameset_left.setDepthFrame(zeros.data()); /* the source changes under a lazy cloud */
				frameset_left.sync();
			}
			fused_cloud.project(fused_frameset, h.mirror != 0);
			fused_frameset.copyColorToHost(images[mode].data());
		}
	}
	FILE* out = std::fopen(argv[2], "wb");
	if (!out) return 2;
	for (auto& im : images) std::fwrite(im.data(), 1, im.size(), out);
	std::fwrite(&cloud_n, sizeof(uint32_t), 1, out);
	std::fwrite(cloud_host.data(), sizeof(float), static_cast<size_t>(cloud_n) * 4, out);
	std::fclose(out);
	cudaStreamDestroy(stream_left);
	cudaStreamDestroy(stream_right);
	return 0;
}
