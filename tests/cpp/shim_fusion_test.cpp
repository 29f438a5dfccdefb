/*
 * shim_fusion_test.cpp — a C++ consumer of the drop-in headers (include/pointcloud_processing/*.h) that
 * replays FusionNode::processSyncedFrames (fusion_node/src/fusion_node.cpp:718-794) and the registration
 * node's deprojection (registration_node/src/registration_node.cpp:313-333) call for call, the way the
 * reference nodes use the library. tests/test_shim_gpu.py feeds it synthetic framesets and compares its
 * output with the CPU oracle.
 *
 *   shim_fusion_test <input.bin> <output.bin> [bench_frames]
 * bench_frames > 0: afterwards the same call sequence runs bench_frames more times with the camera images already on the
 * device (copied D2D into the framesets, like oracle/ref_driver.cu's ref_fuse_bench does for the reference kernels) and
 * one JSON line with the frame rate goes to stdout — bench.py's "shim" leg.
 * input : Header, then depth_left, color_left, depth_right, color_right
 * output: uint32 nL, nR, then the fused rgb8 image, then uint32 n_reg + the registration cloud (float4) of the right camera
 */
#include <chrono>
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
	int32_t mirror, post_filter; /* post_filter: 0 none, 1 gauss, 2 median */
	int32_t out_width, out_height;
	float out_fx, out_fy, out_ppx, out_ppy;
	double right_transform[16]; /* row-major Affine3d of the registration result */
	double fused_transform[16]; /* row-major, already interpolated / prerotated by the caller */
};

#ifdef SHIM_TEST_BACKTRACE /* debugging aid: nvcc -g -DSHIM_TEST_BACKTRACE -Xcompiler -rdynamic */
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void on_segv(int) {
	void* frames[64];
	const int n = backtrace(frames, 64);
	backtrace_symbols_fd(frames, n, STDERR_FILENO);
	_exit(139);
}
#endif

template <typename T>
static std::vector<T> read_vec(FILE* f, size_t n) {
	std::vector<T> v(n);
	if (std::fread(v.data(), sizeof(T), n, f) != n) {
		std::fprintf(stderr, "short read\n");
		std::exit(2);
	}
	return v;
}

int main(int argc, char** argv) {
#ifdef SHIM_TEST_BACKTRACE
	signal(SIGSEGV, on_segv);
#endif
	if (argc != 3 && argc != 4) return 2;
	const int bench_frames = argc == 4 ? std::atoi(argv[3]) : 0;
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

	/* FusionNode::init: Kernels::init_cuda + two non-blocking streams (fusion_node.cpp:204-206) */
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
	intr.model = DISTORTION_BROWN_CONRADY;
	Intrinsics fused_intr = intr;
	fused_intr.width = h.out_width;
	fused_intr.height = h.out_height;
	fused_intr.fx = h.out_fx;
	fused_intr.fy = h.out_fy;
	fused_intr.ppx = h.out_ppx;
	fused_intr.ppy = h.out_ppy;

	/* FusionNode::allocateFrames (fusion_node.cpp:535-563) */
	Frameset frameset_left, frameset_right, fused_frameset;
	frameset_left.setStream(&stream_left);
	frameset_right.setStream(&stream_right);
	frameset_left.allocateDepthFrame(intr);
	frameset_left.allocateColorFrame(intr);
	frameset_right.allocateDepthFrame(intr);
	frameset_right.allocateColorFrame(intr);
	fused_frameset.setStream(&stream_left);
	fused_frameset.allocateColorFrame(fused_intr);
	std::vector<uint8_t> fused(static_cast<size_t>(h.out_width) * h.out_height * 3);

	unsigned n_left = 0, n_right = 0;
	for (int frame = 0; frame < 2; ++frame) { /* twice: the second pass checks that no state leaks between frames */
		/* processSyncedFrames, fusion_node.cpp:718-794 */
		frameset_left.setColorFrame(color_left.data());
		frameset_left.setDepthFrame(depth_left.data());
		frameset_right.setColorFrame(color_right.data());
		frameset_right.setDepthFrame(depth_right.data());
		const std::array<int, 4> roi_left = {0, 0, h.width, h.height};
		frameset_left.filterDepth(h.min_depth, h.max_depth, h.depth_scale, roi_left);
		frameset_right.filterDepth(h.min_depth, h.max_depth, h.depth_scale);
		n_left = frameset_left.getMaskCount();
		n_right = frameset_right.getMaskCount();

		Pointcloud fused_cloud;
		fused_cloud.setStream(frameset_left.getStream());
		fused_cloud.allocate(frameset_left.getMaskCount() + frameset_right.getMaskCount());
		fused_cloud.deproject(frameset_left, h.depth_scale);
		Pointcloud cloud_right;
		cloud_right.setStream(frameset_right.getStream());
		cloud_right.deproject(frameset_right, h.depth_scale);

		cloud_right.transform(h.right_transform);
		cudaStreamSynchronize(*frameset_right.getStream());
		fused_cloud.append(cloud_right);
		fused_cloud.transform(h.fused_transform);
		fused_cloud.project(fused_frameset, h.mirror != 0);
		if (h.post_filter) fused_frameset.filterColor(h.post_filter == 2);
		fused_frameset.copyColorToHost(fused.data());
	}

	if (bench_frames > 0) {
		/* processSyncedFrames (fusion_node.cpp:739-789) per frame, camera images resident on the device: the uploads of
		 * :718-735 are replaced by D2D copies into the framesets and the download of :794 by a D2D copy out, exactly what
		 * ref_fuse_bench times for the reference's own kernels */
		uint16_t *dl = nullptr, *dr = nullptr;
		uint8_t *cl = nullptr, *cr = nullptr, *sink = nullptr;
		cudaMalloc(&dl, pix * 2);
		cudaMalloc(&dr, pix * 2);
		cudaMalloc(&cl, pix * 3);
		cudaMalloc(&cr, pix * 3);
		cudaMalloc(&sink, fused.size());
		cudaMemcpy(dl, depth_left.data(), pix * 2, cudaMemcpyHostToDevice);
		cudaMemcpy(dr, depth_right.data(), pix * 2, cudaMemcpyHostToDevice);
		cudaMemcpy(cl, color_left.data(), pix * 3, cudaMemcpyHostToDevice);
		cudaMemcpy(cr, color_right.data(), pix * 3, cudaMemcpyHostToDevice);
		auto frame = [&]() {
			cudaMemcpyAsync(frameset_left.getColor(), cl, pix * 3, cudaMemcpyDeviceToDevice, stream_left);
			cudaMemcpyAsync(frameset_left.getDepth(), dl, pix * 2, cudaMemcpyDeviceToDevice, stream_left);
			cudaMemcpyAsync(frameset_right.getColor(), cr, pix * 3, cudaMemcpyDeviceToDevice, stream_right);
			cudaMemcpyAsync(frameset_right.getDepth(), dr, pix * 2, cudaMemcpyDeviceToDevice, stream_right);
			const std::array<int, 4> roi_left = {0, 0, h.width, h.height};
			frameset_left.filterDepth(h.min_depth, h.max_depth, h.depth_scale, roi_left);
			frameset_right.filterDepth(h.min_depth, h.max_depth, h.depth_scale);
			Pointcloud fused_cloud;
			fused_cloud.setStream(frameset_left.getStream());
			fused_cloud.allocate(frameset_left.getMaskCount() + frameset_right.getMaskCount());
			fused_cloud.deproject(frameset_left, h.depth_scale);
			Pointcloud cloud_right;
			cloud_right.setStream(frameset_right.getStream());
			cloud_right.deproject(frameset_right, h.depth_scale);
			cloud_right.transform(h.right_transform);
			cudaStreamSynchronize(*frameset_right.getStream());
			fused_cloud.append(cloud_right);
			fused_cloud.transform(h.fused_transform);
			fused_cloud.project(fused_frameset, h.mirror != 0);
			if (h.post_filter) fused_frameset.filterColor(h.post_filter == 2);
			cudaMemcpyAsync(sink, fused_frameset.getColor(), fused.size(), cudaMemcpyDeviceToDevice, stream_left);
			fused_frameset.sync(); /* copyColorToHost ends every frame with this wait (color_frame.cpp:79) */
		};
		for (int i = 0; i < 20; ++i) frame();
		cudaDeviceSynchronize();
		const auto t0 = std::chrono::steady_clock::now();
		for (int i = 0; i < bench_frames; ++i) frame();
		cudaDeviceSynchronize();
		const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		std::vector<uint8_t> check(fused.size());
		cudaMemcpy(check.data(), sink, fused.size(), cudaMemcpyDeviceToHost);
		std::printf("{\"shim_framesets_per_s\": %.1f, \"frames\": %d, \"seconds\": %.6f, \"matches_first_pass\": %s}\n",
		            bench_frames / dt, bench_frames, dt, check == fused ? "true" : "false");
		cudaFree(dl);
		cudaFree(dr);
		cudaFree(cl);
		cudaFree(cr);
		cudaFree(sink);
	}

	/* RegistrationNode::icp deprojection, registration_node.cpp:313-333: depth only, w = 1.0f */
	Frameset source_frameset;
	source_frameset.setStream(&stream_right);
	source_frameset.allocateDepthFrame(intr);
	source_frameset.setDepthFrame(depth_right.data());
	source_frameset.filterDepth(h.min_depth, h.max_depth, h.depth_scale);
	Pointcloud source_cloud;
	source_cloud.setStream(&stream_right);
	source_cloud.deproject(source_frameset, h.depth_scale);
	std::vector<float> reg(static_cast<size_t>(source_cloud.getPointCount()) * 4 + 4);
	source_cloud.copyToHost(reg.data());
	const uint32_t n_reg = source_cloud.getPointCount();

	FILE* out = std::fopen(argv[2], "wb");
	if (!out) return 2;
	const uint32_t counts[2] = {n_left, n_right};
	std::fwrite(counts, sizeof(uint32_t), 2, out);
	std::fwrite(fused.data(), 1, fused.size(), out);
	std::fwrite(&n_reg, sizeof(uint32_t), 1, out);
	std::fwrite(reg.data(), sizeof(float), static_cast<size_t>(n_reg) * 4, out);
	std::fclose(out);
	cudaStreamDestroy(stream_left);
	cudaStreamDestroy(stream_right);
	return 0;
}
