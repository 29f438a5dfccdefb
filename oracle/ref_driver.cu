/*
 * ref_driver.cu — TEST INFRASTRUCTURE ONLY (oracle/_ref).
 *
 * A C-callable driver around the UNMODIFIED reference `Kernels::*` functions
 * (/root/reference/pointcloud_processing/src/kernels.cu, compiled in place for sm_100a by
 * oracle/Makefile). The reference's class layer (Frameset/DepthFrame/ColorFrame/Pointcloud)
 * cannot be built here (needs PCL + Eigen), so this file replays what those classes do, in
 * their order, with raw device buffers:
 *   DepthFrame::filter        depth_frame.cpp:153-182
 *   Pointcloud::deproject     pointcloud.cpp:100-111 -> frameset.cpp:140-156
 *   Pointcloud::transform     pointcloud.cpp:80-93
 *   Pointcloud::append        pointcloud.cpp:117-129
 *   Pointcloud::project       pointcloud.cpp:136-157
 *   Frameset::filterColor     frameset.cpp:194-208
 * in the sequence of FusionNode::processSyncedFrames (fusion_node.cpp:718-794).
 *
 * Two uses: (1) second oracle — pins oracle/dfb_oracle.c against the reference's own compiled
 * arithmetic on a B200; (2) "kernel to beat" timings (device-resident, CUDA events).
 * Nothing in the product links this.
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <npp.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "dfb/dfb_types.h"
#include "intrinsics.h"   /* reference header */
#include "kernels.cuh"    /* reference header */

static_assert(sizeof(Intrinsics) == sizeof(dfb_intrinsics), "Intrinsics layout must match");

#define REF_CHECK(x)                                                                     \
	do {                                                                                 \
		cudaError_t e_ = (x);                                                            \
		if (e_ != cudaSuccess) {                                                         \
			std::fprintf(stderr, "ref_driver: %s failed: %s\n", #x, cudaGetErrorString(e_)); \
			return -1;                                                                   \
		}                                                                                \
	} while (0)

namespace {

struct RefCam {
	uint16_t* d_depth = nullptr;
	uint8_t* d_color = nullptr;
	uint8_t* d_mask = nullptr;
	uint* d_indices = nullptr;
	Intrinsics* d_intr = nullptr;
	uint mask_count = 0;
	size_t pixels = 0;
};

struct RefPipeline {
	dfb_fuse_desc desc;
	cudaStream_t stream = nullptr;
	NppStreamContext npp{};
	RefCam cam[DFB_MAX_CAMERAS];
	float* d_cloud = nullptr;      /* fused cloud, capacity = all pixels */
	float* d_cloud_tmp = nullptr;  /* per-camera cloud before append */
	float* d_transform = nullptr;  /* 16 floats */
	float* d_zbuf = nullptr;
	uint8_t* d_out = nullptr;
	uint8_t* d_out_buf = nullptr;  /* ColorFrame scratch buffer for the post filter */
	Intrinsics* d_out_intr = nullptr;
	size_t out_pixels = 0;
	size_t cap = 0;
	bool ok = false;
};

int pipeline_create(RefPipeline& p, const dfb_fuse_desc* desc) {
	p.desc = *desc;
	REF_CHECK(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
	if (nppGetStreamContext(&p.npp) != NPP_NO_ERROR) return -1;
	p.npp.hStream = p.stream;
	p.cap = 0;
	for (int c = 0; c < desc->n_cameras; ++c) {
		RefCam& rc = p.cam[c];
		rc.pixels = (size_t)desc->cam[c].intr.width * desc->cam[c].intr.height;
		p.cap += rc.pixels;
		REF_CHECK(cudaMalloc(&rc.d_depth, rc.pixels * 2));
		REF_CHECK(cudaMalloc(&rc.d_color, rc.pixels * 3));
		REF_CHECK(cudaMalloc(&rc.d_mask, rc.pixels));
		REF_CHECK(cudaMemset(rc.d_mask, 0, rc.pixels)); /* depth_frame.cpp:46 */
		REF_CHECK(cudaMalloc(&rc.d_indices, rc.pixels * sizeof(uint)));
		REF_CHECK(cudaMalloc(&rc.d_intr, sizeof(Intrinsics)));
		REF_CHECK(cudaMemcpy(rc.d_intr, &desc->cam[c].intr, sizeof(Intrinsics), cudaMemcpyHostToDevice));
	}
	p.out_pixels = (size_t)desc->out_intr.width * desc->out_intr.height;
	REF_CHECK(cudaMalloc(&p.d_cloud, p.cap * 16));
	REF_CHECK(cudaMalloc(&p.d_cloud_tmp, p.cap * 16));
	REF_CHECK(cudaMalloc(&p.d_transform, 64));
	REF_CHECK(cudaMalloc(&p.d_zbuf, p.out_pixels * 4));
	REF_CHECK(cudaMalloc(&p.d_out, p.out_pixels * 3));
	REF_CHECK(cudaMalloc(&p.d_out_buf, p.out_pixels * 3));
	REF_CHECK(cudaMalloc(&p.d_out_intr, sizeof(Intrinsics)));
	REF_CHECK(cudaMemcpy(p.d_out_intr, &desc->out_intr, sizeof(Intrinsics), cudaMemcpyHostToDevice));
	p.ok = true;
	return 0;
}

void pipeline_destroy(RefPipeline& p) {
	for (int c = 0; c < DFB_MAX_CAMERAS; ++c) {
		RefCam& rc = p.cam[c];
		cudaFree(rc.d_depth);
		cudaFree(rc.d_color);
		cudaFree(rc.d_mask);
		cudaFree(rc.d_indices);
		cudaFree(rc.d_intr);
	}
	cudaFree(p.d_cloud);
	cudaFree(p.d_cloud_tmp);
	cudaFree(p.d_transform);
	cudaFree(p.d_zbuf);
	cudaFree(p.d_out);
	cudaFree(p.d_out_buf);
	cudaFree(p.d_out_intr);
	if (p.stream) cudaStreamDestroy(p.stream);
}

/* One frameset, device-resident inputs already in cam[c].d_depth / d_color.
 * post_filter: 0 none, 1 gauss, 2 median. */
int pipeline_run(RefPipeline& p, int post_filter) {
	const dfb_fuse_desc& d = p.desc;
	uint total = 0;
	for (int c = 0; c < d.n_cameras; ++c) {
		RefCam& rc = p.cam[c];
		const dfb_camera_desc& cd = d.cam[c];
		const uint w = (uint)cd.intr.width, h = (uint)cd.intr.height;
		/* DepthFrame::filter, depth_frame.cpp:171-180 */
		Kernels::filter_depth_minmax_npp(rc.d_depth, w, h, cd.min_depth, cd.max_depth, cd.depth_scale, p.npp);
		Kernels::filter_create_mask_npp(rc.d_depth, rc.d_mask, w, h, cd.roi, p.npp);
		Kernels::filter_mask_count_npp(rc.mask_count, rc.d_mask, w, h, p.stream, p.npp);
		Kernels::create_img_indices(rc.d_indices, rc.d_mask, (uint)rc.pixels, p.stream);
		/* Pointcloud::deproject -> Frameset::deprojectToPointcloud (colour exists) */
		float* dst = (c == 0) ? p.d_cloud : p.d_cloud_tmp;
		Kernels::deproject_depth_color(dst, rc.d_depth, rc.d_color, rc.d_intr, rc.d_indices, rc.mask_count,
		                               (uint)rc.pixels, cd.depth_scale, p.stream);
		if (cd.has_transform) {
			/* Pointcloud::transform(Affine3d): setTransform H2D + kernel, pointcloud.cpp:63-93 */
			REF_CHECK(cudaMemcpyAsync(p.d_transform, cd.to_ref, 64, cudaMemcpyHostToDevice, p.stream));
			Kernels::transform(dst, rc.mask_count, p.d_transform, p.stream);
		}
		if (c > 0) {
			REF_CHECK(cudaStreamSynchronize(p.stream)); /* fusion_node.cpp:758 */
			/* Pointcloud::append, pointcloud.cpp:117-129 */
			REF_CHECK(cudaMemcpyAsync(p.d_cloud + 4 * (size_t)total, dst, (size_t)rc.mask_count * 16,
			                          cudaMemcpyDeviceToDevice, p.stream));
		}
		total += rc.mask_count;
	}
	/* fused_cloud.transform(fused_transform), fusion_node.cpp:779 */
	REF_CHECK(cudaMemcpyAsync(p.d_transform, d.fused_transform, 64, cudaMemcpyHostToDevice, p.stream));
	Kernels::transform(p.d_cloud, total, p.d_transform, p.stream);
	/* fused_cloud.project, fusion_node.cpp:784 */
	Kernels::project_pointcloud(p.d_out, p.d_zbuf, (uint)p.out_pixels, p.d_cloud, total, p.d_out_intr, d.mirror != 0,
	                            p.stream);
	if (post_filter == 1)
		Kernels::filter_color_gauss_npp(p.d_out, p.d_out_buf, (uint)d.out_intr.width, (uint)d.out_intr.height, p.stream, p.npp);
	else if (post_filter == 2)
		Kernels::filter_color_median_npp(p.d_out, p.d_out_buf, (uint)d.out_intr.width, (uint)d.out_intr.height, p.stream, p.npp);
	return (int)total;
}

}  // namespace

extern "C" {

int ref_init(int device, int verbose) {
	if (cudaSetDevice(device) != cudaSuccess) return -1;
	if (device == 0) Kernels::init_cuda(verbose != 0); /* kernels.cu:325-351 (hard-codes device 0) */
	return 0;
}

/* Full reference sequence for one frameset given HOST inputs. Optional outputs may be NULL. */
int ref_fuse_frameset(const dfb_fuse_desc* desc, const uint16_t* const* depth, const uint8_t* const* color,
                      uint8_t* fused_out, float* cloud_out, uint32_t* counts_out, float* zbuf_out,
                      uint16_t* const* depth_filtered_out, uint8_t* const* mask_out, uint32_t* const* indices_out,
                      int post_filter) {
	RefPipeline p;
	int rc = -1;
	if (pipeline_create(p, desc) == 0) {
		bool ok = true;
		for (int c = 0; c < desc->n_cameras && ok; ++c) {
			ok = ok && cudaMemcpy(p.cam[c].d_depth, depth[c], p.cam[c].pixels * 2, cudaMemcpyHostToDevice) == cudaSuccess;
			ok = ok && cudaMemcpy(p.cam[c].d_color, color[c], p.cam[c].pixels * 3, cudaMemcpyHostToDevice) == cudaSuccess;
		}
		/* the reference leaves the deproject tail uninitialised at 848x480 (SURVEY App. A.6 i);
		 * make it deterministic zeros */
		ok = ok && cudaMemset(p.d_cloud, 0, p.cap * 16) == cudaSuccess;
		ok = ok && cudaMemset(p.d_cloud_tmp, 0, p.cap * 16) == cudaSuccess;
		if (ok) {
			const int total = pipeline_run(p, post_filter);
			ok = total >= 0 && cudaStreamSynchronize(p.stream) == cudaSuccess && cudaGetLastError() == cudaSuccess;
			if (ok) {
				rc = total;
				if (fused_out) cudaMemcpy(fused_out, p.d_out, p.out_pixels * 3, cudaMemcpyDeviceToHost);
				if (cloud_out) cudaMemcpy(cloud_out, p.d_cloud, (size_t)total * 16, cudaMemcpyDeviceToHost);
				if (zbuf_out) cudaMemcpy(zbuf_out, p.d_zbuf, p.out_pixels * 4, cudaMemcpyDeviceToHost);
				for (int c = 0; c < desc->n_cameras; ++c) {
					if (counts_out) counts_out[c] = p.cam[c].mask_count;
					if (depth_filtered_out && depth_filtered_out[c])
						cudaMemcpy(depth_filtered_out[c], p.cam[c].d_depth, p.cam[c].pixels * 2, cudaMemcpyDeviceToHost);
					if (mask_out && mask_out[c]) cudaMemcpy(mask_out[c], p.cam[c].d_mask, p.cam[c].pixels, cudaMemcpyDeviceToHost);
					if (indices_out && indices_out[c])
						cudaMemcpy(indices_out[c], p.cam[c].d_indices, (size_t)p.cam[c].mask_count * 4, cudaMemcpyDeviceToHost);
				}
			}
		}
	}
	pipeline_destroy(p);
	return rc;
}

/* "Kernel to beat": the reference sequence over a device-resident pool of framesets
 * (depth[F][N][H][W], color[F][N][H][W][3], out[F][Ho][Wo][3]); inputs are copied D2D into the
 * pipeline's frames (stands in for the H2D upload, not counted separately). Returns the mean
 * milliseconds per frameset over `iters` passes, measured with CUDA events. */
int ref_fuse_bench(const dfb_fuse_desc* desc, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                   int n_framesets, int iters, int post_filter, float* ms_per_frameset) {
	RefPipeline p;
	int rc = -1;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (pipeline_create(p, desc) == 0 && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) {
		const int nc = desc->n_cameras;
		const size_t pix = p.cam[0].pixels;
		bool ok = true;
		for (int it = -1; it < iters && ok; ++it) { /* it == -1: warm-up */
			if (it == 0) cudaEventRecord(e0, p.stream);
			for (int f = 0; f < n_framesets && ok; ++f) {
				for (int c = 0; c < nc; ++c) {
					cudaMemcpyAsync(p.cam[c].d_depth, d_depth + ((size_t)f * nc + c) * pix, pix * 2, cudaMemcpyDeviceToDevice, p.stream);
					cudaMemcpyAsync(p.cam[c].d_color, d_color + ((size_t)f * nc + c) * pix * 3, pix * 3, cudaMemcpyDeviceToDevice, p.stream);
				}
				ok = pipeline_run(p, post_filter) >= 0;
				if (d_out) cudaMemcpyAsync(d_out + (size_t)f * p.out_pixels * 3, p.d_out, p.out_pixels * 3, cudaMemcpyDeviceToDevice, p.stream);
			}
		}
		cudaEventRecord(e1, p.stream);
		ok = ok && cudaStreamSynchronize(p.stream) == cudaSuccess && cudaGetLastError() == cudaSuccess;
		if (ok) {
			float ms = 0.f;
			cudaEventElapsedTime(&ms, e0, e1);
			*ms_per_frameset = ms / (float)(iters * n_framesets);
			rc = 0;
		}
	}
	if (e0) cudaEventDestroy(e0);
	if (e1) cudaEventDestroy(e1);
	pipeline_destroy(p);
	return rc;
}

/* kernel_transform alone (KAT 6): points float4[n] host in/out */
int ref_transform(float* points, uint32_t n, const float T[16]) {
	float *d_p = nullptr, *d_t = nullptr;
	cudaStream_t s = nullptr;
	REF_CHECK(cudaMalloc(&d_p, (size_t)n * 16));
	REF_CHECK(cudaMalloc(&d_t, 64));
	REF_CHECK(cudaMemcpy(d_p, points, (size_t)n * 16, cudaMemcpyHostToDevice));
	REF_CHECK(cudaMemcpy(d_t, T, 64, cudaMemcpyHostToDevice));
	Kernels::transform(d_p, n, d_t, s);
	REF_CHECK(cudaDeviceSynchronize());
	REF_CHECK(cudaMemcpy(points, d_p, (size_t)n * 16, cudaMemcpyDeviceToHost));
	cudaFree(d_p);
	cudaFree(d_t);
	return 0;
}

/* kernel_project_pointcloud alone on an arbitrary host cloud */
int ref_project(uint8_t* color_out, float* zbuf_out, const float* points, uint32_t n, const dfb_intrinsics* intr, int mirror) {
	const size_t npix = (size_t)intr->width * intr->height;
	float *d_p = nullptr, *d_z = nullptr;
	uint8_t* d_c = nullptr;
	Intrinsics* d_i = nullptr;
	cudaStream_t s = nullptr;
	REF_CHECK(cudaMalloc(&d_p, (size_t)n * 16 + 16));
	REF_CHECK(cudaMalloc(&d_z, npix * 4));
	REF_CHECK(cudaMalloc(&d_c, npix * 3));
	REF_CHECK(cudaMalloc(&d_i, sizeof(Intrinsics)));
	REF_CHECK(cudaMemcpy(d_p, points, (size_t)n * 16, cudaMemcpyHostToDevice));
	REF_CHECK(cudaMemcpy(d_i, intr, sizeof(Intrinsics), cudaMemcpyHostToDevice));
	Kernels::project_pointcloud(d_c, d_z, (uint)npix, d_p, n, d_i, mirror != 0, s);
	REF_CHECK(cudaDeviceSynchronize());
	if (color_out) REF_CHECK(cudaMemcpy(color_out, d_c, npix * 3, cudaMemcpyDeviceToHost));
	if (zbuf_out) REF_CHECK(cudaMemcpy(zbuf_out, d_z, npix * 4, cudaMemcpyDeviceToHost));
	cudaFree(d_p);
	cudaFree(d_z);
	cudaFree(d_c);
	cudaFree(d_i);
	return 0;
}

/* Frameset::filterColor on a host image (pins the NPP gauss/median semantics) */
int ref_filter_color(uint8_t* color, int width, int height, int median) {
	const size_t bytes = (size_t)width * height * 3;
	uint8_t *d_c = nullptr, *d_b = nullptr;
	cudaStream_t s = nullptr;
	NppStreamContext ctx{};
	if (nppGetStreamContext(&ctx) != NPP_NO_ERROR) return -1;
	ctx.hStream = s;
	REF_CHECK(cudaMalloc(&d_c, bytes));
	REF_CHECK(cudaMalloc(&d_b, bytes));
	REF_CHECK(cudaMemcpy(d_c, color, bytes, cudaMemcpyHostToDevice));
	if (median)
		Kernels::filter_color_median_npp(d_c, d_b, (uint)width, (uint)height, s, ctx);
	else
		Kernels::filter_color_gauss_npp(d_c, d_b, (uint)width, (uint)height, s, ctx);
	REF_CHECK(cudaDeviceSynchronize());
	REF_CHECK(cudaMemcpy(color, d_c, bytes, cudaMemcpyDeviceToHost));
	cudaFree(d_c);
	cudaFree(d_b);
	return 0;
}

/* Frameset::alignDepthToColor (frameset.cpp:180-188) on host images: `aligned` = uint16[colour W*H] */
int ref_align_depth(uint16_t* aligned, const uint16_t* depth, const dfb_intrinsics* depth_intr, const dfb_intrinsics* color_intr,
                    const dfb_extrinsics* depth_to_color, float depth_scale) {
	static_assert(sizeof(Extrinsics) == sizeof(dfb_extrinsics), "Extrinsics layout must match");
	Intrinsics di, ci;
	Extrinsics ex;
	std::memcpy(&di, depth_intr, sizeof(di));
	std::memcpy(&ci, color_intr, sizeof(ci));
	std::memcpy(&ex, depth_to_color, sizeof(ex));
	const int n_depth = di.width * di.height, n_color = ci.width * ci.height;
	uint16_t *d_depth = nullptr;
	unsigned char* d_aligned = nullptr;
	Intrinsics *d_di = nullptr, *d_ci = nullptr;
	Extrinsics* d_ex = nullptr;
	int2* d_map = nullptr;
	cudaStream_t s = nullptr;
	REF_CHECK(cudaMalloc(&d_depth, (size_t)n_depth * 2));
	REF_CHECK(cudaMalloc(&d_aligned, (size_t)n_color * 2));
	REF_CHECK(cudaMalloc(&d_di, sizeof(Intrinsics)));
	REF_CHECK(cudaMalloc(&d_ci, sizeof(Intrinsics)));
	REF_CHECK(cudaMalloc(&d_ex, sizeof(Extrinsics)));
	REF_CHECK(cudaMalloc(&d_map, (size_t)n_depth * 2 * sizeof(int2))); /* DepthFrame::allocateUnaligned, depth_frame.cpp:88 */
	REF_CHECK(cudaMemcpy(d_depth, depth, (size_t)n_depth * 2, cudaMemcpyHostToDevice));
	REF_CHECK(cudaMemcpy(d_di, &di, sizeof(di), cudaMemcpyHostToDevice));
	REF_CHECK(cudaMemcpy(d_ci, &ci, sizeof(ci), cudaMemcpyHostToDevice));
	REF_CHECK(cudaMemcpy(d_ex, &ex, sizeof(ex), cudaMemcpyHostToDevice));
	Kernels::align_depth_to_color(d_aligned, d_depth, depth_scale, di, d_di, d_ex, ci, d_ci, n_depth, n_color, d_map, s);
	REF_CHECK(cudaDeviceSynchronize());
	REF_CHECK(cudaMemcpy(aligned, d_aligned, (size_t)n_color * 2, cudaMemcpyDeviceToHost));
	cudaFree(d_depth);
	cudaFree(d_aligned);
	cudaFree(d_di);
	cudaFree(d_ci);
	cudaFree(d_ex);
	cudaFree(d_map);
	return 0;
}

}  // extern "C"
