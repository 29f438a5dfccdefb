#pragma once
/*
 * pointcloud.h — device float4 cloud: deproject / transform / append / project / copyToHost
 * (reference: pointcloud_processing/include/pointcloud_processing/pointcloud.h:19-62, src/pointcloud.cpp).
 *
 * LAZY by construction: allocate / deproject / transform / append only RECORD what the
 * reference does eagerly (pointcloud.cpp:21-129) — which frameset a segment comes from and which transforms it has
 * seen. project() then issues ONE fused deproject -> transform -> splat launch (dfb_fuser_fuse_images) that never
 * writes the float4 cloud to memory. get() / copyToHost() / set(), a recipe the fused kernel cannot express, or a
 * source frameset that is about to change force materialisation: the recorded operators run one by one through
 * dfb_op_* and produce exactly the cloud the reference's call sequence would have. Device memory comes from a
 * process-wide pool: the per-frame Pointcloud objects of fusion_node (fusion_node.cpp:746-752) cost no cudaMalloc.
 *
 * Differences that a caller can observe, all deliberate:
 *   - project() is deterministic: visibility is resolved with packed 64-bit keys (depth bits, camera of
 *     origin, colour) instead of the reference's racy float atomicMin + plain colour store
 *     (kernels.cu:264-273); the cloud remembers which segment came from which deproject()/append() call so
 *     that equal-depth ties break "first cloud first", as the sequential meaning of the reference does;
 *   - Eigen is optional: the Affine3d overloads exist when <eigen3/Eigen/Geometry> is available, the
 *     row-major double[16] / float[16] overloads always.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "error_check.h"
#include "frameset.h"
#include "kernels.cuh"

#if __has_include(<eigen3/Eigen/Geometry>)
#include <eigen3/Eigen/Geometry>
#define DFB_HAVE_EIGEN 1
#endif

class DFB_API Pointcloud {
public:
	Pointcloud();
	~Pointcloud();
	Pointcloud(const Pointcloud&) = delete;
	Pointcloud& operator=(const Pointcloud&) = delete;

	void setStream(cudaStream_t* stream = nullptr) { this->stream = stream ? *stream : nullptr; }
	void allocate(unsigned point_count);
	void free();
	float* get(); /* forces materialisation */
	unsigned getPointCount() { return point_count; }
	void set(const float* host_points);
	/* row-major 4x4 (what Eigen::Affine3d::matrix() holds, transposed to row-major: pointcloud.cpp:65) */
	void setTransform(const double transform_row_major[16]);
	void setTransform(const float transform_row_major[16]);
	void transform();
	void transform(const double transform_row_major[16]);
	void transform(const float transform_row_major[16]);
#ifdef DFB_HAVE_EIGEN
	void setTransform(const Eigen::Affine3d& t) {
		double m[16];
		for (int r = 0; r < 4; ++r)
			for (int c = 0; c < 4; ++c) m[4 * r + c] = t.matrix()(r, c);
		setTransform(m);
	}
	void transform(const Eigen::Affine3d& t) {
		setTransform(t);
		transform();
	}
#endif
	void deproject(Frameset& frameset, float depth_scale);
	void append(Pointcloud& cloud);
	void project(Frameset& frameset, bool mirror_image = true);
	void copyToHost(float* host_points);

	/* called by a Frameset before it changes (new frame, new filter, free): clouds that only recorded it compute now */
	void sourceWillChange(Frameset* source);

private:
	struct Segment { /* one deproject() / append() worth of points and what happened to them since */
		Frameset* source = nullptr;
		float depth_scale = 0.f;
		uint32_t count = 0;
		int n_transforms = 0;
		float transforms[2][16]; /* row-major, in order of application; a third transform forces materialisation */
	};
	void materialise();
	void record_transform(const float t[16]);
	bool try_fused_project(Frameset& target, bool mirror_image);
	void attach_sources();
	void detach_sources();
	void release_points();
	bool lazy = false;              /* dev_points does not hold the points: `recipe` describes them */
	std::vector<Segment> recipe;    /* parallel to `segments` while lazy */
	size_t points_bytes = 0, keys_bytes = 0; /* pool block sizes */
	cudaEvent_t points_event = nullptr, keys_event = nullptr; /* recorded after the last work given to the blocks */
	cudaStream_t stream = nullptr;
	unsigned point_count = 0;    /* allocated points (what transform / project / copyToHost cover) */
	float* dev_points = nullptr;
	unsigned pointer_offset = 0; /* next free point (the reference counts the same thing in floats/4) */
	float transform_rm[16];      /* row-major float 4x4, rounded from double as Eigen's cast<float>() does */
	uint64_t* dev_keys = nullptr; /* replaces the float z-buffer */
	size_t key_pixels = 0;
	std::vector<uint32_t> segments; /* points per deproject()/append() call, in order */
};
