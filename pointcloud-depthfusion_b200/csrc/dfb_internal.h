/*
 * dfb_internal.h — shared between the translation units of libdfb_b200.so (not installed).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "dfb/dfb.h"
#include "dfb_math.cuh"

/* scratch shared by the eager operators (cloud_ops.cu) and the voxel grid (voxel.cu); grow-only */
struct dfb_workspace {
	uint32_t* d_block_counts = nullptr; /* counts, then exclusive offsets; [capacity + 1] (last = total) */
	size_t capacity = 0;
	uint32_t* h_total = nullptr;        /* pinned */
	cudaEvent_t count_ready = nullptr;  /* dfb_op_filter_depth_begin / dfb_op_count_depth_begin -> dfb_op_filter_depth_count */
	uint32_t* d_count_acc = nullptr;    /* dfb_op_count_depth_begin: accumulator + self-resetting ticket */
	void* d_voxel_tmp = nullptr;        /* sort keys + point indices of dfb_voxel_downsample */
	size_t voxel_tmp_bytes = 0;
};

namespace dfb {

/* ---- error plumbing ---------------------------------------------------------------------------- */
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int require_device();

extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

/* source rows per splat item of the stream kernel (64 columns wide): fuse_stream.cu, the per-item ray table
 * (fuse_column.cu tile_rays_kernel) and its allocation (capi_core.cu) must agree */
#ifndef DFB_STREAM_ITEM_ROWS
#define DFB_STREAM_ITEM_ROWS 64
#endif
constexpr int kStreamItemRows = DFB_STREAM_ITEM_ROWS;
static_assert(kStreamItemRows == 64 || kStreamItemRows == 128, "splat items are 64 x 64 or 64 x 128 source pixels");

/* scan.cu: in-place exclusive prefix sum of u32 words in two launches (see there) */
size_t exclusive_scan_scratch_words(uint32_t n);
int launch_exclusive_scan(uint32_t* d_data, uint32_t n, uint32_t* d_total, uint32_t* d_entry_begin, uint32_t stride,
                          uint32_t* d_scratch, cudaStream_t s);

#define DFB_CUDA(call)                                                      \
	do {                                                                    \
		cudaError_t dfb_e_ = (call);                                        \
		if (dfb_e_ != cudaSuccess) return ::dfb::cuda_fail(dfb_e_, #call, __FILE__, __LINE__); \
	} while (0)

#define DFB_REQUIRE(cond, msg)                  \
	do {                                        \
		if (!(cond)) {                          \
			::dfb::set_error(msg);              \
			return DFB_ERR_INVALID;             \
		}                                       \
	} while (0)

constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;

/* ---- kernel parameter blocks (passed by value: they live in the constant bank) ------------------ */

struct CamParams {
	int width, height;
	int roi_x0, roi_y0, roi_w, roi_h; /* resolved ROI (kernels.cu:379-384), clamped to the frame */
	uint32_t lt, gt;                  /* integer depth thresholds (kernels.cu:357-359) */
	float scale;                      /* depth_scale */
	float ppx, ppy, fx, fy;
	int has_transform;
	Mat34 to_ref;
	/* the same depth test in float form for the packed kernels: valid <=> |float(d) - mid| <= halfw with
	 * mid = (max(lt,1)+gt)/2, halfw = (gt-max(lt,1))/2 (both exact in fp32; halfw < 0 for an empty range) */
	float mid, halfw;
	int roi_all;                      /* ROI covers the whole frame */
};

struct FuseParams {
	int n_cameras;
	int n_out_pixels;
	CamParams cam[DFB_MAX_CAMERAS];
	Mat34 fused;
	ProjParams proj;
};

/* derive CamParams / FuseParams from the public descriptors (host) */
int make_cam_params(const dfb_camera_desc& cd, CamParams& out);
int make_fuse_params(const dfb_fuse_desc& d, FuseParams& out);

/* ---- kernel launchers (defined in the .cu files) ------------------------------------------------- */

/* fused filter->deproject->transform(s)->project->atomicMin over `group` framesets; direct (register) path */
int launch_splat_direct(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color,
                        unsigned long long* d_keys, int group, cudaStream_t s);
/* same arithmetic, lanes run down source columns (coalesced atomics for the 90-degree rotated view);
 * d_xtab = float[n_cameras][row_pitch] deprojection rays (u - ppx)/fx, d_ytab likewise for rows */
struct RayTables {
	const float* x;
	const float* y;
	int x_pitch, y_pitch; /* floats per camera */
	const float* tiles;   /* [camera][64 x 64 tile][64 column rays, 64 row rays]: the stream kernel's copy of the same values */
};
bool splat_column_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color);
int launch_splat_column(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color,
                        unsigned long long* d_keys, int group, cudaStream_t s);
/* persistent CTAs + TMA tiles (cp.async.bulk.tensor) + the same per-pair body; ctas_per_sm <= 0 = default */
bool splat_tiled_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color);
int launch_splat_tiled(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color,
                       unsigned long long* d_keys, int group, int ctas_per_sm, cudaStream_t s);
/* the whole batch (splat of every camera + resolve of every image) in one persistent launch over a ring of
 * `slots` L2-resident key slots; d_counters = fuse_stream_counter_words(n) uint32 (zeroed by the launcher) */
bool fuse_stream_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color, const uint8_t* d_out);
size_t fuse_stream_counter_words(int n_framesets);
size_t fuse_stream_scratch_bytes(); /* extra bytes behind the ALLOCATED key ring (ring_slots slots) */
long long fuse_stream_tickets_per_frameset(const FuseParams& P); /* a launch must stay below 2^23 tickets */
/* where the images of a frameset sit in the depth / colour pools (nullptr = packed [frameset][camera][H][W]) */
struct StreamLayout {
	int imgs_per_fs;
	unsigned char slot_d[DFB_MAX_CAMERAS], slot_c[DFB_MAX_CAMERAS]; /* image f * imgs_per_fs + slot[c] holds camera c */
	size_t depth_stride_bytes, color_stride_bytes;                  /* distance between consecutive images (multiple of 16) */
};
int launch_fuse_stream(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                       unsigned long long* d_keys, int ring_slots, unsigned int* d_counters, int n_framesets, int slots, int lag,
                       bool row_major, const StreamLayout* layout, uint32_t* ring_phase, cudaStream_t s);
/* fills the ray tables with __fdiv_rn(__fsub_rn(i, pp), f) (kernels.cu:72-73), one launch */
int launch_ray_tables(const FuseParams& P, float* d_xtab, float* d_ytab, int x_pitch, int y_pitch, cudaStream_t s);
int launch_tile_rays(const FuseParams& P, float* d_tiles, int tiles_x, int tiles_per_image, cudaStream_t s);
/* keys -> rgb8 and re-arm the keys, over `group` framesets */
int launch_resolve(const FuseParams& P, unsigned long long* d_keys, uint8_t* d_out, int group, cudaStream_t s);
int launch_post_filter_batch(const uint8_t* d_src, uint8_t* d_dst, int width, int height, int n_frames, int median, cudaStream_t s);
int launch_keys_clear(unsigned long long* d_keys, size_t n, cudaStream_t s);

}  // namespace dfb
