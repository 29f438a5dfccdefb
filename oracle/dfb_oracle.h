/*
 * dfb_oracle.h — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's per-frame fusion hot path
 * (MMM-Jasmin/PointCloud-DepthFusion, package pointcloud_processing + the call sequence of
 * fusion_node / registration_node). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or call this library; the product
 * (libdfb_b200.so) never does and has no CPU fallback.
 *
 * Parity pinning: the reference ships no tests, fixtures or golden vectors, so
 * this restatement is pinned against the reference's OWN kernels.cu compiled for sm_100a
 * (oracle/_ref, see oracle/Makefile) and run on a B200: tests/test_ref_kernels_gpu.py on the
 * GPU box, and tests/golden/ fixtures (generated there by tests/golden/make_golden.py) on CPU.
 * Parts with no reference-side pin say "parity unpinned" at their declaration.
 *
 * Every function cites the reference file:line it follows. Arithmetic follows SURVEY.md
 * App. A: fp32, no contraction (-ffp-contract=off) except where nvcc contracts the
 * reference source, which is written out with fmaf().
 */
#ifndef DFB_ORACLE_H
#define DFB_ORACLE_H

#include <stdint.h>

#include "dfb/dfb_synth.h"
#include "dfb/dfb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* kernels.cu:357-359 — thresholds are (Npp16u)(float / float) */
void dfo_thresholds(float min_depth, float max_depth, float depth_scale, uint16_t* lt, uint16_t* gt);

/* kernels.cu:379-384 — ROI resolution; out = {off_x, off_y, size_x, size_y}; additionally clamps an
 * over-wide ROI to the frame (the reference would overrun rows) */
void dfo_resolve_roi(const int32_t roi[4], int width, int height, int32_t out[4]);

/* DepthFrame::filter, depth_frame.cpp:153-182 = filter_depth_minmax_npp (kernels.cu:353-369) +
 * filter_create_mask_npp (:371-395) + filter_mask_count_npp (:501-540) + create_img_indices
 * (:542-547, compactor.cuh:28-68). depth is zeroed in place outside [lt,gt]; mask (255/0) is
 * written inside the ROI only; the count is the whole-mask sum / 255; indices are the
 * ascending pixel indices with mask > 0. Returns mask_count. */
uint32_t dfo_filter_depth(uint16_t* depth, uint8_t* mask, uint32_t* indices, int width, int height,
                          float min_depth, float max_depth, float depth_scale, const int32_t roi[4]);

/* kernel_deproject_depth with colour, kernels.cu:183-204 + :70-79. points = float4[count];
 * w carries bytes {B,G,R,0}. */
void dfo_deproject_color(float* points, const uint16_t* depth, const uint8_t* color, const dfb_intrinsics* intr,
                         const uint32_t* indices, uint32_t count, float depth_scale);

/* kernel_deproject_depth without colour, kernels.cu:206-221 + :81-90; w = 1.0f */
void dfo_deproject(float* points, const uint16_t* depth, const dfb_intrinsics* intr, const uint32_t* indices,
                   uint32_t count, float depth_scale);

/* kernel_transform, kernels.cu:161-181, with nvcc's contraction written out:
 * x' = fl(fma(T2,z, fma(T0,x, fl(T1*y))) + T3); w untouched. T is row-major 4x4. */
void dfo_transform(float* points, uint32_t count, const float T[16]);

/* kernel_project_pointcloud, kernels.cu:239-274, executed SEQUENTIALLY in point order (the
 * meaning of the racy GPU code). color_out (W*H*3) is zeroed and z_buffer (W*H)
 * set to FLT_MAX first, as Kernels::project_pointcloud does (:582-587). */
void dfo_project_sequential(uint8_t* color_out, float* z_buffer, const float* points, uint32_t count,
                            const dfb_intrinsics* intr, int mirror);

/* Pixel index of one point under kernels.cu:247-261; -1 when culled (also culls Z<=0 / NaN, which
 * the reference leaves undefined). */
int32_t dfo_project_index(const float* point4, const dfb_intrinsics* intr, int mirror);

/* Canonical deterministic splat: winner per pixel = minimum 64-bit key
 * (float_bits(Z) << 32) | (cam << 24) | (R << 16) | (G << 8) | B. seg_offsets[n_seg+1] gives
 * the point range of each camera segment. keys (W*H, may be NULL) receives the winning keys
 * (UINT64_MAX = empty); ambiguous (W*H, may be NULL) is set to 1 where two points of the SAME
 * camera tie on Z with different colour (the only place canonical != sequential). */
void dfo_project_canonical(uint8_t* color_out, uint64_t* keys, uint8_t* ambiguous, const float* points,
                           const uint32_t* seg_offsets, int n_seg, const dfb_intrinsics* intr, int mirror);

/* The whole processSyncedFrames sequence (fusion_node.cpp:718-784) for one frameset of
 * desc->n_cameras cameras: filter -> deproject(colour) -> per-camera transform -> append in
 * camera order -> fused transform -> project. depth/color are arrays of per-camera pointers
 * (depth is NOT modified; a scratch copy is filtered). mode 0 = sequential reference
 * semantics, 1 = canonical key semantics. Optional outputs may be NULL:
 * cloud_out (float4[sum of mask counts], post-transform), counts_out[n_cameras],
 * ambiguous (W_out*H_out). Returns total point count, or -1 on allocation failure. */
int64_t dfo_fuse_frameset(const dfb_fuse_desc* desc, const uint16_t* const* depth, const uint8_t* const* color,
                          uint8_t* fused_out, int mode, float* cloud_out, uint32_t* counts_out, uint8_t* ambiguous);

/* Batch of framesets stored as pools depth[F][N_cam][H][W], color[F][N_cam][H][W][3],
 * out[F][H_out][W_out][3] (all cameras must share one resolution). Runs dfo_fuse_frameset over
 * n_threads OpenMP threads, one frameset per thread at a time. Returns 0 on success. */
int dfo_fuse_batch(const dfb_fuse_desc* desc, const uint16_t* depth, const uint8_t* color, uint8_t* out,
                   int n_framesets, int mode, int n_threads);

/* registration_node.cpp:313-333 — setDepthFrame + filterDepth + Pointcloud::deproject (no colour)
 * for one camera; cloud_out is float4[W*H] capacity, w = 1.0f. Returns the point count. */
uint32_t dfo_registration_cloud(const dfb_camera_desc* cam, const uint16_t* depth, float* cloud_out);

/* Frameset::alignDepthToColor (frameset.cpp:180-188, kernels.cu:276-311,137-159,714-747): unaligned depth image
 * (depth intrinsics) -> depth image in the colour camera's frame. `aligned` = uint16[colour W*H]. Returns 0, or -1
 * for the distortion models the reference cannot run (asserted away, or CUDA-libm only) and odd pixel counts. */
int dfo_align_depth_to_color(const uint16_t* depth, const dfb_intrinsics* depth_intr, const dfb_intrinsics* color_intr,
                             const dfb_extrinsics* depth_to_color, float depth_scale, uint16_t* aligned);

/* nppiFilterGauss_8u_C3R 3x3 on the interior (W-2)x(H-2), kernels.cu:632-653. Kernel
 * = separable Gaussian taps exp(-d^2/2)/sum (sigma 1.0, the value NPP documents for a 3x3 mask), fp32
 * accumulation, result TRUNCATED. PARITY UNPINNED in the reference (NPP is closed source); pinned by us
 * against NPP 12.4 on the B200 (tests/test_ref_kernels_gpu.py). */
void dfo_filter_color_gauss(uint8_t* color, int width, int height);

/* nppiFilterMedian_8u_C3R 3x3 on the interior, kernels.cu:594-630. PARITY UNPINNED (see above). */
void dfo_filter_color_median(uint8_t* color, int width, int height);

/* FusionNode::interpolateTransform, fusion_node.cpp:589-604, restated without Eigen
 * (Eigen 3.4.0 Quaterniond(Matrix3d) + slerp; row-major 4x4 double in/out). PARITY UNPINNED
 * (Eigen absent); affects only the INPUT matrices of the hot path. */
void dfo_interpolate_transform(const double left[16], const double right[16], double out[16]);

/* fused_transform.prerotate(AngleAxisd(90deg, UnitZ)), fusion_node.cpp:774-777 */
void dfo_prerotate_z90(double T[16]);

/* Eigen::Affine3d -> row-major float[16], pointcloud.cpp:65 */
void dfo_affine_to_float16(const double T[16], float out[16]);

/* fused intrinsics construction, fusion_node.cpp:155-165 */
void dfo_fused_intrinsics(const dfb_intrinsics* color_left, int vertical, dfb_intrinsics* out);

/* pcl::ApproximateVoxelGrid<PointXYZ> (PCL 1.12.0 approximate_voxel_grid.hpp, restated from the
 * published algorithm; registration.cpp:144-150). in = float4[n] (xyz used), out = float4 capacity
 * n. Returns output count. PARITY UNPINNED (PCL not vendored, no reference test). */
uint32_t dfo_voxel_approximate(const float* in, uint32_t n, float leaf, float* out);

/* Exact voxel grid: one centroid per occupied voxel, sorted by (iz,iy,ix). voxel_ids_out (int32[3] per
 * voxel) may be NULL. Returns voxel count. Used for the set-equality parity criterion (App. A.7). */
uint32_t dfo_voxel_exact(const float* in, uint32_t n, float leaf, float* out, int32_t* voxel_ids_out);

/* synthetic inputs (dfb_synth.h) for one camera image on the CPU */
void dfo_synth_camera(const dfs_scene* scene, const dfs_camera* cam, uint32_t cam_index, uint16_t* depth,
                      uint8_t* color);

int dfo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* DFB_ORACLE_H */
