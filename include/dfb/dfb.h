/*
 * dfb.h — C ABI of libdfb_b200.so: the B200-native (sm_100a) replacement for the device side of
 * the reference's `pointcloud_processing` library on the per-frame fusion hot path.
 *
 * The reference has no FFI layer: its boundary is the C++ class API of a static library
 * (pointcloud_processing/CMakeLists.txt:28). The drop-in C++ headers in
 * include/pointcloud_processing/ (Frameset, Pointcloud, Kernels::init_cuda, Intrinsics) are thin
 * host code over THIS C ABI; each entry point below cites the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes; `stream` is a cudaStream_t passed as void* (NULL = legacy default
 *     stream); device pointers are ordinary CUDA device pointers owned by the caller unless
 *     stated otherwise;
 *   - every function returns 0 on success or a negative dfb_status; dfb_last_error() returns a
 *     thread-local message. (The reference prints and exit(1)s on any CUDA error,
 *     error_check.cpp:10-51; the C++ shim restores that behaviour on top of these codes.);
 *   - there is NO CPU fallback: without a usable sm_100 device every compute entry point fails
 *     with DFB_ERR_NO_DEVICE.
 */
#ifndef DFB_H
#define DFB_H

#include <stddef.h>
#include <stdint.h>

#include "dfb_synth.h"
#include "dfb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DFB_API __attribute__((visibility("default")))
#else
#define DFB_API
#endif

typedef enum dfb_status {
	DFB_OK = 0,
	DFB_ERR_INVALID = -1,   /* bad argument */
	DFB_ERR_CUDA = -2,      /* a CUDA runtime/driver call failed (message in dfb_last_error) */
	DFB_ERR_NO_DEVICE = -3, /* no CUDA device / not sm_100 */
	DFB_ERR_NOMEM = -4,
	DFB_ERR_UNSUPPORTED = -5
} dfb_status;

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Kernels::init_cuda(bool verbose), kernels.cu:325-351 — but the device is selectable (the reference
 * hard-codes device 0, SURVEY App. A.6 vi). Selects `device` for the calling thread and checks
 * that it is compute capability 10.x. */
DFB_API int dfb_init(int device, int verbose);
DFB_API const char* dfb_last_error(void);
DFB_API int dfb_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
DFB_API uint64_t dfb_launch_count(void);

/* pinned host memory for the *_host entry points (cudaHostAlloc / cudaFreeHost) */
DFB_API int dfb_host_alloc(void** ptr, size_t bytes);
DFB_API int dfb_host_free(void* ptr);

/* Device buffers that other processes of the same node can map (one process per GPU): the output rank allocates the
 * gather buffer, exports it, and every rank passes ITS slice of the mapped buffer as d_out of dfb_fuser_fuse_device —
 * the resolve's 16-byte stores then go straight over NVLink into the output rank's memory (peer stores), there is
 * no second pass over the fused frames. handle = 64 opaque bytes (cudaIpcMemHandle_t). */
#define DFB_IPC_HANDLE_BYTES 64
DFB_API int dfb_device_alloc(void** d_ptr, size_t bytes);
DFB_API int dfb_device_free(void* d_ptr);
DFB_API int dfb_ipc_export(void* d_ptr, unsigned char handle[DFB_IPC_HANDLE_BYTES]);
DFB_API int dfb_ipc_open(const unsigned char handle[DFB_IPC_HANDLE_BYTES], void** d_ptr);
DFB_API int dfb_ipc_close(void* d_ptr);

/* ---- host-side glue of fusion_node (pure host math, callable without a GPU) ------------------ */

/* FusionNode::interpolateTransform, fusion_node.cpp:589-604 (row-major 4x4 double) */
DFB_API void dfb_interpolate_transform(const double left[16], const double right[16], double out[16]);
/* fused_transform.prerotate(AngleAxisd(90 deg, UnitZ)), fusion_node.cpp:774-777 */
DFB_API void dfb_prerotate_z90(double T[16]);
/* Eigen::Affine3d -> row-major float[16], Pointcloud::setTransform, pointcloud.cpp:63-66 */
DFB_API void dfb_affine_to_float16(const double T[16], float out[16]);
/* fused-image intrinsics, fusion_node.cpp:155-165 */
DFB_API void dfb_fused_intrinsics(const dfb_intrinsics* color_left, int vertical, dfb_intrinsics* out);
/* FusionNode::cameraInfo2Intrinsics, fusion_node.cpp:570-581: K = row-major 3x3 double, D = 5 doubles;
 * ppx, ppy, fx, fy are truncated through int exactly as the reference does */
DFB_API void dfb_camera_info_to_intrinsics(uint32_t width, uint32_t height, const double K[9], const double D[5],
                                           dfb_intrinsics* out);
/* the virtual camera pose of fusion_node.cpp:168-180 (translation xyz in metres, rotation xyz in degrees) */
DFB_API void dfb_virtual_camera_transform(const double translation[3], const double rotation_deg[3], double out[16]);
/* Wire formats (pure host): one image of an incoming DepthFrameset message into a row of the fuser's host pools —
 * checks size and encoding against the camera, drops the row padding (step > width * bpp), byte-swaps big-endian depth;
 * processSyncedFrames hands msg.data straight to setDepthFrame / setColorFrame (fusion_node.cpp:712-735) and relies on
 * the camera node's exact layout. dst = uint16[height][width] / uint8[height][width][3] (pinned: dfb_host_alloc). */
DFB_API int dfb_wire_unpack_depth(const dfb_image_msg* msg, int width, int height, uint16_t* dst);
DFB_API int dfb_wire_unpack_color(const dfb_image_msg* msg, int width, int height, uint8_t* dst);
/* the header fields of the fused sensor_msgs/Image for a fused frame `data` (fusion_node.cpp:799-808: rgb8, step 3 W) */
DFB_API int dfb_wire_fused_image(const dfb_intrinsics* fused, const uint8_t* data, dfb_image_msg* out);
/* depth thresholds of Kernels::filter_depth_minmax_npp, kernels.cu:357-359 */
DFB_API void dfb_depth_thresholds(float min_depth, float max_depth, float depth_scale, uint16_t* lt, uint16_t* gt);

/* ---- fused batched path: one FusionNode::processSyncedFrames per frameset ------------------- */

/*
 * A dfb_fuser is the plan for one fusion configuration (fusion_node.cpp:718-794 with the two
 * Framesets, the two Pointclouds and the fused Frameset folded into one object). It owns the
 * packed 64-bit key ring (the z-buffer + colour of Pointcloud::project, pointcloud.cpp:136-157)
 * and the staging buffers of the *_host entry point.
 *
 * Pools (all cameras share one resolution W x H):
 *   depth  uint16 [n_framesets][n_cameras][H][W]      (mono16 LE, camera_node.cpp:191-209)
 *   color  uint8  [n_framesets][n_cameras][H][W][3]   (rgb8)
 *   out    uint8  [n_framesets][H_out][W_out][3]      (rgb8, fusion_node.cpp:799-808)
 */
typedef struct dfb_fuser dfb_fuser;

DFB_API int dfb_fuser_create(dfb_fuser** out, const dfb_fuse_desc* desc);
DFB_API int dfb_fuser_destroy(dfb_fuser* f);
/* replace the transforms (registration publishes a new right->left transform at 0.5 Hz,
 * fusion_node.cpp:610-613); intrinsics / resolution must not change */
DFB_API int dfb_fuser_update(dfb_fuser* f, const dfb_fuse_desc* desc);
/* framesets per launch group (0 = automatic: the key ring of a group is sized to stay L2-resident) */
DFB_API int dfb_fuser_set_group(dfb_fuser* f, int framesets_per_group);
DFB_API int dfb_fuser_get_group(const dfb_fuser* f);
/* kernel variant: 0 = automatic, 1 = row traversal (direct loads), 2 = column traversal (direct loads),
 * 3 = column traversal over TMA tiles in shared memory, 4 = the whole batch (every splat and every resolve)
 * in one persistent launch over a ring of L2-resident key slots; all produce identical keys */
DFB_API int dfb_fuser_set_variant(dfb_fuser* f, int variant);

/* Frameset::filterColor (frameset.cpp:194-208), which fusion_node applies to every fused frame
 * (fusion_node.cpp:789), folded into the batch: 0 = none (default), 1 = 3x3 gauss, 2 = 3x3 median —
 * NPP-exact like dfb_op_filter_color. Applies to dfb_fuser_fuse_device / _fuse_host. */
DFB_API int dfb_fuser_set_post_filter(dfb_fuser* f, int mode);

/* Fuse n_framesets framesets whose pools are already in device memory; asynchronous on `stream`.
 * Replaces setDepthFrame/setColorFrame .. Pointcloud::project for each frameset
 * (fusion_node.cpp:718-784) minus the uploads. */
/* (A fuser owns one key ring: calls on the same fuser are serialised on the device, also when they are given
 * different streams — each call waits for the previous one. Use one fuser per concurrent stream.) */
DFB_API int dfb_fuser_fuse_device(dfb_fuser* f, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                                  int n_framesets, void* stream);

/* dfb_fuser_fuse_device with CUDA events around every kernel launch (adds event overhead: use it to
 * attribute time, not to measure throughput). Synchronises `stream`. ms[0] = sum of splat kernel
 * times, ms[1] = sum of resolve kernel times, launches[0..1] = number of launches of each. */
DFB_API int dfb_fuser_profile_device(dfb_fuser* f, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                                     int n_framesets, void* stream, float ms[2], int launches[2]);

/* Same with HOST pools: uploads (Frameset::set*Frame), fuses, downloads
 * (Frameset::copyColorToHost, fusion_node.cpp:794), pipelined over internal streams; returns
 * when h_out is complete. Pinned host memory (dfb_host_alloc) is strongly recommended. */
DFB_API int dfb_fuser_fuse_host(dfb_fuser* f, const uint16_t* h_depth, const uint8_t* h_color, uint8_t* h_out,
                                int n_framesets);
/* One frameset whose cameras live in SEPARATE device buffers (d_depth[c], d_color[c], c < n_cameras) — the layout of
 * the reference's Frameset objects (depth_frame.cpp:34, color_frame.cpp:33). This is what Pointcloud::project
 * (pointcloud.cpp:136-157) of the drop-in classes calls once it has recorded deproject / transform / append. */
DFB_API int dfb_fuser_fuse_images(dfb_fuser* f, const uint16_t* const* d_depth, const uint8_t* const* d_color, uint8_t* d_out,
                                  void* stream);

/* ---- eager operators: one per reference Kernels::* wrapper (used by the C++ shim classes) ---- */

/* scratch for dfb_op_filter_depth: a few KB of device + pinned host memory */
typedef struct dfb_workspace dfb_workspace;
DFB_API int dfb_workspace_create(dfb_workspace** out);
DFB_API int dfb_workspace_destroy(dfb_workspace* ws);

/* DepthFrame::filter, depth_frame.cpp:153-182 = Kernels::filter_depth_minmax_npp (kernels.cu:353-369)
 * + filter_create_mask_npp (:371-395) + filter_mask_count_npp (:501-540) + create_img_indices
 * (:542-547) in ONE pass + an ordered compaction. d_depth is thresholded in place, d_mask (255/0)
 * is written inside the ROI only, d_indices receives the ascending valid pixel indices and
 * *mask_count the number of mask>0 pixels. Synchronises `stream` (the reference does too). */
DFB_API int dfb_op_filter_depth(dfb_workspace* ws, uint16_t* d_depth, uint8_t* d_mask, uint32_t* d_indices, int width,
                                int height, float min_depth, float max_depth, float depth_scale, const int32_t roi[4],
                                uint32_t* mask_count, void* stream);
/* The same in two steps, for callers that filter several cameras before they need a count: _begin queues the kernels and
 * the read-back on `stream` and returns, _count waits for that read-back only. (Frameset::filterDepth uses this: the
 * left and the right camera of fusion_node.cpp:739-741 overlap, the wait moves into getMaskCount().) */
DFB_API int dfb_op_filter_depth_begin(dfb_workspace* ws, uint16_t* d_depth, uint8_t* d_mask, uint32_t* d_indices, int width,
                                      int height, float min_depth, float max_depth, float depth_scale, const int32_t roi[4],
                                      void* stream);
DFB_API int dfb_op_filter_depth_count(dfb_workspace* ws, uint32_t* mask_count);
/* d_indices == NULL in dfb_op_filter_depth_begin skips the index list (depth, mask and count are complete); this builds
 * it afterwards from the mask (create_img_indices, kernels.cu:542-547) for whoever needs it after all. */
DFB_API int dfb_op_filter_depth_indices(dfb_workspace* ws, const uint8_t* d_mask, uint32_t* d_indices, int width, int height,
                                        void* stream);

/* Kernels::deproject_depth_color (kernels.cu:549-558) when d_color != NULL, else
 * Kernels::deproject_depth (:560-568). d_points = float4[count]. (The reference's floor() grid bug,
 * SURVEY App. A.6 i, is not replicated: all `count` points are written.) */
DFB_API int dfb_op_deproject(float* d_points, const uint16_t* d_depth, const uint8_t* d_color,
                             const dfb_intrinsics* intr, const uint32_t* d_indices, uint32_t count, float depth_scale,
                             void* stream);

/* Kernels::transform, kernels.cu:570-575 (T = row-major 4x4 on the HOST; passed by value to the kernel) */
DFB_API int dfb_op_transform(float* d_points, uint32_t count, const float T[16], void* stream);

/* Kernels::project_pointcloud, kernels.cu:577-592, deterministic: d_keys = uint64[W*H] scratch
 * (replaces the float z-buffer; must be all-ones on entry, is all-ones again on return),
 * segments = camera of origin for the tie-break (point i of segment s gets camera id s). */
DFB_API int dfb_op_project(uint8_t* d_color_out, uint64_t* d_keys, const float* d_points, const uint32_t* seg_counts,
                           int n_segments, const dfb_intrinsics* out_intr, int mirror, void* stream);
/* fill a key buffer with all-ones (initial state required by dfb_op_project) */
DFB_API int dfb_op_keys_clear(uint64_t* d_keys, size_t n_pixels, void* stream);

/* Kernels::filter_color_gauss_npp (kernels.cu:632-653, median=0) / filter_color_median_npp
 * (:594-630, median=1): 3x3 on the interior (W-2)x(H-2); d_scratch = W*H*3 bytes. */
DFB_API int dfb_op_filter_color(uint8_t* d_color, uint8_t* d_scratch, int width, int height, int median, void* stream);

/* Frameset::alignDepthToColor (frameset.cpp:180-188 -> Kernels::align_depth_to_color, kernels.cu:714-747,
 * 276-311): unaligned depth image (depth intrinsics) -> depth image in the colour camera's frame,
 * d_aligned = uint16[colour W*H] (even pixel count). Intrinsics / extrinsics are HOST structs.
 * Bit-identical to the reference kernels, including the pixel-pair quirk of its 32-bit atomicMin. */
DFB_API int dfb_op_align_depth_to_color(uint16_t* d_aligned, const uint16_t* d_depth, const dfb_intrinsics* depth_intr,
                                        const dfb_intrinsics* color_intr, const dfb_extrinsics* depth_to_color,
                                        float depth_scale, void* stream);

/* ---- registration feed (registration_node.cpp:313-333, registration.cpp:144-150) ------------ */

/* filterDepth + deproject (no colour) of one depth image in one ordered pass:
 * d_points = float4[W*H] capacity, {X,Y,Z,1.0f}, ascending source pixel order (the layout
 * pcl::PointXYZ expects, registration_node.cpp:328-333). d_depth is not modified. */
DFB_API int dfb_registration_cloud(dfb_workspace* ws, const dfb_camera_desc* cam, const uint16_t* d_depth,
                                   float* d_points, uint32_t* count, void* stream);

/* GPU voxel grid replacing Registration::voxelFilter (registration.cpp:144-150): one centroid per
 * occupied voxel of edge `leaf`, output sorted by voxel (iz,iy,ix). d_out = float4[n] capacity. */
DFB_API int dfb_voxel_downsample(dfb_workspace* ws, const float* d_points, uint32_t n, float leaf, float* d_out,
                                 uint32_t* n_out, void* stream);
/* The same filter with PCL's OWN semantics: pcl::ApproximateVoxelGrid<PointXYZ> (PCL 1.12 approximate_voxel_grid.hpp) —
 * 512-entry hash history, float centroids accumulated in visiting order, runs emitted when a colliding voxel evicts
 * them and the open entries in table order at the end: the cloud fast_gicp receives from the reference
 * (registration.cpp:144-150), point for point in the same order, possibly with several centroids per voxel.
 * d_out = float4[n] capacity (w = 1.0f). */
DFB_API int dfb_voxel_downsample_approximate(dfb_workspace* ws, const float* d_points, uint32_t n, float leaf, float* d_out,
                                             uint32_t* n_out, void* stream);

/* ---- synthetic inputs (dfb_synth.h) generated on the device --------------------------------- */

/* scenes[n_framesets], cams[n_cameras] are HOST arrays; fills the pools described above */
DFB_API int dfb_synth_generate(const dfs_scene* scenes, int n_framesets, const dfs_camera* cams, int n_cameras,
                               uint16_t* d_depth, uint8_t* d_color, void* stream);

/* ---- self tests / microbenchmarks (device) --------------------------------------------------- */

/* compares the shared-reciprocal division used by the projection with __fdiv_rn on n pseudo-random
 * operand triples drawn from the projection's domain; returns the number of mismatching bit patterns */
DFB_API int64_t dfb_selftest_division(uint64_t n, uint64_t seed);

/* L2 atomic throughput probe: n_ops 64-bit atomicMin into a buffer of n_slots keys.
 * pattern 0 = contiguous (lane i -> slot base+i), 1 = strided by `stride` slots per lane,
 * 2 = pseudo-random. Returns milliseconds for `iters` launches (CUDA events), <0 on error. */
DFB_API float dfb_probe_atomics(uint64_t n_slots, uint64_t n_ops, int pattern, uint32_t stride, int iters);

/* Self-test hook for the prefix sum under the registration feed's sorts (csrc/scan.cu): in-place exclusive scan of
 * n u32 words at d_data on `stream`, *total = their sum (mod 2^32). Synchronous. */
DFB_API int dfb_selftest_exclusive_scan(uint32_t* d_data, uint32_t n, uint32_t* total, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFB_H */
