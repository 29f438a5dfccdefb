/*
 * dfb_types.h — plain-old-data shared by the C ABI (dfb.h), the CPU oracle and the
 * synthetic input generator.
 *
 * dfb_intrinsics is byte-identical to the reference's `struct Intrinsics`
 * (pointcloud_processing/include/pointcloud_processing/intrinsics.h:20-30), which is itself a
 * byte copy of librealsense's rs2_intrinsics (camera_node/src/camera_node.cpp:87-92): the
 * layout is ABI, 48 bytes, ppx at offset 8, fx at offset 16.
 * dfb_extrinsics mirrors `struct Extrinsics` (intrinsics.h:35-39).
 */
#ifndef DFB_TYPES_H
#define DFB_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_MAX_CAMERAS 8

/* enum Distortion, intrinsics.h:6-15 (values only; the float4 deproject overloads and the
 * projection kernel ignore the model, kernels.cu:70-90,239-274) */
enum {
	DFB_DISTORTION_NONE = 0,
	DFB_DISTORTION_MODIFIED_BROWN_CONRADY = 1,
	DFB_DISTORTION_INVERSE_BROWN_CONRADY = 2,
	DFB_DISTORTION_FTHETA = 3,
	DFB_DISTORTION_BROWN_CONRADY = 4,
	DFB_DISTORTION_KANNALA_BRANDT4 = 5,
	DFB_DISTORTION_COUNT = 6
};

typedef struct dfb_intrinsics {
	int32_t width;
	int32_t height;
	float ppx;
	float ppy;
	float fx;
	float fy;
	int32_t model;
	float coeffs[5];
} dfb_intrinsics;

typedef struct dfb_extrinsics {
	float rotation[9];
	float translation[3];
} dfb_extrinsics;

/* One input camera of a fusion: everything FusionNode::processSyncedFrames feeds into
 * Frameset::filterDepth / Pointcloud::deproject / Pointcloud::transform for that camera
 * (fusion_node/src/fusion_node.cpp:739-757). */
typedef struct dfb_camera_desc {
	dfb_intrinsics intr;  /* intrinsics of the (colour-aligned) depth frame, fusion_node.cpp:546,549 */
	float depth_scale;    /* metres per raw unit (0.001 for a D455) */
	float min_depth;      /* metres, kernels.cu:357 */
	float max_depth;      /* metres, kernels.cu:359 */
	int32_t roi[4];       /* left, top, width, height; negative = full (kernels.cu:379-384) */
	int32_t has_transform; /* 0: cloud is not transformed before the fused transform (left camera) */
	float to_ref[16];     /* row-major 4x4 as Pointcloud::setTransform uploads it (pointcloud.cpp:63-75) */
} dfb_camera_desc;

/* One fusion configuration = the parameters of one processSyncedFrames call
 * (fusion_node.cpp:718-794), generalised to n_cameras inputs appended in camera order. */
typedef struct dfb_fuse_desc {
	int32_t n_cameras;
	int32_t mirror;              /* Pointcloud::project(frameset, mirror_image), pointcloud.cpp:136 */
	dfb_camera_desc cam[DFB_MAX_CAMERAS];
	float fused_transform[16];   /* row-major 4x4 applied to the whole fused cloud, fusion_node.cpp:779 */
	dfb_intrinsics out_intr;     /* intrinsics of the fused colour frame, fusion_node.cpp:155-165 */
} dfb_fuse_desc;

/* The fields of a sensor_msgs/Image that matter to the path: what camera_node puts into the depth_image / color_image
 * members of camera_interfaces/msg/DepthFrameset.msg:1-7 (camera_node.cpp:191-209: "mono16" little-endian, step = 2 W;
 * "rgb8", step = 3 W) and what fusion_node publishes as the fused image (fusion_node.cpp:799-808). `data` is borrowed. */
typedef struct dfb_image_msg {
	uint32_t height;
	uint32_t width;
	const char* encoding; /* "mono16" (also "16UC1") or "rgb8" */
	uint8_t is_bigendian;
	uint32_t step;        /* bytes per row, >= width * bytes per pixel */
	const uint8_t* data;
	size_t data_size;     /* >= step * height */
} dfb_image_msg;

#ifdef __cplusplus
}
#endif
#endif /* DFB_TYPES_H */
