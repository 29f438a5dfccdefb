#pragma once
/*
 * intrinsics.h — camera model PODs of the pointcloud_processing interface
 * (reference: pointcloud_processing/include/pointcloud_processing/intrinsics.h:6-39).
 * The layouts are ABI: Intrinsics is byte-copied from librealsense's rs2_intrinsics by the camera node
 * (camera_node/src/camera_node.cpp:87-92) and is the same 48 bytes as dfb_intrinsics in the C ABI.
 */
#include "dfb/dfb_types.h"

enum Distortion {
	DISTORTION_NONE = DFB_DISTORTION_NONE,
	DISTORTION_MODIFIED_BROWN_CONRADY = DFB_DISTORTION_MODIFIED_BROWN_CONRADY,
	DISTORTION_INVERSE_BROWN_CONRADY = DFB_DISTORTION_INVERSE_BROWN_CONRADY,
	DISTORTION_FTHETA = DFB_DISTORTION_FTHETA,
	DISTORTION_BROWN_CONRADY = DFB_DISTORTION_BROWN_CONRADY,
	DISTORTION_KANNALA_BRANDT4 = DFB_DISTORTION_KANNALA_BRANDT4,
	DISTORTION_COUNT = DFB_DISTORTION_COUNT
};

struct Intrinsics {
	int width;
	int height;
	float ppx;
	float ppy;
	float fx;
	float fy;
	Distortion model;
	float coeffs[5];
};

struct Extrinsics {
	float rotation[9];
	float translation[3];
};

static_assert(sizeof(Intrinsics) == sizeof(dfb_intrinsics), "Intrinsics must match the C ABI layout");
static_assert(sizeof(Extrinsics) == sizeof(dfb_extrinsics), "Extrinsics must match the C ABI layout");
