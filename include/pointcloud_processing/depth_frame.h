#pragma once
/*
 * depth_frame.h — device depth image, validity mask, index list and intrinsics
 * (reference: pointcloud_processing/include/pointcloud_processing/depth_frame.h:21-186, src/depth_frame.cpp).
 * filter() = threshold -> mask -> count -> ordered index list in one C-ABI call (dfb_op_filter_depth).
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <sys/types.h>

#include <string>

#include "error_check.h"
#include "intrinsics.h"

struct dfb_workspace;

class DFB_API DepthFrame {
public:
	DepthFrame();
	~DepthFrame();
	DepthFrame(const DepthFrame&) = delete;
	DepthFrame& operator=(const DepthFrame&) = delete;

	void allocate(const Intrinsics intrinsics, cudaStream_t& stream);
	/* Unaligned depth for Frameset::alignDepthToColor (only used when align_frames=true; the camera node aligns by
	 * default). fusion_node allocates these unconditionally (fusion_node.cpp:548,551). */
	void allocateUnaligned(const Intrinsics intrinsics_unaligned, const Extrinsics extrinsics, cudaStream_t& stream);
	void free();
	bool exists() { return is_allocated; }
	uint16_t* get() { return dev_depth; }
	uint16_t* getUnaligned() { return dev_depth_unaligned; }
	uint8_t* getMask() { return dev_mask; }
	Intrinsics getIntrinsics() { return intrinsics; }
	Intrinsics* getDevIntrinsics() { return dev_intrinsics; }
	Intrinsics getIntrinsicsUnaligned() { return intrinsics_unaligned; }
	Intrinsics* getDevIntrinsicsUnaligned() { return dev_intrinsics_unaligned; }
	Extrinsics getExtrinsics() { return extrinsics; }
	Extrinsics* getDevExtrinsics() { return dev_extrinsics; }
	uint* getIndices(); /* builds the index list of the last filter() if nobody has asked for it yet */
	unsigned getPixelCount() { return pixel_count; }
	unsigned getPixelCountUnaligned() { return pixel_count_unaligned; }
	unsigned getMaskCount(); /* waits for the count of the last filter() if it is still in flight */
	/* parameters of the last filter() call (the fused path applies the same thresholds itself) */
	struct FilterState {
		bool valid = false, full_roi = false;
		float min_depth = 0.f, max_depth = 0.f, depth_scale = 0.f;
	};
	const FilterState& lastFilter() const { return last_filter; }
	void copyToDevice(const uint16_t* frame);
	void copyToHost(uint16_t* host_frame);
	void copyUnalignedToDevice(const uint16_t* frame_unaligned);
	void filter(const float min_depth, const float max_depth, const float depth_scale, const int roi[4]);
	void setSaveImages(bool save_images, std::string filepath = "", std::string fileprefix = "");

private:
	cudaStream_t stream = nullptr;
	bool is_allocated = false;
	bool save_images = false;
	std::string filepath, fileprefix;
	unsigned pixel_count = 0;
	unsigned pixel_count_unaligned = 0;
	Intrinsics intrinsics{};
	Intrinsics intrinsics_unaligned{};
	Intrinsics* dev_intrinsics = nullptr;
	Intrinsics* dev_intrinsics_unaligned = nullptr;
	Extrinsics extrinsics{};
	Extrinsics* dev_extrinsics = nullptr;
	uint16_t* dev_depth = nullptr;
	uint16_t* dev_depth_unaligned = nullptr;
	unsigned mask_count = 0;
	uint8_t* dev_mask = nullptr;
	uint* dev_indices = nullptr;
	dfb_workspace* workspace = nullptr;
	FilterState last_filter;
	bool count_pending = false;
	bool indices_stale = false;
};
