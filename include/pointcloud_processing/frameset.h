#pragma once
/*
 * frameset.h — a depth + colour frame pair on one CUDA stream
 * (reference: pointcloud_processing/include/pointcloud_processing/frameset.h:20-145, src/frameset.cpp).
 * Same public surface as the reference minus NPP (getNppContext is gone: no NPP underneath).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <array>
#include <string>
#include <vector>

#include "color_frame.h"
#include "depth_frame.h"
#include "error_check.h"
#include "intrinsics.h"

class DFB_API Frameset {
public:
	Frameset();
	~Frameset();
	Frameset(const Frameset&) = delete;
	Frameset& operator=(const Frameset&) = delete;

	void setProfiling(bool profiling) { this->profiling = profiling; }
	cudaStream_t* getStream() { return &stream; }
	void setStream(cudaStream_t* stream = nullptr);
	void sync() { cudaStreamSynchronize(stream); }
	void allocateDepthFrame(const Intrinsics& intrinsics);
	void allocateDepthFrameUnaligned(const Intrinsics& intrinsics_unaligned, const Extrinsics& extrinsics);
	void allocateColorFrame(const Intrinsics& intrinsics);
	void setDepthFrame(const uint16_t* depth_frame);
	void setDepthFrameUnaligned(const uint16_t* depth_frame_unaligned);
	void setColorFrame(const uint8_t* color_frame);
	uint16_t* getDepth();
	uint16_t* getDepthUnaligned();
	uint8_t* getColor();
	uint8_t* getMask();
	Intrinsics getColorIntrinsics() { return color_frame.getIntrinsics(); }
	Intrinsics getDepthIntrinsics() { return depth_frame.getIntrinsics(); }
	Intrinsics* getColorDevIntrinsics() { return color_frame.getDevIntrinsics(); }
	Intrinsics* getDepthDevIntrinsics() { return depth_frame.getDevIntrinsics(); }
	unsigned getMaskCount() { return depth_frame.getMaskCount(); }
	void setSaveImages(bool save_images, std::string filepath, std::string fileprefix = "") {
		color_frame.setSaveImages(save_images, filepath, fileprefix);
		depth_frame.setSaveImages(save_images, filepath, fileprefix);
	}
	void filterDepth(const float min_depth, const float max_depth, const float depth_scale,
	                 const std::array<int, 4> roi = {-1, -1, -1, -1});
	void deprojectToPointcloud(float* dev_points, const float depth_scale);
	void copyColorToHost(uint8_t* host_color_frame);
	void copyDepthToHost(uint16_t* host_depth_frame);
	/* GPU depth->colour alignment of the unaligned depth frame (used when align_frames is set, fusion_node.cpp:443-457);
	 * bit-identical to the reference kernels, for every colour distortion model the reference handles (F-theta included). */
	void alignDepthToColor(const float depth_scale);
	void filterColor(bool median = false);

	/* accessors used by Pointcloud */
	DepthFrame& depth() { return depth_frame; }
	ColorFrame& color() { return color_frame; }
	/* lazy clouds that recorded this frameset as a source (see pointcloud.h); they materialise before it changes */
	void addDependent(class Pointcloud* cloud);
	void removeDependent(class Pointcloud* cloud);

private:
	bool profiling = false;
	cudaStream_t stream = nullptr;
	cudaEvent_t start = nullptr, stop = nullptr;
	DepthFrame depth_frame;
	ColorFrame color_frame;
	std::vector<class Pointcloud*> dependents;
	void sourceWillChange();
	void timerStart();
	void timerEnd(std::string msg);
};
