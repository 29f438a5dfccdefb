#pragma once
/*
 * color_frame.h — device RGB8 image + scratch buffer + intrinsics
 * (reference: pointcloud_processing/include/pointcloud_processing/color_frame.h:18-95, src/color_frame.cpp).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "error_check.h"
#include "intrinsics.h"

class DFB_API ColorFrame {
public:
	ColorFrame();
	~ColorFrame();
	ColorFrame(const ColorFrame&) = delete;
	ColorFrame& operator=(const ColorFrame&) = delete;

	void allocate(const Intrinsics intrinsics, cudaStream_t& stream);
	void free();
	bool exists() { return is_allocated; }
	uint8_t* get() { return dev_color; }
	uint8_t* getBuffer() { return dev_color_buffer; }
	Intrinsics getIntrinsics() { return intrinsics; }
	Intrinsics* getDevIntrinsics() { return dev_intrinsics; }
	unsigned getPixelCount() { return pixel_count; }
	void copyToDevice(const uint8_t* frame);
	void copyToHost(uint8_t* host_frame);
	/* The reference dumps PNGs through PCL when enabled (debug only, color_frame.cpp:60-64,82-86). PCL is
	 * not a dependency here: the flag is stored, nothing is written. */
	void setSaveImages(bool save_images, std::string filepath = "", std::string fileprefix = "");

private:
	cudaStream_t stream = nullptr;
	bool is_allocated = false;
	bool save_images = false;
	std::string filepath, fileprefix;
	unsigned pixel_count = 0;
	Intrinsics intrinsics{};
	Intrinsics* dev_intrinsics = nullptr;
	uint8_t* dev_color = nullptr;
	uint8_t* dev_color_buffer = nullptr;
};
