/*
 * pointcloud_processing_shim.cpp — the reference's C++ class surface (Frameset, DepthFrame, ColorFrame,
 * Pointcloud, Kernels::init_cuda, checkError*) implemented over the C ABI of this library
 * (include/dfb/dfb.h). Host code only: memory management + calls into dfb_op_*; every kernel lives behind
 * the C ABI. Each method cites the reference method it stands in for.
 *
 * Error convention of the reference (error_check.cpp:10-51): print and exit(EXIT_FAILURE).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>

#include "dfb/dfb.h"
#include "pointcloud_processing/pointcloud.h"

/* ---- error_check.cpp ------------------------------------------------------------------------------ */

void checkError(const char* msg, bool sync) {
	if (sync) cudaDeviceSynchronize();
	checkError(cudaGetLastError(), msg, false);
}

void checkError(cudaError_t error, const char* msg, bool sync) {
	if (sync) cudaDeviceSynchronize();
	if (error != cudaSuccess) {
		std::cerr << "Cuda error: " << msg << " failed: " << cudaGetErrorString(error) << std::endl;
		std::exit(EXIT_FAILURE);
	}
}

void checkNppError(int status, const char* msg, bool sync) {
	if (sync) cudaDeviceSynchronize();
	if (status != 0) {
		std::cerr << "Cuda error: " << msg << " failed: " << status << std::endl;
		std::exit(EXIT_FAILURE);
	}
}

void checkDfbError(int status, const char* msg) {
	if (status != DFB_OK) {
		std::cerr << "Cuda error: " << msg << " failed: " << dfb_last_error() << std::endl;
		std::exit(EXIT_FAILURE);
	}
}

/* ---- Kernels::init_cuda (kernels.cu:325-351) ------------------------------------------------------- */

namespace Kernels {
bool cuda_initialized = false;

void init_cuda(int device, bool verbose) {
	const int rc = dfb_init(device, verbose ? 1 : 0);
	if (rc != DFB_OK) {
		std::cout << "Error initializing CUDA device #" << device << ": " << dfb_last_error() << std::endl;
		std::exit(1);
	}
	cuda_initialized = true;
	if (verbose) std::cout << "| Threads per block:  128 (4 compute warps per persistent CTA)" << std::endl;
}

void init_cuda(bool verbose) { init_cuda(0, verbose); }
}  // namespace Kernels

/* ---- ColorFrame (color_frame.cpp) ------------------------------------------------------------------ */

ColorFrame::ColorFrame() {}
ColorFrame::~ColorFrame() { free(); }

void ColorFrame::allocate(const Intrinsics intr, cudaStream_t& s) { /* color_frame.cpp:26-44 */
	free();
	stream = s;
	intrinsics = intr;
	pixel_count = static_cast<unsigned>(intr.width * intr.height);
	const size_t bytes = static_cast<size_t>(pixel_count) * 3;
	checkError(cudaMalloc(&dev_color, bytes), "ColorFrame cudaMalloc dev_color");
	checkError(cudaMalloc(&dev_color_buffer, bytes), "ColorFrame cudaMalloc dev_color_buffer");
	checkError(cudaMalloc(&dev_intrinsics, sizeof(Intrinsics)), "ColorFrame cudaMalloc dev_intrinsics");
	checkError(cudaMemcpyAsync(dev_intrinsics, &intrinsics, sizeof(Intrinsics), cudaMemcpyHostToDevice, stream),
	           "ColorFrame cudaMemcpy intrinsics to dev_intrinsics");
	is_allocated = true;
}

void ColorFrame::free() {
	cudaFree(dev_color);
	cudaFree(dev_color_buffer);
	cudaFree(dev_intrinsics);
	dev_color = dev_color_buffer = nullptr;
	dev_intrinsics = nullptr;
	is_allocated = false;
}

void ColorFrame::copyToDevice(const uint8_t* frame) { /* color_frame.cpp:58-69 */
	checkError(cudaMemcpyAsync(dev_color, frame, static_cast<size_t>(pixel_count) * 3, cudaMemcpyHostToDevice, stream),
	           "ColorFrame cudaMemcpy frame to dev_color");
}

void ColorFrame::copyToHost(uint8_t* host_frame) { /* color_frame.cpp:75-87 */
	checkError(cudaMemcpyAsync(host_frame, dev_color, static_cast<size_t>(pixel_count) * 3, cudaMemcpyDeviceToHost, stream),
	           "ColorFrame cudaMemcpy dev_color to host_frame");
	cudaStreamSynchronize(stream);
}

void ColorFrame::setSaveImages(bool save, std::string path, std::string prefix) {
	save_images = save;
	if (!path.empty()) filepath = path;
	if (!prefix.empty()) fileprefix = prefix;
}

/* ---- DepthFrame (depth_frame.cpp) ------------------------------------------------------------------ */

DepthFrame::DepthFrame() {}

DepthFrame::~DepthFrame() {
	free();
	cudaFree(dev_depth_unaligned);
	cudaFree(dev_intrinsics_unaligned);
	cudaFree(dev_extrinsics);
	if (workspace) dfb_workspace_destroy(workspace);
}

void DepthFrame::allocate(const Intrinsics intr, cudaStream_t& s) { /* depth_frame.cpp:23-52 */
	free();
	stream = s;
	intrinsics = intr;
	pixel_count = static_cast<unsigned>(intr.width * intr.height);
	checkError(cudaMalloc(&dev_depth, static_cast<size_t>(pixel_count) * sizeof(uint16_t)), "DepthFrame cudaMalloc dev_depth");
	checkError(cudaMalloc(&dev_intrinsics, sizeof(Intrinsics)), "DepthFrame cudaMalloc dev_intrinsics");
	checkError(cudaMemcpyAsync(dev_intrinsics, &intrinsics, sizeof(Intrinsics), cudaMemcpyHostToDevice, stream),
	           "DepthFrame cudaMemcpy intrinsics to dev_intrinsics");
	checkError(cudaMalloc(&dev_indices, static_cast<size_t>(pixel_count) * sizeof(uint)), "DepthFrame cudaMalloc dev_indices");
	checkError(cudaMalloc(&dev_mask, static_cast<size_t>(pixel_count)), "DepthFrame cudaMalloc dev_mask");
	/* the mask is cleared here and never again: pixels outside a later ROI keep their old state (depth_frame.cpp:46) */
	checkError(cudaMemsetAsync(dev_mask, 0, static_cast<size_t>(pixel_count), stream), "DepthFrame cudaMemset dev_mask");
	if (!workspace) checkDfbError(dfb_workspace_create(&workspace), "DepthFrame dfb_workspace_create");
	is_allocated = true;
}

void DepthFrame::allocateUnaligned(const Intrinsics intr, const Extrinsics extr, cudaStream_t& s) { /* depth_frame.cpp:60-90 */
	stream = s;
	intrinsics_unaligned = intr;
	extrinsics = extr;
	cudaFree(dev_depth_unaligned);
	cudaFree(dev_intrinsics_unaligned);
	cudaFree(dev_extrinsics);
	pixel_count_unaligned = static_cast<unsigned>(intr.width * intr.height);
	checkError(cudaMalloc(&dev_depth_unaligned, static_cast<size_t>(pixel_count_unaligned) * sizeof(uint16_t)),
	           "DepthFrame cudaMalloc dev_depth_unaligned");
	checkError(cudaMalloc(&dev_intrinsics_unaligned, sizeof(Intrinsics)), "DepthFrame cudaMalloc dev_intrinsics_unaligned");
	checkError(cudaMemcpyAsync(dev_intrinsics_unaligned, &intrinsics_unaligned, sizeof(Intrinsics), cudaMemcpyHostToDevice, stream),
	           "DepthFrame cudaMemcpy intrinsics_unaligned");
	checkError(cudaMalloc(&dev_extrinsics, sizeof(Extrinsics)), "DepthFrame cudaMalloc dev_extrinsics");
	checkError(cudaMemcpyAsync(dev_extrinsics, &extrinsics, sizeof(Extrinsics), cudaMemcpyHostToDevice, stream),
	           "DepthFrame cudaMemcpy extrinsics");
}

void DepthFrame::free() { /* unlike depth_frame.cpp:95-102 this also releases the mask and the index list */
	cudaFree(dev_depth);
	cudaFree(dev_intrinsics);
	cudaFree(dev_mask);
	cudaFree(dev_indices);
	dev_depth = nullptr;
	dev_intrinsics = nullptr;
	dev_mask = nullptr;
	dev_indices = nullptr;
	mask_count = 0;
	is_allocated = false;
}

void DepthFrame::copyToDevice(const uint16_t* frame) { /* depth_frame.cpp:108-120 */
	checkError(cudaMemcpyAsync(dev_depth, frame, static_cast<size_t>(pixel_count) * sizeof(uint16_t), cudaMemcpyHostToDevice, stream),
	           "DepthFrame cudaMemcpy frame to dev_depth");
}

void DepthFrame::copyToHost(uint16_t* host_frame) { /* depth_frame.cpp:126-132 */
	checkError(cudaMemcpyAsync(host_frame, dev_depth, static_cast<size_t>(pixel_count) * sizeof(uint16_t), cudaMemcpyDeviceToHost, stream),
	           "DepthFrame cudaMemcpy dev_depth to host_frame");
	cudaStreamSynchronize(stream);
}

void DepthFrame::copyUnalignedToDevice(const uint16_t* frame) { /* depth_frame.cpp:138-145 */
	checkError(cudaMemcpyAsync(dev_depth_unaligned, frame, static_cast<size_t>(pixel_count_unaligned) * sizeof(uint16_t),
	                           cudaMemcpyHostToDevice, stream),
	           "DepthFrame cudaMemcpy frame_unaligned to dev_depth_unaligned");
}

void DepthFrame::filter(const float min_depth, const float max_depth, const float depth_scale, const int roi[4]) {
	/* depth_frame.cpp:153-182: minmax threshold, mask, count (with the reference's stream sync) and index list */
	uint32_t count = 0;
	const int32_t r[4] = {roi[0], roi[1], roi[2], roi[3]};
	checkDfbError(dfb_op_filter_depth(workspace, dev_depth, dev_mask, reinterpret_cast<uint32_t*>(dev_indices), intrinsics.width,
	                                  intrinsics.height, min_depth, max_depth, depth_scale, r, &count, stream),
	              "DepthFrame filter");
	mask_count = count;
	if (mask_count == 0) std::cout << "Warning: empty mask!" << std::endl; /* kernels.cu:533-535 */
}

void DepthFrame::setSaveImages(bool save, std::string path, std::string prefix) {
	save_images = save;
	if (!path.empty()) filepath = path;
	fileprefix = prefix;
}

/* ---- Frameset (frameset.cpp) ------------------------------------------------------------------------ */

Frameset::Frameset() {
	cudaEventCreate(&start);
	cudaEventCreate(&stop);
}

Frameset::~Frameset() {
	if (start) cudaEventDestroy(start);
	if (stop) cudaEventDestroy(stop);
}

void Frameset::setStream(cudaStream_t* s) { stream = s ? *s : nullptr; } /* frameset.cpp:27-32, no NPP context needed */
void Frameset::allocateDepthFrame(const Intrinsics& intr) { depth_frame.allocate(intr, stream); }
void Frameset::allocateDepthFrameUnaligned(const Intrinsics& intr, const Extrinsics& extr) { depth_frame.allocateUnaligned(intr, extr, stream); }
void Frameset::allocateColorFrame(const Intrinsics& intr) { color_frame.allocate(intr, stream); }
void Frameset::setDepthFrame(const uint16_t* f) { depth_frame.copyToDevice(f); }
void Frameset::setDepthFrameUnaligned(const uint16_t* f) { depth_frame.copyUnalignedToDevice(f); }
void Frameset::setColorFrame(const uint8_t* f) { color_frame.copyToDevice(f); }
uint16_t* Frameset::getDepth() { return depth_frame.get(); }
uint16_t* Frameset::getDepthUnaligned() { return depth_frame.getUnaligned(); }
uint8_t* Frameset::getColor() { return color_frame.get(); }
uint8_t* Frameset::getMask() { return depth_frame.getMask(); }

void Frameset::filterDepth(const float min_depth, const float max_depth, const float depth_scale, const std::array<int, 4> roi) {
	timerStart();
	depth_frame.filter(min_depth, max_depth, depth_scale, roi.data()); /* frameset.cpp:125-131 */
	timerEnd("filter");
}

void Frameset::deprojectToPointcloud(float* dev_points, const float depth_scale) { /* frameset.cpp:140-156 */
	timerStart();
	const Intrinsics intr = depth_frame.getIntrinsics();
	dfb_intrinsics ci;
	std::memcpy(&ci, &intr, sizeof(ci));
	checkDfbError(dfb_op_deproject(dev_points, depth_frame.get(), color_frame.exists() ? color_frame.get() : nullptr, &ci,
	                               reinterpret_cast<const uint32_t*>(depth_frame.getIndices()), depth_frame.getMaskCount(), depth_scale,
	                               stream),
	              "Frameset deprojectToPointcloud");
	timerEnd("deproject");
}

void Frameset::copyColorToHost(uint8_t* host) { color_frame.copyToHost(host); }
void Frameset::copyDepthToHost(uint16_t* host) { depth_frame.copyToHost(host); }

void Frameset::alignDepthToColor(const float depth_scale) { /* frameset.cpp:180-188 */
	static_assert(sizeof(Extrinsics) == sizeof(dfb_extrinsics), "Extrinsics must stay layout-compatible with rs2_extrinsics");
	const Intrinsics depth_intr = depth_frame.getIntrinsicsUnaligned();
	const Intrinsics color_intr = color_frame.getIntrinsics();
	const Extrinsics extr = depth_frame.getExtrinsics();
	checkDfbError(dfb_op_align_depth_to_color(depth_frame.get(), depth_frame.getUnaligned(),
	                                          reinterpret_cast<const dfb_intrinsics*>(&depth_intr),
	                                          reinterpret_cast<const dfb_intrinsics*>(&color_intr),
	                                          reinterpret_cast<const dfb_extrinsics*>(&extr), depth_scale, stream),
	              "Frameset alignDepthToColor");
}

void Frameset::filterColor(bool median) { /* frameset.cpp:194-208 */
	const Intrinsics intr = color_frame.getIntrinsics();
	checkDfbError(dfb_op_filter_color(color_frame.get(), color_frame.getBuffer(), intr.width, intr.height, median ? 1 : 0, stream),
	              "Frameset filterColor");
}

void Frameset::timerStart() {
	if (profiling) cudaEventRecord(start, stream);
}

void Frameset::timerEnd(std::string msg) { /* frameset.cpp:222-237: "| <msg>:   <ms> ms" */
	if (!profiling) return;
	cudaEventRecord(stop, stream);
	cudaEventSynchronize(stop);
	float ms = 0.f;
	cudaEventElapsedTime(&ms, start, stop);
	msg = "| " + msg + ":";
	if (msg.size() < 32) msg.append(32 - msg.size(), ' ');
	std::cout << msg << std::setprecision(4) << ms << " ms" << std::endl;
}

/* ---- Pointcloud (pointcloud.cpp) ---------------------------------------------------------------------- */

Pointcloud::Pointcloud() {
	for (int i = 0; i < 16; ++i) transform_rm[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

Pointcloud::~Pointcloud() { free(); }

void Pointcloud::allocate(unsigned count) { /* pointcloud.cpp:21-35 */
	point_count = count;
	cudaFree(dev_points);
	dev_points = nullptr;
	checkError(cudaMalloc(&dev_points, static_cast<size_t>(count ? count : 1) * 4 * sizeof(float)), "Pointcloud cudaMalloc dev_points");
	/* the reference leaves the cloud uninitialised; zeros (Z = 0 -> culled by project) make a partly filled cloud deterministic */
	checkError(cudaMemsetAsync(dev_points, 0, static_cast<size_t>(count ? count : 1) * 4 * sizeof(float), stream),
	           "Pointcloud cudaMemset dev_points");
	pointer_offset = 0;
	segments.clear();
}

void Pointcloud::free() { /* pointcloud.cpp:40-45 */
	cudaFree(dev_points);
	cudaFree(dev_keys);
	dev_points = nullptr;
	dev_keys = nullptr;
	key_pixels = 0;
	point_count = 0;
	pointer_offset = 0;
	segments.clear();
}

void Pointcloud::set(const float* host_points) { /* pointcloud.cpp:51-57 */
	checkError(cudaMemcpyAsync(dev_points, host_points, static_cast<size_t>(point_count) * 4 * sizeof(float), cudaMemcpyHostToDevice, stream),
	           "Pointcloud cudaMemcpy host_points to dev_points");
	pointer_offset = point_count;
	segments.assign(1, point_count);
}

void Pointcloud::setTransform(const double t[16]) { /* pointcloud.cpp:63-75: Affine3d -> cast<float>() -> row-major */
	dfb_affine_to_float16(t, transform_rm);
}

void Pointcloud::setTransform(const float t[16]) { std::memcpy(transform_rm, t, sizeof(transform_rm)); }

void Pointcloud::transform() { /* pointcloud.cpp:80-83: the whole allocated cloud */
	checkDfbError(dfb_op_transform(dev_points, point_count, transform_rm, stream), "Pointcloud transform");
}

void Pointcloud::transform(const double t[16]) {
	setTransform(t);
	transform();
}

void Pointcloud::transform(const float t[16]) {
	setTransform(t);
	transform();
}

void Pointcloud::deproject(Frameset& frameset, float depth_scale) { /* pointcloud.cpp:100-111 */
	const unsigned n = frameset.getMaskCount();
	if (point_count == 0) allocate(n);
	if (n > point_count) {
		std::cout << "error: deprojecting more points than allocated" << std::endl;
		return;
	}
	frameset.sync();
	/* like the reference, deproject always writes at the START of the cloud (frameset.cpp:144,150 get dev_points) */
	frameset.deprojectToPointcloud(dev_points, depth_scale);
	if (*frameset.getStream() != stream) frameset.sync(); /* the cloud's own stream consumes the points next */
	pointer_offset += n;
	if (segments.empty())
		segments.push_back(n);
	else
		segments[0] = n;
}

void Pointcloud::append(Pointcloud& cloud) { /* pointcloud.cpp:117-129 */
	const unsigned n = cloud.getPointCount();
	if (pointer_offset >= point_count || pointer_offset + n > point_count) {
		std::cout << "error: appending more points to pointcloud than allocated" << std::endl;
		return;
	}
	checkError(cudaMemcpyAsync(dev_points + static_cast<size_t>(pointer_offset) * 4, cloud.get(), static_cast<size_t>(n) * 4 * sizeof(float),
	                           cudaMemcpyDeviceToDevice, stream),
	           "Pointcloud cudaMemcpy cloud.get() to dev_points + pointer_offset");
	pointer_offset += n;
	segments.push_back(n);
}

void Pointcloud::project(Frameset& frameset, bool mirror_image) { /* pointcloud.cpp:136-157 */
	const Intrinsics intr = frameset.getColorIntrinsics();
	const size_t pixels = static_cast<size_t>(intr.width) * intr.height;
	cudaStream_t fs = *frameset.getStream();
	if (key_pixels != pixels) {
		cudaFree(dev_keys);
		dev_keys = nullptr;
		checkError(cudaMalloc(&dev_keys, pixels * sizeof(uint64_t)), "Pointcloud cudaMalloc dev_keys");
		checkDfbError(dfb_op_keys_clear(dev_keys, pixels, fs), "Pointcloud keys clear");
		key_pixels = pixels;
	}
	/* camera of origin per point = index of the deproject()/append() call that produced it; anything allocated
	 * beyond the filled part belongs to the last segment (zeros: culled) */
	std::vector<uint32_t> seg(segments);
	if (seg.empty()) seg.push_back(0);
	uint64_t filled = 0;
	for (uint32_t c : seg) filled += c;
	if (filled < point_count) seg.back() += static_cast<uint32_t>(point_count - filled);
	if (seg.size() > DFB_MAX_CAMERAS) { /* more appends than tie-break ids: merge the tail */
		for (size_t i = DFB_MAX_CAMERAS; i < seg.size(); ++i) seg[DFB_MAX_CAMERAS - 1] += seg[i];
		seg.resize(DFB_MAX_CAMERAS);
	}
	if (fs != stream) cudaStreamSynchronize(stream); /* the reference launches project on the FRAMESET's stream too */
	dfb_intrinsics ci;
	std::memcpy(&ci, &intr, sizeof(ci));
	checkDfbError(dfb_op_project(frameset.getColor(), dev_keys, dev_points, seg.data(), static_cast<int>(seg.size()), &ci,
	                             mirror_image ? 1 : 0, fs),
	              "Pointcloud project");
}

void Pointcloud::copyToHost(float* host_points) { /* pointcloud.cpp:163-169 */
	checkError(cudaMemcpyAsync(host_points, dev_points, static_cast<size_t>(point_count) * 4 * sizeof(float), cudaMemcpyDeviceToHost, stream),
	           "cudaMemcpy cam.dev_points to cloud_ptr");
	cudaStreamSynchronize(stream);
}
