/*
 * pointcloud_processing_shim.cpp — the reference's C++ class surface (Frameset, DepthFrame, ColorFrame,
 * Pointcloud, Kernels::init_cuda, checkError*) implemented over the C ABI of this library
 * (include/dfb/dfb.h). Host code only: memory management + calls into dfb_op_*; every kernel lives behind
 * the C ABI. Each method cites the reference method it stands in for.
 *
 * Error convention of the reference (error_check.cpp:10-51): print and exit(EXIT_FAILURE).
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <mutex>

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

/* ---- device memory pool ------------------------------------------------------------------------------
 * fusion_node builds two Pointcloud objects on the stack for every frame (fusion_node.cpp:746-752); in the reference
 * each costs cudaMalloc + cudaFree (pointcloud.cpp:21-45), which serialise the whole device. Here freed blocks are
 * parked together with an event of the stream that used them last and handed to the next request of a similar
 * size, which first makes ITS stream wait for that event. Steady state: no allocation calls at all. */
namespace {
struct PooledBlock {
	void* ptr;
	size_t bytes;
	cudaEvent_t last_use;
	int device;
};
std::mutex g_pool_mutex;
std::vector<PooledBlock> g_pool;
size_t g_pool_bytes = 0;
constexpr size_t kPoolCap = 2ull << 30; /* parked bytes; beyond this blocks are really freed */

void* pool_get(size_t bytes, cudaStream_t stream, size_t* got) {
	int dev = 0;
	cudaGetDevice(&dev);
	{
		std::lock_guard<std::mutex> lk(g_pool_mutex);
		int best = -1;
		for (int i = 0; i < (int)g_pool.size(); ++i)
			if (g_pool[i].device == dev && g_pool[i].bytes >= bytes && g_pool[i].bytes <= 2 * bytes + (1u << 20) &&
			    (best < 0 || g_pool[i].bytes < g_pool[best].bytes))
				best = i;
		if (best >= 0) {
			PooledBlock b = g_pool[best];
			g_pool.erase(g_pool.begin() + best);
			g_pool_bytes -= b.bytes;
			if (b.last_use) {
				cudaStreamWaitEvent(stream, b.last_use, 0);
				cudaEventDestroy(b.last_use);
			}
			*got = b.bytes;
			return b.ptr;
		}
	}
	void* p = nullptr;
	const size_t rounded = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1); /* similar requests share blocks */
	checkError(cudaMalloc(&p, rounded), "device pool cudaMalloc");
	*got = rounded;
	return p;
}

/* last_use: an event recorded after the last operation that touched the block (may be null: never used). The pool takes
 * it over. Nothing here touches a stream: clouds may outlive the streams they were given (the reference's cudaFree does
 * not care either). */
void pool_put(void* ptr, size_t bytes, cudaEvent_t last_use) {
	if (!ptr) {
		if (last_use) cudaEventDestroy(last_use);
		return;
	}
	int dev = 0;
	cudaGetDevice(&dev);
	std::lock_guard<std::mutex> lk(g_pool_mutex);
	if (g_pool_bytes + bytes > kPoolCap) {
		if (last_use) {
			cudaEventSynchronize(last_use);
			cudaEventDestroy(last_use);
		}
		cudaFree(ptr);
		return;
	}
	g_pool.emplace_back(PooledBlock{ptr, bytes, last_use, dev});
	g_pool_bytes += bytes;
}

/* note that `stream` has just been given work on a pooled block */
void mark_use(cudaEvent_t* ev, cudaStream_t stream) {
	if (!*ev) cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
	cudaEventRecord(*ev, stream);
}

/* One fuser per (cameras, resolutions): its key ring, counters and ray tables live as long as the process. */
struct CachedFuser {
	int n_cameras, w, h, out_w, out_h, device;
	dfb_fuser* fuser;
};
std::vector<CachedFuser> g_fusers;

/* updates (or creates) the fuser of this geometry and queues ONE frameset on it, all under the cache's lock: clouds of
 * different threads may share a geometry, a fuser takes one call at a time */
int fuse_with_cached_fuser(const dfb_fuse_desc& d, const uint16_t* const* depth, const uint8_t* const* color, uint8_t* out,
                           cudaStream_t stream) {
	int dev = 0;
	cudaGetDevice(&dev);
	std::lock_guard<std::mutex> lk(g_pool_mutex);
	dfb_fuser* fuser = nullptr;
	for (auto& c : g_fusers)
		if (c.device == dev && c.n_cameras == d.n_cameras && c.w == d.cam[0].intr.width && c.h == d.cam[0].intr.height &&
		    c.out_w == d.out_intr.width && c.out_h == d.out_intr.height) {
			fuser = c.fuser;
			const int rc = dfb_fuser_update(fuser, &d);
			if (rc != DFB_OK) return rc;
			break;
		}
	if (!fuser) {
		const int rc = dfb_fuser_create(&fuser, &d);
		if (rc != DFB_OK) return rc;
		g_fusers.push_back({d.n_cameras, d.cam[0].intr.width, d.cam[0].intr.height, d.out_intr.width, d.out_intr.height, dev, fuser});
	}
	return dfb_fuser_fuse_images(fuser, depth, color, out, stream);
}
}  // namespace

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
	if (verbose) std::cout << "| Threads per block:  576 (16 compute warps + producer + retirer per persistent CTA)" << std::endl;
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
	const int32_t r[4] = {roi[0], roi[1], roi[2], roi[3]};
	/* depth, mask and count now; the index list (two more kernels) only if somebody reads it — the lazy Pointcloud fuses
	 * straight from the depth frame */
	checkDfbError(dfb_op_filter_depth_begin(workspace, dev_depth, dev_mask, nullptr, intrinsics.width, intrinsics.height, min_depth,
	                                        max_depth, depth_scale, r, stream),
	              "DepthFrame filter");
	indices_stale = true;
	/* the reference waits for the count here (kernels.cu:529); we wait where it is first needed — getMaskCount() — so that
	 * the filters of several cameras (fusion_node.cpp:739-741) overlap on their streams */
	count_pending = true;
	/* ROI as kernels.cu:379-384 resolves it: the fused path may stand in for this call only when it covers the frame
	 * (a partial ROI leaves the mask outside it as it was, depth_frame.cpp:46, which only the materialised mask knows) */
	const int off_x = roi[0] < 0 ? 0 : roi[0], off_y = roi[1] < 0 ? 0 : roi[1];
	const int size_x = (roi[2] < 0 || off_x + roi[2] > intrinsics.width) ? intrinsics.width : roi[2];
	const int size_y = (roi[3] < 0 || off_y + roi[3] > intrinsics.height) ? intrinsics.height : roi[3];
	last_filter.valid = true;
	last_filter.full_roi = off_x == 0 && off_y == 0 && size_x == intrinsics.width && size_y == intrinsics.height;
	last_filter.min_depth = min_depth;
	last_filter.max_depth = max_depth;
	last_filter.depth_scale = depth_scale;
}

uint* DepthFrame::getIndices() {
	if (indices_stale) {
		checkDfbError(dfb_op_filter_depth_indices(workspace, dev_mask, reinterpret_cast<uint32_t*>(dev_indices), intrinsics.width,
		                                          intrinsics.height, stream),
		              "DepthFrame filter (index list)");
		indices_stale = false;
	}
	return dev_indices;
}

unsigned DepthFrame::getMaskCount() {
	if (count_pending) {
		uint32_t count = 0;
		checkDfbError(dfb_op_filter_depth_count(workspace, &count), "DepthFrame filter (count)");
		count_pending = false;
		mask_count = count;
		if (mask_count == 0) std::cout << "Warning: empty mask!" << std::endl; /* kernels.cu:533-535 */
	}
	return mask_count;
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

void Frameset::addDependent(Pointcloud* cloud) {
	if (std::find(dependents.begin(), dependents.end(), cloud) == dependents.end()) dependents.push_back(cloud);
}

void Frameset::removeDependent(Pointcloud* cloud) {
	dependents.erase(std::remove(dependents.begin(), dependents.end(), cloud), dependents.end());
}

void Frameset::sourceWillChange() {
	while (!dependents.empty()) dependents.back()->sourceWillChange(this); /* each call removes the cloud from the list */
}

Frameset::~Frameset() {
	sourceWillChange();
	if (start) cudaEventDestroy(start);
	if (stop) cudaEventDestroy(stop);
}

void Frameset::setStream(cudaStream_t* s) { stream = s ? *s : nullptr; } /* frameset.cpp:27-32, no NPP context needed */
void Frameset::allocateDepthFrame(const Intrinsics& intr) {
	sourceWillChange();
	depth_frame.allocate(intr, stream);
}
void Frameset::allocateDepthFrameUnaligned(const Intrinsics& intr, const Extrinsics& extr) { depth_frame.allocateUnaligned(intr, extr, stream); }
void Frameset::allocateColorFrame(const Intrinsics& intr) {
	sourceWillChange();
	color_frame.allocate(intr, stream);
}
void Frameset::setDepthFrame(const uint16_t* f) {
	sourceWillChange();
	depth_frame.copyToDevice(f);
}
void Frameset::setDepthFrameUnaligned(const uint16_t* f) { depth_frame.copyUnalignedToDevice(f); }
void Frameset::setColorFrame(const uint8_t* f) {
	sourceWillChange();
	color_frame.copyToDevice(f);
}
uint16_t* Frameset::getDepth() { return depth_frame.get(); }
uint16_t* Frameset::getDepthUnaligned() { return depth_frame.getUnaligned(); }
uint8_t* Frameset::getColor() { return color_frame.get(); }
uint8_t* Frameset::getMask() { return depth_frame.getMask(); }

void Frameset::filterDepth(const float min_depth, const float max_depth, const float depth_scale, const std::array<int, 4> roi) {
	sourceWillChange();
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
	sourceWillChange();
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
	sourceWillChange();
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

/* ---- Pointcloud (pointcloud.cpp) ----------------------------------------------------------------------
 * A cloud is either LAZY — `recipe` lists, per deproject()/append() call, the source frameset and the transforms seen
 * since — or MATERIALISED: dev_points holds float4 points exactly as the reference's operators leave them. */

Pointcloud::Pointcloud() {
	for (int i = 0; i < 16; ++i) transform_rm[i] = (i % 5 == 0) ? 1.0f : 0.0f;
}

Pointcloud::~Pointcloud() { free(); }

void Pointcloud::attach_sources() {
	for (auto& seg : recipe)
		if (seg.source) seg.source->addDependent(this);
}

void Pointcloud::detach_sources() {
	for (auto& seg : recipe)
		if (seg.source) seg.source->removeDependent(this);
}

void Pointcloud::release_points() {
	pool_put(dev_points, points_bytes, points_event);
	dev_points = nullptr;
	points_event = nullptr;
	points_bytes = 0;
}

void Pointcloud::allocate(unsigned count) { /* pointcloud.cpp:21-35: nothing is allocated until somebody needs the points */
	detach_sources();
	release_points();
	point_count = count;
	pointer_offset = 0;
	segments.clear();
	recipe.clear();
	lazy = true;
}

void Pointcloud::free() { /* pointcloud.cpp:40-45 */
	detach_sources();
	release_points();
	pool_put(dev_keys, keys_bytes, keys_event);
	dev_keys = nullptr;
	keys_event = nullptr;
	keys_bytes = 0;
	key_pixels = 0;
	point_count = 0;
	pointer_offset = 0;
	segments.clear();
	recipe.clear();
	lazy = false;
}

/* run the recorded operators the way the reference runs them: deproject at the segment's place, then its transforms */
void Pointcloud::materialise() {
	if (!lazy) return;
	lazy = false;
	dev_points = static_cast<float*>(pool_get(static_cast<size_t>(point_count ? point_count : 1) * 4 * sizeof(float), stream, &points_bytes));
	/* the reference leaves the cloud uninitialised; zeros (Z = 0 -> culled by project) make a partly filled cloud deterministic */
	checkError(cudaMemsetAsync(dev_points, 0, static_cast<size_t>(point_count ? point_count : 1) * 4 * sizeof(float), stream),
	           "Pointcloud cudaMemset dev_points");
	size_t off = 0;
	for (auto& seg : recipe) {
		if (seg.source && seg.count) {
			cudaStream_t fs = *seg.source->getStream();
			if (fs != stream) cudaStreamSynchronize(stream); /* the memset above, before another stream writes the points */
			seg.source->deprojectToPointcloud(dev_points + off * 4, seg.depth_scale);
			if (fs != stream) seg.source->sync(); /* the cloud's own stream consumes the points next */
			for (int t = 0; t < seg.n_transforms; ++t)
				checkDfbError(dfb_op_transform(dev_points + off * 4, seg.count, seg.transforms[t], stream), "Pointcloud transform");
		}
		off += seg.count;
	}
	if (off < point_count) { /* transform() covers the whole allocation (pointcloud.cpp:80-83), also its unfilled tail */
		const Segment* last = recipe.empty() ? nullptr : &recipe.back();
		for (int t = 0; last && t < last->n_transforms; ++t)
			checkDfbError(dfb_op_transform(dev_points + off * 4, static_cast<uint32_t>(point_count - off), last->transforms[t], stream),
			              "Pointcloud transform");
	}
	mark_use(&points_event, stream);
	detach_sources();
	recipe.clear();
}

void Pointcloud::sourceWillChange(Frameset* source) {
	(void)source;
	materialise(); /* also drops this cloud from every source's list */
	for (auto& seg : recipe) seg.source = nullptr;
}

float* Pointcloud::get() {
	materialise();
	return dev_points;
}

void Pointcloud::set(const float* host_points) { /* pointcloud.cpp:51-57 */
	detach_sources();
	recipe.clear();
	if (lazy || !dev_points) {
		lazy = false;
		dev_points = static_cast<float*>(pool_get(static_cast<size_t>(point_count ? point_count : 1) * 4 * sizeof(float), stream, &points_bytes));
	}
	checkError(cudaMemcpyAsync(dev_points, host_points, static_cast<size_t>(point_count) * 4 * sizeof(float), cudaMemcpyHostToDevice, stream),
	           "Pointcloud cudaMemcpy host_points to dev_points");
	mark_use(&points_event, stream);
	pointer_offset = point_count;
	segments.assign(1, point_count);
}

void Pointcloud::setTransform(const double t[16]) { /* pointcloud.cpp:63-75: Affine3d -> cast<float>() -> row-major */
	dfb_affine_to_float16(t, transform_rm);
}

void Pointcloud::setTransform(const float t[16]) { std::memcpy(transform_rm, t, sizeof(transform_rm)); }

void Pointcloud::record_transform(const float t[16]) {
	for (auto& seg : recipe) {
		if (seg.n_transforms == 2) { /* more than the fused kernel's camera + fused transform: compute what we have */
			materialise();
			return;
		}
	}
	for (auto& seg : recipe) std::memcpy(seg.transforms[seg.n_transforms++], t, 16 * sizeof(float));
}

void Pointcloud::transform() { /* pointcloud.cpp:80-83: the whole allocated cloud */
	if (lazy) {
		uint64_t filled = 0;
		for (auto& seg : recipe) filled += seg.count;
		if (filled == point_count && !recipe.empty()) {
			record_transform(transform_rm);
			if (lazy) return;
		} else {
			materialise(); /* an unfilled tail is transformed too: only the eager operators know what it holds */
		}
	}
	checkDfbError(dfb_op_transform(dev_points, point_count, transform_rm, stream), "Pointcloud transform");
	mark_use(&points_event, stream);
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
	if (lazy && recipe.size() <= 1) {
		/* like the reference, deproject always writes at the START of the cloud (frameset.cpp:144,150 get dev_points) */
		detach_sources();
		Segment seg;
		seg.source = &frameset;
		seg.depth_scale = depth_scale;
		seg.count = n;
		recipe.assign(1, seg);
		attach_sources();
	} else {
		materialise();
		frameset.deprojectToPointcloud(dev_points, depth_scale);
		if (*frameset.getStream() != stream) frameset.sync(); /* the cloud's own stream consumes the points next */
		mark_use(&points_event, *frameset.getStream());
	}
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
	uint64_t filled = 0, other_filled = 0;
	for (auto& seg : recipe) filled += seg.count;
	for (auto& seg : cloud.recipe) other_filled += seg.count;
	if (lazy && cloud.lazy && filled == pointer_offset && other_filled == n) {
		/* the reference copies the points as they are NOW: so does the recipe (transforms by value) */
		detach_sources();
		for (auto& seg : cloud.recipe) recipe.push_back(seg);
		attach_sources();
	} else {
		materialise();
		checkError(cudaMemcpyAsync(dev_points + static_cast<size_t>(pointer_offset) * 4, cloud.get(), static_cast<size_t>(n) * 4 * sizeof(float),
		                           cudaMemcpyDeviceToDevice, stream),
		           "Pointcloud cudaMemcpy cloud.get() to dev_points + pointer_offset");
		mark_use(&points_event, stream);
		mark_use(&cloud.points_event, stream);
	}
	pointer_offset += n;
	segments.push_back(n);
}

/* The recorded sequence as ONE launch, when it has the shape processSyncedFrames gives it (fusion_node.cpp:746-784):
 * every segment a full-frame filtered frameset with colour, at most one transform of its own and the same final
 * transform for all, the allocation exactly filled. Returns false when the eager operators have to do it. */
bool Pointcloud::try_fused_project(Frameset& target, bool mirror_image) {
	if (!lazy || recipe.empty() || recipe.size() > DFB_MAX_CAMERAS) return false;
	uint64_t filled = 0;
	const int nt_last = recipe[0].n_transforms;
	if (nt_last < 1) return false;
	const float* fused = recipe[0].transforms[nt_last - 1];
	dfb_fuse_desc d;
	std::memset(&d, 0, sizeof(d));
	d.n_cameras = static_cast<int>(recipe.size());
	const uint16_t* depth_ptr[DFB_MAX_CAMERAS];
	const uint8_t* color_ptr[DFB_MAX_CAMERAS];
	for (size_t i = 0; i < recipe.size(); ++i) {
		Segment& seg = recipe[i];
		Frameset* fs = seg.source;
		if (!fs || fs == &target || !fs->depth().exists() || !fs->color().exists()) return false;
		const DepthFrame::FilterState& flt = fs->depth().lastFilter();
		if (!flt.valid || !flt.full_roi || flt.depth_scale != seg.depth_scale) return false;
		if (seg.count != fs->getMaskCount()) return false;
		if (seg.n_transforms < 1 || seg.n_transforms > 2) return false;
		if (std::memcmp(seg.transforms[seg.n_transforms - 1], fused, 16 * sizeof(float)) != 0) return false;
		const Intrinsics di = fs->getDepthIntrinsics(), ci = fs->getColorIntrinsics();
		if (di.width != ci.width || di.height != ci.height) return false;
		dfb_camera_desc& c = d.cam[i];
		std::memcpy(&c.intr, &di, sizeof(c.intr));
		c.depth_scale = seg.depth_scale;
		c.min_depth = flt.min_depth;
		c.max_depth = flt.max_depth;
		c.roi[0] = c.roi[1] = c.roi[2] = c.roi[3] = -1;
		c.has_transform = seg.n_transforms == 2;
		for (int k = 0; k < 16; ++k) c.to_ref[k] = c.has_transform ? seg.transforms[0][k] : ((k % 5 == 0) ? 1.0f : 0.0f);
		depth_ptr[i] = fs->depth().get();
		color_ptr[i] = fs->color().get();
		filled += seg.count;
	}
	if (filled != point_count) return false;
	std::memcpy(d.fused_transform, fused, 16 * sizeof(float));
	const Intrinsics out = target.getColorIntrinsics();
	std::memcpy(&d.out_intr, &out, sizeof(d.out_intr));
	d.mirror = mirror_image ? 1 : 0;
	cudaStream_t fs = *target.getStream();
	if (fs != stream) cudaStreamSynchronize(stream); /* the reference launches project on the FRAMESET's stream too */
	const int rc = fuse_with_cached_fuser(d, depth_ptr, color_ptr, target.getColor(), fs);
	if (rc == DFB_ERR_INVALID) return false; /* a description the fused path does not take: the eager operators decide (loudly) */
	checkDfbError(rc, "Pointcloud project (fused)");
	return true;
}

void Pointcloud::project(Frameset& frameset, bool mirror_image) { /* pointcloud.cpp:136-157 */
	if (try_fused_project(frameset, mirror_image)) return;
	materialise();
	const Intrinsics intr = frameset.getColorIntrinsics();
	const size_t pixels = static_cast<size_t>(intr.width) * intr.height;
	cudaStream_t fs = *frameset.getStream();
	if (key_pixels != pixels) {
		pool_put(dev_keys, keys_bytes, keys_event);
		keys_event = nullptr;
		dev_keys = static_cast<uint64_t*>(pool_get(pixels * sizeof(uint64_t), fs, &keys_bytes));
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
	mark_use(&points_event, fs);
	mark_use(&keys_event, fs);
}

void Pointcloud::copyToHost(float* host_points) { /* pointcloud.cpp:163-169 */
	materialise();
	checkError(cudaMemcpyAsync(host_points, dev_points, static_cast<size_t>(point_count) * 4 * sizeof(float), cudaMemcpyDeviceToHost, stream),
	           "cudaMemcpy cam.dev_points to cloud_ptr");
	cudaStreamSynchronize(stream);
}
