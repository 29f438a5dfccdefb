/*
 * capi_core.cu — lifecycle, error plumbing, parameter derivation and the dfb_fuser plan
 * (batched fused path) of the C ABI declared in include/dfb/dfb.h.
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "dfb_internal.h"

namespace dfb {

std::atomic<uint64_t> g_launches{0};

namespace {
thread_local std::string t_error;
std::mutex g_dev_mutex;
int g_dev_checked[64] = {0}; /* 0 unknown, 1 ok, -1 unsupported */
}  // namespace

void set_error(const std::string& msg) { t_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
	char buf[512];
	std::snprintf(buf, sizeof(buf), "Cuda error: %s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
	t_error = buf;
	return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? DFB_ERR_NO_DEVICE
	       : (e == cudaErrorMemoryAllocation)                         ? DFB_ERR_NOMEM
	                                                                  : DFB_ERR_CUDA;
}

/* No CPU fallback: every compute entry point starts here. */
int require_device() {
	int dev = -1;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) {
		cudaGetLastError();
		set_error(std::string("no usable CUDA device: ") + cudaGetErrorString(e) +
		          " (libdfb_b200 has no CPU fallback)");
		return DFB_ERR_NO_DEVICE;
	}
	if (dev < 0 || dev >= 64) return DFB_OK;
	std::lock_guard<std::mutex> lk(g_dev_mutex);
	if (g_dev_checked[dev] == 0) {
		int major = 0;
		e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
		g_dev_checked[dev] = (e == cudaSuccess && major == 10) ? 1 : -1;
	}
	if (g_dev_checked[dev] < 0) {
		set_error("device is not compute capability 10.x: libdfb_b200 carries sm_100a code only");
		return DFB_ERR_NO_DEVICE;
	}
	return DFB_OK;
}

static void copy_mat34(const float* T16, Mat34& out) {
	for (int i = 0; i < 12; ++i) out.m[i] = T16[i];
}

int make_cam_params(const dfb_camera_desc& cd, CamParams& out) {
	DFB_REQUIRE(cd.intr.width > 0 && cd.intr.height > 0 && cd.intr.width <= 32768 && cd.intr.height <= 32768,
	            "camera resolution must be in [1, 32768]");
	DFB_REQUIRE(cd.depth_scale > 0.0f, "depth_scale must be positive");
	out.width = cd.intr.width;
	out.height = cd.intr.height;
	/* kernels.cu:379-384, then clamped to the frame (the reference would overrun rows) */
	int off_x = cd.roi[0] < 0 ? 0 : cd.roi[0];
	int off_y = cd.roi[1] < 0 ? 0 : cd.roi[1];
	int size_x = (cd.roi[2] < 0 || off_x + cd.roi[2] > out.width) ? out.width : cd.roi[2];
	int size_y = (cd.roi[3] < 0 || off_y + cd.roi[3] > out.height) ? out.height : cd.roi[3];
	if (off_x > out.width) off_x = out.width;
	if (off_y > out.height) off_y = out.height;
	if (off_x + size_x > out.width) size_x = out.width - off_x;
	if (off_y + size_y > out.height) size_y = out.height - off_y;
	out.roi_x0 = off_x;
	out.roi_y0 = off_y;
	out.roi_w = size_x;
	out.roi_h = size_y;
	uint16_t lt, gt;
	dfb_depth_thresholds(cd.min_depth, cd.max_depth, cd.depth_scale, &lt, &gt);
	out.lt = lt;
	out.gt = gt;
	out.scale = cd.depth_scale;
	out.ppx = cd.intr.ppx;
	out.ppy = cd.intr.ppy;
	out.fx = cd.intr.fx;
	out.fy = cd.intr.fy;
	out.has_transform = cd.has_transform ? 1 : 0;
	copy_mat34(cd.to_ref, out.to_ref);
	const uint32_t lt1 = out.lt < 1u ? 1u : out.lt; /* mask = depth > 0 after thresholding, kernels.cu:392 */
	out.mid = 0.5f * (float)(lt1 + out.gt);
	out.halfw = (out.gt >= lt1) ? 0.5f * (float)(out.gt - lt1) : -1.0f;
	out.roi_all = (off_x == 0 && off_y == 0 && size_x == out.width && size_y == out.height) ? 1 : 0;
	/* Supported numeric domain. Inside it every intermediate of the fused arithmetic is finite (|fx*X| stays
	 * below 1e19), which the packed kernels rely on to test image bounds on integers; the reference itself
	 * is undefined for non-finite projections. */
	DFB_REQUIRE(std::isfinite(cd.depth_scale) && cd.depth_scale <= 1.5f, "depth_scale must be finite and <= 1.5 m/unit");
	DFB_REQUIRE(std::isfinite(cd.intr.fx) && std::isfinite(cd.intr.fy) && std::isfinite(cd.intr.ppx) && std::isfinite(cd.intr.ppy) &&
	                std::fabs(cd.intr.fx) > 0.0f && std::fabs(cd.intr.fy) > 0.0f,
	            "camera intrinsics must be finite with non-zero focal lengths");
	DFB_REQUIRE(((float)cd.intr.width + std::fabs(cd.intr.ppx)) / std::fabs(cd.intr.fx) <= 1.0e4f &&
	                ((float)cd.intr.height + std::fabs(cd.intr.ppy)) / std::fabs(cd.intr.fy) <= 1.0e4f,
	            "camera intrinsics outside the supported range: |u - ppx| / fx must stay below 1e4");
	if (cd.has_transform) {
		for (int r = 0; r < 3; ++r) {
			for (int c = 0; c < 3; ++c)
				DFB_REQUIRE(std::fabs(cd.to_ref[4 * r + c]) <= 16.0f, "camera transform: linear part entries must be finite and <= 16");
			DFB_REQUIRE(std::fabs(cd.to_ref[4 * r + 3]) <= 1.0e6f, "camera transform: translation must be finite and <= 1e6 m");
		}
	}
	return DFB_OK;
}

int make_fuse_params(const dfb_fuse_desc& d, FuseParams& P) {
	DFB_REQUIRE(d.n_cameras >= 1 && d.n_cameras <= DFB_MAX_CAMERAS, "n_cameras must be in [1, DFB_MAX_CAMERAS]");
	std::memset(&P, 0, sizeof(P));
	P.n_cameras = d.n_cameras;
	for (int c = 0; c < d.n_cameras; ++c) {
		int rc = make_cam_params(d.cam[c], P.cam[c]);
		if (rc != DFB_OK) return rc;
		DFB_REQUIRE(d.cam[c].intr.width == d.cam[0].intr.width && d.cam[c].intr.height == d.cam[0].intr.height,
		            "all cameras of a fuser must share one resolution");
	}
	DFB_REQUIRE(d.out_intr.width > 0 && d.out_intr.height > 0 && d.out_intr.width <= 32768 && d.out_intr.height <= 32768,
	            "fused image resolution must be in [1, 32768]");
	for (int r = 0; r < 3; ++r) {
		for (int c = 0; c < 3; ++c)
			DFB_REQUIRE(std::fabs(d.fused_transform[4 * r + c]) <= 16.0f, "fused transform: linear part entries must be finite and <= 16");
		DFB_REQUIRE(std::fabs(d.fused_transform[4 * r + 3]) <= 1.0e6f, "fused transform: translation must be finite and <= 1e6 m");
	}
	DFB_REQUIRE(std::fabs(d.out_intr.fx) <= 1.0e6f && std::fabs(d.out_intr.fy) <= 1.0e6f && std::fabs(d.out_intr.ppx) <= 1.0e6f &&
	                std::fabs(d.out_intr.ppy) <= 1.0e6f,
	            "fused image intrinsics must be finite and <= 1e6");
	P.n_out_pixels = d.out_intr.width * d.out_intr.height;
	copy_mat34(d.fused_transform, P.fused);
	P.proj.ppx = d.out_intr.ppx;
	P.proj.ppy = d.out_intr.ppy;
	P.proj.fx = d.out_intr.fx;
	P.proj.fy = d.out_intr.fy;
	P.proj.wf = (float)d.out_intr.width;
	P.proj.hf = (float)d.out_intr.height;
	P.proj.width = d.out_intr.width;
	P.proj.height = d.out_intr.height;
	P.proj.mirror = d.mirror ? 1 : 0;
	return DFB_OK;
}

}  // namespace dfb

using namespace dfb;

/* ------------------------------------------------------------------------------------------------ */

/* what is known about a key ring between calls: how many frameset slots were ALLOCATED (the stream kernel parks rejected
 * atomics behind them, wherever the current group size ends) and the stream kernel's epoch phase (launch_fuse_stream):
 * 0 = KEY_EMPTY everywhere */
struct RingState {
	int alloc_slots = 0;
	uint32_t phase = 0;
};

struct dfb_fuser {
	dfb_fuse_desc desc;
	FuseParams P;
	int device = 0;
	int group = 1;
	int variant = 0;      /* requested: 0 auto, 1 row, 2 column */
	int auto_variant = 1; /* what pick_variant chose for this configuration */
	size_t in_pix = 0, out_pix = 0;
	unsigned long long* d_keys = nullptr;
	int keys_group = 0; /* framesets the key ring holds */
	/* single-launch stream kernel (variant 4): ring of `stream_slots` key slots + dependency counters */
	int stream_slots = 2, stream_lag = 1;
	RingState ring; /* of d_keys */
	unsigned int* d_counters = nullptr;
	size_t counters_cap = 0;
	/* deprojection ray tables (x = (u-ppx)/fx per column, y per row, per camera); rebuilt lazily on the
	 * caller's stream after create/update. Tables replaced by an update stay alive until destroy: a fuse
	 * call on another stream may still be reading them. */
	float* d_rays = nullptr;
	int x_pitch = 0, y_pitch = 0;
	bool rays_dirty = true;
	std::vector<float*> retired_rays;
	/* host pipeline */
	static constexpr int kStages = 3;
	struct Stage {
		cudaStream_t s = nullptr;
		uint16_t* d_depth = nullptr;
		uint8_t* d_color = nullptr;
		uint8_t* d_out = nullptr;
		unsigned long long* d_keys = nullptr;
		RingState ring;
		unsigned int* d_counters = nullptr;
		uint8_t* d_tmp = nullptr; /* unfiltered frames of the chunk when a post filter is set */
	} stage[kStages];
	int stage_group = 0;
	/* Frameset::filterColor (fusion_node.cpp:789) folded into the batch: 0 none, 1 gauss, 2 median */
	int post_filter = 0;
	static constexpr int kFilterChunk = 128; /* framesets fused into the scratch before one filter launch */
	uint8_t* d_filter_tmp = nullptr;
	/* dfb_fuser_fuse_images: one packed frameset, used when the caller's per-camera buffers cannot be described to the
	 * stream kernel as a strided pool */
	uint16_t* d_gather_depth = nullptr;
	uint8_t* d_gather_color = nullptr;
	/* The key ring, the counters and the ray tables belong to ONE call at a time. Calls may arrive on different
	 * streams: each call first makes its stream wait for the event the previous call recorded when it was queued. */
	cudaEvent_t last_call = nullptr;
};

static int fuser_enter(dfb_fuser* f, cudaStream_t s) {
	if (!f->last_call) DFB_CUDA(cudaEventCreateWithFlags(&f->last_call, cudaEventDisableTiming));
	else DFB_CUDA(cudaStreamWaitEvent(s, f->last_call, 0));
	return DFB_OK;
}

static int fuser_leave(dfb_fuser* f, cudaStream_t s, int rc) {
	if (f->last_call) cudaEventRecord(f->last_call, s);
	return rc;
}

/* Which traversal gives warp-contiguous atomics? Project the centre pixel of camera 0 and its right / lower
 * neighbours at mid range (host double math is enough: this only picks a schedule, never a result). */
static int pick_variant(const dfb_fuse_desc& d) {
	const dfb_camera_desc& c = d.cam[0];
	const double z = 0.5 * ((double)c.min_depth + (double)c.max_depth);
	auto out_index = [&](double u, double v) {
		double p[3] = {z * (u - c.intr.ppx) / c.intr.fx, z * (v - c.intr.ppy) / c.intr.fy, z};
		const float* T = d.fused_transform;
		double q[3];
		for (int r = 0; r < 3; ++r) q[r] = T[4 * r] * p[0] + T[4 * r + 1] * p[1] + T[4 * r + 2] * p[2] + T[4 * r + 3];
		if (!(q[2] > 0)) return 0.0;
		const double px = d.out_intr.ppx + d.out_intr.fx * q[0] / q[2];
		const double py = d.out_intr.ppy + d.out_intr.fy * q[1] / q[2];
		return py * d.out_intr.width + px;
	};
	const double uc = 0.5 * c.intr.width, vc = 0.5 * c.intr.height;
	const double base = out_index(uc, vc);
	const double du = std::fabs(out_index(uc + 1, vc) - base), dv = std::fabs(out_index(uc, vc + 1) - base);
	return (dv < du) ? 2 : 1; /* 2 = column traversal, 1 = row traversal */
}

static int auto_group(size_t out_pix) {
	/* keep one group's keys (8 B / output pixel) well inside the 126 MB L2 */
	const size_t budget = 48u << 20;
	size_t g = budget / (out_pix * 8);
	if (g < 1) g = 1;
	if (g > 32) g = 32;
	return (int)g;
}

static int key_ring_framesets(const dfb_fuser* f) { return f->group > f->stream_slots ? f->group : f->stream_slots; }

static int fuser_ensure_counters(unsigned int** p, size_t* cap, int n_framesets, cudaStream_t s) {
	const size_t need = fuse_stream_counter_words(n_framesets);
	if (*p && *cap >= need) return DFB_OK;
	if (*p) {
		DFB_CUDA(cudaStreamSynchronize(s));
		DFB_CUDA(cudaFree(*p));
		*p = nullptr;
	}
	DFB_CUDA(cudaMalloc(p, need * sizeof(unsigned int)));
	*cap = need;
	return DFB_OK;
}

static int fuser_ensure_keys(dfb_fuser* f, cudaStream_t s) {
	if (f->keys_group >= key_ring_framesets(f) && f->d_keys) return DFB_OK;
	if (f->d_keys) {
		DFB_CUDA(cudaStreamSynchronize(s));
		DFB_CUDA(cudaFree(f->d_keys));
		f->d_keys = nullptr;
	}
	const size_t n = f->out_pix * (size_t)key_ring_framesets(f);
	DFB_CUDA(cudaMalloc(&f->d_keys, n * sizeof(unsigned long long) + fuse_stream_scratch_bytes()));
	f->keys_group = key_ring_framesets(f);
	f->ring.alloc_slots = f->keys_group;
	f->ring.phase = 0;
	return launch_keys_clear(f->d_keys, n, s);
}

/* the two-launch variants (and anything else that wants KEY_EMPTY everywhere) after the stream kernel has left its
 * epoch state in the ring */
static int ring_make_clean(dfb_fuser* f, unsigned long long* d_keys, RingState* ring, cudaStream_t s) {
	if (ring->phase == 0) return DFB_OK;
	ring->phase = 0;
	return launch_keys_clear(d_keys, f->out_pix * (size_t)ring->alloc_slots, s);
}

static void fuser_free_stages(dfb_fuser* f) {
	for (auto& st : f->stage) {
		if (st.s) cudaStreamSynchronize(st.s);
		cudaFree(st.d_depth);
		cudaFree(st.d_color);
		cudaFree(st.d_out);
		cudaFree(st.d_keys);
		cudaFree(st.d_counters);
		cudaFree(st.d_tmp);
		st.d_tmp = nullptr;
		st.d_counters = nullptr;
		st.d_depth = nullptr;
		st.d_color = nullptr;
		st.d_out = nullptr;
		st.d_keys = nullptr;
	}
	f->stage_group = 0;
}

static int fuser_ensure_stages(dfb_fuser* f) {
	if (f->stage_group == f->group) return DFB_OK;
	fuser_free_stages(f);
	const size_t g = (size_t)f->group;
	for (auto& st : f->stage) {
		if (!st.s) DFB_CUDA(cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking));
		DFB_CUDA(cudaMalloc(&st.d_depth, g * f->desc.n_cameras * f->in_pix * 2));
		DFB_CUDA(cudaMalloc(&st.d_color, g * f->desc.n_cameras * f->in_pix * 3));
		DFB_CUDA(cudaMalloc(&st.d_out, g * f->out_pix * 3));
		const size_t ring = (size_t)key_ring_framesets(f);
		DFB_CUDA(cudaMalloc(&st.d_keys, ring * f->out_pix * 8 + fuse_stream_scratch_bytes()));
		DFB_CUDA(cudaMalloc(&st.d_counters, fuse_stream_counter_words(f->group) * sizeof(unsigned int)));
		st.ring.alloc_slots = (int)ring;
		st.ring.phase = 0;
		int rc = launch_keys_clear(st.d_keys, ring * f->out_pix, st.s);
		if (rc != DFB_OK) return rc;
	}
	f->stage_group = f->group;
	return DFB_OK;
}

static int fuser_ensure_rays(dfb_fuser* f, cudaStream_t s) {
	if (!f->rays_dirty && f->d_rays) return DFB_OK;
	if (f->d_rays) f->retired_rays.push_back(f->d_rays);
	f->d_rays = nullptr;
	f->x_pitch = (f->desc.cam[0].intr.width + 63) / 64 * 64;
	f->y_pitch = (f->desc.cam[0].intr.height + kStreamItemRows - 1) / kStreamItemRows * kStreamItemRows; /* whole items: the stream kernel bulk-copies them */
	const size_t n = (size_t)f->desc.n_cameras * (f->x_pitch + f->y_pitch);
	const int tiles_x = f->x_pitch / 64, tiles_per_image = tiles_x * (f->y_pitch / kStreamItemRows);
	DFB_CUDA(cudaMalloc(&f->d_rays, (n + (size_t)f->desc.n_cameras * tiles_per_image * (64 + kStreamItemRows)) * sizeof(float)));
	float* ytab = f->d_rays + (size_t)f->desc.n_cameras * f->x_pitch;
	int rc = launch_ray_tables(f->P, f->d_rays, ytab, f->x_pitch, f->y_pitch, s);
	if (rc != DFB_OK) return rc;
	rc = launch_tile_rays(f->P, f->d_rays + n, tiles_x, tiles_per_image, s);
	if (rc != DFB_OK) return rc;
	f->rays_dirty = false;
	return DFB_OK;
}

static RayTables fuser_rays(const dfb_fuser* f) {
	const size_t nc = (size_t)f->desc.n_cameras;
	return RayTables{f->d_rays, f->d_rays + nc * f->x_pitch, f->x_pitch, f->y_pitch, f->d_rays + nc * (f->x_pitch + f->y_pitch)};
}

static int env_int(const char* name, int fallback) {
	const char* e = std::getenv(name);
	return e ? std::atoi(e) : fallback;
}

/* the stream kernel walks tiles by row (lane = pixel pair of one source row) when source rows map to output rows, by
 * column otherwise, so that a warp's atomics are contiguous; DFB_STREAM_ROW_MAJOR=0/1 overrides (both give the same keys) */
static bool stream_row_major(const dfb_fuser* f) {
	static const int forced = env_int("DFB_STREAM_ROW_MAJOR", -1);
	return forced >= 0 ? forced != 0 : f->auto_variant == 1;
}

/* requested variant, or automatic: the single-launch stream kernel whenever the pools fit its tensor maps (it walks
 * tiles by column or by row, see stream_row_major); unsupported requests degrade 4 -> 3 -> 2 -> 1 (all variants
 * produce identical keys) */
static int effective_variant(const dfb_fuser* f, const uint16_t* d_depth, const uint8_t* d_color, const uint8_t* d_out) {
	int v = f->variant ? f->variant : 4;
	if (v == 4 && !fuse_stream_supported(f->P, d_depth, d_color, d_out)) v = 3;
	if (v == 3 && !splat_tiled_supported(f->P, d_depth, d_color)) v = 2;
	if (v == 2 && !splat_column_supported(f->P, d_depth, d_color)) v = 1;
	return v;
}

static int launch_splat(dfb_fuser* f, int v, const uint16_t* d_depth, const uint8_t* d_color, unsigned long long* d_keys,
                        int g, cudaStream_t s) {
	if (v == 1) return launch_splat_direct(f->P, d_depth, d_color, d_keys, g, s);
	int rc = fuser_ensure_rays(f, s);
	if (rc != DFB_OK) return rc;
	const RayTables rays = fuser_rays(f);
	if (v == 3) {
		static const int ctas = env_int("DFB_TILED_CTAS_PER_SM", 0);
		return launch_splat_tiled(f->P, rays, d_depth, d_color, d_keys, g, ctas, s);
	}
	return launch_splat_column(f->P, rays, d_depth, d_color, d_keys, g, s);
}

/* the whole device-side sequence for n framesets on stream s; d_counters is only used by the stream kernel */
static int fuse_groups(dfb_fuser* f, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                       unsigned long long* d_keys, RingState* ring, unsigned int* d_counters, int n, cudaStream_t s) {
	const size_t in_stride = (size_t)f->desc.n_cameras * f->in_pix;
	const int v = effective_variant(f, d_depth, d_color, d_out);
	if (v == 4) {
		int rc = fuser_ensure_rays(f, s);
		if (rc != DFB_OK) return rc;
		const RayTables rays = fuser_rays(f);
		/* one launch per <= 2^23 tickets (the kernel's ticket counters are floats, exact below 2^24): ~9000 720p pairs */
		const long long per_fs = fuse_stream_tickets_per_frameset(f->P);
		const int chunk = (int)std::min<long long>(n, std::max<long long>(1, (1ll << 23) / per_fs - 16));
		for (int f0 = 0; f0 < n; f0 += chunk) {
			const int g = (n - f0 < chunk) ? (n - f0) : chunk;
			rc = launch_fuse_stream(f->P, rays, d_depth + (size_t)f0 * in_stride, d_color + (size_t)f0 * in_stride * 3,
			                        d_out + (size_t)f0 * f->out_pix * 3, d_keys, ring->alloc_slots, d_counters, g, f->stream_slots,
			                        f->stream_lag, stream_row_major(f), nullptr, &ring->phase, s);
			if (rc != DFB_OK) return rc;
		}
		return DFB_OK;
	}
	{
		int rc = ring_make_clean(f, d_keys, ring, s);
		if (rc != DFB_OK) return rc;
	}
	for (int f0 = 0; f0 < n; f0 += f->group) {
		const int g = (n - f0 < f->group) ? (n - f0) : f->group;
		int rc = launch_splat(f, v, d_depth + (size_t)f0 * in_stride, d_color + (size_t)f0 * in_stride * 3, d_keys, g, s);
		if (rc != DFB_OK) return rc;
		rc = launch_resolve(f->P, d_keys, d_out + (size_t)f0 * f->out_pix * 3, g, s);
		if (rc != DFB_OK) return rc;
	}
	return DFB_OK;
}

extern "C" {

int dfb_version(void) { return 100; }

const char* dfb_last_error(void) { return t_error.c_str(); }

uint64_t dfb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int dfb_init(int device, int verbose) {
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		cudaGetLastError();
		set_error(std::string("Error initializing CUDA: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "no device") +
		          " (libdfb_b200 has no CPU fallback)");
		return DFB_ERR_NO_DEVICE;
	}
	if (device < 0 || device >= count) {
		set_error("Error initializing CUDA device #" + std::to_string(device));
		return DFB_ERR_INVALID;
	}
	DFB_CUDA(cudaSetDevice(device));
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	if (verbose) {
		cudaDeviceProp prop;
		DFB_CUDA(cudaGetDeviceProperties(&prop, device));
		/* same banner as Kernels::init_cuda, kernels.cu:343-350 */
		std::printf("+-- Cuda\n| Device:             #%d %s SM %d.%d\n| Multiprocessors:    %d\n| Max threads per MP: %d\n",
		            device, prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.maxThreadsPerMultiProcessor);
	}
	return DFB_OK;
}

int dfb_host_alloc(void** ptr, size_t bytes) {
	DFB_REQUIRE(ptr != nullptr, "ptr is NULL");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	DFB_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
	return DFB_OK;
}

int dfb_host_free(void* ptr) {
	if (ptr) DFB_CUDA(cudaFreeHost(ptr));
	return DFB_OK;
}

int dfb_device_alloc(void** d_ptr, size_t bytes) {
	DFB_REQUIRE(d_ptr != nullptr, "d_ptr is NULL");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	DFB_CUDA(cudaMalloc(d_ptr, bytes));
	return DFB_OK;
}

int dfb_device_free(void* d_ptr) {
	if (d_ptr) DFB_CUDA(cudaFree(d_ptr));
	return DFB_OK;
}

int dfb_ipc_export(void* d_ptr, unsigned char handle[DFB_IPC_HANDLE_BYTES]) {
	static_assert(sizeof(cudaIpcMemHandle_t) == DFB_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
	DFB_REQUIRE(d_ptr && handle, "NULL argument");
	cudaIpcMemHandle_t h;
	DFB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
	std::memcpy(handle, &h, sizeof(h));
	return DFB_OK;
}

int dfb_ipc_open(const unsigned char handle[DFB_IPC_HANDLE_BYTES], void** d_ptr) {
	DFB_REQUIRE(d_ptr && handle, "NULL argument");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	cudaIpcMemHandle_t h;
	std::memcpy(&h, handle, sizeof(h));
	DFB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return DFB_OK;
}

int dfb_ipc_close(void* d_ptr) {
	if (d_ptr) DFB_CUDA(cudaIpcCloseMemHandle(d_ptr));
	return DFB_OK;
}

int dfb_fuser_create(dfb_fuser** out, const dfb_fuse_desc* desc) {
	DFB_REQUIRE(out && desc, "NULL argument");
	*out = nullptr;
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	dfb_fuser* f = new (std::nothrow) dfb_fuser();
	if (!f) return DFB_ERR_NOMEM;
	f->desc = *desc;
	rc = make_fuse_params(*desc, f->P);
	if (rc != DFB_OK) {
		delete f;
		return rc;
	}
	cudaGetDevice(&f->device);
	f->in_pix = (size_t)desc->cam[0].intr.width * desc->cam[0].intr.height;
	f->out_pix = (size_t)desc->out_intr.width * desc->out_intr.height;
	f->group = auto_group(f->out_pix);
	f->auto_variant = pick_variant(*desc);
	{ /* stream kernel: a ring of key slots that stays L2-resident. A frameset's resolve is queued `lag` framesets
	   * behind its splat and its slot is reused `slots` framesets later, so slots - 1 framesets of work separate the
	   * two dependencies; small images need more slots for the same slack in work items. Measured on a B200: about
	   * 44 MB of keys stay in L2 next to the streaming inputs (720p: 6 slots -> 0.41 GB of DRAM writes per 128
	   * framesets, 8 slots -> 1.07 GB, 10 -> 1.86 GB and the throughput drops). */
		const size_t slot_bytes = f->out_pix * 8;
		int slots = (int)((60u << 20) / slot_bytes);
		slots = slots < 1 ? 1 : (slots > 16 ? 16 : slots);
		f->stream_slots = env_int("DFB_STREAM_SLOTS", slots);
		if (f->stream_slots < 1) f->stream_slots = 1;
		/* lag: tools/ring_sweep.sh on a B200 — 8 slots: lag 3 beats 4 by 1.3 %, 7 slots: 2 beats 3, 16 slots: 6 = 8,
		 * a 4K frame (one 66 MB slot): lag 0 beats two slots with lag 1 by 6 % */
		f->stream_lag = env_int("DFB_STREAM_LAG", f->stream_slots * 3 / 8);
	}
	*out = f;
	return DFB_OK;
}

int dfb_fuser_destroy(dfb_fuser* f) {
	if (!f) return DFB_OK;
	fuser_free_stages(f);
	for (auto& st : f->stage)
		if (st.s) cudaStreamDestroy(st.s);
	cudaFree(f->d_keys);
	cudaFree(f->d_counters);
	cudaFree(f->d_rays);
	cudaFree(f->d_filter_tmp);
	cudaFree(f->d_gather_depth);
	cudaFree(f->d_gather_color);
	if (f->last_call) cudaEventDestroy(f->last_call);
	for (float* p : f->retired_rays) cudaFree(p);
	delete f;
	return DFB_OK;
}

int dfb_fuser_update(dfb_fuser* f, const dfb_fuse_desc* desc) {
	DFB_REQUIRE(f && desc, "NULL argument");
	DFB_REQUIRE(desc->n_cameras == f->desc.n_cameras && desc->cam[0].intr.width == f->desc.cam[0].intr.width &&
	                desc->cam[0].intr.height == f->desc.cam[0].intr.height &&
	                desc->out_intr.width == f->desc.out_intr.width && desc->out_intr.height == f->desc.out_intr.height,
	            "dfb_fuser_update cannot change camera count or resolutions");
	FuseParams P;
	int rc = make_fuse_params(*desc, P);
	if (rc != DFB_OK) return rc;
	bool same_rays = true; /* the ray tables depend on the cameras' ppx / ppy / fx / fy only */
	for (int c = 0; c < desc->n_cameras; ++c)
		same_rays = same_rays && desc->cam[c].intr.ppx == f->desc.cam[c].intr.ppx && desc->cam[c].intr.ppy == f->desc.cam[c].intr.ppy &&
		            desc->cam[c].intr.fx == f->desc.cam[c].intr.fx && desc->cam[c].intr.fy == f->desc.cam[c].intr.fy;
	f->desc = *desc;
	f->P = P;
	f->auto_variant = pick_variant(*desc);
	if (!same_rays) f->rays_dirty = true;
	if (f->retired_rays.size() > 16) { /* bound the parked tables: make sure nothing reads them, then free */
		DFB_CUDA(cudaDeviceSynchronize());
		for (float* p : f->retired_rays) cudaFree(p);
		f->retired_rays.clear();
	}
	return DFB_OK;
}

int dfb_fuser_set_group(dfb_fuser* f, int g) {
	DFB_REQUIRE(f != nullptr, "NULL fuser");
	DFB_REQUIRE(g >= 0 && g <= 4096, "framesets_per_group out of range");
	f->group = (g == 0) ? auto_group(f->out_pix) : g;
	return DFB_OK;
}

int dfb_fuser_get_group(const dfb_fuser* f) { return f ? f->group : DFB_ERR_INVALID; }

int dfb_fuser_set_variant(dfb_fuser* f, int variant) {
	DFB_REQUIRE(f != nullptr, "NULL fuser");
	DFB_REQUIRE(variant >= 0 && variant <= 4, "variant must be 0 (auto), 1 (row), 2 (column), 3 (tiled TMA) or 4 (stream)");
	f->variant = variant;
	return DFB_OK;
}

int dfb_fuser_fuse_device(dfb_fuser* f, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                          int n_framesets, void* stream) {
	DFB_REQUIRE(f && d_depth && d_color && d_out, "NULL argument");
	DFB_REQUIRE(n_framesets >= 0, "n_framesets must be >= 0");
	DFB_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 3u) == 0, "d_out must be 4-byte aligned");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	if (n_framesets == 0) return DFB_OK;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	rc = fuser_enter(f, s);
	if (rc != DFB_OK) return rc;
	rc = fuser_ensure_keys(f, s);
	if (rc != DFB_OK) return rc;
	rc = fuser_ensure_counters(&f->d_counters, &f->counters_cap, n_framesets, s);
	if (rc != DFB_OK) return rc;
	if (!f->post_filter) return fuser_leave(f, s, fuse_groups(f, d_depth, d_color, d_out, f->d_keys, &f->ring, f->d_counters, n_framesets, s));
	/* post filter: fuse a chunk into the scratch, filter scratch -> d_out (every byte written), next chunk */
	const size_t frame = f->out_pix * 3, in_stride = (size_t)f->desc.n_cameras * f->in_pix;
	if (!f->d_filter_tmp) DFB_CUDA(cudaMalloc(&f->d_filter_tmp, (size_t)dfb_fuser::kFilterChunk * frame));
	for (int f0 = 0; f0 < n_framesets; f0 += dfb_fuser::kFilterChunk) {
		const int g = n_framesets - f0 < dfb_fuser::kFilterChunk ? n_framesets - f0 : dfb_fuser::kFilterChunk;
		rc = fuse_groups(f, d_depth + (size_t)f0 * in_stride, d_color + (size_t)f0 * in_stride * 3, f->d_filter_tmp, f->d_keys, &f->ring,
		                 f->d_counters, g, s);
		if (rc != DFB_OK) return rc;
		rc = launch_post_filter_batch(f->d_filter_tmp, d_out + (size_t)f0 * frame, f->desc.out_intr.width, f->desc.out_intr.height, g,
		                              f->post_filter == 2, s);
		if (rc != DFB_OK) return rc;
	}
	return fuser_leave(f, s, DFB_OK);
}

/* One frameset whose cameras sit in SEPARATE device buffers — the layout the reference's Frameset objects have
 * (depth_frame.cpp:34, color_frame.cpp:33) and what the lazy Pointcloud of the drop-in classes hands over. One or two
 * cameras are described to the stream kernel as a strided two-image pool (no copy); anything else is packed into a
 * pool of the fuser first. */
int dfb_fuser_fuse_images(dfb_fuser* f, const uint16_t* const* d_depth, const uint8_t* const* d_color, uint8_t* d_out, void* stream) {
	DFB_REQUIRE(f && d_depth && d_color && d_out, "NULL argument");
	DFB_REQUIRE((reinterpret_cast<uintptr_t>(d_out) & 3u) == 0, "d_out must be 4-byte aligned");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	const int nc = f->desc.n_cameras;
	for (int c = 0; c < nc; ++c) DFB_REQUIRE(d_depth[c] && d_color[c], "NULL camera image");
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	rc = fuser_enter(f, s);
	if (rc != DFB_OK) return rc;
	rc = fuser_ensure_keys(f, s);
	if (rc != DFB_OK) return rc;
	rc = fuser_ensure_counters(&f->d_counters, &f->counters_cap, 1, s);
	if (rc != DFB_OK) return rc;
	const size_t depth_bytes = f->in_pix * 2, color_bytes = f->in_pix * 3;
	if (nc <= 2 && !f->post_filter && (f->variant == 0 || f->variant == 4)) {
		StreamLayout L{};
		L.imgs_per_fs = nc;
		auto plan = [&](const void* a, const void* b, size_t bytes, unsigned char* slot, size_t* stride, const void** base) {
			const uintptr_t pa = reinterpret_cast<uintptr_t>(a), pb = nc == 2 ? reinterpret_cast<uintptr_t>(b) : pa + bytes;
			const uintptr_t lo = pa < pb ? pa : pb, hi = pa < pb ? pb : pa;
			*base = reinterpret_cast<const void*>(lo);
			*stride = nc == 2 ? (size_t)(hi - lo) : ((bytes + 15) & ~(size_t)15);
			slot[0] = pa <= pb ? 0 : 1;
			slot[1] = pa <= pb ? 1 : 0;
			return (*stride % 16 == 0) && *stride >= bytes && *stride < (1ull << 40);
		};
		const void *base_d = nullptr, *base_c = nullptr;
		const bool ok = plan(d_depth[0], nc == 2 ? d_depth[1] : nullptr, depth_bytes, L.slot_d, &L.depth_stride_bytes, &base_d) &&
		                plan(d_color[0], nc == 2 ? d_color[1] : nullptr, color_bytes, L.slot_c, &L.color_stride_bytes, &base_c);
		if (ok && fuse_stream_supported(f->P, static_cast<const uint16_t*>(base_d), static_cast<const uint8_t*>(base_c), d_out)) {
			rc = fuser_ensure_rays(f, s);
			if (rc != DFB_OK) return rc;
			const RayTables rays = fuser_rays(f);
			return fuser_leave(f, s, launch_fuse_stream(f->P, rays, static_cast<const uint16_t*>(base_d), static_cast<const uint8_t*>(base_c),
			                                            d_out, f->d_keys, f->ring.alloc_slots, f->d_counters, 1, f->stream_slots,
			                                            f->stream_lag, stream_row_major(f), &L, &f->ring.phase, s));
		}
	}
	if (!f->d_gather_depth) {
		DFB_CUDA(cudaMalloc(&f->d_gather_depth, (size_t)nc * depth_bytes));
		DFB_CUDA(cudaMalloc(&f->d_gather_color, (size_t)nc * color_bytes));
	}
	for (int c = 0; c < nc; ++c) {
		DFB_CUDA(cudaMemcpyAsync(f->d_gather_depth + (size_t)c * f->in_pix, d_depth[c], depth_bytes, cudaMemcpyDeviceToDevice, s));
		DFB_CUDA(cudaMemcpyAsync(f->d_gather_color + (size_t)c * color_bytes, d_color[c], color_bytes, cudaMemcpyDeviceToDevice, s));
	}
	return dfb_fuser_fuse_device(f, f->d_gather_depth, f->d_gather_color, d_out, 1, stream);
}

int dfb_fuser_set_post_filter(dfb_fuser* f, int mode) {
	DFB_REQUIRE(f != nullptr, "NULL argument");
	DFB_REQUIRE(mode >= 0 && mode <= 2, "post filter: 0 = none, 1 = gauss, 2 = median");
	DFB_REQUIRE(mode == 0 || (f->out_pix * 3) % 4 == 0, "post filter needs a fused frame size that is a multiple of 4 bytes");
	f->post_filter = mode;
	return DFB_OK;
}

int dfb_fuser_profile_device(dfb_fuser* f, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                             int n_framesets, void* stream, float ms[2], int launches[2]) {
	DFB_REQUIRE(f && d_depth && d_color && d_out && ms && launches, "NULL argument");
	DFB_REQUIRE(n_framesets > 0, "n_framesets must be > 0");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	rc = fuser_ensure_keys(f, s);
	if (rc != DFB_OK) return rc;
	const int v = effective_variant(f, d_depth, d_color, d_out);
	if (v == 4) { /* one kernel does everything: report it as the "splat" entry */
		rc = fuser_ensure_counters(&f->d_counters, &f->counters_cap, n_framesets, s);
		if (rc != DFB_OK) return rc;
		cudaEvent_t e0, e1;
		DFB_CUDA(cudaEventCreate(&e0));
		DFB_CUDA(cudaEventCreate(&e1));
		const uint64_t l0 = g_launches.load();
		DFB_CUDA(cudaEventRecord(e0, s));
		rc = fuse_groups(f, d_depth, d_color, d_out, f->d_keys, &f->ring, f->d_counters, n_framesets, s);
		if (rc != DFB_OK) return rc;
		DFB_CUDA(cudaEventRecord(e1, s));
		DFB_CUDA(cudaStreamSynchronize(s));
		DFB_CUDA(cudaEventElapsedTime(&ms[0], e0, e1));
		ms[1] = 0.f;
		launches[0] = (int)(g_launches.load() - l0);
		launches[1] = 0;
		cudaEventDestroy(e0);
		cudaEventDestroy(e1);
		return DFB_OK;
	}
	rc = ring_make_clean(f, f->d_keys, &f->ring, s);
	if (rc != DFB_OK) return rc;
	const int n_groups = (n_framesets + f->group - 1) / f->group;
	std::vector<cudaEvent_t> ev((size_t)n_groups * 3);
	for (auto& e : ev) DFB_CUDA(cudaEventCreate(&e));
	const size_t in_stride = (size_t)f->desc.n_cameras * f->in_pix;
	int gi = 0;
	for (int f0 = 0; f0 < n_framesets; f0 += f->group, ++gi) {
		const int g = (n_framesets - f0 < f->group) ? (n_framesets - f0) : f->group;
		DFB_CUDA(cudaEventRecord(ev[gi * 3 + 0], s));
		rc = launch_splat(f, v, d_depth + (size_t)f0 * in_stride, d_color + (size_t)f0 * in_stride * 3, f->d_keys, g, s);
		if (rc != DFB_OK) return rc;
		DFB_CUDA(cudaEventRecord(ev[gi * 3 + 1], s));
		rc = launch_resolve(f->P, f->d_keys, d_out + (size_t)f0 * f->out_pix * 3, g, s);
		if (rc != DFB_OK) return rc;
		DFB_CUDA(cudaEventRecord(ev[gi * 3 + 2], s));
	}
	DFB_CUDA(cudaStreamSynchronize(s));
	ms[0] = ms[1] = 0.f;
	for (int i = 0; i < n_groups; ++i) {
		float a = 0.f, b = 0.f;
		DFB_CUDA(cudaEventElapsedTime(&a, ev[i * 3 + 0], ev[i * 3 + 1]));
		DFB_CUDA(cudaEventElapsedTime(&b, ev[i * 3 + 1], ev[i * 3 + 2]));
		ms[0] += a;
		ms[1] += b;
	}
	launches[0] = launches[1] = n_groups;
	for (auto& e : ev) cudaEventDestroy(e);
	return DFB_OK;
}

int dfb_fuser_fuse_host(dfb_fuser* f, const uint16_t* h_depth, const uint8_t* h_color, uint8_t* h_out,
                        int n_framesets) {
	DFB_REQUIRE(f && h_depth && h_color && h_out, "NULL argument");
	DFB_REQUIRE(n_framesets >= 0, "n_framesets must be >= 0");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	if (n_framesets == 0) return DFB_OK;
	rc = fuser_ensure_stages(f);
	if (rc != DFB_OK) return rc;
	if (f->rays_dirty || !f->d_rays) { /* the stages run on different streams: build the tables before any of them */
		rc = fuser_ensure_rays(f, f->stage[0].s);
		if (rc != DFB_OK) return rc;
		DFB_CUDA(cudaStreamSynchronize(f->stage[0].s));
	}
	const size_t in_stride = (size_t)f->desc.n_cameras * f->in_pix;
	int chunk = 0;
	for (int f0 = 0; f0 < n_framesets; f0 += f->group, ++chunk) {
		const int g = (n_framesets - f0 < f->group) ? (n_framesets - f0) : f->group;
		dfb_fuser::Stage& st = f->stage[chunk % dfb_fuser::kStages];
		/* Frameset::setDepthFrame / setColorFrame (depth_frame.cpp:117, color_frame.cpp:66) */
		DFB_CUDA(cudaMemcpyAsync(st.d_depth, h_depth + (size_t)f0 * in_stride, (size_t)g * in_stride * 2,
		                         cudaMemcpyHostToDevice, st.s));
		DFB_CUDA(cudaMemcpyAsync(st.d_color, h_color + (size_t)f0 * in_stride * 3, (size_t)g * in_stride * 3,
		                         cudaMemcpyHostToDevice, st.s));
		if (f->post_filter) {
			if (!st.d_tmp) DFB_CUDA(cudaMalloc(&st.d_tmp, (size_t)f->stage_group * f->out_pix * 3));
			rc = fuse_groups(f, st.d_depth, st.d_color, st.d_tmp, st.d_keys, &st.ring, st.d_counters, g, st.s);
			if (rc == DFB_OK)
				rc = launch_post_filter_batch(st.d_tmp, st.d_out, f->desc.out_intr.width, f->desc.out_intr.height, g, f->post_filter == 2, st.s);
		} else {
			rc = fuse_groups(f, st.d_depth, st.d_color, st.d_out, st.d_keys, &st.ring, st.d_counters, g, st.s);
		}
		if (rc != DFB_OK) return rc;
		/* Frameset::copyColorToHost (color_frame.cpp:75-81) */
		DFB_CUDA(cudaMemcpyAsync(h_out + (size_t)f0 * f->out_pix * 3, st.d_out, (size_t)g * f->out_pix * 3,
		                         cudaMemcpyDeviceToHost, st.s));
	}
	for (auto& st : f->stage) DFB_CUDA(cudaStreamSynchronize(st.s));
	return DFB_OK;
}

}  // extern "C"
