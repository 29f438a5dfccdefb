/*
 * dfb_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY (see dfb_oracle.h).
 *
 * Plain C99 restatement of the reference hot path. Build with
 *   gcc -std=gnu99 -O2 -ffp-contract=off -fno-fast-math -fopenmp
 * (-O3 -march=native for the timed CPU baseline; -ffp-contract=off ALWAYS — the only fused
 * multiply-adds are the explicit fmaf() calls that restate nvcc's contraction of
 * kernel_transform).
 */
#include "dfb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * depth filter (kernels.cu:353-395, 501-547)
 * ---------------------------------------------------------------------------------------- */

void dfo_thresholds(float min_depth, float max_depth, float depth_scale, uint16_t* lt, uint16_t* gt) {
	/* kernels.cu:357,359: static_cast<Npp16u>(min_depth / depth_scale), all operands float.
	 * Out-of-range float->u16 is UB in C; clamp (the reference never reaches it at its configs). */
	float a = min_depth / depth_scale;
	float b = max_depth / depth_scale;
	if (!(a > 0.0f)) a = 0.0f;
	if (a > 65535.0f) a = 65535.0f;
	if (!(b > 0.0f)) b = 0.0f;
	if (b > 65535.0f) b = 65535.0f;
	*lt = (uint16_t)a;
	*gt = (uint16_t)b;
}

void dfo_resolve_roi(const int32_t roi[4], int width, int height, int32_t out[4]) {
	/* kernels.cu:379-384 */
	int off_x = (roi[0] < 0) ? 0 : roi[0];
	int off_y = (roi[1] < 0) ? 0 : roi[1];
	int size_x = (roi[2] < 0 || (off_x + roi[2]) > width) ? width : roi[2];
	int size_y = (roi[3] < 0 || (off_y + roi[3]) > height) ? height : roi[3];
	/* clamp instead of overrunning rows (reference bug, not replicated) */
	if (off_x > width) off_x = width;
	if (off_y > height) off_y = height;
	if (off_x + size_x > width) size_x = width - off_x;
	if (off_y + size_y > height) size_y = height - off_y;
	out[0] = off_x;
	out[1] = off_y;
	out[2] = size_x;
	out[3] = size_y;
}

uint32_t dfo_filter_depth(uint16_t* depth, uint8_t* mask, uint32_t* indices, int width, int height,
                          float min_depth, float max_depth, float depth_scale, const int32_t roi[4]) {
	uint16_t lt, gt;
	int32_t r[4];
	dfo_thresholds(min_depth, max_depth, depth_scale, &lt, &gt);
	dfo_resolve_roi(roi, width, height, r);
	const size_t n = (size_t)width * (size_t)height;
	/* nppiThreshold_LTValGTVal_16u_C1IR(lt -> 0, gt -> 0), whole image, in place (kernels.cu:366-368) */
	for (size_t i = 0; i < n; ++i) {
		const uint16_t d = depth[i];
		if (d < lt || d > gt) depth[i] = 0;
	}
	/* nppiCompareC_16u_C1R(> 0) inside the ROI only (kernels.cu:392-394) */
	for (int y = r[1]; y < r[1] + r[3]; ++y)
		for (int x = r[0]; x < r[0] + r[2]; ++x) {
			const size_t i = (size_t)y * (size_t)width + (size_t)x;
			mask[i] = (depth[i] > 0) ? 255 : 0;
		}
	/* nppiSum_8u_C1R over the whole mask / 255 (kernels.cu:521,532), and compactor::compact writing
	 * the ascending pixel index of every mask > 0 (compactor.cuh:28-68) */
	uint64_t sum = 0;
	uint32_t k = 0;
	for (size_t i = 0; i < n; ++i) {
		sum += mask[i];
		if (mask[i] > 0) indices[k++] = (uint32_t)i;
	}
	return (uint32_t)(sum / 255u);
}

/* ------------------------------------------------------------------------------------------
 * deproject (kernels.cu:70-90, 183-221)
 * ---------------------------------------------------------------------------------------- */

static inline void deproject_xyz(float* p, uint32_t img_idx, uint16_t d, const dfb_intrinsics* intr, float depth_scale) {
	const int v = (int)img_idx / intr->width; /* kernels.cu:193-194 */
	const int u = (int)img_idx - v * intr->width;
	const float z = depth_scale * (float)d;   /* kernels.cu:202 */
	const float x = ((float)u - intr->ppx) / intr->fx; /* kernels.cu:72-73 */
	const float y = ((float)v - intr->ppy) / intr->fy;
	p[0] = z * x;
	p[1] = z * y;
	p[2] = z;
}

void dfo_deproject_color(float* points, const uint16_t* depth, const uint8_t* color, const dfb_intrinsics* intr,
                         const uint32_t* indices, uint32_t count, float depth_scale) {
	for (uint32_t i = 0; i < count; ++i) {
		const uint32_t img_idx = indices[i];
		float* p = points + 4 * (size_t)i;
		deproject_xyz(p, img_idx, depth[img_idx], intr, depth_scale);
		/* kernels.cu:198: bytes {B,G,R,0} reinterpreted as a float */
		const uint8_t b4[4] = {color[(size_t)img_idx * 3 + 2], color[(size_t)img_idx * 3 + 1], color[(size_t)img_idx * 3], 0};
		memcpy(p + 3, b4, 4);
	}
}

void dfo_deproject(float* points, const uint16_t* depth, const dfb_intrinsics* intr, const uint32_t* indices,
                   uint32_t count, float depth_scale) {
	for (uint32_t i = 0; i < count; ++i) {
		const uint32_t img_idx = indices[i];
		float* p = points + 4 * (size_t)i;
		deproject_xyz(p, img_idx, depth[img_idx], intr, depth_scale);
		p[3] = 1.0f; /* kernels.cu:89 */
	}
}

/* ------------------------------------------------------------------------------------------
 * transform (kernels.cu:161-181; contraction per SURVEY App. A.3)
 * ---------------------------------------------------------------------------------------- */

static inline void transform_xyz(float* p, const float* T) {
	const float x = p[0], y = p[1], z = p[2];
	/* nvcc -fmad=true SASS: FMUL T1*y ; FFMA T0*x + . ; FFMA T2*z + . ; FADD + T3 */
	const float nx = fmaf(T[2], z, fmaf(T[0], x, T[1] * y)) + T[3];
	const float ny = fmaf(T[6], z, fmaf(T[4], x, T[5] * y)) + T[7];
	const float nz = fmaf(T[10], z, fmaf(T[8], x, T[9] * y)) + T[11];
	p[0] = nx;
	p[1] = ny;
	p[2] = nz;
}

void dfo_transform(float* points, uint32_t count, const float T[16]) {
	for (uint32_t i = 0; i < count; ++i) transform_xyz(points + 4 * (size_t)i, T);
}

/* ------------------------------------------------------------------------------------------
 * project (kernels.cu:239-274)
 * ---------------------------------------------------------------------------------------- */

int32_t dfo_project_index(const float* p, const dfb_intrinsics* intr, int mirror) {
	const float X = p[0], Y = p[1], Z = p[2];
	if (!(Z > 0.0f)) return -1; /* canonical divergence: reference undefined for Z<=0 (App. A.5) */
	/* kernels.cu:247-248: ppx + fx * X / Z  (left-to-right: (fx*X)/Z; the add cannot contract) */
	const float ix = intr->ppx + (intr->fx * X) / Z;
	const float iy = intr->ppy + (intr->fy * Y) / Z;
	/* kernels.cu:249-250: static_cast<int>(ix + 0.5f) truncates toward zero; bounds :253.
	 * trunc(t) >= 0 <=> t > -1 ; trunc(t) <= W-1 <=> t < W — tested in float so that huge / NaN
	 * values never reach the (undefined on CPU) float->int conversion. */
	const float tx = ix + 0.5f;
	const float ty = iy + 0.5f;
	if (!(tx > -1.0f && tx < (float)intr->width && ty > -1.0f && ty < (float)intr->height)) return -1;
	const int px = (int)tx;
	const int py = (int)ty;
	return py * intr->width + (mirror ? (intr->width - 1 - px) : px); /* kernels.cu:258-261 */
}

void dfo_project_sequential(uint8_t* color_out, float* z_buffer, const float* points, uint32_t count,
                            const dfb_intrinsics* intr, int mirror) {
	const size_t npix = (size_t)intr->width * (size_t)intr->height;
	memset(color_out, 0, npix * 3); /* kernels.cu:582 */
	for (size_t i = 0; i < npix; ++i) z_buffer[i] = FLT_MAX; /* kernels.cu:585-587 */
	for (uint32_t i = 0; i < count; ++i) {
		const float* p = points + 4 * (size_t)i;
		const int32_t idx = dfo_project_index(p, intr, mirror);
		if (idx < 0) continue;
		const float z = p[2];
		/* kernels.cu:264-267 with atomicMinFloat (:118-126) executed one point at a time */
		const float old = z_buffer[idx];
		z_buffer[idx] = fminf(z, old);
		if (old > 0 && old <= z) continue;
		uint8_t bgr[4];
		memcpy(bgr, p + 3, 4);
		color_out[(size_t)idx * 3 + 0] = bgr[2]; /* kernels.cu:270-273 */
		color_out[(size_t)idx * 3 + 1] = bgr[1];
		color_out[(size_t)idx * 3 + 2] = bgr[0];
	}
}

void dfo_project_canonical(uint8_t* color_out, uint64_t* keys, uint8_t* ambiguous, const float* points,
                           const uint32_t* seg_offsets, int n_seg, const dfb_intrinsics* intr, int mirror) {
	const size_t npix = (size_t)intr->width * (size_t)intr->height;
	uint64_t* k = keys ? keys : (uint64_t*)malloc(npix * sizeof(uint64_t));
	if (!k) return;
	for (size_t i = 0; i < npix; ++i) k[i] = UINT64_MAX;
	for (int s = 0; s < n_seg; ++s) {
		for (uint32_t i = seg_offsets[s]; i < seg_offsets[s + 1]; ++i) {
			const float* p = points + 4 * (size_t)i;
			const int32_t idx = dfo_project_index(p, intr, mirror);
			if (idx < 0) continue;
			uint32_t zbits;
			uint8_t bgr[4];
			memcpy(&zbits, p + 2, 4);
			memcpy(bgr, p + 3, 4);
			const uint64_t key = ((uint64_t)zbits << 32) | ((uint64_t)(uint32_t)s << 24) | ((uint64_t)bgr[2] << 16) |
			                     ((uint64_t)bgr[1] << 8) | (uint64_t)bgr[0];
			if (key < k[idx]) k[idx] = key;
		}
	}
	for (size_t i = 0; i < npix; ++i) {
		const uint64_t key = k[i];
		if (key == UINT64_MAX) {
			color_out[i * 3 + 0] = color_out[i * 3 + 1] = color_out[i * 3 + 2] = 0;
		} else {
			color_out[i * 3 + 0] = (uint8_t)(key >> 16);
			color_out[i * 3 + 1] = (uint8_t)(key >> 8);
			color_out[i * 3 + 2] = (uint8_t)key;
		}
	}
	if (ambiguous) {
		/* second pass: a pixel is ambiguous iff another point shares the winner's (Z bits, camera)
		 * with a different colour — the only case where "first in order" != "minimum key" */
		memset(ambiguous, 0, npix);
		for (int s = 0; s < n_seg; ++s)
			for (uint32_t i = seg_offsets[s]; i < seg_offsets[s + 1]; ++i) {
				const float* p = points + 4 * (size_t)i;
				const int32_t idx = dfo_project_index(p, intr, mirror);
				if (idx < 0) continue;
				uint32_t zbits;
				uint8_t bgr[4];
				memcpy(&zbits, p + 2, 4);
				memcpy(bgr, p + 3, 4);
				const uint64_t key = ((uint64_t)zbits << 32) | ((uint64_t)(uint32_t)s << 24) | ((uint64_t)bgr[2] << 16) |
				                     ((uint64_t)bgr[1] << 8) | (uint64_t)bgr[0];
				if ((key >> 24) == (k[idx] >> 24) && key != k[idx]) ambiguous[idx] = 1;
			}
	}
	if (!keys) free(k);
}

/* ------------------------------------------------------------------------------------------
 * whole-frameset sequence (fusion_node.cpp:718-784)
 * ---------------------------------------------------------------------------------------- */

int64_t dfo_fuse_frameset(const dfb_fuse_desc* desc, const uint16_t* const* depth, const uint8_t* const* color,
                          uint8_t* fused_out, int mode, float* cloud_out, uint32_t* counts_out, uint8_t* ambiguous) {
	const int nc = desc->n_cameras;
	size_t cap = 0, max_pix = 0;
	for (int c = 0; c < nc; ++c) {
		const size_t n = (size_t)desc->cam[c].intr.width * (size_t)desc->cam[c].intr.height;
		cap += n;
		if (n > max_pix) max_pix = n;
	}
	const size_t out_pix = (size_t)desc->out_intr.width * (size_t)desc->out_intr.height;
	uint16_t* d_scratch = (uint16_t*)malloc(max_pix * sizeof(uint16_t));
	uint8_t* mask = (uint8_t*)malloc(max_pix);
	uint32_t* indices = (uint32_t*)malloc(max_pix * sizeof(uint32_t));
	float* cloud = (float*)malloc(cap * 4 * sizeof(float));
	float* zbuf = (mode == 0) ? (float*)malloc(out_pix * sizeof(float)) : NULL;
	uint32_t seg[DFB_MAX_CAMERAS + 1];
	int64_t total = -1;
	if (!d_scratch || !mask || !indices || !cloud || (mode == 0 && !zbuf)) goto done;

	seg[0] = 0;
	for (int c = 0; c < nc; ++c) {
		const dfb_camera_desc* cam = &desc->cam[c];
		const size_t n = (size_t)cam->intr.width * (size_t)cam->intr.height;
		memcpy(d_scratch, depth[c], n * sizeof(uint16_t)); /* setDepthFrame, fusion_node.cpp:727,735 */
		memset(mask, 0, n);                                  /* fresh DepthFrame, depth_frame.cpp:46 */
		/* filterDepth, fusion_node.cpp:739-741 */
		const uint32_t cnt = dfo_filter_depth(d_scratch, mask, indices, cam->intr.width, cam->intr.height,
		                                      cam->min_depth, cam->max_depth, cam->depth_scale, cam->roi);
		/* deproject into the fused cloud at the running offset: left at [0,nL), right appended
		 * (fusion_node.cpp:746-752,763; pointcloud.cpp:100-129) */
		float* seg_pts = cloud + 4 * (size_t)seg[c];
		dfo_deproject_color(seg_pts, d_scratch, color[c], &cam->intr, indices, cnt, cam->depth_scale);
		/* cloud_right.transform(m_right_transform), fusion_node.cpp:757 */
		if (cam->has_transform) dfo_transform(seg_pts, cnt, cam->to_ref);
		seg[c + 1] = seg[c] + cnt;
		if (counts_out) counts_out[c] = cnt;
	}
	total = seg[nc];
	/* fused_cloud.transform(fused_transform), fusion_node.cpp:779 */
	dfo_transform(cloud, (uint32_t)total, desc->fused_transform);
	/* fused_cloud.project(m_fused_frameset, m_mirror_image), fusion_node.cpp:784 */
	if (mode == 0) {
		dfo_project_sequential(fused_out, zbuf, cloud, (uint32_t)total, &desc->out_intr, desc->mirror);
		if (ambiguous) {
			uint8_t* tmp = (uint8_t*)malloc(out_pix * 3);
			if (tmp) {
				dfo_project_canonical(tmp, NULL, ambiguous, cloud, seg, nc, &desc->out_intr, desc->mirror);
				free(tmp);
			}
		}
	} else {
		dfo_project_canonical(fused_out, NULL, ambiguous, cloud, seg, nc, &desc->out_intr, desc->mirror);
	}
	if (cloud_out) memcpy(cloud_out, cloud, (size_t)total * 4 * sizeof(float));
done:
	free(d_scratch);
	free(mask);
	free(indices);
	free(cloud);
	free(zbuf);
	return total;
}

int dfo_max_threads(void) {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

int dfo_fuse_batch(const dfb_fuse_desc* desc, const uint16_t* depth, const uint8_t* color, uint8_t* out,
                   int n_framesets, int mode, int n_threads) {
	const int nc = desc->n_cameras;
	const size_t pix = (size_t)desc->cam[0].intr.width * (size_t)desc->cam[0].intr.height;
	const size_t out_bytes = (size_t)desc->out_intr.width * (size_t)desc->out_intr.height * 3;
	for (int c = 1; c < nc; ++c)
		if ((size_t)desc->cam[c].intr.width * (size_t)desc->cam[c].intr.height != pix) return -2;
	int failed = 0;
	if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
	omp_set_dynamic(0); /* the caller's count is the count: neither OMP_NUM_THREADS nor OMP_DYNAMIC may shrink the team */
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
	for (int f = 0; f < n_framesets; ++f) {
		const uint16_t* dptr[DFB_MAX_CAMERAS];
		const uint8_t* cptr[DFB_MAX_CAMERAS];
		for (int c = 0; c < nc; ++c) {
			dptr[c] = depth + ((size_t)f * nc + c) * pix;
			cptr[c] = color + ((size_t)f * nc + c) * pix * 3;
		}
		if (dfo_fuse_frameset(desc, dptr, cptr, out + (size_t)f * out_bytes, mode, NULL, NULL, NULL) < 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
			failed = 1;
		}
	}
	return failed ? -1 : 0;
}

uint32_t dfo_registration_cloud(const dfb_camera_desc* cam, const uint16_t* depth, float* cloud_out) {
	const size_t n = (size_t)cam->intr.width * (size_t)cam->intr.height;
	uint16_t* d_scratch = (uint16_t*)malloc(n * sizeof(uint16_t));
	uint8_t* mask = (uint8_t*)calloc(n, 1);
	uint32_t* indices = (uint32_t*)malloc(n * sizeof(uint32_t));
	uint32_t cnt = 0;
	if (d_scratch && mask && indices) {
		memcpy(d_scratch, depth, n * sizeof(uint16_t)); /* registration_node.cpp:315,317 */
		cnt = dfo_filter_depth(d_scratch, mask, indices, cam->intr.width, cam->intr.height, cam->min_depth,
		                       cam->max_depth, cam->depth_scale, cam->roi); /* :316,318 */
		dfo_deproject(cloud_out, d_scratch, &cam->intr, indices, cnt, cam->depth_scale); /* :322,324 */
	}
	free(d_scratch);
	free(mask);
	free(indices);
	return cnt;
}

/* ------------------------------------------------------------------------------------------
 * post-splat image filters (kernels.cu:594-653) — NPP semantics, parity unpinned in the reference
 * ---------------------------------------------------------------------------------------- */

void dfo_filter_color_gauss(uint8_t* color, int width, int height) {
	if (width < 3 || height < 3) return;
	const size_t bytes = (size_t)width * (size_t)height * 3;
	uint8_t* src = (uint8_t*)malloc(bytes);
	if (!src) return;
	memcpy(src, color, bytes); /* D2D copy into the scratch buffer, kernels.cu:644-647 */
	/* NPP_MASK_SIZE_3_X_3 Gaussian: sigma = 0.4 + (3/2)*0.6 = 1.0 (comment at kernels.cu:649);
	 * separable taps g = exp(-d^2/2)/sum */
	const float e = expf(-0.5f);
	const float s = 1.0f + 2.0f * e;
	const float g0 = 1.0f / s, g1 = e / s;
	const float w[3] = {g1, g0, g1};
	for (int y = 1; y < height - 1; ++y)
		for (int x = 1; x < width - 1; ++x)
			for (int c = 0; c < 3; ++c) {
				float acc = 0.0f;
				for (int dy = -1; dy <= 1; ++dy)
					for (int dx = -1; dx <= 1; ++dx)
						acc += w[dy + 1] * w[dx + 1] * (float)src[((size_t)(y + dy) * width + (x + dx)) * 3 + c];
				int v = (int)acc; /* NPP truncates (pinned against NPP 12.4 on a B200) */
				if (v > 255) v = 255;
				color[((size_t)y * width + x) * 3 + c] = (uint8_t)v;
			}
	free(src);
}

void dfo_filter_color_median(uint8_t* color, int width, int height) {
	if (width < 3 || height < 3) return;
	const size_t bytes = (size_t)width * (size_t)height * 3;
	uint8_t* src = (uint8_t*)malloc(bytes);
	if (!src) return;
	memcpy(src, color, bytes); /* kernels.cu:616-620 */
	for (int y = 1; y < height - 1; ++y)
		for (int x = 1; x < width - 1; ++x)
			for (int c = 0; c < 3; ++c) {
				uint8_t v[9];
				int k = 0;
				for (int dy = -1; dy <= 1; ++dy)
					for (int dx = -1; dx <= 1; ++dx) v[k++] = src[((size_t)(y + dy) * width + (x + dx)) * 3 + c];
				for (int i = 1; i < 9; ++i) { /* insertion sort */
					const uint8_t t = v[i];
					int j = i - 1;
					while (j >= 0 && v[j] > t) {
						v[j + 1] = v[j];
						--j;
					}
					v[j + 1] = t;
				}
				/* Reference quirk, replicated because it is the observable output: filter_color_median_npp
				 * passes the UN-offset destination pointer (kernels.cu:625: `dev_color_npp`, while the source is
				 * `dev_color_buffer_npp + roi_offset`), so the median of the window centred on (y, x) lands
				 * on (y-1, x-1): the filtered image is shifted up-left by one pixel and the last two rows
				 * and columns keep the splat output. Pinned against the reference wrapper + NPP 12.4 on a B200. */
				color[((size_t)(y - 1) * width + (x - 1)) * 3 + c] = v[4];
			}
	free(src);
}

/* ------------------------------------------------------------------------------------------
 * host-side transform glue (fusion_node.cpp:155-165, 589-604, 774-777; pointcloud.cpp:65)
 * ---------------------------------------------------------------------------------------- */

static void quat_from_mat3(const double* T, double q[4] /* x y z w */) {
	/* Eigen 3.4 quaternionbase_assign_impl<Matrix3,3,3>::run (Shoemake) on the linear part */
#define M(r, c) T[4 * (r) + (c)]
	double t = M(0, 0) + M(1, 1) + M(2, 2);
	if (t > 0.0) {
		t = sqrt(t + 1.0);
		q[3] = 0.5 * t;
		t = 0.5 / t;
		q[0] = (M(2, 1) - M(1, 2)) * t;
		q[1] = (M(0, 2) - M(2, 0)) * t;
		q[2] = (M(1, 0) - M(0, 1)) * t;
	} else {
		int i = 0;
		if (M(1, 1) > M(0, 0)) i = 1;
		if (M(2, 2) > M(i, i)) i = 2;
		const int j = (i + 1) % 3;
		const int k = (j + 1) % 3;
		t = sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
		q[i] = 0.5 * t;
		t = 0.5 / t;
		q[3] = (M(k, j) - M(j, k)) * t;
		q[j] = (M(j, i) + M(i, j)) * t;
		q[k] = (M(k, i) + M(i, k)) * t;
	}
#undef M
}

static void quat_slerp(const double a[4], const double b[4], double t, double out[4]) {
	/* Eigen 3.4 QuaternionBase::slerp */
	const double one = 1.0 - DBL_EPSILON;
	const double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
	const double absD = fabs(d);
	double scale0, scale1;
	if (absD >= one) {
		scale0 = 1.0 - t;
		scale1 = t;
	} else {
		const double theta = acos(absD);
		const double sinTheta = sin(theta);
		scale0 = sin((1.0 - t) * theta) / sinTheta;
		scale1 = sin(t * theta) / sinTheta;
	}
	if (d < 0.0) scale1 = -scale1;
	for (int i = 0; i < 4; ++i) out[i] = scale0 * a[i] + scale1 * b[i];
}

static void quat_to_mat3(const double q[4], double R[9]) {
	/* Eigen 3.4 QuaternionBase::toRotationMatrix */
	const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
	const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
	const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
	const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
	R[0] = 1.0 - (tyy + tzz);
	R[1] = txy - twz;
	R[2] = txz + twy;
	R[3] = txy + twz;
	R[4] = 1.0 - (txx + tzz);
	R[5] = tyz - twx;
	R[6] = txz - twy;
	R[7] = tyz + twx;
	R[8] = 1.0 - (txx + tyy);
}

void dfo_interpolate_transform(const double left[16], const double right[16], double out[16]) {
	double ql[4], qr[4], qf[4], R[9];
	quat_from_mat3(left, ql);  /* fusion_node.cpp:593 (rotation() of a rigid transform == linear part) */
	quat_from_mat3(right, qr); /* :594 */
	quat_slerp(ql, qr, 0.5, qf); /* :596 */
	quat_to_mat3(qf, R);
	double t[3];
	for (int i = 0; i < 3; ++i) t[i] = left[4 * i + 3] * 0.5 + right[4 * i + 3] * 0.5; /* :600 */
	/* Identity.prerotate(q).pretranslate(t), :601-602 */
	for (int r = 0; r < 3; ++r) {
		for (int c = 0; c < 3; ++c) out[4 * r + c] = R[3 * r + c];
		out[4 * r + 3] = t[r];
	}
	out[12] = out[13] = out[14] = 0.0;
	out[15] = 1.0;
	if (qr[3] < 0.0) { /* :603 — Affine inverse: linear^-1 by cofactors, translation = -linear^-1 * t */
		double inv[9];
		const double a = R[0], b = R[1], c = R[2], d = R[3], e = R[4], f = R[5], g = R[6], h = R[7], i = R[8];
		const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
		const double id = 1.0 / det;
		inv[0] = (e * i - f * h) * id;
		inv[1] = (c * h - b * i) * id;
		inv[2] = (b * f - c * e) * id;
		inv[3] = (f * g - d * i) * id;
		inv[4] = (a * i - c * g) * id;
		inv[5] = (c * d - a * f) * id;
		inv[6] = (d * h - e * g) * id;
		inv[7] = (b * g - a * h) * id;
		inv[8] = (a * e - b * d) * id;
		for (int r = 0; r < 3; ++r) {
			for (int cc = 0; cc < 3; ++cc) out[4 * r + cc] = inv[3 * r + cc];
			out[4 * r + 3] = -(inv[3 * r + 0] * t[0] + inv[3 * r + 1] * t[1] + inv[3 * r + 2] * t[2]);
		}
	}
}

void dfo_prerotate_z90(double T[16]) {
	/* AngleAxisd(deg2rad(90.), UnitZ).toRotationMatrix(): [[c,-s,0],[s,c,0],[0,0,1]] applied to the
	 * affine 3x4 block (linear AND translation), fusion_node.cpp:776 */
	const double ang = 90.0 * 3.14159265358979323846 / 180.0;
	const double c = cos(ang), s = sin(ang);
	for (int col = 0; col < 4; ++col) {
		const double r0 = T[col], r1 = T[4 + col];
		T[col] = c * r0 - s * r1;
		T[4 + col] = s * r0 + c * r1;
	}
}

void dfo_affine_to_float16(const double T[16], float out[16]) {
	for (int i = 0; i < 16; ++i) out[i] = (float)T[i]; /* .cast<float>(), pointcloud.cpp:65 */
}

void dfo_fused_intrinsics(const dfb_intrinsics* color_left, int vertical, dfb_intrinsics* out) {
	*out = *color_left; /* fusion_node.cpp:155 */
	if (vertical) {     /* :156-162 */
		out->width = color_left->height;
		out->height = color_left->width;
		out->fx = color_left->fy;
		out->fy = color_left->fx;
	}
	out->ppx = (float)(out->width / 2); /* integer division, :164-165 */
	out->ppy = (float)(out->height / 2);
}

/* ------------------------------------------------------------------------------------------
 * depth -> colour alignment (Frameset::alignDepthToColor, frameset.cpp:180-188 ->
 * Kernels::align_depth_to_color, kernels.cu:714-747: kernel_map_depth_to_color :292-303 with
 * kernel_transfer_pixels :137-159, kernel_align_depth_to_color :276-300, kernel_replace_to_zero :305-311).
 * Every multiply-add below is the contraction nvcc 12.9 chose for the reference, read from the SASS of the
 * compiled reference kernel (oracle/_ref), and pinned against that kernel on the B200.
 * ---------------------------------------------------------------------------------------- */

/* cvt.rzi.s32.f32: saturating, NaN -> 0 */
static inline int32_t f2i_trunc_sat(float v) {
	if (v != v) return 0;
	if (v >= 2147483648.0f) return INT32_MAX;
	if (v <= -2147483648.0f) return INT32_MIN;
	return (int32_t)v;
}

/* the two corners a depth pixel maps to; returns 0 when the depth is zero (mapped = {-1,-1}) */
static int map_depth_pixel(const dfb_intrinsics* di, const dfb_intrinsics* ci, const dfb_extrinsics* ex, float depth_val, int dx,
                           int dy, int corner, int32_t out_xy[2]) {
	if (depth_val == 0.0f) {
		out_xy[0] = out_xy[1] = -1;
		return 0;
	}
	const float shift = corner ? 0.5f : -0.5f;
	/* deproject_pixel_to_point, kernels.cu:46-68 */
	float x = ((float)dx + shift - di->ppx) / di->fx;
	float y = ((float)dy + shift - di->ppy) / di->fy;
	if (di->model == DFB_DISTORTION_INVERSE_BROWN_CONRADY) {
		const float* c = di->coeffs;
		const float r2 = fmaf(x, x, y * y);
		float f = fmaf(r2, c[0], 1.0f);
		f = fmaf(r2, r2 * c[1], f);
		f = fmaf(r2, r2 * (r2 * c[4]), f);
		const float a = x * (c[3] + c[3]), b = x * (c[2] + c[2]);
		const float gx = fmaf(x, x + x, r2), gy = fmaf(y, y + y, r2);
		const float ux = fmaf(c[3], gx, fmaf(y, b, x * f));
		const float uy = fmaf(c[2], gy, fmaf(y, f, y * a));
		x = ux;
		y = uy;
	}
	const float p0 = depth_val * x, p1 = depth_val * y, p2 = depth_val;
	/* transform_point_to_point, kernels.cu:128-136 (column-major rotation) */
	const float* R = ex->rotation;
	const float* t = ex->translation;
	const float q0 = fmaf(p2, R[6], fmaf(p0, R[0], p1 * R[3])) + t[0];
	const float q1 = fmaf(p2, R[7], fmaf(p0, R[1], p1 * R[4])) + t[1];
	const float q2 = fmaf(p2, R[8], fmaf(p0, R[2], p1 * R[5])) + t[2];
	/* project_point_to_pixel, kernels.cu:92-116 */
	float px = q0 / q2, py = q1 / q2;
	if (ci->model == DFB_DISTORTION_MODIFIED_BROWN_CONRADY) {
		const float* c = ci->coeffs;
		const float r2 = fmaf(px, px, py * py);
		float f = fmaf(r2, c[0], 1.0f);
		f = fmaf(r2, r2 * c[1], f);
		f = fmaf(r2, r2 * (r2 * c[4]), f);
		const float xf = px * f, yf = py * f;
		const float ddx = fmaf(c[3], fmaf(xf, xf + xf, r2), fmaf(yf, xf * (c[2] + c[2]), xf));
		const float ddy = fmaf(c[2], fmaf(yf, yf + yf, r2), fmaf(yf, xf * (c[3] + c[3]), yf));
		px = ddx;
		py = ddy;
	}
	out_xy[0] = f2i_trunc_sat(fmaf(ci->fx, px, ci->ppx) + 0.5f);
	out_xy[1] = f2i_trunc_sat(fmaf(ci->fy, py, ci->ppy) + 0.5f);
	return 1;
}

int dfo_align_depth_to_color(const uint16_t* depth, const dfb_intrinsics* di, const dfb_intrinsics* ci, const dfb_extrinsics* ex,
                             float depth_scale, uint16_t* aligned) {
	/* the reference asserts these away (kernels.cu:48-49) or needs CUDA's tanf/atanf (FTHETA, :107-112) */
	if (di->model == DFB_DISTORTION_MODIFIED_BROWN_CONRADY || di->model == DFB_DISTORTION_FTHETA || ci->model == DFB_DISTORTION_FTHETA) return -1;
	const size_t n_color = (size_t)ci->width * (size_t)ci->height;
	if (n_color & 1u) return -1; /* the reference updates pixel PAIRS through 32-bit atomics */
	uint32_t* words = (uint32_t*)aligned;
	memset(aligned, 0xff, n_color * 2);
	for (int dy = 0; dy < di->height; ++dy)
		for (int dx = 0; dx < di->width; ++dx) {
			const uint16_t d = depth[(size_t)dy * di->width + dx];
			const float depth_val = (float)d * depth_scale;
			int32_t p0[2], p1[2];
			map_depth_pixel(di, ci, ex, depth_val, dx, dy, 0, p0);
			map_depth_pixel(di, ci, ex, depth_val, dx, dy, 1, p1);
			if (p0[0] < 0 || p0[1] < 0 || p1[0] >= ci->width || p1[1] >= ci->height) continue;
			/* kernels.cu:293-299: the value is widened to d<<16|d and min-ed into the 32-bit word that holds the
			 * target pixel AND its horizontal neighbour (a reference quirk that is kept: both pixels of a pair
			 * end up with the smallest depth that touched either of them) */
			const uint32_t v = ((uint32_t)d << 16) | d;
			for (int yy = p0[1]; yy <= p1[1]; ++yy)
				for (int xx = p0[0]; xx <= p1[0]; ++xx) {
					const size_t w = ((size_t)yy * ci->width + xx) / 2;
					if (v < words[w]) words[w] = v;
				}
		}
	for (size_t i = 0; i < n_color; ++i)
		if (aligned[i] == 0xffff) aligned[i] = 0;
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * voxel downsample (registration.cpp:144-150) — parity unpinned
 * ---------------------------------------------------------------------------------------- */

typedef struct {
	int ix, iy, iz, count;
	float c[3];
} dfo_he;

uint32_t dfo_voxel_approximate(const float* in, uint32_t n, float leaf, float* out) {
	enum { HIST = 512 };
	dfo_he hist[HIST];
	memset(hist, 0, sizeof(hist));
	const float inv = 1.0f / leaf;
	uint32_t op = 0;
	for (uint32_t i = 0; i < n; ++i) {
		const float* p = in + 4 * (size_t)i;
		const int ix = (int)floorf(p[0] * inv);
		const int iy = (int)floorf(p[1] * inv);
		const int iz = (int)floorf(p[2] * inv);
		const unsigned hash = (unsigned)(ix * 7171 + iy * 3079 + iz * 4231) & (HIST - 1);
		dfo_he* h = &hist[hash];
		if (h->count && (ix != h->ix || iy != h->iy || iz != h->iz)) {
			float* o = out + 4 * (size_t)op++;
			const float cnt = (float)h->count;
			o[0] = h->c[0] / cnt;
			o[1] = h->c[1] / cnt;
			o[2] = h->c[2] / cnt;
			o[3] = 1.0f;
			h->count = 0;
			h->c[0] = h->c[1] = h->c[2] = 0.0f;
		}
		h->ix = ix;
		h->iy = iy;
		h->iz = iz;
		h->count++;
		h->c[0] += p[0];
		h->c[1] += p[1];
		h->c[2] += p[2];
	}
	for (int b = 0; b < HIST; ++b) {
		dfo_he* h = &hist[b];
		if (h->count) {
			float* o = out + 4 * (size_t)op++;
			const float cnt = (float)h->count;
			o[0] = h->c[0] / cnt;
			o[1] = h->c[1] / cnt;
			o[2] = h->c[2] / cnt;
			o[3] = 1.0f;
		}
	}
	return op;
}

typedef struct {
	int32_t iz, iy, ix;
	uint32_t src;
} dfo_vkey;

static int vkey_cmp(const void* a, const void* b) {
	const dfo_vkey* x = (const dfo_vkey*)a;
	const dfo_vkey* y = (const dfo_vkey*)b;
	if (x->iz != y->iz) return x->iz < y->iz ? -1 : 1;
	if (x->iy != y->iy) return x->iy < y->iy ? -1 : 1;
	if (x->ix != y->ix) return x->ix < y->ix ? -1 : 1;
	return x->src < y->src ? -1 : (x->src > y->src);
}

uint32_t dfo_voxel_exact(const float* in, uint32_t n, float leaf, float* out, int32_t* voxel_ids_out) {
	if (n == 0) return 0;
	dfo_vkey* keys = (dfo_vkey*)malloc((size_t)n * sizeof(dfo_vkey));
	if (!keys) return 0;
	const float inv = 1.0f / leaf;
	for (uint32_t i = 0; i < n; ++i) {
		const float* p = in + 4 * (size_t)i;
		keys[i].ix = (int)floorf(p[0] * inv);
		keys[i].iy = (int)floorf(p[1] * inv);
		keys[i].iz = (int)floorf(p[2] * inv);
		keys[i].src = i;
	}
	qsort(keys, n, sizeof(dfo_vkey), vkey_cmp);
	uint32_t nv = 0;
	uint32_t i = 0;
	while (i < n) {
		uint32_t j = i;
		double sx = 0, sy = 0, sz = 0;
		while (j < n && keys[j].ix == keys[i].ix && keys[j].iy == keys[i].iy && keys[j].iz == keys[i].iz) {
			const float* p = in + 4 * (size_t)keys[j].src;
			sx += p[0];
			sy += p[1];
			sz += p[2];
			++j;
		}
		const double cnt = (double)(j - i);
		float* o = out + 4 * (size_t)nv;
		o[0] = (float)(sx / cnt);
		o[1] = (float)(sy / cnt);
		o[2] = (float)(sz / cnt);
		o[3] = 1.0f;
		if (voxel_ids_out) {
			voxel_ids_out[3 * (size_t)nv + 0] = keys[i].ix;
			voxel_ids_out[3 * (size_t)nv + 1] = keys[i].iy;
			voxel_ids_out[3 * (size_t)nv + 2] = keys[i].iz;
		}
		++nv;
		i = j;
	}
	free(keys);
	return nv;
}

/* ------------------------------------------------------------------------------------------
 * synthetic inputs
 * ---------------------------------------------------------------------------------------- */

void dfo_synth_camera(const dfs_scene* scene, const dfs_camera* cam, uint32_t cam_index, uint16_t* depth,
                      uint8_t* color) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
	for (int v = 0; v < cam->height; ++v)
		for (int u = 0; u < cam->width; ++u) {
			const size_t i = (size_t)v * (size_t)cam->width + (size_t)u;
			dfs_pixel(scene, cam, cam_index, u, v, depth + i, color + 3 * i);
		}
}
