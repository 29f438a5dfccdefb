/*
 * splat_core.cuh — the per-pixel-pair body shared by the fast splat kernels (fuse_column.cu: direct global
 * loads; fuse_tiled.cu: TMA + shared-memory tiles). Same arithmetic as fuse_direct.cu / dfb_math.cuh, bit
 * for bit, restructured around what a B200 measured as the limits of this path:
 *
 *  - Packed fp32. Blackwell's FFMA2 / FMUL2 / FADD2 (`__ffma2_rn` etc.: IEEE round-to-nearest per component,
 *    scalar constants broadcast from uniform registers) process TWO pixels per instruction with identical
 *    results, so a thread walks its 16-pixel row segment in pixel PAIRS.
 *  - Divisions. x = (u-ppx)/fx depends on the column only and comes from a table built once per plan with
 *    the same IEEE division; the two projection divisions share one reciprocal (dfb_math.cuh divide_pair,
 *    packed here). Operands outside the guarded fast-path range are not approximated: the thread re-does
 *    that pixel through the scalar __fdiv_rn path (splat_pixel_slow).
 *  - Atomics. One RED.MIN.64 per valid in-bounds point; key address computed in 32-bit byte offsets.
 */
#pragma once
#include <cstdlib>
#include <utility>

#include "dfb_internal.h"

namespace dfb {
namespace {

constexpr int kSeg = 16; /* pixels per thread: 2 x 16 B of depth + 3 x 16 B of colour */

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
	/* read-once data: non-coherent path, do not allocate in L1 */
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
	             : "l"(p));
	return r;
}

__device__ __forceinline__ float2 bcast(float a) { return make_float2(a, a); }

/* transform_point (dfb_math.cuh) on two points at once: x' = fl(fma(T2,z, fma(T0,x, fl(T1*y))) + T3) */
__device__ __forceinline__ void transform_pair(float2& x, float2& y, float2& z, const Mat34& T) {
	const float2 nx = __fadd2_rn(__ffma2_rn(bcast(T.m[2]), z, __ffma2_rn(bcast(T.m[0]), x, __fmul2_rn(bcast(T.m[1]), y))), bcast(T.m[3]));
	const float2 ny = __fadd2_rn(__ffma2_rn(bcast(T.m[6]), z, __ffma2_rn(bcast(T.m[4]), x, __fmul2_rn(bcast(T.m[5]), y))), bcast(T.m[7]));
	const float2 nz = __fadd2_rn(__ffma2_rn(bcast(T.m[10]), z, __ffma2_rn(bcast(T.m[8]), x, __fmul2_rn(bcast(T.m[9]), y))), bcast(T.m[11]));
	x = nx;
	y = ny;
	z = nz;
}

/* bytes [3k, 3k+3) of a 48-byte run held in 12 words -> [B, G, R, cam] (cam id in byte 0 of `cam`) */
template <int kPix>
__device__ __forceinline__ uint32_t rgb_of(const uint32_t (&w)[12], uint32_t cam) {
	constexpr int b = kPix * 3;
	constexpr int wi = b >> 2, sh = b & 3, wn = (wi + 1 < 12) ? wi + 1 : wi;
	if (sh == 0) return __byte_perm(w[wi], cam, 0x4012);
	if (sh == 1) return __byte_perm(w[wi], cam, 0x4123);
	if (sh == 2) return __byte_perm(__byte_perm(w[wi], w[wn], 0x2234), cam, 0x4210);
	return __byte_perm(__byte_perm(w[wi], w[wn], 0x3345), cam, 0x4210);
}

/* scalar, fully general version of one pixel (identical to fuse_direct.cu splat_one + filters); used for
 * the rare pixels whose division operands fall outside the packed fast path */
__device__ __noinline__ void splat_pixel_slow(const FuseParams& P, int cam, const uint16_t* __restrict__ depth_img,
                                              const uint8_t* __restrict__ color_img, int u, int v,
                                              unsigned long long* __restrict__ keys, uint32_t key_bias = 0) {
	const CamParams& C = P.cam[cam];
	/* bounds and ROI BEFORE touching memory: the partner column of a split pair (u + 32) may lie beyond the image */
	if (!((unsigned)u < (unsigned)C.width && (unsigned)v < (unsigned)C.height)) return;
	if (!((unsigned)(u - C.roi_x0) < (unsigned)C.roi_w && (unsigned)(v - C.roi_y0) < (unsigned)C.roi_h)) return;
	const size_t i = (size_t)v * C.width + u;
	const uint32_t d = depth_img[i];
	if (!(d >= C.lt && d <= C.gt && d > 0u)) return;
	const float z = __fmul_rn(C.scale, (float)d);
	float X = __fmul_rn(z, deproject_ray(u, C.ppx, C.fx));
	float Y = __fmul_rn(z, deproject_ray(v, C.ppy, C.fy));
	float Z = z;
	if (C.has_transform) transform_point(X, Y, Z, C.to_ref);
	transform_point(X, Y, Z, P.fused);
	int idx;
	if (!project_point(X, Y, Z, P.proj, idx)) return;
	const uint32_t rgb = ((uint32_t)color_img[i * 3] << 16) | ((uint32_t)color_img[i * 3 + 1] << 8) | color_img[i * 3 + 2];
	atomicMin(keys + idx, make_key_biased(Z, ((uint32_t)cam << 24) | rgb, key_bias));
}

/* What the hot loop reads, split by lifetime so that the compiler can keep it in uniform registers:
 *   SplatShared — the same for every camera of a plan (fused transform, projection, key addressing): always at
 *                 a fixed constant-bank address;
 *   SplatCam    — per camera: a fixed address when a launch handles one camera (fuse_column / fuse_tiled), a
 *                 warp-uniform index into an array in the stream kernel. */
struct SplatShared {
	Mat34 Tf;               /* fused transform, fusion_node.cpp:779 */
	float fx, fy, ppx, ppy; /* fused-image projection */
	int out_w, out_h;
	uint32_t row_bytes;     /* out_w * 8: key bytes per output row */
	int32_t px_step;        /* +8, or -8 when mirrored */
	uint32_t px_origin;     /* 0, or (out_w - 1) * 8 when mirrored */
	int n_cameras;          /* images per frameset in the pools */
	uint32_t n_out_pixels;
	int width, height;      /* camera image (all cameras of a plan share one resolution) */
	int x_pitch, y_base, y_pitch; /* ray tables: x rows at cam*x_pitch, y rows at y_base + cam*y_pitch (floats) */
	int zero;               /* always 0: see the tail of splat_pair */
};

struct SplatCam {
	Mat34 Tc;                /* camera -> reference, fusion_node.cpp:757 */
	float scale, mid, halfw; /* depth scale; float form of the depth range test */
	int roi_x0, roi_y0, roi_w, roi_h;
	uint32_t cam;            /* camera id (tie-break byte of the key) */
	int x_off, y_off;        /* offsets of this camera's rows in the ray tables */
	int has_t, roi_all;      /* run-time copies of the two template switches (kHasT / kRoiAll == 2) */
};

/* NOTE on the packed intrinsics: unlike __fmul_rn / __fadd_rn, a __fadd2_rn fed by a __fmul2_rn IS contracted into FFMA2 by
 * nvcc 12.9 (seen in tools/pipe_probe.cu). The body below never adds to a bare product where the reference does not use an
 * FMA itself, and the parity tests would catch it; keep it that way. */

/* one pixel pair (u, u+1) of row v. Returns true when the pair must be redone through the scalar path.
 * kHasT / kRoiAll: 0 or 1 = compile-time, 2 = read K.has_t / K.roi_all (a warp-uniform branch). */
/* Where the keys of the current tile go. key address = base + origin + py*row_bytes + px*px_step.
 * dump != 0: BRANCHLESS mode — a point that is invalid or falls outside the image still issues its RED, but
 * into a scratch area at base + dump (a per-lane address nobody reads). ptxas cannot predicate a 64-bit RED
 * and wraps each one in a branch + convergence barrier, which cuts a thread's 8 independent pixel pairs into
 * 16 tiny basic blocks that the scheduler cannot overlap; with unconditional REDs the whole segment is one
 * block. The extra L2 atomics are cheap here (measured: removing ALL atomics gains ~5 %). */
struct KeyTarget {
	char* base;
	uint32_t origin; /* slot offset + (mirrored ? (out_w-1)*8 : 0) */
	uint32_t dump;   /* 0 = predicated mode */
	uint32_t bias;   /* added to the key's high word: the ring epoch of the stream kernel (dfb_math.cuh), else 0 */
};

/* one pixel pair (u, u + du) of row v; sets `bit` in `slow` when the pair must be redone through the scalar path.
 * kHasT / kRoiAll: 0 or 1 = compile-time, 2 = read K.has_t / K.roi_all (a warp-uniform branch). */
template <int kHasT, int kRoiAll, int kDump = 0>
__device__ __forceinline__ void splat_pair(const SplatShared& G, const SplatCam& K, uint32_t dpair, float2 xr, float yr, uint32_t key_lo0,
                                           uint32_t key_lo1, int u, bool row_in_roi, const KeyTarget& T, uint32_t& slow, uint32_t bit,
                                           int du = 1) {
	/* DepthFrame::filter folded in (kernels.cu:357-368,392): lt <= d <= gt and d > 0, in exact float form
	 * (|float(d) - mid| <= halfw, tested below) */
	const float2 fd = make_float2((float)(dpair & 0xFFFFu), (float)(dpair >> 16));
	const float2 m = __fadd2_rn(fd, bcast(-K.mid));
	/* kernels.cu:202,75-77: z = scale*d, X = z*x, Y = z*y */
	float2 Z = __fmul2_rn(bcast(K.scale), fd);
	float2 X = __fmul2_rn(Z, xr);
	float2 Y = __fmul2_rn(Z, bcast(yr));
	if (kHasT == 1 || (kHasT == 2 && K.has_t)) transform_pair(X, Y, Z, K.Tc); /* fusion_node.cpp:757 */
	transform_pair(X, Y, Z, G.Tf);            /* fusion_node.cpp:779 */
	/* kernels.cu:247-248: ix = ppx + (fx*X)/Z — divide_pair's fast path, two pixels wide. Guard: Z and both
	 * |numerators| of both pixels >= 2^-40 (upper ends are bounded by the plan's input validation), one
	 * 6-way minimum. */
	const float2 ax = __fmul2_rn(bcast(G.fx), X);
	const float2 ay = __fmul2_rn(bcast(G.fy), Y);
	const float g0 = fminf(fminf(Z.x, Z.y), fabsf(ax.x));
	const float g1 = fminf(fminf(fabsf(ay.x), fabsf(ax.y)), fabsf(ay.y));
	const float g = fminf(g0, g1);
	float2 r = make_float2(rcp_seed(Z.x), rcp_seed(Z.y));
	const float2 nZ = make_float2(-Z.x, -Z.y);
	const float2 e = __ffma2_rn(nZ, r, bcast(1.0f));
	r = __ffma2_rn(r, e, r);
	float2 q = __fmul2_rn(ax, r);
	float2 t = __ffma2_rn(nZ, q, ax);
	const float2 qx = __ffma2_rn(r, t, q);
	q = __fmul2_rn(ay, r);
	t = __ffma2_rn(nZ, q, ay);
	const float2 qy = __ffma2_rn(r, t, q);
	/* kernels.cu:249-261: px = trunc(ix + 0.5f), 0 <= px < W, idx = py*W + (mirror ? W-1-px : px). Every value is
	 * finite here (plan creation bounds the inputs), so the bounds test runs on the converted integers. */
	const float2 tx = __fadd2_rn(__fadd2_rn(bcast(G.ppx), qx), bcast(0.5f));
	const float2 ty = __fadd2_rn(__fadd2_rn(bcast(G.ppy), qy), bcast(0.5f));
	const int px0 = __float2int_rz(tx.x), px1 = __float2int_rz(tx.y);
	const int py0 = __float2int_rz(ty.x), py1 = __float2int_rz(ty.y);
	const uint32_t off0 = (uint32_t)py0 * G.row_bytes + (T.origin + (uint32_t)(px0 * G.px_step));
	const uint32_t off1 = (uint32_t)py1 * G.row_bytes + (T.origin + (uint32_t)(px1 * G.px_step));
	const unsigned long long key0 = make_key_biased(Z.x, key_lo0, T.bias), key1 = make_key_biased(Z.y, key_lo1, T.bias);
	/* DepthFrame::filter folded in, divide_pair's guard, image bounds — as PREDICATES, non-short-circuit on purpose.
	 * G.zero is ALWAYS 0: the warp-uniform, never-taken branch on it is a scheduling fence. The form of this tail decides
	 * how ptxas lays out the whole straight-line segment (8 pairs, ~700 instructions), measured on a B200, C2:
	 *   this form                                                    157 k framesets/s  (2 SEL per pair)
	 *   the same without the branch                                  if-converted into select chains, 12 SEL per pair
	 *   the whole tail as one PTX block (setp.and chains, 81 instead of 88 instructions per pair)   144 k
	 * The branch costs one uniform predicate test per pair and changes no result. */
	bool v0 = fabsf(m.x) <= K.halfw, v1 = fabsf(m.y) <= K.halfw;
	if (kRoiAll == 0 || (kRoiAll == 2 && !K.roi_all)) {
		v0 = v0 && row_in_roi && (unsigned)(u - K.roi_x0) < (unsigned)K.roi_w;
		v1 = v1 && row_in_roi && (unsigned)(u + du - K.roi_x0) < (unsigned)K.roi_w; /* the pair is (u, u + du) */
	}
	const bool fast = g >= 0x1p-40f;
	bool ok0 = v0 & fast & ((unsigned)px0 < (unsigned)G.out_w) & ((unsigned)py0 < (unsigned)G.out_h);
	bool ok1 = v1 & fast & ((unsigned)px1 < (unsigned)G.out_w) & ((unsigned)py1 < (unsigned)G.out_h);
	if (G.zero & 1) {
		ok0 = ok0 & ((px0 | py0) == 0);
		ok1 = ok1 & ((px1 | py1) == 0);
	}
	if (kDump) { /* branchless: redirect instead of skipping (RED.MIN.64) */
		const uint32_t o0 = ok0 ? off0 : T.dump, o1 = ok1 ? off1 : T.dump;
		asm volatile("red.global.min.u64 [%0], %1;" ::"l"(T.base + o0), "l"(key0) : "memory");
		asm volatile("red.global.min.u64 [%0], %1;" ::"l"(T.base + o1), "l"(key1) : "memory");
	} else { /* one predicated RED.MIN.64 per pixel (ptxas turns the predicate into a branch around the RED) */
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p red.global.min.u64 [%0], %1;\n\t}" ::"l"(T.base + off0), "l"(key0), "r"((int)ok0) : "memory");
		asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p red.global.min.u64 [%0], %1;\n\t}" ::"l"(T.base + off1), "l"(key1), "r"((int)ok1) : "memory");
	}
	slow |= (!fast & (v0 | v1)) ? bit : 0u;
}

template <int kHasT, int kRoiAll, int... kI>
__device__ __forceinline__ uint32_t splat_segment(std::integer_sequence<int, kI...>, const SplatShared& G, const SplatCam& K,
                                                  const uint32_t (&dw)[8], const uint32_t (&cw)[12], const float2 (&xr)[8],
                                                  float yr, int u0, bool row_in_roi, const KeyTarget& T) {
	uint32_t slow = 0;
	(splat_pair<kHasT, kRoiAll>(G, K, dw[kI], xr[kI], yr, rgb_of<2 * kI>(cw, K.cam), rgb_of<2 * kI + 1>(cw, K.cam), u0 + 2 * kI,
	                            row_in_roi, T, slow, 1u << kI),
	 ...);
	return slow;
}

/* same, with the 16 column rays of the segment read from shared memory (address xs, 8-byte aligned) right
 * where each pair needs them: no ray registers live across the tile */
template <int kHasT, int kRoiAll, int kDump, int... kI>
__device__ __forceinline__ uint32_t splat_segment_lds(std::integer_sequence<int, kI...>, const SplatShared& G, const SplatCam& K,
                                                      const uint32_t (&dw)[8], const uint32_t (&cw)[12], uint32_t xs, float yr,
                                                      int u0, bool row_in_roi, const KeyTarget& T) {
	auto ray_pair = [xs](int k) {
		float2 v;
		asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(xs + 8u * (uint32_t)k));
		return v;
	};
	uint32_t slow = 0;
	(splat_pair<kHasT, kRoiAll, kDump>(G, K, dw[kI], ray_pair(kI), yr, rgb_of<2 * kI>(cw, K.cam), rgb_of<2 * kI + 1>(cw, K.cam),
	                                   u0 + 2 * kI, row_in_roi, T, slow, 1u << kI),
	 ...);
	return slow;
}

inline SplatShared make_splat_shared(const FuseParams& P) {
	SplatShared G;
	G.Tf = P.fused;
	G.fx = P.proj.fx;
	G.fy = P.proj.fy;
	G.ppx = P.proj.ppx;
	G.ppy = P.proj.ppy;
	G.out_w = P.proj.width;
	G.out_h = P.proj.height;
	G.row_bytes = (uint32_t)P.proj.width * 8u;
	G.px_step = P.proj.mirror ? -8 : 8;
	G.px_origin = P.proj.mirror ? (uint32_t)(P.proj.width - 1) * 8u : 0u;
	G.n_cameras = P.n_cameras;
	G.n_out_pixels = (uint32_t)P.n_out_pixels;
	G.width = P.cam[0].width;
	G.height = P.cam[0].height;
	G.x_pitch = G.y_base = G.y_pitch = 0;
	G.zero = 0;
	return G;
}

inline SplatCam make_splat_cam(const FuseParams& P, int c, const RayTables& rays) {
	const CamParams& C = P.cam[c];
	SplatCam K;
	K.Tc = C.to_ref;
	K.scale = C.scale;
	K.mid = C.mid;
	K.halfw = C.halfw;
	K.roi_x0 = C.roi_x0;
	K.roi_y0 = C.roi_y0;
	K.roi_w = C.roi_w;
	K.roi_h = C.roi_h;
	K.cam = (uint32_t)c;
	K.x_off = c * rays.x_pitch;
	K.y_off = (int)(rays.y - rays.x) + c * rays.y_pitch;
	K.has_t = C.has_transform;
	K.roi_all = C.roi_all;
	return K;
}

}  // namespace
}  // namespace dfb
