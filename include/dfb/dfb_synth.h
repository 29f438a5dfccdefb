/*
 * dfb_synth.h — deterministic synthetic RealSense-style framesets.
 *
 * Header-only, compiled both by gcc (CPU generator used by the tests) and by nvcc (GPU
 * generator used by bench.py), and the two MUST produce identical bytes. To get that:
 *   - counter-based RNG (splitmix64 of seed/camera/pixel): no sequential state;
 *   - the ray cast uses only IEEE double + - * / sqrt, every operation written through the
 *     DFS_* macros so that neither nvcc (-fmad) nor gcc (-ffp-contract) can fuse a multiply
 *     into an add. On the device the macros are the __d*_rn intrinsics; on the host they are
 *     plain operators inside functions compiled with fp-contract=off.
 *
 * Input format produced = what camera_node publishes (camera_node/src/camera_node.cpp:191-209):
 * depth `mono16` little-endian row-major uint16[H][W], colour `rgb8` uint8[H][W][3].
 *
 * Scene (all in the reference camera's frame = camera 0, +Z forward): a tilted back wall
 * (plane) plus a sphere ("person"). Each camera c has a pose `cam_to_ref` (row-major 4x4,
 * the same matrix the fusion applies to that camera's cloud); its depth image is the
 * analytic ray cast of the scene from that pose, quantised to `depth_unit` with +-noise_lsb
 * uniform noise. Colour is a 24-bit hash of (camera, pixel, seed) so every source pixel is
 * distinguishable in the fused image.
 */
#ifndef DFB_SYNTH_H
#define DFB_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define DFS_HD __host__ __device__ static inline
#else
#define DFS_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define DFS_MUL(a, b) __dmul_rn((a), (b))
#define DFS_ADD(a, b) __dadd_rn((a), (b))
#define DFS_SUB(a, b) __dsub_rn((a), (b))
#define DFS_DIV(a, b) __ddiv_rn((a), (b))
#define DFS_SQRT(a) __dsqrt_rn((a))
#define DFS_NOCONTRACT
#else
#include <math.h>
#define DFS_MUL(a, b) ((a) * (b))
#define DFS_ADD(a, b) ((a) + (b))
#define DFS_SUB(a, b) ((a) - (b))
#define DFS_DIV(a, b) ((a) / (b))
#define DFS_SQRT(a) sqrt((a))
#if defined(__GNUC__) && !defined(__clang__) && !defined(__CUDACC__)
#define DFS_NOCONTRACT __attribute__((optimize("fp-contract=off")))
#else
#define DFS_NOCONTRACT
#endif
#endif

typedef struct dfs_camera {
	int32_t width, height;
	float ppx, ppy, fx, fy;
	float cam_to_ref[16]; /* row-major 4x4 */
} dfs_camera;

typedef struct dfs_scene {
	double wall_p0[3];   /* a point on the wall */
	double wall_n[3];    /* wall normal (need not be unit) */
	double sphere_c[3];  /* sphere centre */
	double sphere_r;     /* sphere radius; <=0 disables the sphere */
	double inv_depth_unit; /* raw units per metre (1000 for depth_scale 0.001) */
	int32_t noise_lsb;   /* uniform integer noise in [-noise_lsb, +noise_lsb] */
	uint32_t hole_q16;   /* probability of an invalid (raw 0) pixel, in 1/65536 */
	uint64_t seed;       /* per-frameset seed: 0xD4550000 + frameset index */
} dfs_scene;

DFS_HD uint64_t dfs_mix64(uint64_t x) {
	x += 0x9E3779B97F4A7C15ULL;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
	return x ^ (x >> 31);
}

/* one 64-bit hash per (seed, camera, pixel) */
DFS_HD uint64_t dfs_pixel_hash(uint64_t seed, uint32_t cam, uint32_t pixel) {
	return dfs_mix64(dfs_mix64(seed) ^ (((uint64_t)cam << 40) | (uint64_t)pixel));
}

DFS_HD DFS_NOCONTRACT double dfs_dot3(const double* a, const double* b) {
	return DFS_ADD(DFS_ADD(DFS_MUL(a[0], b[0]), DFS_MUL(a[1], b[1])), DFS_MUL(a[2], b[2]));
}

/* Distance along the pixel ray (== camera-frame Z, since the ray direction has z = 1) of the
 * first scene hit, or a negative number when the ray hits nothing in front of the camera. */
DFS_HD DFS_NOCONTRACT double dfs_raycast(const dfs_scene* sc, const dfs_camera* cam, int u, int v) {
	const double dx = DFS_DIV(DFS_SUB((double)u, (double)cam->ppx), (double)cam->fx);
	const double dy = DFS_DIV(DFS_SUB((double)v, (double)cam->ppy), (double)cam->fy);
	const float* T = cam->cam_to_ref;
	double D[3], O[3];
	for (int r = 0; r < 3; ++r) {
		D[r] = DFS_ADD(DFS_ADD(DFS_MUL((double)T[4 * r + 0], dx), DFS_MUL((double)T[4 * r + 1], dy)),
		               (double)T[4 * r + 2]);
		O[r] = (double)T[4 * r + 3];
	}
	double best = -1.0;
	/* wall: n.(O + l*D - p0) = 0 */
	{
		double rel[3] = {DFS_SUB(sc->wall_p0[0], O[0]), DFS_SUB(sc->wall_p0[1], O[1]), DFS_SUB(sc->wall_p0[2], O[2])};
		const double nd = dfs_dot3(sc->wall_n, D);
		if (nd != 0.0) {
			const double l = DFS_DIV(dfs_dot3(sc->wall_n, rel), nd);
			if (l > 0.0) best = l;
		}
	}
	/* sphere: |O + l*D - C|^2 = r^2 */
	if (sc->sphere_r > 0.0) {
		double oc[3] = {DFS_SUB(O[0], sc->sphere_c[0]), DFS_SUB(O[1], sc->sphere_c[1]), DFS_SUB(O[2], sc->sphere_c[2])};
		const double a = dfs_dot3(D, D);
		const double b = dfs_dot3(D, oc);
		const double c = DFS_SUB(dfs_dot3(oc, oc), DFS_MUL(sc->sphere_r, sc->sphere_r));
		const double disc = DFS_SUB(DFS_MUL(b, b), DFS_MUL(a, c));
		if (disc >= 0.0) {
			const double l = DFS_DIV(DFS_SUB(-b, DFS_SQRT(disc)), a);
			if (l > 0.0 && (best < 0.0 || l < best)) best = l;
		}
	}
	return best;
}

/* One pixel of one camera: raw depth + RGB. */
DFS_HD DFS_NOCONTRACT void dfs_pixel(const dfs_scene* sc, const dfs_camera* cam, uint32_t cam_index, int u, int v,
                                     uint16_t* depth_out, uint8_t* rgb_out) {
	const uint32_t pixel = (uint32_t)v * (uint32_t)cam->width + (uint32_t)u;
	const uint64_t h = dfs_pixel_hash(sc->seed, cam_index, pixel);
	rgb_out[0] = (uint8_t)(h & 0xFF);
	rgb_out[1] = (uint8_t)((h >> 8) & 0xFF);
	rgb_out[2] = (uint8_t)((h >> 16) & 0xFF);
	uint16_t d = 0;
	const uint32_t hole_draw = (uint32_t)((h >> 24) & 0xFFFF);
	if (hole_draw >= sc->hole_q16) {
		const double l = dfs_raycast(sc, cam, u, v);
		if (l > 0.0) {
			/* floor(l*inv_unit + 0.5) for positive values via truncation */
			const double q = DFS_ADD(DFS_MUL(l, sc->inv_depth_unit), 0.5);
			long long raw = (q < 1.0e9) ? (long long)q : 1000000000LL;
			if (sc->noise_lsb > 0) {
				const uint32_t span = 2u * (uint32_t)sc->noise_lsb + 1u;
				raw += (long long)((uint32_t)((h >> 40) & 0xFFFFF) % span) - (long long)sc->noise_lsb;
			}
			if (raw < 0) raw = 0;
			if (raw > 65535) raw = 65535;
			d = (uint16_t)raw;
		}
	}
	*depth_out = d;
}

#endif /* DFB_SYNTH_H */
