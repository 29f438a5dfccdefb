/*
 * fuse_stream.cu — the whole batch in ONE persistent, warp-specialised launch: the splat of every camera of
 * every frameset AND the resolve of every fused image, streamed through a small ring of L2-resident key slots.
 *
 * Why (ncu / CUDA events on a B200, profiles/): with one launch per camera and group (fuse_column.cu,
 * fuse_tiled.cu + resolve_kernel) a launch lasts ~20 us; a quarter of it is ramp-up and tail, the three
 * kernels of a group cannot overlap, and a group cannot grow without pushing the key ring (8 B per output
 * pixel and frameset) out of L2.
 *
 *   work list, in order:   S(0) .. S(lag-1) | S(f) R(f-lag) | ..   R(n-lag) .. R(n-1)
 *       S(f) = the input tiles (64 x 32 source pixels) of all cameras of frameset f   -> atomicMin into slot f % slots
 *       R(f) = the output tiles (4096 pixels) of fused image f: key -> rgb8, re-arm key <- slot f % slots
 *   Items are CLAIMED from one global ticket counter, always in list order, so an item only ever depends on
 *   items with smaller tickets, which are held by CTAs that are running: no deadlock whatever the residency.
 *   Two counters per frameset carry the dependencies:
 *       R(f) needs every S(f) tile done           (done_splat[f]         == splat tiles per frameset)
 *       S(f) needs R(f - slots) done              (done_resolve[f-slots] == resolve tiles per frameset)
 *
 * CTA = 4 compute warps + 1 manager warp (one lane active), kStages-deep ring of shared-memory stages:
 *   manager : claims tickets, decodes them, waits for the item's dependency (ld.acquire.gpu, cached per
 *             frameset), publishes the item into a stage and, for a splat tile, starts its four TMA loads
 *             (cp.async.bulk.tensor, mbarrier full[s] with expect_tx); retires finished stages: waits on
 *             done[s] (one arrival per compute warp) and bumps the frameset counter with red.release.gpu —
 *             cumulative over everything the compute warps wrote before they arrived. It is an event loop:
 *             it never blocks on a dependency without retiring its own finished items, which is what the
 *             dependency may be waiting for.
 *   compute : wait full[s] -> read the item -> five conflict-free LDS.128 (splat) -> the shared per-pair body
 *             (splat_core.cuh) or the resolve of 4096 pixels -> arrive on done[s]. No global waits, no fences,
 *             no CTA-wide barriers: all long latencies live in the manager.
 * Every wait is bounded (trap, never a hung GPU). Every key is identical to the other variants.
 */
#include <cstdio>

#include "splat_core.cuh"
#include "tma_utils.cuh"

namespace dfb {
namespace {

constexpr int kComputeWarps = kTileW / kSeg; /* 4 */
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = kComputeThreads + 64; /* + the producer warp and the retirer warp */
#ifndef DFB_STREAM_STAGES
#define DFB_STREAM_STAGES 3
#endif
constexpr int kStages = DFB_STREAM_STAGES;
/* a stage = the TMA tile (tma_utils.cuh) + the deprojection rays of the tile: 64 column rays + 32 row rays */
constexpr uint32_t kRayXOff = kStageBytes;                /* 256 B */
constexpr uint32_t kRayYOff = kStageBytes + kTileW * 4;   /* 128 B */
constexpr uint32_t kStageTx = kStageBytes + (kTileW + kTileH) * 4;
constexpr uint32_t kStageStride = 11 * 1024;              /* keeps every depth tile 1024-byte aligned */
constexpr uint32_t kSmemBytes = kStages * kStageStride + 1024;
constexpr int kResolveTile = 4096; /* output pixels per resolve item: 128 threads x 4 pixels x 8 rounds */
constexpr int kTicketBatchDefault = 1; /* tickets claimed per atomicAdd (DFB_STREAM_BATCH) */
constexpr uint32_t kDumpWords = 1u << 16; /* 512 KB of scratch keys for the branchless REDs */

struct StreamCams {
	SplatCam K[DFB_MAX_CAMERAS];
};

struct StreamParams {
	int n_framesets, n_cameras;
	int tiles_x, tiles_per_image;
	int splat_per_fs;   /* n_cameras * tiles_per_image */
	int resolve_per_fs; /* ceil(n_out_pixels / kResolveTile) */
	int lag, slots;
	int total_items;
	uint32_t n_out_pixels;
	uint32_t slot_bytes; /* bytes between key slots */
	uint32_t dump_off;   /* byte offset of the RED scratch area from the ring base */
	int batch;   /* tickets claimed per atomicAdd */
	int profile; /* DFB_STREAM_PROFILE: a few CTAs print where their cycles went */
};

/* a claimed work item, decoded by the manager and published in shared memory */
struct Item {
	int kind; /* 0 splat, 1 resolve, -1 none (ticket past the end of the list) */
	int f;    /* frameset */
	int a, b; /* splat: camera, tile index inside the image; resolve: tile index, unused */
};

__device__ __forceinline__ Item decode_item(const StreamParams& Q, int ticket) {
	Item it;
	it.kind = -1;
	it.f = it.a = it.b = 0;
	if (ticket >= Q.total_items) return it;
	unsigned i = (unsigned)ticket;
	const unsigned S = (unsigned)Q.splat_per_fs, R = (unsigned)Q.resolve_per_fs;
	const unsigned head = S * (unsigned)Q.lag; /* rounds 0..lag-1: splat only */
	unsigned j;
	if (i < head) {
		it.kind = 0;
		it.f = (int)(i / S);
		j = i - (unsigned)it.f * S;
	} else {
		i -= head;
		const unsigned per = S + R;
		const unsigned mid = per * (unsigned)(Q.n_framesets - Q.lag); /* rounds lag..n-1: splat(r) then resolve(r-lag) */
		if (i < mid) {
			const unsigned r = i / per;
			j = i - r * per;
			if (j < S) {
				it.kind = 0;
				it.f = (int)r + Q.lag;
			} else {
				it.kind = 1;
				it.f = (int)r;
				j -= S;
			}
		} else { /* rounds n..n+lag-1: resolve only */
			i -= mid;
			const unsigned r = i / R;
			it.kind = 1;
			it.f = Q.n_framesets - Q.lag + (int)r;
			j = i - r * R;
		}
	}
	if (it.kind == 0) {
		it.a = (int)(j / (unsigned)Q.tiles_per_image);
		it.b = (int)(j - (unsigned)it.a * (unsigned)Q.tiles_per_image);
	} else {
		it.a = (int)j;
	}
	return it;
}

/* Dependency counters are polled with a strong RELAXED load. The producer side (retirer warp) bumps a counter only
 * after a gpu-scope fence has drained the CTA's atomics / stores to L2, and every consumer access to the keys goes
 * to L2 as well (RED, ld.global.cg, st.global.cg: nothing is ever cached in L1), so observing the final count is
 * enough; the consumers start strictly later (mbarrier release/acquire inside the CTA). An ld.acquire.gpu here
 * costs ~2000 cycles a piece (it invalidates the SM's L1) and was 50 % of the producer's time. */
__device__ __forceinline__ unsigned int ld_counter(const unsigned int* p) {
	unsigned int v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void resolve_item(const StreamParams& Q, int f, int j, unsigned long long* __restrict__ keys,
                                             uint8_t* __restrict__ out, int tid) {
	const size_t n = Q.n_out_pixels;
	unsigned long long* k = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(keys) + (size_t)(f % Q.slots) * Q.slot_bytes);
	uint8_t* o = out + (size_t)f * n * 3;
	const size_t tile0 = (size_t)j * kResolveTile;
#pragma unroll 2
	for (int r = 0; r < kResolveTile / (kComputeThreads * 4); ++r) {
		const size_t i0 = tile0 + (size_t)r * (kComputeThreads * 4) + (size_t)tid * 4;
		if (i0 + 4 <= n) {
			ulonglong2* kp = reinterpret_cast<ulonglong2*>(k + i0);
			const ulonglong2 a = __ldcg(kp), b = __ldcg(kp + 1);
			const ulonglong2 empty = make_ulonglong2(KEY_EMPTY, KEY_EMPTY);
			__stcg(kp, empty);
			__stcg(kp + 1, empty);
			const uint32_t c0 = (a.x == KEY_EMPTY) ? 0u : (uint32_t)a.x, c1 = (a.y == KEY_EMPTY) ? 0u : (uint32_t)a.y;
			const uint32_t c2 = (b.x == KEY_EMPTY) ? 0u : (uint32_t)b.x, c3 = (b.y == KEY_EMPTY) ? 0u : (uint32_t)b.y;
			uint32_t* op = reinterpret_cast<uint32_t*>(o + i0 * 3);
			__stcs(op, __byte_perm(c0, c1, 0x6012));
			__stcs(op + 1, __byte_perm(c1, c2, 0x5601));
			__stcs(op + 2, __byte_perm(c2, c3, 0x4560));
		} else {
			for (size_t i = i0; i < n && i < i0 + 4; ++i) {
				const unsigned long long key = __ldcg(k + i);
				__stcg(k + i, KEY_EMPTY);
				const uint32_t c = (key == KEY_EMPTY) ? 0u : (uint32_t)key;
				o[i * 3 + 0] = (uint8_t)(c >> 16);
				o[i * 3 + 1] = (uint8_t)(c >> 8);
				o[i * 3 + 2] = (uint8_t)c;
			}
		}
	}
}

/* the splat of one thread's 16-pixel row segment with camera constants K: picks the straight-line instance
 * (camera transform yes/no, ROI test needed yes/no) once per tile, so that the 8 pixel pairs of the segment
 * form ONE basic block (branchless REDs, see KeyTarget) the scheduler can interleave */
__device__ __forceinline__ uint32_t splat_tile(const SplatShared& G, const SplatCam& K, uint32_t xs, float yr,
                                               const uint32_t (&dw)[8], const uint32_t (&cw)[12], int u0, int v, int tile_u0,
                                               int tile_v0, const KeyTarget& T) {
	constexpr auto seq = std::make_integer_sequence<int, kSeg / 2>{};
	const bool row_in_roi = (unsigned)(v - K.roi_y0) < (unsigned)K.roi_h;
	const bool inside = K.roi_all || (tile_u0 >= K.roi_x0 && tile_u0 + kTileW <= K.roi_x0 + K.roi_w && tile_v0 >= K.roi_y0 &&
	                                  tile_v0 + kTileH <= K.roi_y0 + K.roi_h);
	if (K.has_t)
		return inside ? splat_segment_lds<1, 1, 1>(seq, G, K, dw, cw, xs, yr, u0, row_in_roi, T)
		              : splat_segment_lds<1, 0, 1>(seq, G, K, dw, cw, xs, yr, u0, row_in_roi, T);
	return inside ? splat_segment_lds<0, 1, 1>(seq, G, K, dw, cw, xs, yr, u0, row_in_roi, T)
	              : splat_segment_lds<0, 0, 1>(seq, G, K, dw, cw, xs, yr, u0, row_in_roi, T);
}

/* counters: [0] ticket, [1] unused, [2 .. 2+n) done_splat, [2+n .. 2+2n) done_resolve */
__global__ void __launch_bounds__(kThreads, 4)
fuse_stream_kernel(const __grid_constant__ SplatShared G, const __grid_constant__ StreamCams cams,
                   const __grid_constant__ FuseParams P, const __grid_constant__ StreamParams Q,
                   const __grid_constant__ CUtensorMap map_depth, const __grid_constant__ CUtensorMap map_color,
                   const float* __restrict__ ray_tables, const uint16_t* __restrict__ depth,
                   const uint8_t* __restrict__ color, unsigned long long* __restrict__ keys, uint8_t* __restrict__ out,
                   unsigned int* __restrict__ counters) {
	extern __shared__ uint8_t smem_raw[];
	__shared__ __align__(8) unsigned long long full_bar[kStages], done_bar[kStages];
	__shared__ Item items[kStages];
	__shared__ int s_counts[3]; /* published, retired, finished: producer <-> retirer */
	volatile int* s_published = &s_counts[0];
	volatile int* s_retired = &s_counts[1];
	volatile int* s_finished = &s_counts[2];
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	if (tid == 0) {
		for (int s = 0; s < kStages; ++s) {
			mbar_init(smem_u32(&full_bar[s]), 1);
			mbar_init(smem_u32(&done_bar[s]), kComputeWarps);
		}
		mbar_fence_init();
		s_counts[0] = s_counts[1] = s_counts[2] = 0;
	}
	__syncthreads(); /* the only CTA-wide barrier */

	if (warp == kComputeWarps) {
		/* -------------------------------- producer (one lane of warp 4) -------------------------------- */
		if (lane != 0) return;
		const unsigned int* done_splat = counters + 2;
		const unsigned int* done_resolve = counters + 2 + Q.n_framesets;
		const uint64_t policy = l2_evict_first_policy();
		int published = 0;
		int next_ticket = 0, batch_end = 0; /* claimed-but-undecoded tickets [next_ticket, batch_end) */
		int known_resolved = -1, known_splatted = -1;
		unsigned long long idle = 0;
		long long t_claim = 0, t_pub = 0, t_dep = 0, t_stage = 0, t0 = clock64(), t1;
		while (true) {
			t1 = clock64();
			if (next_ticket == batch_end) {
				next_ticket = (int)atomicAdd(&counters[0], (unsigned)Q.batch);
				batch_end = next_ticket + Q.batch;
			}
			const Item it = decode_item(Q, next_ticket++);
			t_claim += clock64() - t1;
			t1 = clock64();
			/* a stage is reusable once its previous item has been retired (signalled) */
			while (published - *s_retired >= kStages) {
				__nanosleep(64);
				if (++idle > (1ull << 26)) __trap();
			}
			idle = 0;
			t_stage += clock64() - t1;
			t1 = clock64();
			/* the item's dependency: checked once per frameset, then cached */
			if (it.kind == 0) {
				const int need = it.f - Q.slots; /* key slot re-armed? */
				if (need > known_resolved) {
					while (ld_counter(&done_resolve[need]) < (unsigned)Q.resolve_per_fs) {
						__nanosleep(200);
						if (++idle > (1ull << 24)) __trap(); /* a protocol bug must be an error, not a hung GPU */
					}
					known_resolved = need;
				}
			} else if (it.kind == 1) {
				if (it.f > known_splatted) { /* every camera of frameset f splatted? */
					while (ld_counter(&done_splat[it.f]) < (unsigned)Q.splat_per_fs) {
						__nanosleep(200);
						if (++idle > (1ull << 24)) __trap();
					}
					known_splatted = it.f;
				}
			}
			idle = 0;
			t_dep += clock64() - t1;
			t1 = clock64();
			const int s = published % kStages;
			items[s] = it;
			const uint32_t bar = smem_u32(&full_bar[s]);
			if (it.kind == 0) {
				const int ty = it.b / Q.tiles_x;
				const int u0 = (it.b - ty * Q.tiles_x) * kTileW, v0 = ty * kTileH;
				const int image = it.f * Q.n_cameras + it.a;
				const uint32_t dst = base + (uint32_t)s * kStageStride;
				mbar_expect_tx(bar, kStageTx);
				tma_load_3d(dst, &map_depth, u0, v0, image, bar, policy);
				tma_load_3d(dst + kDepthBytes, &map_color, u0 * 3, v0, image, bar, policy);
				tma_load_3d(dst + kDepthBytes + kColorBoxBytes, &map_color, u0 * 3 + 64, v0, image, bar, policy);
				tma_load_3d(dst + kDepthBytes + 2 * kColorBoxBytes, &map_color, u0 * 3 + 128, v0, image, bar, policy);
				bulk_load_1d(dst + kRayXOff, ray_tables + it.a * G.x_pitch + u0, kTileW * 4, bar);
				bulk_load_1d(dst + kRayYOff, ray_tables + G.y_base + it.a * G.y_pitch + v0, kTileH * 4, bar);
			} else {
				mbar_arrive(bar); /* resolve item or terminator: nothing to load */
			}
			t_pub += clock64() - t1;
			if (it.kind < 0) break; /* the terminator: compute warps stop there and never report it done */
			++published;
			__threadfence_block();
			*s_published = published;
		}
		__threadfence_block();
		*s_finished = 1;
		if (Q.profile && (blockIdx.x % 149) == 0)
			printf("cta %4d producer: items %d total %lld claim+decode %lld wait-stage %lld wait-dependency %lld publish %lld\n",
			       blockIdx.x, published, clock64() - t0, t_claim, t_stage, t_dep, t_pub);
		return;
	}
	if (warp == kComputeWarps + 1) {
		/* -------------------------------- retirer (one lane of warp 5) --------------------------------- */
		if (lane != 0) return;
		unsigned int* done_splat = counters + 2;
		unsigned int* done_resolve = counters + 2 + Q.n_framesets;
		int retired = 0;
		unsigned long long idle = 0;
		long long t_wait = 0, t_signal = 0, t0 = clock64(), t1;
		int fences = 0;
		while (true) {
			const int pub = *s_published;
			if (retired == pub) {
				if (*s_finished && retired == *s_published) break;
				__nanosleep(64);
				if (++idle > (1ull << 26)) __trap();
				continue;
			}
			idle = 0;
			t1 = clock64();
			/* the compute warps finish items in publication order: block on the oldest, then take every further one
			 * that is already done, and pay ONE gpu-scope fence for the batch */
			mbar_wait(smem_u32(&done_bar[retired % kStages]), (uint32_t)(retired / kStages) & 1u);
			int n = 1;
			while (retired + n < pub && mbar_test(smem_u32(&done_bar[(retired + n) % kStages]), (uint32_t)((retired + n) / kStages) & 1u)) ++n;
			t_wait += clock64() - t1;
			t1 = clock64();
			__threadfence(); /* cumulative: everything the compute warps wrote before arriving on done[] */
			for (int i = 0; i < n; ++i) {
				const Item it = items[(retired + i) % kStages];
				atomicAdd(it.kind == 0 ? &done_splat[it.f] : &done_resolve[it.f], 1u);
			}
			retired += n;
			++fences;
			__threadfence_block();
			*s_retired = retired;
			t_signal += clock64() - t1;
		}
		if (Q.profile && (blockIdx.x % 149) == 0)
			printf("cta %4d retirer: items %d fences %d total %lld wait-done %lld fence+signal %lld\n", blockIdx.x, retired, fences,
			       clock64() - t0, t_wait, t_signal);
		return;
	}

	/* ---------------------------------------- compute warps ---------------------------------------- */
	const TileReadOffsets off = tile_read_offsets((uint32_t)lane, (uint32_t)warp);
	long long c_wait = 0, c_splat = 0, c_resolve = 0, c0 = clock64();
	int n_splat = 0, n_resolve = 0;
	for (int k = 0;; ++k) {
		const int s = k % kStages;
		long long tc = clock64();
		mbar_wait(smem_u32(&full_bar[s]), (uint32_t)(k / kStages) & 1u);
		c_wait += clock64() - tc;
		tc = clock64();
		const Item it = items[s];
		if (it.kind < 0) break;
		if (it.kind == 0) {
			const uint32_t stage = base + (uint32_t)s * kStageStride;
			const uint4 d0 = lds128(stage + off.d[0]), d1 = lds128(stage + off.d[1]);
			const uint4 c0 = lds128(stage + off.c[0]), c1 = lds128(stage + off.c[1]), c2 = lds128(stage + off.c[2]);
			float yr;
			asm volatile("ld.shared.f32 %0, [%1];" : "=f"(yr) : "r"(stage + kRayYOff + 4u * (uint32_t)lane));
			const uint32_t xs = stage + kRayXOff + (uint32_t)warp * (kSeg * 4);
			const int ty = it.b / Q.tiles_x;
			const int v = ty * kTileH + lane;
			const int u0 = (it.b - ty * Q.tiles_x) * kTileW + warp * kSeg;
			const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
			const uint32_t cw[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
			const uint32_t slot_off = (uint32_t)(it.f % Q.slots) * Q.slot_bytes;
			unsigned long long* kslot = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(keys) + slot_off);
			/* invalid / out-of-image points dump their RED into a per-thread word of the scratch area behind the ring */
			const KeyTarget T{reinterpret_cast<char*>(keys), slot_off + G.px_origin,
			                  Q.dump_off + (((uint32_t)blockIdx.x * kComputeThreads + (uint32_t)tid) & (kDumpWords - 1)) * 8u};
			const int tu0 = u0 - warp * kSeg, tv0 = ty * kTileH;
			/* One copy of the tile body per camera index: with a STATIC index the camera's constants sit at fixed
			 * constant-bank addresses and feed FFMA2 straight from uniform registers; a run-time index costs an
			 * LDC + R2UR per use (~18 % more instructions, measured in SASS). Cameras >= 2 take the generic copy. */
			uint32_t slow;
			switch (it.a) {
				case 0: slow = splat_tile(G, cams.K[0], xs, yr, dw, cw, u0, v, tu0, tv0, T); break;
				case 1: slow = splat_tile(G, cams.K[1], xs, yr, dw, cw, u0, v, tu0, tv0, T); break;
				default: slow = splat_tile(G, cams.K[it.a], xs, yr, dw, cw, u0, v, tu0, tv0, T); break;
			}
			if (slow) { /* rare: operands outside the guarded fast-division range */
				const size_t img = ((size_t)it.f * Q.n_cameras + it.a) * ((size_t)G.width * G.height);
				while (slow) {
					const int j = __ffs(slow) - 1;
					slow &= slow - 1;
					splat_pixel_slow(P, it.a, depth + img, color + img * 3, u0 + 2 * j, v, kslot);
					splat_pixel_slow(P, it.a, depth + img, color + img * 3, u0 + 2 * j + 1, v, kslot);
				}
			}
			c_splat += clock64() - tc;
			++n_splat;
		} else {
			resolve_item(Q, it.f, it.a, keys, out, tid);
			c_resolve += clock64() - tc;
			++n_resolve;
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(smem_u32(&done_bar[s])); /* release: everything this warp wrote for the item */
	}
	if (Q.profile && (blockIdx.x % 149) == 0 && lane == 0 && warp == 1)
		printf("cta %4d compute: total %lld wait-full %lld splat %lld (%d tiles) resolve %lld (%d tiles)\n", blockIdx.x, clock64() - c0,
		       c_wait, c_splat, n_splat, c_resolve, n_resolve);
}

}  // namespace

bool fuse_stream_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color, const uint8_t* d_out) {
	return pools_fit_tensor_maps(P.cam[0].width, d_depth, d_color) && ((size_t)P.n_out_pixels * 8 < (1ull << 32)) &&
	       (P.n_out_pixels % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 3u) == 0);
}

size_t fuse_stream_counter_words(int n_framesets) { return 2 + 2 * (size_t)n_framesets; }

/* bytes the stream kernel needs BEHIND the `slots` key slots (scratch for the branchless REDs) */
size_t fuse_stream_scratch_bytes() { return (size_t)kDumpWords * 8; }

/* d_keys holds `slots` frameset slots (all-ones on entry and on exit); d_counters = fuse_stream_counter_words(n) */
int launch_fuse_stream(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                       unsigned long long* d_keys, unsigned int* d_counters, int n_framesets, int slots, int lag,
                       int ctas_per_sm, cudaStream_t s) {
	const int w = P.cam[0].width, h = P.cam[0].height;
	alignas(64) CUtensorMap md, mc;
	int rc = make_pool_tensor_maps(d_depth, d_color, w, h, (uint64_t)n_framesets * P.n_cameras, &md, &mc);
	if (rc != DFB_OK) return rc;
	StreamParams Q;
	Q.n_framesets = n_framesets;
	Q.n_cameras = P.n_cameras;
	Q.tiles_x = (w + kTileW - 1) / kTileW;
	Q.tiles_per_image = Q.tiles_x * ((h + kTileH - 1) / kTileH);
	Q.splat_per_fs = Q.tiles_per_image * P.n_cameras;
	Q.resolve_per_fs = (P.n_out_pixels + kResolveTile - 1) / kResolveTile;
	if (slots < 1) slots = 1;
	if (lag > slots - 1) lag = slots - 1; /* R(f-lag) must be queued before S(f-lag+slots) */
	if (lag > n_framesets) lag = n_framesets;
	if (lag < 0) lag = 0;
	Q.lag = lag;
	Q.slots = slots;
	Q.n_out_pixels = (uint32_t)P.n_out_pixels;
	Q.slot_bytes = (uint32_t)P.n_out_pixels * 8u;
	Q.dump_off = Q.slot_bytes * (uint32_t)slots;
	DFB_REQUIRE((unsigned long long)P.n_out_pixels * 8ull * (unsigned long long)slots + fuse_stream_scratch_bytes() < (1ull << 32),
	            "key ring too large for 32-bit offsets: use fewer slots");
	const long long total = (long long)n_framesets * (Q.splat_per_fs + Q.resolve_per_fs);
	DFB_REQUIRE(total < (1ll << 30), "too many work items for one launch: split the batch");
	Q.total_items = (int)total;
	{
		const char* e = std::getenv("DFB_STREAM_PROFILE");
		Q.profile = e ? std::atoi(e) : 0;
		const char* b = std::getenv("DFB_STREAM_BATCH");
		Q.batch = b ? std::atoi(b) : kTicketBatchDefault;
		if (Q.batch < 1) Q.batch = 1;
	}
	StreamCams cams;
	for (int c = 0; c < DFB_MAX_CAMERAS; ++c) cams.K[c] = make_splat_cam(P, c < P.n_cameras ? c : 0, rays);
	SplatShared G = make_splat_shared(P);
	G.x_pitch = rays.x_pitch;
	G.y_base = (int)(rays.y - rays.x);
	G.y_pitch = rays.y_pitch;

	static int max_ctas = 0, sms = 0;
	if (max_ctas == 0) {
		DFB_CUDA(cudaFuncSetAttribute(fuse_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
		int dev = 0;
		DFB_CUDA(cudaGetDevice(&dev));
		DFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
		DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_ctas, fuse_stream_kernel, kThreads, kSmemBytes));
		if (max_ctas < 1) max_ctas = 1;
	}
	const int per_sm = (ctas_per_sm > 0 && ctas_per_sm < max_ctas) ? ctas_per_sm : max_ctas;
	long long grid = (long long)sms * per_sm;
	if (grid > total) grid = total;
	DFB_CUDA(cudaMemsetAsync(d_counters, 0, fuse_stream_counter_words(n_framesets) * sizeof(unsigned int), s));
	fuse_stream_kernel<<<(unsigned)grid, kThreads, kSmemBytes, s>>>(G, cams, P, Q, md, mc, rays.x, d_depth, d_color, d_keys, d_out,
	                                                                 d_counters);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

}  // namespace dfb
