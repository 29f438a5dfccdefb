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
 *       S(f) = the input items (64 x 64 source pixels) of all cameras of frameset f   -> atomicMin into slot f % slots
 *       R(f) = the output tiles (2560 pixels) of fused image f: key -> rgb8, re-arm key <- slot f % slots
 *   Items are CLAIMED from ticket counters in list order (16 counters, group g hands out tickets g, g+16, ..), so
 *   an item only ever depends on items with smaller tickets, which are held by CTAs that are running or are next
 *   in line of their group: no deadlock whatever the residency. Two counters per frameset carry the dependencies:
 *       R(f) needs every S(f) tile done           (done_splat[f]         == splat tiles per frameset)
 *       S(f) needs R(f - slots) done              (done_resolve[f-slots] == resolve tiles per frameset)
 *
 * CTA = 4 compute warps + producer warp + retirer warp (one lane each), kStages shared-memory stages:
 *   producer: waits for a free stage, THEN claims a ticket (claiming late keeps the set of claimed-but-unsignalled
 *             tickets — the window the ring's slack of slots-1 framesets must cover — small), decodes it, checks the
 *             item's dependency (counters polled one item ahead with ld.relaxed.gpu) and starts the bulk copies of
 *             the item: four TMA tensor boxes + two ray rows for a splat tile, the 20 KB key tile for a resolve item
 *             (mbarrier full[s] with expect_tx).
 *   compute : wait full[s] -> five conflict-free LDS.128 -> the shared per-pair body (splat_core.cuh), or keys from
 *             shared memory -> rgb8, staged through the dead key tile and stored as whole 16-byte pieces -> arrive on
 *             done[] and free[s]. No global waits, no fences.
 *   retirer : runs behind the pipeline (queue of kRetireDepth items): re-arms the keys of a finished resolve item with one
 *             bulk store (shared -> global) of a constant KEY_EMPTY tile, then one fence.acq_rel.gpu per kSignalBatch
 *             items — cumulative over everything the compute warps wrote before they arrived — and the frameset counters.
 * Lessons measured on the way (profiles/r01_stream_ncu.md): same-address atomics are served at ~1 per 10 cycles by
 * the L2 — one global ticket counter capped the kernel at 130 k framesets/s; ptxas warp-aggregates every integer
 * atom.add and waits for the result at once (the ticket counters are floats); a persisting-L2 window for the keys
 * makes things worse; more than 6 key slots (44 MB at 720p) spill the ring to DRAM.
 * Every wait is bounded (trap, never a hung GPU). Every key is identical to the other variants.
 */
#if DFB_STREAM_V1 /* round-1 kernel, kept for A/B measurements only: -DDFB_STREAM_V1=1 */
#include <cstdio>

#include "splat_core.cuh"
#include "tma_utils.cuh"

namespace dfb {
namespace {

constexpr int kComputeWarps = kTileW / kSeg; /* 4 */
constexpr int kComputeThreads = kComputeWarps * 32;
#ifndef DFB_STREAM_SETMAXNREG
#define DFB_STREAM_SETMAXNREG 0
#endif
/* + the producer warp and the retirer warp; with register hand-over two more (idle) warps complete the second
 * warpgroup, because setmaxnreg must be executed by all four warps of a warpgroup */
constexpr int kThreads = kComputeThreads + (DFB_STREAM_SETMAXNREG ? 128 : 64);
#ifndef DFB_STREAM_FORMAL_ACQUIRE
#define DFB_STREAM_FORMAL_ACQUIRE 0 /* 1 = a fence.acq_rel.gpu after every newly observed dependency count (PTX-model acquire for
                                      the relaxed polls). Measured cost: 158 k -> 150 k framesets/s. Off by default: every access
                                      to the keys is an L2-level operation (RED, bulk copy, bulk store; nothing is cached in L1),
                                      the retirer orders them before its counter update with a gpu-scope fence, and the consumer
                                      issues its accesses only after the poll has returned the final count (see ld_counter). */
#endif
#ifndef DFB_COLS_DUMP
#define DFB_COLS_DUMP 1 /* column (straight-line) body: 1 = branch-free atomics, 0 = predicated (measured: 158 k -> 140 k: sixteen
                           branches cut the 8-pair basic block into pieces) */
#endif
#ifndef DFB_ROWS_SPLIT
#define DFB_ROWS_SPLIT 1 /* row-major pairing: 1 = columns (lane, lane + 32), 0 = columns (2 lane, 2 lane + 1) */
#endif
#ifndef DFB_ROWS_DUMP
#define DFB_ROWS_DUMP 0 /* row-major loop body: 0 = predicated atomics (rejected points issue nothing), 1 = branch-free atomics into a
                           scratch word. In the LOOP form the two branches per pair cost less than the wasted atomics: un-rotated C2
                           142.3 k -> 145.1 k, profile P1 147.4 k -> 152.6 k, sparse 153.8 k -> 159.3 k framesets/s. (The straight-line
                           column body needs the branch-free form: its eight pairs are one basic block.) */
#endif
#define DFB_PRAGMA_STR(x) _Pragma(#x)
#define DFB_PRAGMA_UNROLL(n) DFB_PRAGMA_STR(unroll n)
#ifndef DFB_ROWS_UNROLL
#define DFB_ROWS_UNROLL 1 /* row-major body: 8 = straight-line, 1 / 2 / 4 = a loop over the rows. Measured on the un-rotated C2:
                             139.4 k (8), 140.9 k (4), 142.2 k (2), 142.4 k (1) framesets/s — ILP across a thread's pairs buys nothing,
                             the loop is 1/8 of the code and needs 62 registers */
#endif
#ifndef DFB_STREAM_STAGES
#define DFB_STREAM_STAGES 2
#endif
constexpr int kStages = DFB_STREAM_STAGES;
/* A splat item is 64 x 64 source pixels: two 32-row halves (lane = row inside a half) that arrive with ONE set of
 * bulk copies — a 64x64 uint16 depth box (SWIZZLE_128B) + three 64-byte x 64-row colour boxes (SWIZZLE_64B) + 64 column
 * rays + 64 row rays. The per-item cost of the producer (claim ~1300, decode ~450, six copy issues ~850 cycles), of the
 * retirer (one gpu-scope fence) and of the barriers is paid per 4096 pixels instead of 2048: with 32-row items the
 * producer was 91 % busy and the compute warps starved whenever a cheap resolve item came by. */
constexpr int kItemRows = 2 * kTileH;
constexpr uint32_t kItemDepthBytes = kTileW * 2 * kItemRows;                 /* 8192 */
constexpr uint32_t kItemColorBoxBytes = 64 * kItemRows;                      /* 4096 */
constexpr uint32_t kItemBytes = kItemDepthBytes + 3 * kItemColorBoxBytes;    /* 20480 */
constexpr uint32_t kRayXOff = kItemBytes;                 /* 256 B */
constexpr uint32_t kRayYOff = kItemBytes + kTileW * 4;    /* 256 B */
constexpr uint32_t kStageTx = kItemBytes + (kTileW + kItemRows) * 4;
constexpr int kResolveTile = 2560; /* output pixels per resolve item: 128 threads x 4 pixels x 5 rounds; its keys (20 KB) are
                                      bulk-copied into the stage like an input tile */
constexpr uint32_t kStageStride = 21 * 1024;              /* >= max(input item + rays, key tile); keeps every depth box
                                                             1024-byte aligned */
static_assert(kResolveTile * 8 <= (int)kStageStride && kStageTx <= kStageStride, "stage too small");
/* behind the stages: one key tile of KEY_EMPTY, the source of the bulk stores that re-arm the ring */
constexpr uint32_t kRearmOff = kStages * kStageStride;
constexpr uint32_t kRearmBytes = kResolveTile * 8; /* a resolve tile is re-armed by kResolveTile * 8 / kRearmBytes bulk stores (one) */
static_assert((kResolveTile * 8) % kRearmBytes == 0, "re-arm chunking");
constexpr uint32_t kSmemBytes = kStages * kStageStride + kRearmBytes + 1024;
#ifndef DFB_STREAM_RETIRE_DEPTH
#define DFB_STREAM_RETIRE_DEPTH 4
#endif
constexpr int kRetireDepth = DFB_STREAM_RETIRE_DEPTH; /* items the retirer may lag behind the compute warps */
#if DFB_STREAM_SETMAXNREG
constexpr int kManagerRegs = 32, kComputeRegs = 96; /* setmaxnreg targets: 8 x 64 = 4 x 96 + 4 x 32 */
#endif
constexpr int kTicketGroups = 16;  /* independent ticket counters: group g hands out tickets g, g + 16, g + 32, .. */
constexpr int kCounterStride = 32; /* words between counters: one 128-byte line each */
constexpr uint32_t kDumpWords = 1u << 16; /* 512 KB of scratch keys for the branchless REDs */

/* cycle accounting of the three roles (printed for a few CTAs when DFB_STREAM_PROFILE=1): compiled in only with
 * -DDFB_STREAM_INSTRUMENT=1, otherwise every timestamp is the constant 0 and the bookkeeping folds away */
#ifndef DFB_STREAM_INSTRUMENT
#define DFB_STREAM_INSTRUMENT 0
#endif
__device__ __forceinline__ long long prof_clock() { return DFB_STREAM_INSTRUMENT ? clock64() : 0ll; }

struct StreamCams {
	SplatCam K[DFB_MAX_CAMERAS];
};

/* n / d for a divisor fixed at launch: q = hi64(n * ceil(2^64 / d)), exact for every 32-bit n (n * d < 2^64). The
 * producer decodes every ticket on ONE thread, where a hardware-less 32-bit division is a ~150-cycle dependent chain. */
struct FastDiv {
	unsigned long long m;
	unsigned d;
	__host__ void init(unsigned div) {
		d = div ? div : 1u;
		m = d == 1u ? 0ull : (~0ull / d + 1ull);
	}
	__device__ __forceinline__ unsigned div(unsigned n) const { return d == 1u ? n : (unsigned)__umul64hi((unsigned long long)n, m); }
};

struct StreamParams {
	FastDiv by_splat, by_resolve, by_round, by_tiles, by_tiles_x; /* splat_per_fs, resolve_per_fs, their sum, tiles_per_image, tiles_x */
	int n_framesets, n_cameras;
	int tiles_x, tiles_per_image;
	int splat_per_fs;   /* n_cameras * tiles_per_image */
	int resolve_per_fs; /* ceil(n_out_pixels / kResolveTile) */
	int lag, slots;
	int total_items;
	uint32_t n_out_pixels;
	uint32_t slot_bytes; /* bytes between key slots */
	uint32_t dump_off;   /* byte offset of the RED scratch area from the ring base */
	int groups;  /* ticket groups in use: min(kTicketGroups, grid) — every group needs at least one CTA */
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
		it.f = (int)Q.by_splat.div(i);
		j = i - (unsigned)it.f * S;
	} else {
		i -= head;
		const unsigned per = S + R;
		const unsigned mid = per * (unsigned)(Q.n_framesets - Q.lag); /* rounds lag..n-1: splat(r) then resolve(r-lag) */
		if (i < mid) {
			const unsigned r = Q.by_round.div(i);
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
			const unsigned r = Q.by_resolve.div(i);
			it.kind = 1;
			it.f = Q.n_framesets - Q.lag + (int)r;
			j = i - r * R;
		}
	}
	if (it.kind == 0) {
		it.a = (int)Q.by_tiles.div(j);
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
__device__ __forceinline__ float claim_ticket(float* p, float n) {
	float v;
	asm volatile("atom.relaxed.gpu.global.add.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "f"(n) : "memory");
	return v;
}

__device__ __forceinline__ unsigned int ld_counter(const unsigned int* p) {
	unsigned int v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

/* fence.proxy.async: orders this thread's (acquired) view of generic-proxy writes — the REDs of other CTAs,
 * published through the frameset counter — before the bulk copies it issues next through the async proxy */
__device__ __forceinline__ void fence_acquire_then_async() {
	asm volatile("fence.proxy.async;" ::: "memory");
}

/* resolve of one key tile that the producer has bulk-copied into the stage: key -> rgb8 (kernels.cu:262-271, colour of
 * the nearest point, black where nothing landed). The L2 latency of the key reads lives in the TMA pipeline; the keys in
 * L2 are re-armed by the retirer (one bulk store per item). */
__device__ __forceinline__ void resolve_item(const StreamParams& Q, int f, int j, uint32_t stage, uint32_t rearm_src,
                                             unsigned long long* __restrict__ keys, uint8_t* __restrict__ out, int tid, long long* prof) {
	const size_t n = Q.n_out_pixels;
	long long tp = prof_clock();
	uint8_t* o = out + (size_t)f * n * 3;
	const size_t tile0 = (size_t)j * kResolveTile;
	/* keys -> packed rgb words in registers; then, through the (now dead) key tile in shared memory, out to global as
	 * whole 16-byte pieces: 12 bytes per thread stored directly are three partial sectors per lane and tripled the write
	 * requests on the L1TEX -> crossbar port */
	constexpr int kRounds = kResolveTile / (kComputeThreads * 4);
	uint32_t w[3 * kRounds];
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const uint32_t l0 = (uint32_t)r * (kComputeThreads * 4) + (uint32_t)tid * 4; /* pixel inside the tile */
		uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
		if (tile0 + l0 + 4 <= n) { /* n is a multiple of 4 (fuse_stream_supported) */
			const uint4 a = lds128(stage + l0 * 8u), b = lds128(stage + l0 * 8u + 16u);
			/* a key is empty iff its high word (the Z bits) is all ones */
			c0 = (a.y == 0xFFFFFFFFu) ? 0u : a.x;
			c1 = (a.w == 0xFFFFFFFFu) ? 0u : a.z;
			c2 = (b.y == 0xFFFFFFFFu) ? 0u : b.x;
			c3 = (b.w == 0xFFFFFFFFu) ? 0u : b.z;
		}
		w[3 * r + 0] = __byte_perm(c0, c1, 0x6012);
		w[3 * r + 1] = __byte_perm(c1, c2, 0x5601);
		w[3 * r + 2] = __byte_perm(c2, c3, 0x4560);
	}
	prof[0] += prof_clock() - tp;
	tp = prof_clock();
	asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory"); /* every key of the tile has been read */
	prof[1] += prof_clock() - tp;
	tp = prof_clock();
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const uint32_t word = (uint32_t)r * (kComputeThreads * 3) + (uint32_t)tid * 3; /* stride 3 words: conflict-free */
#pragma unroll
		for (int k = 0; k < 3; ++k) asm volatile("st.shared.u32 [%0], %1;" ::"r"(stage + 4u * (word + k)), "r"(w[3 * r + k]) : "memory");
	}
	asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
	prof[2] += prof_clock() - tp;
	tp = prof_clock();
	{
		const size_t count = n - tile0 < (size_t)kResolveTile ? n - tile0 : (size_t)kResolveTile;
		const uint32_t bytes = (uint32_t)count * 3u; /* a multiple of 4; out is 16-byte aligned and so is tile0 * 3 */
		uint8_t* dst = o + tile0 * 3;
		for (uint32_t i = (uint32_t)tid * 16u; i + 16u <= bytes; i += (uint32_t)kComputeThreads * 16u) {
			const uint4 v = lds128(stage + i);
			__stcs(reinterpret_cast<uint4*>(dst + i), v);
		}
		if (tid < (int)((bytes & 15u) >> 2)) { /* the last tile of an image may end on a 4-byte boundary */
			const uint32_t i = (bytes & ~15u) + (uint32_t)tid * 4u;
			uint32_t v;
			asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(stage + i));
			__stcs(reinterpret_cast<uint32_t*>(dst + i), v);
		}
	}
	prof[3] += prof_clock() - tp;
	tp = prof_clock();
	prof[4] += prof_clock() - tp;
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
		return inside ? splat_segment_lds<1, 1, DFB_COLS_DUMP>(seq, G, K, dw, cw, xs, yr, u0, row_in_roi, T)
		              : splat_segment_lds<1, 0, DFB_COLS_DUMP>(seq, G, K, dw, cw, xs, yr, u0, row_in_roi, T);
	return inside ? splat_segment_lds<0, 1, DFB_COLS_DUMP>(seq, G, K, dw, cw, xs, yr, u0, row_in_roi, T)
	              : splat_segment_lds<0, 0, DFB_COLS_DUMP>(seq, G, K, dw, cw, xs, yr, u0, row_in_roi, T);
}

/* ---- row-major traversal (un-rotated views) ----------------------------------------------------------------------
 * When source rows map to output ROWS, a warp's atomics are contiguous if its 32 lanes sit side by side on ONE source
 * row: lane = pixel pair (2 lane, 2 lane + 1), warp w of the half owns rows 8w .. 8w+7 and a thread visits one pair per
 * row (8 pairs, like the column form). Operands come out of the same swizzled TMA tile with 4-byte reads:
 *   depth pair  row r, bytes [4 lane, +4)         16-byte chunk (lane >> 2) ^ (r & 7)                     (SWIZZLE_128B)
 *   colour      row r, bytes [6 lane, +6) = two aligned words j0 = (6 lane) >> 2 and j0 + 1; word j sits in box (4j) >> 6,
 *               chunk ((4j & 63) >> 4) ^ ((r >> 1) & 3)                                                   (SWIZZLE_64B)
 * r & 7 == k and (r >> 1) & 3 == (k >> 1) & 3 for the k-th row of a warp (rows start at multiples of 8), so the XOR terms
 * are compile-time constants. */
#if DFB_ROWS_SPLIT
/* Row-major with the pair SPLIT across the tile: lane owns columns (lane, lane + 32) of every row, so that each of the two
 * atomics of a pair covers 32 CONSECUTIVE output pixels (with the (2 lane, 2 lane + 1) pairing each atomic touches every
 * other pixel: twice the sectors). Costs two 2-byte depth reads and four colour words per pair instead of one + two. */
struct RowLaneSplit {
	uint32_t d_off[2];    /* depth halfwords of columns lane, lane + 32 (before the swizzle XOR) */
	uint32_t c_off[2][2]; /* colour words j, j + 1 that hold bytes [3c, 3c + 3) of columns c = lane, lane + 32 */
	uint32_t sh[2];       /* 8 * (3c & 3): funnel shift that brings R to byte 0 */
};

__device__ __forceinline__ RowLaneSplit row_lane_split(uint32_t lane) {
	RowLaneSplit L;
#pragma unroll
	for (int h = 0; h < 2; ++h) {
		const uint32_t c = lane + 32u * (uint32_t)h;
		L.d_off[h] = (((2u * c) >> 4) << 4) | ((2u * c) & 15u);
		const uint32_t p = 3u * c;
		L.sh[h] = 8u * (p & 3u);
#pragma unroll
		for (int i = 0; i < 2; ++i) {
			uint32_t q = 4u * ((p >> 2) + (uint32_t)i);
			if (q >= 3u * kTileW) q -= 4u; /* the word after the row's last one (column 63 only): re-read the last word instead — its
			                                  bytes are shifted out — rather than touch memory outside the colour boxes */
			L.c_off[h][i] = kItemDepthBytes + (q >> 6) * kItemColorBoxBytes + (((q & 63u) >> 4) << 4) + (q & 15u);
		}
	}
	return L;
}

template <int kHasT, int kRoiAll>
__device__ __forceinline__ uint32_t splat_rows_split_loop(const SplatShared& G, const SplatCam& K, uint32_t stage, uint32_t row0,
                                                          const RowLaneSplit& L, float2 xr, int u, int v0, const KeyTarget& T) {
	uint32_t slow = 0;
	DFB_PRAGMA_UNROLL(DFB_ROWS_UNROLL)
	for (int k = 0; k < 8; ++k) {
		const uint32_t r = row0 + (uint32_t)k;
		const uint32_t dx = (uint32_t)k << 4, cx = (uint32_t)((k >> 1) & 3) << 4;
		const uint32_t drow = stage + r * 128u, crow = stage + r * 64u;
		uint32_t da, db, a0, a1, b0, b1;
		float yr;
		asm volatile("ld.shared.u16 %0, [%1];" : "=r"(da) : "r"(drow + (L.d_off[0] ^ dx)));
		asm volatile("ld.shared.u16 %0, [%1];" : "=r"(db) : "r"(drow + (L.d_off[1] ^ dx)));
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(crow + (L.c_off[0][0] ^ cx)));
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a1) : "r"(crow + (L.c_off[0][1] ^ cx)));
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b0) : "r"(crow + (L.c_off[1][0] ^ cx)));
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b1) : "r"(crow + (L.c_off[1][1] ^ cx)));
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(yr) : "r"(stage + kRayYOff + 4u * r));
		const uint32_t dpair = da | (db << 16);
		const uint32_t lo0 = __byte_perm(__funnelshift_r(a0, a1, L.sh[0]), K.cam, 0x4012);
		const uint32_t lo1 = __byte_perm(__funnelshift_r(b0, b1, L.sh[1]), K.cam, 0x4012);
		const int v = v0 + k;
		const bool row_in_roi = (unsigned)(v - K.roi_y0) < (unsigned)K.roi_h;
		splat_pair<kHasT, kRoiAll, DFB_ROWS_DUMP>(G, K, dpair, xr, yr, lo0, lo1, u, row_in_roi, T, slow, 1u << k, 32);
	}
	return slow;
}

__device__ __forceinline__ uint32_t splat_tile_rows_split(const SplatShared& G, const SplatCam& K, uint32_t stage, uint32_t row0,
                                                          const RowLaneSplit& L, float2 xr, int u, int v0, int tile_u0, int tile_v0,
                                                          const KeyTarget& T) {
	const bool inside = K.roi_all || (tile_u0 >= K.roi_x0 && tile_u0 + kTileW <= K.roi_x0 + K.roi_w && tile_v0 >= K.roi_y0 &&
	                                  tile_v0 + kTileH <= K.roi_y0 + K.roi_h);
	if (K.has_t)
		return inside ? splat_rows_split_loop<1, 1>(G, K, stage, row0, L, xr, u, v0, T) : splat_rows_split_loop<1, 0>(G, K, stage, row0, L, xr, u, v0, T);
	return inside ? splat_rows_split_loop<0, 1>(G, K, stage, row0, L, xr, u, v0, T) : splat_rows_split_loop<0, 0>(G, K, stage, row0, L, xr, u, v0, T);
}

#else /* the (2 lane, 2 lane + 1) pairing, kept for comparison */
struct RowLane {
	uint32_t d_off;      /* (lane >> 2) << 4 | (lane & 3) << 2: depth word inside the row, before the swizzle XOR */
	uint32_t c_off[2];   /* colour words j0, j0 + 1: box offset + chunk << 4 + byte inside the chunk, before the XOR */
	uint32_t sel0, sel1; /* PRMT selectors that gather [B, G, R, -] of pixel 0 / 1 from the two colour words */
};

__device__ __forceinline__ RowLane row_lane(uint32_t lane) {
	RowLane L;
	L.d_off = ((lane >> 2) << 4) | ((lane & 3u) << 2);
#pragma unroll
	for (int i = 0; i < 2; ++i) {
		const uint32_t q = 4u * (((6u * lane) >> 2) + (uint32_t)i); /* byte position of the word inside the 192-byte row */
		L.c_off[i] = kItemDepthBytes + (q >> 6) * kItemColorBoxBytes + (((q & 63u) >> 4) << 4) + (q & 15u);
	}
	const bool odd = (lane & 1u) != 0; /* the pair starts at byte 0 (even lane) or 2 (odd lane) of word j0 */
	L.sel0 = odd ? 0x3234u : 0x3012u;
	L.sel1 = odd ? 0x3567u : 0x3345u;
	return L;
}

template <int kHasT, int kRoiAll, int... kK>
__device__ __forceinline__ uint32_t splat_rows_lds(std::integer_sequence<int, kK...>, const SplatShared& G, const SplatCam& K,
                                                   uint32_t stage, uint32_t row0, const RowLane& L, float2 xr, int u, int v0,
                                                   const KeyTarget& T) {
	auto one = [&](auto kc) -> uint32_t {
		constexpr int k = decltype(kc)::value;
		const uint32_t r = row0 + (uint32_t)k;
		uint32_t dpair, wa, wb;
		float yr;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dpair) : "r"(stage + r * 128u + (L.d_off ^ ((uint32_t)k << 4))));
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa) : "r"(stage + r * 64u + (L.c_off[0] ^ ((uint32_t)((k >> 1) & 3) << 4))));
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb) : "r"(stage + r * 64u + (L.c_off[1] ^ ((uint32_t)((k >> 1) & 3) << 4))));
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(yr) : "r"(stage + kRayYOff + 4u * r));
		const uint32_t lo0 = __byte_perm(__byte_perm(wa, wb, L.sel0), K.cam, 0x4210);
		const uint32_t lo1 = __byte_perm(__byte_perm(wa, wb, L.sel1), K.cam, 0x4210);
		const int v = v0 + k;
		const bool row_in_roi = (unsigned)(v - K.roi_y0) < (unsigned)K.roi_h;
		uint32_t sl = 0; splat_pair<kHasT, kRoiAll, 1>(G, K, dpair, xr, yr, lo0, lo1, u, row_in_roi, T, sl, 1u << k); return sl;
	};
	uint32_t slow = 0;
	((slow |= one(std::integral_constant<int, kK>{})), ...);
	return slow;
}

/* the same as a LOOP over the 8 rows (run-time k); the default, see DFB_ROWS_UNROLL */
template <int kHasT, int kRoiAll>
__device__ __forceinline__ uint32_t splat_rows_loop(const SplatShared& G, const SplatCam& K, uint32_t stage, uint32_t row0,
                                                    const RowLane& L, float2 xr, int u, int v0, const KeyTarget& T) {
	uint32_t slow = 0;
	DFB_PRAGMA_UNROLL(DFB_ROWS_UNROLL)
	for (int k = 0; k < 8; ++k) {
		const uint32_t r = row0 + (uint32_t)k;
		uint32_t dpair, wa, wb;
		float yr;
		const uint32_t cx = (uint32_t)((k >> 1) & 3) << 4;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dpair) : "r"(stage + r * 128u + (L.d_off ^ ((uint32_t)k << 4))));
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wa) : "r"(stage + r * 64u + (L.c_off[0] ^ cx)));
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wb) : "r"(stage + r * 64u + (L.c_off[1] ^ cx)));
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(yr) : "r"(stage + kRayYOff + 4u * r));
		const uint32_t lo0 = __byte_perm(__byte_perm(wa, wb, L.sel0), K.cam, 0x4210);
		const uint32_t lo1 = __byte_perm(__byte_perm(wa, wb, L.sel1), K.cam, 0x4210);
		const int v = v0 + k;
		const bool row_in_roi = (unsigned)(v - K.roi_y0) < (unsigned)K.roi_h;
		splat_pair<kHasT, kRoiAll, DFB_ROWS_DUMP>(G, K, dpair, xr, yr, lo0, lo1, u, row_in_roi, T, slow, 1u << k);
	}
	return slow;
}

__device__ __forceinline__ uint32_t splat_tile_rows(const SplatShared& G, const SplatCam& K, uint32_t stage, uint32_t row0,
                                                    const RowLane& L, float2 xr, int u, int v0, int tile_u0, int tile_v0,
                                                    const KeyTarget& T) {
	constexpr auto seq = std::make_integer_sequence<int, 8>{};
	const bool inside = K.roi_all || (tile_u0 >= K.roi_x0 && tile_u0 + kTileW <= K.roi_x0 + K.roi_w && tile_v0 >= K.roi_y0 &&
	                                  tile_v0 + kTileH <= K.roi_y0 + K.roi_h);
#if DFB_ROWS_UNROLL < 8
	(void)seq;
	if (K.has_t)
		return inside ? splat_rows_loop<1, 1>(G, K, stage, row0, L, xr, u, v0, T) : splat_rows_loop<1, 0>(G, K, stage, row0, L, xr, u, v0, T);
	return inside ? splat_rows_loop<0, 1>(G, K, stage, row0, L, xr, u, v0, T) : splat_rows_loop<0, 0>(G, K, stage, row0, L, xr, u, v0, T);
#else
	if (K.has_t)
		return inside ? splat_rows_lds<1, 1>(seq, G, K, stage, row0, L, xr, u, v0, T) : splat_rows_lds<1, 0>(seq, G, K, stage, row0, L, xr, u, v0, T);
	return inside ? splat_rows_lds<0, 1>(seq, G, K, stage, row0, L, xr, u, v0, T) : splat_rows_lds<0, 0>(seq, G, K, stage, row0, L, xr, u, v0, T);
#endif
}

#endif

/* Counters, one per 128-byte line (same-address atomics are served at ~1 per 10 cycles by the L2, and neighbours in a
 * line share that unit): line g < kTicketGroups = ticket counter of CTA group g (float), line kTicketGroups + 2f =
 * done_splat[f], line kTicketGroups + 2f + 1 = done_resolve[f]. */
#ifndef DFB_STREAM_MIN_CTAS
#define DFB_STREAM_MIN_CTAS 3
#endif
/* kRowMajor: which way a warp walks a tile — lane = source row (views rotated by 90 degrees: the default vertical
 * image) or lane = pixel pair of one source row (un-rotated views); chosen so that a warp's atomics are contiguous */
template <bool kRowMajor>
__global__ void __launch_bounds__(kThreads, DFB_STREAM_MIN_CTAS)
fuse_stream_kernel(const __grid_constant__ SplatShared G, const __grid_constant__ StreamCams cams,
                   const __grid_constant__ FuseParams P, const __grid_constant__ StreamParams Q,
                   const __grid_constant__ CUtensorMap map_depth, const __grid_constant__ CUtensorMap map_color,
                   const float* __restrict__ ray_tables, const uint16_t* __restrict__ depth,
                   const uint8_t* __restrict__ color, unsigned long long* __restrict__ keys, uint8_t* __restrict__ out,
                   unsigned int* __restrict__ counters) {
	extern __shared__ uint8_t smem_raw[];
	/* full: producer (+ TMA bytes) -> compute; done: compute -> retirer; free: retirer -> producer. Every wait inside
	 * the CTA is a hardware-suspended mbarrier wait: spin loops on shared-memory flags were ~45 % of all issued
	 * instructions (ncu) and took issue slots from the compute warps. */
	__shared__ __align__(8) unsigned long long full_bar[kStages], free_bar[kStages], done_bar[kRetireDepth];
	__shared__ Item items[kStages];
	/* The retirer runs BEHIND the pipeline: a stage is recycled as soon as the compute warps have left it (free_bar),
	 * while the item's global signal waits in a queue of kRetireDepth entries for the retirer's next fence. done
	 * barriers are indexed by item number modulo kRetireDepth, so the retirer may lag that many items without a
	 * parity alias; s_retired tells the producer how far it got. */
	__shared__ int retire_queue[kRetireDepth]; /* 2 * frameset + kind, -1 = terminator */
	__shared__ int retire_tile[kRetireDepth];  /* resolve items: tile index (the retirer re-arms its keys) */
	__shared__ int s_retired_count;
	volatile int* s_retired = &s_retired_count;
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	if (tid == 0) {
		for (int s = 0; s < kStages; ++s) {
			mbar_init(smem_u32(&full_bar[s]), 1);
			mbar_init(smem_u32(&free_bar[s]), kComputeWarps);
		}
		for (int s = 0; s < kRetireDepth; ++s) mbar_init(smem_u32(&done_bar[s]), kComputeWarps);
		s_retired_count = 0;
		mbar_fence_init();
	}
	for (uint32_t i = (uint32_t)tid * 16u; i < kRearmBytes; i += (uint32_t)kThreads * 16u)
		asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(base + kRearmOff + i), "r"(0xFFFFFFFFu) : "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the bulk stores read this buffer through the async proxy */
	__syncthreads(); /* the only CTA-wide barrier */

	if (warp == kComputeWarps) {
		/* -------------------------------- producer (one lane of warp 4) -------------------------------- */
#if DFB_STREAM_SETMAXNREG
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kManagerRegs));
#endif
		if (lane != 0) return;
		const unsigned int* done_ctr = counters + kTicketGroups * kCounterStride; /* [2f] splat, [2f+1] resolve */
		auto done_splat = [&](int f) { return done_ctr + (size_t)(2 * f) * kCounterStride; };
		auto done_resolve = [&](int f) { return done_ctr + (size_t)(2 * f + 1) * kCounterStride; };
		const uint64_t policy = l2_evict_first_policy();
#ifdef DFB_KEYS_EVICT_LAST
		const uint64_t keep_policy = l2_evict_last_policy();
#endif
		int published = 0, retired_seen = 0;
		/* Tickets are claimed ONE ITEM AHEAD from the counter of this CTA's group; group g owns tickets g, g + G,
		 * g + 2G, .. Measured with ONE global counter: the L2 serves same-address atomics at ~1 per 10 cycles, which
		 * capped the whole kernel at 1520 tickets x 10 cycles per frameset = 130 k framesets/s whatever else changed,
		 * and the claim latency (592 requests queued) exceeded a whole item period. Progress: the smallest unfinished
		 * ticket is either held by a CTA whose dependencies (smaller tickets) are all finished, or it is the next
		 * ticket of its group, whose CTAs hold only finished tickets and are free to claim it.
		 * The counter is a FLOAT (exact below 2^24; the launcher splits larger batches): ptxas warp-aggregates every
		 * integer atom.add — vote + ATOMG + SHFL of the result — and the SHFL waits for the round trip on the spot;
		 * a float add is left alone and its result stays in flight until its first use. */
		const int group = (int)(blockIdx.x % (unsigned)Q.groups);
		float* ticket_counter = reinterpret_cast<float*>(counters) + group * kCounterStride;
		auto global_ticket = [&](float local) { return (int)local * Q.groups + group; };
#ifndef DFB_STREAM_CLAIM_LATE
#define DFB_STREAM_CLAIM_LATE 1
#endif
#if !DFB_STREAM_CLAIM_LATE
		int ticket = global_ticket(claim_ticket(ticket_counter, 1.0f));
		float ahead = claim_ticket(ticket_counter, 1.0f);
#endif
		long long t_depS = 0;
		int n_depS = 0, n_depR = 0;
		int known_resolved = -1, known_splatted = -1;
#if DFB_STREAM_FORMAL_ACQUIRE
		/* the counters are polled with relaxed loads; before the first item that RELIES on a newly confirmed count is
		 * published, one fence.acq_rel.gpu turns the observation into an acquire (it pairs with the retirer's fence +
		 * atomicAdd release). One fence covers everything confirmed so far: at most two per frameset and CTA. */
		int fenced_resolved = -1, fenced_splatted = -1;
#endif
		/* look-ahead polls: every iteration loads the counters of the NEXT framesets this CTA will depend on and
		 * consumes them one item later, so that a satisfied dependency costs no L2 round trip when it is checked */
		int pre_r_idx = -1, pre_s_idx = -1;
		unsigned pre_r_val = 0, pre_s_val = 0;
		unsigned long long idle = 0;
		long long t_claim = 0, t_pub = 0, t_dep = 0, t_stage = 0, t0 = prof_clock(), t1;
		while (true) {
#if DFB_STREAM_CLAIM_LATE
			/* claim LATE: a ticket is only taken once a stage is free for it, so that a CTA never sits on tickets it
			 * cannot start — every held ticket widens the window the dependency distance (slots, lag) must cover */
			t1 = prof_clock();
			if (published >= kStages) mbar_wait(smem_u32(&free_bar[published % kStages]), (uint32_t)(published / kStages - 1) & 1u);
			while (published - retired_seen >= kRetireDepth) { /* rare: the retirer is a whole queue behind */
				retired_seen = *s_retired;
				if (published - retired_seen < kRetireDepth) break;
				__nanosleep(200);
				if (++idle > (1ull << 24)) __trap();
			}
			t_stage += prof_clock() - t1;
			t1 = prof_clock();
			const Item it = decode_item(Q, global_ticket(claim_ticket(ticket_counter, 1.0f)));
			if (pre_r_idx >= 0 && pre_r_val >= (unsigned)Q.resolve_per_fs && pre_r_idx > known_resolved) known_resolved = pre_r_idx;
			if (pre_s_idx >= 0 && pre_s_val >= (unsigned)Q.splat_per_fs && pre_s_idx > known_splatted) known_splatted = pre_s_idx;
			t_claim += prof_clock() - t1;
#else
			t1 = prof_clock();
			const Item it = decode_item(Q, ticket);
			ticket = global_ticket(ahead);
			if (it.kind >= 0) ahead = claim_ticket(ticket_counter, 1.0f);
			if (pre_r_idx >= 0 && pre_r_val >= (unsigned)Q.resolve_per_fs && pre_r_idx > known_resolved) known_resolved = pre_r_idx;
			if (pre_s_idx >= 0 && pre_s_val >= (unsigned)Q.splat_per_fs && pre_s_idx > known_splatted) known_splatted = pre_s_idx;
			t_claim += prof_clock() - t1;
			t1 = prof_clock();
			/* a stage is reusable once the compute warps have left its previous item */
			if (published >= kStages) mbar_wait(smem_u32(&free_bar[published % kStages]), (uint32_t)(published / kStages - 1) & 1u);
			while (published - retired_seen >= kRetireDepth) { /* rare: the retirer is a whole queue behind */
				retired_seen = *s_retired;
				if (published - retired_seen < kRetireDepth) break;
				__nanosleep(200);
				if (++idle > (1ull << 24)) __trap();
			}
			t_stage += prof_clock() - t1;
#endif
			t1 = prof_clock();
			/* the item's dependency: checked once per frameset, then cached */
			if (DFB_STREAM_INSTRUMENT && (Q.profile & 2)) { /* EXPERIMENT (instrumented builds only): ignore dependencies —
			                                                  wrong results — to measure the ceiling without them */
			} else if (it.kind == 0) {
				const int need = it.f - Q.slots; /* key slot re-armed? */
				if (need > known_resolved) {
					const long long td = prof_clock();
					while (ld_counter(done_resolve(need)) < (unsigned)Q.resolve_per_fs) {
						__nanosleep(100);
						if (++idle > (1ull << 24)) __trap(); /* a protocol bug must be an error, not a hung GPU */
					}
					known_resolved = need;
					t_depS += prof_clock() - td;
					++n_depS;
				}
#if DFB_STREAM_FORMAL_ACQUIRE
				if (need > fenced_resolved) {
					asm volatile("fence.acq_rel.gpu;" ::: "memory");
					fenced_resolved = known_resolved;
					fenced_splatted = known_splatted;
				}
#endif
			} else if (it.kind == 1) {
				if (it.f > known_splatted) { /* every camera of frameset f splatted? */
					while (ld_counter(done_splat(it.f)) < (unsigned)Q.splat_per_fs) {
						__nanosleep(100);
						if (++idle > (1ull << 24)) __trap();
					}
					known_splatted = it.f;
					++n_depR;
				}
#if DFB_STREAM_FORMAL_ACQUIRE
				if (it.f > fenced_splatted) {
					asm volatile("fence.acq_rel.gpu;" ::: "memory");
					fenced_resolved = known_resolved;
					fenced_splatted = known_splatted;
				}
#endif
				fence_acquire_then_async();
			}
			/* next look-ahead (results are consumed at the top of the next iteration) */
			pre_r_idx = (known_resolved + 1 < Q.n_framesets && !(DFB_STREAM_INSTRUMENT && (Q.profile & 2))) ? known_resolved + 1 : -1;
			if (pre_r_idx >= 0) pre_r_val = ld_counter(done_resolve(pre_r_idx));
			pre_s_idx = (known_splatted + 1 < Q.n_framesets && !(DFB_STREAM_INSTRUMENT && (Q.profile & 2))) ? known_splatted + 1 : -1;
			if (pre_s_idx >= 0) pre_s_val = ld_counter(done_splat(pre_s_idx));
			idle = 0;
			t_dep += prof_clock() - t1;
			t1 = prof_clock();
			const int s = published % kStages;
			items[s] = it;
			retire_queue[published % kRetireDepth] = it.kind < 0 ? -1 : 2 * it.f + it.kind;
			retire_tile[published % kRetireDepth] = it.a;
			const uint32_t bar = smem_u32(&full_bar[s]);
			if (it.kind == 0) {
				const int ty = (int)Q.by_tiles_x.div((unsigned)it.b);
				const int u0 = (it.b - ty * Q.tiles_x) * kTileW, v0 = ty * kItemRows;
				const int image = it.f * Q.n_cameras + it.a;
				const uint32_t dst = base + (uint32_t)s * kStageStride;
				mbar_expect_tx(bar, kStageTx);
				tma_load_3d(dst, &map_depth, u0, v0, image, bar, policy);
				tma_load_3d(dst + kItemDepthBytes, &map_color, u0 * 3, v0, image, bar, policy);
				tma_load_3d(dst + kItemDepthBytes + kItemColorBoxBytes, &map_color, u0 * 3 + 64, v0, image, bar, policy);
				tma_load_3d(dst + kItemDepthBytes + 2 * kItemColorBoxBytes, &map_color, u0 * 3 + 128, v0, image, bar, policy);
				bulk_load_1d(dst + kRayXOff, ray_tables + it.a * G.x_pitch + u0, kTileW * 4, bar);
				bulk_load_1d(dst + kRayYOff, ray_tables + G.y_base + it.a * G.y_pitch + v0, kItemRows * 4, bar);
			} else if (it.kind == 1) { /* the key tile of the resolve item */
				const uint32_t first = (uint32_t)it.a * kResolveTile;
				const uint32_t count = Q.n_out_pixels - first < (uint32_t)kResolveTile ? Q.n_out_pixels - first : (uint32_t)kResolveTile;
				const char* src = reinterpret_cast<const char*>(keys) + (size_t)(it.f % Q.slots) * Q.slot_bytes + (size_t)first * 8;
				mbar_expect_tx(bar, count * 8u);
#ifdef DFB_KEYS_EVICT_LAST
				bulk_load_1d_hint(base + (uint32_t)s * kStageStride, src, count * 8u, bar, keep_policy);
#else
				bulk_load_1d(base + (uint32_t)s * kStageStride, src, count * 8u, bar);
#endif
			} else {
				mbar_arrive(bar); /* terminator: nothing to load */
			}
			t_pub += prof_clock() - t1;
			if (it.kind < 0) break; /* the terminator: the compute warps pass it on to the retirer and stop */
			++published;
		}
		if (DFB_STREAM_INSTRUMENT && (Q.profile & 1) && (blockIdx.x % 149) == 0)
			printf("cta %4d producer: items %d total %lld claim+decode %lld wait-stage %lld wait-dependency %lld (splat-on-resolve %lld, %d+%d checks) publish %lld\n",
			       blockIdx.x, published, prof_clock() - t0, t_claim, t_stage, t_dep, t_depS, n_depS, n_depR, t_pub);
		return;
	}
	if (warp == kComputeWarps + 1) {
		/* -------------------------------- retirer (one lane of warp 5) --------------------------------- */
#if DFB_STREAM_SETMAXNREG
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kManagerRegs));
#endif
		if (lane != 0) return;
		unsigned int* done_ctr = counters + kTicketGroups * kCounterStride; /* [2f] splat, [2f+1] resolve */
		int retired = 0;
		unsigned long long idle = 0;
		long long t_wait = 0, t_signal = 0, t0 = prof_clock(), t1;
		int fences = 0;
		/* Finished items free their stage at once; their global signals are BATCHED: one gpu-scope fence (~3000
		 * cycles while the SM's atomics drain) covers up to kSignalBatch items, whose per-frameset counts sit in a
		 * tiny pending list. The delay is far below the distance of the dependencies (lag framesets). */
#ifndef DFB_STREAM_SIGNAL_BATCH
#define DFB_STREAM_SIGNAL_BATCH 2
#endif
		constexpr int kPend = 4, kSignalBatch = DFB_STREAM_SIGNAL_BATCH;
		int pend_idx[kPend], pend_cnt[kPend], n_pend = 0, pend_items = 0;
		auto flush = [&]() {
			/* the re-arm bulk stores of the pending resolve items must have landed before their frameset is reported */
			asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
			asm volatile("fence.proxy.async;" ::: "memory");
			asm volatile("fence.acq_rel.gpu;" ::: "memory"); /* cumulative: everything the compute warps wrote before arriving on done[] */
#pragma unroll
			for (int i = 0; i < kPend; ++i)
				if (i < n_pend) atomicAdd(done_ctr + (size_t)pend_idx[i] * kCounterStride, (unsigned)pend_cnt[i]);
			n_pend = pend_items = 0;
			++fences;
		};
		while (true) {
			t1 = prof_clock();
			const uint32_t bar = smem_u32(&done_bar[retired % kRetireDepth]), parity = (uint32_t)(retired / kRetireDepth) & 1u;
			/* wait for the oldest outstanding item (the compute warps finish items in publication order); whenever
			 * ~1 us passes without one, the pending signals are flushed */
			while (!mbar_try_wait_ns(bar, parity, 1000u)) {
				if (pend_items) {
					t_wait += prof_clock() - t1;
					t1 = prof_clock();
					flush();
					t_signal += prof_clock() - t1;
					t1 = prof_clock();
				}
				__nanosleep(200);
				if (++idle > (1ull << 24)) __trap();
			}
			idle = 0;
			t_wait += prof_clock() - t1;
			t1 = prof_clock();
			const int idx = retire_queue[retired % kRetireDepth];
			if (idx < 0) break; /* terminator */
			if (idx & 1) { /* a resolve item: its key tile has been consumed from shared memory — re-arm the keys in L2 for the
			                  slot's next frameset with ONE bulk store of the constant KEY_EMPTY tile (shared -> global). With
			                  per-thread stores in the compute warps the 20 KB per item went through the L1TEX -> crossbar port
			                  the atomics keep 70 % busy; issued here, its ~1800-cycle latency is off their critical path. */
				const int rf = idx >> 1;
				const size_t tile0 = (size_t)retire_tile[retired % kRetireDepth] * kResolveTile;
				const size_t count = (size_t)Q.n_out_pixels - tile0 < (size_t)kResolveTile ? (size_t)Q.n_out_pixels - tile0 : (size_t)kResolveTile;
				char* k = reinterpret_cast<char*>(keys) + (size_t)(rf % Q.slots) * Q.slot_bytes + tile0 * 8;
				for (uint32_t done = 0, bytes = (uint32_t)(count * 8); done < bytes; done += kRearmBytes) {
					const uint32_t piece = bytes - done < kRearmBytes ? bytes - done : kRearmBytes;
					asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(k + done), "r"(base + kRearmOff), "r"(piece) : "memory");
				}
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			}
			++retired;
			*s_retired = retired;
			int slot = -1;
#pragma unroll
			for (int i = 0; i < kPend; ++i)
				if (i < n_pend && pend_idx[i] == idx) slot = i;
			if (slot < 0 && n_pend == kPend) flush();
#pragma unroll
			for (int i = 0; i < kPend; ++i) {
				if (slot < 0 && i == n_pend) {
					pend_idx[i] = idx;
					pend_cnt[i] = 0;
					slot = i;
				}
			}
			if (slot == n_pend) ++n_pend;
#pragma unroll
			for (int i = 0; i < kPend; ++i)
				if (i == slot) ++pend_cnt[i];
			if (++pend_items >= kSignalBatch) flush();
			t_signal += prof_clock() - t1;
		}
		if (pend_items) flush();
		if (DFB_STREAM_INSTRUMENT && (Q.profile & 1) && (blockIdx.x % 149) == 0)
			printf("cta %4d retirer: items %d fences %d total %lld wait-done %lld fence+signal %lld\n", blockIdx.x, retired, fences,
			       prof_clock() - t0, t_wait, t_signal);
		return;
	}

#if DFB_STREAM_SETMAXNREG
	if (warp > kComputeWarps + 1) { /* the two filler warps of the manager warpgroup */
		asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kManagerRegs));
		return;
	}
#endif
	/* ---------------------------------------- compute warps ---------------------------------------- */
#if DFB_STREAM_SETMAXNREG
	/* register hand-over (4 CTAs/SM = 80 registers at launch): the two single-lane manager warps shrink to kManagerRegs
	 * and the four compute warps grow to kComputeRegs (6 x 80 >= 4 x 96 + 2 x 40). The instructions sit INSIDE the role
	 * branches, which never rejoin: ptxas applies the smallest limit to everything downstream of a merge point. */
	asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kComputeRegs));
#endif
	/* byte offsets, inside a stage, of the five 16-byte pieces a thread (lane = row of a half, warp = 16-pixel column
	 * chunk) reads, with the TMA swizzles undone; both swizzle patterns repeat every 8 rows, so the second half is the
	 * same offsets plus 32 rows */
	TileReadOffsets off;
	{
		const uint32_t r = (uint32_t)lane, w = (uint32_t)warp;
		off.d[0] = r * 128u + (((2u * w + 0u) ^ (r & 7u)) << 4);
		off.d[1] = r * 128u + (((2u * w + 1u) ^ (r & 7u)) << 4);
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const uint32_t p = 48u * w + 16u * k; /* byte position inside the 192-byte colour row */
			const uint32_t box = p >> 6, chunk = (p & 63u) >> 4;
			off.c[k] = kItemDepthBytes + box * kItemColorBoxBytes + r * 64u + ((chunk ^ ((r >> 1) & 3u)) << 4);
		}
	}
#ifdef DFB_KEYS_EVICT_LAST
	const uint64_t keep_policy = l2_evict_last_policy();
#else
	const uint64_t keep_policy = 0;
#endif
#if DFB_ROWS_SPLIT
	const RowLaneSplit rls = row_lane_split((uint32_t)lane);
#else
	const RowLane rl = row_lane((uint32_t)lane);
#endif
	long long c_wait = 0, c_splat = 0, c_resolve = 0, c0 = prof_clock();
	long long r_prof[5] = {0, 0, 0, 0, 0}; /* resolve phases: read+convert, barrier, stage+barrier, store, re-arm wait */
	int n_splat = 0, n_resolve = 0;
	for (int k = 0;; ++k) {
		const int s = k % kStages;
		long long tc = prof_clock();
		mbar_wait(smem_u32(&full_bar[s]), (uint32_t)(k / kStages) & 1u);
		c_wait += prof_clock() - tc;
		tc = prof_clock();
		const Item it = items[s];
		if (it.kind < 0) { /* terminator: hand it to the retirer */
			if (lane == 0) mbar_arrive(smem_u32(&done_bar[k % kRetireDepth]));
			break;
		}
		if (it.kind == 0) {
			const uint32_t stage = base + (uint32_t)s * kStageStride;
			const int ty = (int)Q.by_tiles_x.div((unsigned)it.b);
			const int u0 = (it.b - ty * Q.tiles_x) * kTileW + warp * kSeg;
			const uint32_t xs = stage + kRayXOff + (uint32_t)warp * (kSeg * 4);
			const uint32_t slot_off = (uint32_t)(it.f % Q.slots) * Q.slot_bytes;
			unsigned long long* kslot = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(keys) + slot_off);
			/* invalid / out-of-image points dump their RED into a per-thread word of the scratch area behind the ring */
			const KeyTarget T{reinterpret_cast<char*>(keys), slot_off + G.px_origin,
			                  Q.dump_off + (((uint32_t)blockIdx.x * kComputeThreads + (uint32_t)tid) & (kDumpWords - 1)) * 8u};
#pragma unroll 1
			for (int half = 0; half < kItemRows / kTileH; ++half) {
				const int tv0 = ty * kItemRows + half * kTileH;
				if (tv0 >= G.height) break; /* the item hangs over the bottom of the image: nothing but zero fill there */
				const int tu0 = u0 - warp * kSeg;
				uint32_t slow;
				if constexpr (kRowMajor) {
					const uint32_t row0 = (uint32_t)half * kTileH + (uint32_t)warp * 8u; /* this warp's 8 rows inside the item */
					const int v0 = tv0 + warp * 8;
#if DFB_ROWS_SPLIT
					float2 xr;
					asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xr.x) : "r"(stage + kRayXOff + 4u * (uint32_t)lane));
					asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xr.y) : "r"(stage + kRayXOff + 4u * (uint32_t)(lane + 32)));
					const int u = tu0 + lane, du = 32;
					switch (it.a) {
						case 0: slow = splat_tile_rows_split(G, cams.K[0], stage, row0, rls, xr, u, v0, tu0, tv0, T); break;
						case 1: slow = splat_tile_rows_split(G, cams.K[1], stage, row0, rls, xr, u, v0, tu0, tv0, T); break;
						default: slow = splat_tile_rows_split(G, cams.K[it.a], stage, row0, rls, xr, u, v0, tu0, tv0, T); break;
					}
#else
					float2 xr;
					asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(xr.x), "=f"(xr.y) : "r"(stage + kRayXOff + 8u * (uint32_t)lane));
					const int u = tu0 + 2 * lane, du = 1;
					switch (it.a) {
						case 0: slow = splat_tile_rows(G, cams.K[0], stage, row0, rl, xr, u, v0, tu0, tv0, T); break;
						case 1: slow = splat_tile_rows(G, cams.K[1], stage, row0, rl, xr, u, v0, tu0, tv0, T); break;
						default: slow = splat_tile_rows(G, cams.K[it.a], stage, row0, rl, xr, u, v0, tu0, tv0, T); break;
					}
#endif
					if (slow) { /* rare: operands outside the guarded fast-division range; bit k = the pair in row v0 + k */
						const size_t img = ((size_t)it.f * Q.n_cameras + it.a) * ((size_t)G.width * G.height);
						while (slow) {
							const int j = __ffs(slow) - 1;
							slow &= slow - 1;
							splat_pixel_slow(P, it.a, depth + img, color + img * 3, u, v0 + j, kslot);
							splat_pixel_slow(P, it.a, depth + img, color + img * 3, u + du, v0 + j, kslot);
						}
					}
				} else {
					const uint32_t hd = stage + (uint32_t)half * (kTileH * 128u), hc = stage + (uint32_t)half * (kTileH * 64u);
					const uint4 d0 = lds128(hd + off.d[0]), d1 = lds128(hd + off.d[1]);
					const uint4 c0 = lds128(hc + off.c[0]), c1 = lds128(hc + off.c[1]), c2 = lds128(hc + off.c[2]);
					float yr;
					asm volatile("ld.shared.f32 %0, [%1];" : "=f"(yr) : "r"(stage + kRayYOff + 4u * (uint32_t)(half * kTileH + lane)));
					const int v = tv0 + lane;
					const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
					const uint32_t cw[12] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
					/* One copy of the tile body per camera index: with a STATIC index the camera's constants sit at fixed
					 * constant-bank addresses and feed FFMA2 straight from uniform registers; a run-time index costs an
					 * LDC + R2UR per use (~18 % more instructions, measured in SASS). Cameras >= 2 take the generic copy. */
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
				}
			}
			c_splat += prof_clock() - tc;
			++n_splat;
		} else {
			resolve_item(Q, it.f, it.a, base + (uint32_t)s * kStageStride, base + kRearmOff, keys, out, tid, r_prof);
			c_resolve += prof_clock() - tc;
			++n_resolve;
		}
		__syncwarp();
		if (lane == 0) {
			mbar_arrive(smem_u32(&done_bar[k % kRetireDepth])); /* release: everything this warp wrote for the item */
			mbar_arrive(smem_u32(&free_bar[s]));                 /* and it no longer reads the stage */
		}
	}
	if (DFB_STREAM_INSTRUMENT && (Q.profile & 1) && (blockIdx.x % 149) == 0 && lane == 0 && (warp == 1 || warp == 0))
		printf("cta %4d compute (warp %d): total %lld wait-full %lld splat %lld (%d tiles) resolve %lld (%d tiles: read %lld bar %lld stage %lld store %lld rearm-wait %lld)\n",
		       blockIdx.x, warp, prof_clock() - c0, c_wait, c_splat, n_splat, c_resolve, n_resolve, r_prof[0], r_prof[1], r_prof[2], r_prof[3], r_prof[4]);
}

}  // namespace

bool fuse_stream_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color, const uint8_t* d_out) {
	return pools_fit_tensor_maps(P.cam[0].width, d_depth, d_color) && ((size_t)P.n_out_pixels * 8 < (1ull << 32)) &&
	       (P.n_out_pixels % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15u) == 0) && (((size_t)P.n_out_pixels * 3) % 16 == 0);
}

size_t fuse_stream_counter_words(int n_framesets) { return (size_t)kCounterStride * (kTicketGroups + 2 * (size_t)n_framesets); }

/* bytes the stream kernel needs BEHIND the `slots` key slots (scratch for the branchless REDs) */
size_t fuse_stream_scratch_bytes() { return (size_t)kDumpWords * 8; }

/* d_keys holds `slots` frameset slots (all-ones on entry and on exit); d_counters = fuse_stream_counter_words(n) */
long long fuse_stream_tickets_per_frameset(const FuseParams& P) {
	return (long long)((P.cam[0].width * (long long)P.cam[0].height + 2047) / 2048 + 64) * P.n_cameras + (P.n_out_pixels + 1023) / 1024 + 8;
}

int launch_fuse_stream(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                       unsigned long long* d_keys, int ring_slots, unsigned int* d_counters, int n_framesets, int slots, int lag,
                       bool row_major, const StreamLayout* layout, cudaStream_t s) {
	DFB_REQUIRE(layout == nullptr, "the round-1 stream kernel only takes packed pools");
	const char* ce = std::getenv("DFB_STREAM_CTAS_PER_SM");
	const int ctas_per_sm = ce ? std::atoi(ce) : 0;
	const int w = P.cam[0].width, h = P.cam[0].height;
	alignas(64) CUtensorMap md, mc;
	int rc = make_pool_tensor_maps(d_depth, d_color, w, h, (uint64_t)n_framesets * P.n_cameras, &md, &mc, kItemRows);
	if (rc != DFB_OK) return rc;
	StreamParams Q;
	Q.n_framesets = n_framesets;
	Q.n_cameras = P.n_cameras;
	Q.tiles_x = (w + kTileW - 1) / kTileW;
	Q.tiles_per_image = Q.tiles_x * ((h + kItemRows - 1) / kItemRows);
	Q.splat_per_fs = Q.tiles_per_image * P.n_cameras;
	Q.resolve_per_fs = (P.n_out_pixels + kResolveTile - 1) / kResolveTile;
	Q.by_splat.init((unsigned)Q.splat_per_fs);
	Q.by_resolve.init((unsigned)Q.resolve_per_fs);
	Q.by_round.init((unsigned)(Q.splat_per_fs + Q.resolve_per_fs));
	Q.by_tiles.init((unsigned)Q.tiles_per_image);
	Q.by_tiles_x.init((unsigned)Q.tiles_x);
	if (slots < 1) slots = 1;
	if (lag > slots - 1) lag = slots - 1; /* R(f-lag) must be queued before S(f-lag+slots) */
	if (lag > n_framesets) lag = n_framesets;
	if (lag < 0) lag = 0;
	Q.lag = lag;
	Q.slots = slots;
	Q.n_out_pixels = (uint32_t)P.n_out_pixels;
	Q.slot_bytes = (uint32_t)P.n_out_pixels * 8u;
	if (slots > ring_slots) slots = ring_slots;
	Q.slots = slots;
	Q.dump_off = Q.slot_bytes * (uint32_t)ring_slots;
	DFB_REQUIRE((unsigned long long)P.n_out_pixels * 8ull * (unsigned long long)ring_slots + fuse_stream_scratch_bytes() < (1ull << 32),
	            "key ring too large for 32-bit offsets: use fewer slots");
	const long long total = (long long)n_framesets * (Q.splat_per_fs + Q.resolve_per_fs);
	DFB_REQUIRE(total + 65536 < (1ll << 24) * kTicketGroups, "too many work items for one launch (float ticket counters): split the batch");
	Q.total_items = (int)total;
	{
		const char* e = std::getenv("DFB_STREAM_PROFILE");
		Q.profile = e ? std::atoi(e) : 0;
	}
	StreamCams cams;
	for (int c = 0; c < DFB_MAX_CAMERAS; ++c) cams.K[c] = make_splat_cam(P, c < P.n_cameras ? c : 0, rays);
	SplatShared G = make_splat_shared(P);
	G.x_pitch = rays.x_pitch;
	G.y_base = (int)(rays.y - rays.x);
	G.y_pitch = rays.y_pitch;

	static int max_ctas = 0, sms = 0;
	if (max_ctas == 0) {
		DFB_CUDA(cudaFuncSetAttribute(fuse_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
		DFB_CUDA(cudaFuncSetAttribute(fuse_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
		int dev = 0, a = 0, b = 0;
		DFB_CUDA(cudaGetDevice(&dev));
		DFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
		DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, fuse_stream_kernel<false>, kThreads, kSmemBytes));
		DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, fuse_stream_kernel<true>, kThreads, kSmemBytes));
		max_ctas = a < b ? a : b;
		if (max_ctas < 1) max_ctas = 1;
	}
	const int per_sm = (ctas_per_sm > 0 && ctas_per_sm < max_ctas) ? ctas_per_sm : max_ctas;
	long long grid = (long long)sms * per_sm;
	if (grid > total) grid = total;
	Q.groups = grid < kTicketGroups ? (int)grid : kTicketGroups;
	DFB_CUDA(cudaMemsetAsync(d_counters, 0, fuse_stream_counter_words(n_framesets) * sizeof(unsigned int), s));
	/* (A persisting-L2 access-policy window over the key ring was measured and rejected: 126 k -> 107 k framesets/s,
	 * 2.6 GB of DRAM writes per 128 framesets. The ring stays in L2 on its own as long as it is <= ~46 MB.) */
	if (row_major)
		fuse_stream_kernel<true><<<(unsigned)grid, kThreads, kSmemBytes, s>>>(G, cams, P, Q, md, mc, rays.x, d_depth, d_color, d_keys, d_out,
		                                                                      d_counters);
	else
		fuse_stream_kernel<false><<<(unsigned)grid, kThreads, kSmemBytes, s>>>(G, cams, P, Q, md, mc, rays.x, d_depth, d_color, d_keys,
		                                                                       d_out, d_counters);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

}  // namespace dfb
#endif /* DFB_STREAM_V1 */
