/*
 * fuse_stream.cu — the whole batch in ONE persistent, warp-specialised launch: the splat of every camera of
 * every frameset AND the resolve of every fused image, streamed through a small ring of L2-resident key slots.
 *
 *   work list, by round r = 0 .. n + lag - 1:   S(r) R(r - lag)
 *       S(f) = the input items (64 x 64 source pixels) of all cameras of frameset f   -> atomicMin into slot f % slots
 *       R(f) = the output tiles (2048 pixels) of fused image f: key -> rgb8                <- slot f % slots
 *   A slot is NOT re-armed after every use: the keys of its e-th use since the last re-arm carry an epoch bias in their
 *   high word that puts them below everything earlier uses left behind (dfb_math.cuh key_epoch_bias), the resolve
 *   takes a key as present iff it lies below the epoch's limit, and only every third use ends with a re-arm.
 *   (every round has S + R ticket numbers; those of S(f >= n) and R(f < 0) are empty and skipped.) Items are
 *   CLAIMED from ticket counters in list order (16 counters, group g hands out tickets g, g+16, ..), so an item only
 *   ever depends on items with smaller tickets, which are held by CTAs that are running — a CTA picks its group when
 *   it STARTS (arrival counter), so every group that has handed out a ticket has a live CTA — or are next in line of
 *   their group: no deadlock whatever the residency. Two counters per frameset carry the dependencies:
 *       R(f) needs every S(f) item done           (done_splat[f]         == S)
 *       S(f) needs R(f - slots) done              (done_resolve[f-slots] == R)
 *
 * kCtasPerSm CTAs per SM, each kW compute warps + ONE producer warp + ONE retirer warp (one lane each) around a ring
 * of kStages shared-memory stages (default 2 x (8 + 2) warps x 3 stages):
 *   producer: tickets are claimed ONE AHEAD (the atomic's round trip overlaps the previous item), decoded with two
 *             exact multiply-shift divisions, the item's dependency is checked (counters polled with ld.relaxed.gpu one
 *             item ahead; the first item that relies on a newly observed count is preceded by ONE fence.acq_rel.gpu,
 *             which makes the observation an acquire in the PTX memory model, plus a proxy fence before key tiles are
 *             bulk-copied) and the bulk copies are started: four TMA tensor boxes + one 512-byte row of rays for a splat
 *             item, the 16 KB key tile for a resolve item (mbarrier full[s], expect_tx).
 *   compute : every item is 8 UNITS (splat: 16 columns x 32 rows; resolve: 256 pixels, warp-local, no CTA barrier);
 *             warp w owns units w, w + kW, .. of every item, waits for the stage, runs them and arrives ONCE, on
 *             done[item % kRetireDepth] — the retirer's signal and, for the producer, the release of the stage.
 *             On a slot's third use warp 0 re-arms the keys of a resolve item as soon as its tile is in the stage: one
 *             bulk store (shared -> global) of a constant KEY_EMPTY tile, complete before the warp reports the item.
 *   retirer : runs behind the pipeline: gathers every item that has finished and publishes them with one
 *             red.release.gpu per frameset counter (MEMBAR.ALL.GPU + RED, ~3000 cycles while the SM's atomics drain;
 *             cumulative over everything the compute warps wrote before they arrived).
 * What round 2 measured about this structure (profiles/r02_stream_ncu.md): the splat body alone, run repeatedly over a
 * resident stage, sustains ~270 k framesets/s with 12 warps per SM and does not get faster with 16 or 20; the whole
 * kernel landed at 150-160 k whatever the split (170 k with the ring epochs) (1 x 16 warps x 6 stages with warps pulling units from a shared counter,
 * 2 x 8 x 3, 3 x 4 x 2), because everything between two dependent framesets — stages in flight, the retirer's fence, the
 * counter round trip — has to fit into the slack of the key ring (Little's law: in-flight tickets + latency x rate <
 * distance between S(f) and R(f - slots)); more ring slots buy slack until the ring leaves L2.
 * Every wait is bounded (trap, never a hung GPU). Every key is identical to the other variants.
 */
#include <climits>
#include <cstdio>
#include <mutex>

#include "splat_core.cuh"
#include "tma_utils.cuh"

namespace dfb {
namespace {

/* CTAs per SM x compute warps per CTA x stages per CTA. One producer thread publishes an item in ~2000 cycles (claim,
 * decode, five bulk-copy issues, one acquire fence per frameset), an SM consumes one every ~1500: with ONE CTA per SM
 * the producer is the bottleneck (measured: 134 k framesets/s whatever the number of compute warps or stages), so an SM
 * runs two half-size CTAs, each with its own producer, retirer and ticket stream. */
#ifndef DFB_STREAM_CTAS
#define DFB_STREAM_CTAS 2
#endif
#ifndef DFB_STREAM_WARPS
#define DFB_STREAM_WARPS (DFB_STREAM_CTAS == 1 ? 16 : DFB_STREAM_CTAS == 2 ? 8 : 4)
#endif
#ifndef DFB_STREAM_STAGES
#define DFB_STREAM_STAGES (DFB_STREAM_ITEM_ROWS == 128 ? 2 : DFB_STREAM_CTAS == 1 ? 6 : DFB_STREAM_CTAS == 2 ? 3 : 2)
#endif
constexpr int kCtasPerSm = DFB_STREAM_CTAS;



constexpr int kW = DFB_STREAM_WARPS; /* compute warps per CTA */
constexpr int kStages = DFB_STREAM_STAGES;
constexpr int kThreads = (kW + 2) * 32; /* + producer warp + retirer warp */
constexpr int kItemRows = kStreamItemRows;          /* source rows of a splat item */
constexpr int kSplatUnits = 4 * (kItemRows / kTileH); /* units per splat item: 16 columns x 32 rows (column form), 8 rows (row-major) */
constexpr int kResolveUnits = 8;                    /* units per resolve item */
constexpr int kRetireDepth = 2 * kStages; /* items the retirer may lag behind the producer */
static_assert(kSplatUnits % kW == 0 && kResolveUnits % kW == 0 && kW % 4 == 0,
              "warp w owns units w, w + kW, .. of every item (same column group, kW / 4 row groups further down)");

/* A splat item is 64 x 64 source pixels that arrive with ONE set of bulk copies — a 64x64 uint16 depth box
 * (SWIZZLE_128B) + three 64-byte x 64-row colour boxes (SWIZZLE_64B) + 64 column rays + 64 row rays. */
constexpr uint32_t kItemDepthBytes = kTileW * 2 * kItemRows;                 /* 8192 */
constexpr uint32_t kItemColorBoxBytes = 64 * kItemRows;                      /* 4096 */
constexpr uint32_t kItemBytes = kItemDepthBytes + 3 * kItemColorBoxBytes;    /* 20480 */
constexpr uint32_t kRayXOff = kItemBytes;                 /* 256 B */
constexpr uint32_t kRayYOff = kItemBytes + kTileW * 4;    /* 256 B */
constexpr uint32_t kStageTx = kItemBytes + (kTileW + kItemRows) * 4;
constexpr int kResolveUnit = 256;                     /* output pixels per resolve unit: 32 lanes x 4 pixels x 2 rounds */
constexpr int kResolveTile = kResolveUnits * kResolveUnit;   /* 2048: its keys (16 KB) are bulk-copied into the stage like an input tile */
constexpr uint32_t kStageStride = (kItemRows == 64 ? 21 : 41) * 1024;          /* >= max(input item + rays, key tile); keeps every depth box 1024-byte aligned */
static_assert(kResolveTile * 8 <= (int)kStageStride && kStageTx <= kStageStride, "stage too small");
static_assert(kRayYOff == kRayXOff + kTileW * 4, "column rays and row rays arrive as one 512-byte copy");
/* behind the stages: one key tile of KEY_EMPTY, the source of the bulk stores that re-arm the ring */
constexpr uint32_t kRearmOff = kStages * kStageStride;
constexpr uint32_t kRearmBytes = kResolveTile * 8;
constexpr uint32_t kSmemBytes = kStages * kStageStride + kRearmBytes + 1024;

constexpr int kTicketGroups = 16;  /* independent ticket counters: group g hands out tickets g, g + 16, g + 32, .. */
constexpr int kCounterStride = 32; /* words between counters: one 128-byte line each */
constexpr int kFirstFramesetLine = 1 + kTicketGroups; /* line 0 = CTA arrival counter, 1..16 = ticket counters */
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

/* n / d for a divisor fixed at launch, exact for EVERY 32-bit n (Granlund-Montgomery round-up method): the producer
 * decodes every ticket on one thread, where a hardware-less division is a ~150-cycle dependent chain. */
struct Div32 {
	uint32_t m, s, one;
	__host__ void init(uint32_t d) {
		m = s = 0;
		one = d <= 1u;
		if (one) return;
		uint32_t l = 0;
		while ((1ull << l) < d) ++l; /* ceil(log2 d) >= 1 */
		m = (uint32_t)((((1ull << l) - d) << 32) / d + 1ull);
		s = l - 1u;
	}
	__device__ __forceinline__ uint32_t div(uint32_t n) const {
		const uint32_t t = __umulhi(m, n);
		return one ? n : (t + ((n - t) >> 1)) >> s;
	}
};

struct StreamParams {
	Div32 by_round, by_tiles, by_tiles_x; /* splat_per_fs + resolve_per_fs, tiles_per_image, tiles_x */
	Div32 by_slots;
	uint32_t phase0; /* ring phase of the launch's first frameset: frameset f uses slot (phase0 + f) % slots in epoch
	                  * ((phase0 + f) / slots) % kKeyEpochs (dfb_math.cuh); the fuser carries the phase from launch to launch */
	int n_framesets, n_cameras;
	int tiles_x, tiles_per_image;
	int splat_per_fs;   /* n_cameras * tiles_per_image */
	int resolve_per_fs; /* ceil(n_out_pixels / kResolveTile) */
	int lag, slots;
	int total_tickets;  /* (n_framesets + lag) * (splat_per_fs + resolve_per_fs) */
	uint32_t n_out_pixels;
	uint32_t slot_bytes; /* bytes between key slots */
	uint32_t dump_off;   /* byte offset of the RED scratch area from the ring base (behind the ALLOCATED ring) */
	/* where image (frameset f, camera c) sits: depth pool image f * imgs_per_fs + slot_d[c], colour pool image
	 * f * imgs_per_fs + slot_c[c]; a packed pool has slot[c] = c and images W*H elements apart, cameras that live in
	 * separate allocations (Frameset objects) are described as a 2-image pool with the allocations' distance as stride */
	int imgs_per_fs;
	unsigned char slot_d[DFB_MAX_CAMERAS], slot_c[DFB_MAX_CAMERAS];
	size_t depth_img_stride, color_img_stride; /* in elements (uint16) / bytes */
	int groups;  /* ticket groups in use: min(kTicketGroups, grid) */
	int profile; /* DFB_STREAM_PROFILE (instrumented builds): a few CTAs print where their cycles went */
	int repeat;  /* DFB_EXPERIMENTS builds: the splat body of every unit runs this many times (idempotent: same keys) */
};

/* a claimed work item, decoded by the producer and published in shared memory (one 16-byte store / load) */
struct __align__(16) Item {
	short kind;                /* 0 splat, 1 resolve, -1 terminator, -2 empty ticket */
	unsigned char slot, epoch; /* ring slot of the frameset and which use of the slot since its last re-arm this is */
	int f;                     /* frameset */
	int a, b; /* splat: camera, tile index inside the image; resolve: tile index, unused */
};

__device__ __forceinline__ Item decode_item(const StreamParams& Q, int ticket) {
	Item it;
	it.kind = -1;
	it.slot = it.epoch = 0;
	it.f = it.a = it.b = 0;
	if (ticket >= Q.total_tickets) return it;
	const uint32_t per = (uint32_t)(Q.splat_per_fs + Q.resolve_per_fs);
	const uint32_t r = Q.by_round.div((uint32_t)ticket);
	uint32_t j = (uint32_t)ticket - r * per;
	if (j < (uint32_t)Q.splat_per_fs) {
		it.f = (int)r;
		it.kind = it.f < Q.n_framesets ? 0 : -2;
		it.a = (int)Q.by_tiles.div(j);
		it.b = (int)(j - (uint32_t)it.a * (uint32_t)Q.tiles_per_image);
	} else {
		it.f = (int)r - Q.lag;
		it.kind = it.f >= 0 ? 1 : -2;
		it.a = (int)(j - (uint32_t)Q.splat_per_fs);
	}
	if (it.kind >= 0) {
		const uint32_t g = Q.phase0 + (uint32_t)it.f, round = Q.by_slots.div(g);
		it.slot = (unsigned char)(g - round * (uint32_t)Q.slots);
		it.epoch = (unsigned char)(round % (uint32_t)kKeyEpochs);
	}
	return it;
}

/* The ticket counter is a FLOAT (exact below 2^24; the launcher splits larger batches): ptxas warp-aggregates every
 * integer atom.add — vote + ATOMG + SHFL of the result — and the SHFL waits for the round trip on the spot; a float
 * add is left alone and its result stays in flight until its first use. */
__device__ __forceinline__ float claim_ticket(float* p, float n) {
	float v;
	asm volatile("atom.relaxed.gpu.global.add.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "f"(n) : "memory");
	return v;
}

/* Dependency counters are polled with a strong RELAXED load (an ld.acquire.gpu costs ~2000 cycles a piece: it
 * invalidates the SM's L1); the producer turns a newly observed final count into an acquire with one fence. */
__device__ __forceinline__ unsigned int ld_counter(const unsigned int* p) {
	unsigned int v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ int ld_shared_volatile(const int* p) {
	int v;
	asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
	return v;
}

/* resolve of one 256-pixel unit of a key tile that the producer has bulk-copied into the stage: key -> rgb8
 * (kernels.cu:262-271, colour of the nearest point, black where nothing landed). The L2 latency of the key reads
 * lives in the TMA pipeline; the keys in L2 are re-armed by the retirer (one bulk store per item). Warp-local: the
 * packed words go through the unit's own (now dead) keys in shared memory and out to global as whole 16-byte pieces
 * — 12 bytes per thread stored directly are three partial sectors per lane. */
__device__ __forceinline__ void resolve_unit(const StreamParams& Q, int f, int tile, int unit, uint32_t stage,
                                             uint8_t* __restrict__ out, int lane, uint32_t limit) {
	const size_t n = Q.n_out_pixels;
	const size_t px0 = (size_t)tile * kResolveTile + (size_t)unit * kResolveUnit;
	if (px0 >= n) return;
	const uint32_t ubase = stage + (uint32_t)unit * (kResolveUnit * 8u);
	constexpr int kRounds = kResolveUnit / 128;
	static_assert(kRounds == 2, "the straight-line form below is written for 256-pixel units");
	if (px0 + kResolveUnit <= n) {
		/* a whole unit (every one but the last of an image): straight-line, all four key loads in flight together, no
		 * per-round conditions — the general form below spent most of its time waiting for one LDS after the other */
		uint8_t* dst = out + (size_t)f * n * 3 + px0 * 3;
		const uint32_t kaddr = ubase + (uint32_t)lane * 32u;
		const uint4 a0 = lds128(kaddr), b0 = lds128(kaddr + 16u), a1 = lds128(kaddr + 1024u), b1 = lds128(kaddr + 1040u);
		const uint32_t p0 = (a0.y >= limit) ? 0u : a0.x, p1 = (a0.w >= limit) ? 0u : a0.z;
		const uint32_t p2 = (b0.y >= limit) ? 0u : b0.x, p3 = (b0.w >= limit) ? 0u : b0.z;
		const uint32_t q0 = (a1.y >= limit) ? 0u : a1.x, q1 = (a1.w >= limit) ? 0u : a1.z;
		const uint32_t q2 = (b1.y >= limit) ? 0u : b1.x, q3 = (b1.w >= limit) ? 0u : b1.z;
		__syncwarp(); /* every key of the unit has been read */
		const uint32_t waddr = ubase + (uint32_t)lane * 12u; /* stride 3 words: conflict-free */
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(waddr), "r"(__byte_perm(p0, p1, 0x6012)) : "memory");
		asm volatile("st.shared.u32 [%0+4], %1;" ::"r"(waddr), "r"(__byte_perm(p1, p2, 0x5601)) : "memory");
		asm volatile("st.shared.u32 [%0+8], %1;" ::"r"(waddr), "r"(__byte_perm(p2, p3, 0x4560)) : "memory");
		asm volatile("st.shared.u32 [%0+384], %1;" ::"r"(waddr), "r"(__byte_perm(q0, q1, 0x6012)) : "memory");
		asm volatile("st.shared.u32 [%0+388], %1;" ::"r"(waddr), "r"(__byte_perm(q1, q2, 0x5601)) : "memory");
		asm volatile("st.shared.u32 [%0+392], %1;" ::"r"(waddr), "r"(__byte_perm(q2, q3, 0x4560)) : "memory");
		__syncwarp();
		/* 768 bytes = 48 pieces of 16: every lane stores one, the lower half-warp a second one */
		const uint32_t oaddr = ubase + (uint32_t)lane * 16u;
		const uint4 v0 = lds128(oaddr);
		uint4 v1 = make_uint4(0u, 0u, 0u, 0u);
		if (lane < 16) v1 = lds128(oaddr + 512u);
		__stcs(reinterpret_cast<uint4*>(dst + lane * 16), v0);
		if (lane < 16) __stcs(reinterpret_cast<uint4*>(dst + 512 + lane * 16), v1);
		return;
	}
	uint32_t w[3 * kRounds];
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const uint32_t l0 = (uint32_t)r * 128u + (uint32_t)lane * 4u; /* pixel inside the unit */
		uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
		if (px0 + l0 + 4 <= n) { /* n is a multiple of 4 (fuse_stream_supported) */
			const uint4 a = lds128(ubase + l0 * 8u), b = lds128(ubase + l0 * 8u + 16u);
			/* a key belongs to this use of the slot iff its high word lies below the epoch's limit; KEY_EMPTY and what
			 * earlier uses left behind lie at or above it (dfb_math.cuh) */
			c0 = (a.y >= limit) ? 0u : a.x;
			c1 = (a.w >= limit) ? 0u : a.z;
			c2 = (b.y >= limit) ? 0u : b.x;
			c3 = (b.w >= limit) ? 0u : b.z;
		}
		w[3 * r + 0] = __byte_perm(c0, c1, 0x6012);
		w[3 * r + 1] = __byte_perm(c1, c2, 0x5601);
		w[3 * r + 2] = __byte_perm(c2, c3, 0x4560);
	}
	__syncwarp(); /* every key of the unit has been read */
#pragma unroll
	for (int r = 0; r < kRounds; ++r) {
		const uint32_t word = (uint32_t)r * 96u + (uint32_t)lane * 3u; /* stride 3 words: conflict-free */
#pragma unroll
		for (int k = 0; k < 3; ++k) asm volatile("st.shared.u32 [%0], %1;" ::"r"(ubase + 4u * (word + k)), "r"(w[3 * r + k]) : "memory");
	}
	__syncwarp();
	const size_t count = n - px0 < (size_t)kResolveUnit ? n - px0 : (size_t)kResolveUnit;
	const uint32_t bytes = (uint32_t)count * 3u; /* a multiple of 4; out is 16-byte aligned and so is px0 * 3 */
	uint8_t* dst = out + (size_t)f * n * 3 + px0 * 3;
	for (uint32_t i = (uint32_t)lane * 16u; i + 16u <= bytes; i += 512u) {
		const uint4 v = lds128(ubase + i);
		__stcs(reinterpret_cast<uint4*>(dst + i), v);
	}
	if (lane < (int)((bytes & 15u) >> 2)) { /* the last unit of an image may end on a 4-byte boundary */
		const uint32_t i = (bytes & ~15u) + (uint32_t)lane * 4u;
		uint32_t v;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(ubase + i));
		__stcs(reinterpret_cast<uint32_t*>(dst + i), v);
	}
}

/* the splat of one thread's 16-pixel row segment with camera constants K: picks the straight-line instance
 * (camera transform yes/no, ROI test needed yes/no) once per unit, so that the 8 pixel pairs of the segment
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

/* ---- row-major traversal (un-rotated views) ----------------------------------------------------------------------
 * When source rows map to output ROWS, a warp's atomics are contiguous if its 32 lanes sit side by side on ONE source
 * row: a unit is 8 rows of the item, a lane owns columns (lane, lane + 32) of every row — each of the two atomics of a
 * pair covers 32 CONSECUTIVE output pixels — and visits one pair per row. Operands come out of the same swizzled
 * TMA tile with small reads:
 *   depth   row r, halfword c           16-byte chunk (2c >> 4) ^ (r & 7)                                 (SWIZZLE_128B)
 *   colour  row r, bytes [3c, 3c + 3) = two aligned words j = 3c >> 2 and j + 1; word j sits in box (4j) >> 6,
 *           chunk ((4j & 63) >> 4) ^ ((r >> 1) & 3)                                                       (SWIZZLE_64B)
 * r & 7 == k and (r >> 1) & 3 == (k >> 1) & 3 for the k-th row of a unit (units start at multiples of 8 rows).
 * The body is a LOOP over the rows with predicated atomics (62 registers; an unrolled, branch-free form like the column
 * body's was measured slower here: un-rotated C2 139.4 k vs 142.4 k framesets/s). */
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
#pragma unroll 1
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
		splat_pair<kHasT, kRoiAll, 0>(G, K, dpair, xr, yr, lo0, lo1, u, row_in_roi, T, slow, 1u << k, 32);
	}
	return slow;
}

__device__ __forceinline__ uint32_t splat_tile_rows_split(const SplatShared& G, const SplatCam& K, uint32_t stage, uint32_t row0,
                                                          const RowLaneSplit& L, float2 xr, int u, int v0, int tile_u0, int tile_v0,
                                                          const KeyTarget& T) {
	const bool inside = K.roi_all || (tile_u0 >= K.roi_x0 && tile_u0 + kTileW <= K.roi_x0 + K.roi_w && tile_v0 >= K.roi_y0 &&
	                                  tile_v0 + 8 <= K.roi_y0 + K.roi_h);
	if (K.has_t)
		return inside ? splat_rows_split_loop<1, 1>(G, K, stage, row0, L, xr, u, v0, T) : splat_rows_split_loop<1, 0>(G, K, stage, row0, L, xr, u, v0, T);
	return inside ? splat_rows_split_loop<0, 1>(G, K, stage, row0, L, xr, u, v0, T) : splat_rows_split_loop<0, 0>(G, K, stage, row0, L, xr, u, v0, T);
}

/* column form, thread (unit, lane): offsets from the stage of its two depth pieces, three colour pieces (TMA swizzles
 * undone: both patterns repeat every 8 rows), its row ray and its 16 column rays. ~45 integer instructions that used to
 * run per item and warp; now two LDS.128 */
constexpr uint32_t kLaneTblHalf = kW * 32 * 16;
__device__ __forceinline__ void lane_offsets(uint32_t unit, uint32_t lane, uint4& oa, uint4& ob) {
	const uint32_t w = unit & 3u, half = unit >> 2, r = lane;
	const uint32_t hd = half * (kTileH * 128u) + r * 128u, sw = r & 7u;
	const uint32_t hc = kItemDepthBytes + half * (kTileH * 64u) + r * 64u, cs = (r >> 1) & 3u;
	uint32_t co[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const uint32_t p = 48u * w + 16u * (uint32_t)k; /* byte position inside the 192-byte colour row */
		co[k] = hc + (p >> 6) * kItemColorBoxBytes + ((((p & 63u) >> 4) ^ cs) << 4);
	}
	oa = make_uint4(hd + (((2u * w) ^ sw) << 4), hd + (((2u * w + 1u) ^ sw) << 4), co[0], co[1]);
	ob = make_uint4(co[2], kRayYOff + 4u * (half * kTileH + r), kRayXOff + w * (kSeg * 4), 0u);
}

/* one splat unit of item `it` that sits in `stage`. Column form: unit = 16 columns (unit & 3) x 32 rows (unit >> 2),
 * lane = row; row-major form: unit = rows 8 unit .. 8 unit + 7, lane = columns (lane, lane + 32). */
template <bool kRowMajor>
__device__ __forceinline__ void splat_unit(const SplatShared& G, const StreamCams& cams, const FuseParams& P, const StreamParams& Q,
                                           const Item& it, int unit, uint32_t stage, const uint16_t* __restrict__ depth,
                                           const uint8_t* __restrict__ color, unsigned long long* __restrict__ keys, int lane,
                                           uint32_t dump_slot, uint32_t lane_tbl) {
	const int ty = (int)Q.by_tiles_x.div((unsigned)it.b);
	const int tu0 = (it.b - ty * Q.tiles_x) * kTileW;
	const uint32_t slot_off = (uint32_t)it.slot * Q.slot_bytes;
	const uint32_t bias = key_epoch_bias(it.epoch);
	unsigned long long* kslot = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(keys) + slot_off);
	/* invalid / out-of-image points dump their RED into a per-thread word of the scratch area behind the ring */
	const KeyTarget T{reinterpret_cast<char*>(keys), slot_off + G.px_origin, Q.dump_off + dump_slot * 8u, bias};
	uint32_t slow;
	if constexpr (kRowMajor) {
		const RowLaneSplit rls = row_lane_split((uint32_t)lane);
		const uint32_t row0 = (uint32_t)unit * 8u;
		const int v0 = ty * kItemRows + unit * 8;
		if (v0 >= G.height) return; /* the item hangs over the bottom of the image: nothing but zero fill there */
		float2 xr;
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xr.x) : "r"(stage + kRayXOff + 4u * (uint32_t)lane));
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xr.y) : "r"(stage + kRayXOff + 4u * (uint32_t)(lane + 32)));
		const int u = tu0 + lane, du = 32;
		switch (it.a) {
			case 0: slow = splat_tile_rows_split(G, cams.K[0], stage, row0, rls, xr, u, v0, tu0, v0, T); break;
			case 1: slow = splat_tile_rows_split(G, cams.K[1], stage, row0, rls, xr, u, v0, tu0, v0, T); break;
			default: slow = splat_tile_rows_split(G, cams.K[it.a], stage, row0, rls, xr, u, v0, tu0, v0, T); break;
		}
		if (slow) { /* rare: operands outside the guarded fast-division range; bit k = the pair in row v0 + k */
			const uint16_t* dimg = depth + (size_t)(it.f * Q.imgs_per_fs + Q.slot_d[it.a]) * Q.depth_img_stride;
			const uint8_t* cimg = color + (size_t)(it.f * Q.imgs_per_fs + Q.slot_c[it.a]) * Q.color_img_stride;
			while (slow) {
				const int j = __ffs(slow) - 1;
				slow &= slow - 1;
				splat_pixel_slow(P, it.a, dimg, cimg, u, v0 + j, kslot, bias);
				splat_pixel_slow(P, it.a, dimg, cimg, u + du, v0 + j, kslot, bias);
			}
		}
	} else {
		const uint32_t w = (uint32_t)unit & 3u, half = (uint32_t)unit >> 2;
		const int tv0 = ty * kItemRows + (int)half * kTileH;
		if (tv0 >= G.height) return;
		const int u0 = tu0 + (int)w * kSeg;
		/* the five 16-byte pieces of this thread's segment with the TMA swizzles undone, its row ray and its column rays:
		 * offsets from the stage that depend on (unit, lane) only — computed once per kernel (lane_offsets) */
		const uint4 oa = lds128(lane_tbl), ob = lds128(lane_tbl + kLaneTblHalf);
		/* the table holds the warp's FIRST unit (unit % kW); its further units lie whole row groups further down */
		const uint32_t rows_down = ((uint32_t)unit / kW) * (kW / 4) * kTileH;
		const uint32_t sd = stage + rows_down * 128u, sc = stage + rows_down * 64u;
		const uint4 d0 = lds128(sd + oa.x), d1 = lds128(sd + oa.y);
		uint4 c[3];
		c[0] = lds128(sc + oa.z);
		c[1] = lds128(sc + oa.w);
		c[2] = lds128(sc + ob.x);
		float yr;
		asm volatile("ld.shared.f32 %0, [%1];" : "=f"(yr) : "r"(stage + ob.y + rows_down * 4u));
		const uint32_t xs = stage + ob.z;
		const int v = tv0 + lane;
		const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
		const uint32_t cw[12] = {c[0].x, c[0].y, c[0].z, c[0].w, c[1].x, c[1].y, c[1].z, c[1].w, c[2].x, c[2].y, c[2].z, c[2].w};
		/* One copy of the body per camera index: with a STATIC index the camera's constants sit at fixed constant-bank
		 * addresses and feed FFMA2 straight from uniform registers; a run-time index costs an LDC + R2UR per use (~18 %
		 * more instructions, measured in SASS). Cameras >= 2 take the generic copy. */
		switch (it.a) {
			case 0: slow = splat_tile(G, cams.K[0], xs, yr, dw, cw, u0, v, tu0, tv0, T); break;
			case 1: slow = splat_tile(G, cams.K[1], xs, yr, dw, cw, u0, v, tu0, tv0, T); break;
			default: slow = splat_tile(G, cams.K[it.a], xs, yr, dw, cw, u0, v, tu0, tv0, T); break;
		}
		if (slow) { /* rare: operands outside the guarded fast-division range */
			const uint16_t* dimg = depth + (size_t)(it.f * Q.imgs_per_fs + Q.slot_d[it.a]) * Q.depth_img_stride;
			const uint8_t* cimg = color + (size_t)(it.f * Q.imgs_per_fs + Q.slot_c[it.a]) * Q.color_img_stride;
			while (slow) {
				const int j = __ffs(slow) - 1;
				slow &= slow - 1;
				splat_pixel_slow(P, it.a, dimg, cimg, u0 + 2 * j, v, kslot, bias);
				splat_pixel_slow(P, it.a, dimg, cimg, u0 + 2 * j + 1, v, kslot, bias);
			}
		}
	}
}

/* kRowMajor: which way a warp walks a tile — lane = source row (views rotated by 90 degrees: the default vertical
 * image) or lane = pixel of one source row (un-rotated views); chosen so that a warp's atomics are contiguous */
template <bool kRowMajor>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
fuse_stream_kernel(const __grid_constant__ SplatShared G, const __grid_constant__ StreamCams cams,
                   const __grid_constant__ FuseParams P, const __grid_constant__ StreamParams Q,
                   const __grid_constant__ CUtensorMap map_depth, const __grid_constant__ CUtensorMap map_color,
                   const float* __restrict__ tile_rays, const uint16_t* __restrict__ depth,
                   const uint8_t* __restrict__ color, unsigned long long* __restrict__ keys, uint8_t* __restrict__ out,
                   unsigned int* __restrict__ counters) {
	extern __shared__ uint8_t smem_raw[];
	/* full: producer (+ TMA bytes) -> compute; done: compute (kW warps) -> retirer AND producer (stage recycling).
	 * done barriers are indexed by item number modulo kRetireDepth, so the retirer may lag that many items behind the
	 * producer without a parity alias; s_retired tells the producer how far it got. Every wait inside the CTA is a
	 * hardware-suspended mbarrier wait. */
	__shared__ __align__(8) unsigned long long full_bar[kStages], done_bar[kRetireDepth];
	__shared__ Item items[kStages];
	__shared__ int retire_info[kRetireDepth]; /* 2 * frameset + kind of the item: the counter its completion goes to */
	__shared__ __align__(16) uint4 s_lane_tbl[2][kW * 32]; /* column form: lane_offsets() of every compute thread */
	__shared__ int s_retired_count;            /* items the retirer has taken over */
	__shared__ int s_final_item;               /* number of real items of this CTA, INT_MAX until the producer knows */
#if DFB_STREAM_INSTRUMENT
	__shared__ long long pub_time[kRetireDepth]; /* when the producer published the item (latency accounting) */
#endif
	const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	if (tid == 0) {
		for (int s = 0; s < kStages; ++s) {
			mbar_init(smem_u32(&full_bar[s]), 1);
		}
		for (int s = 0; s < kRetireDepth; ++s) mbar_init(smem_u32(&done_bar[s]), kW);
		s_retired_count = 0;
		s_final_item = INT_MAX;
		mbar_fence_init();
	}
	for (uint32_t i = (uint32_t)tid * 16u; i < kRearmBytes; i += (uint32_t)kThreads * 16u)
		asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(base + kRearmOff + i), "r"(0xFFFFFFFFu) : "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the bulk stores read this buffer through the async proxy */
	if (!kRowMajor && warp < kW) lane_offsets((uint32_t)warp, (uint32_t)lane, s_lane_tbl[0][tid], s_lane_tbl[1][tid]);
	__syncthreads(); /* the only CTA-wide barrier */

	unsigned int* const done_ctr = counters + kFirstFramesetLine * kCounterStride; /* [2f] splat, [2f+1] resolve */

	if (warp == kW) {
		/* -------------------------------- producer (one lane) -------------------------------- */
		if (lane != 0) return;
		auto done_splat = [&](int f) { return done_ctr + (size_t)(2 * f) * kCounterStride; };
		auto done_resolve = [&](int f) { return done_ctr + (size_t)(2 * f + 1) * kCounterStride; };
		const uint64_t policy = l2_evict_first_policy();
		/* the ticket group is picked in ARRIVAL order: the CTAs that actually run cover the groups, however many of
		 * the grid's CTAs are resident (other kernels may share the GPU) */
		const int group = (int)(atomicAdd(counters, 1u) % (unsigned)Q.groups);
		float* ticket_counter = reinterpret_cast<float*>(counters) + (1 + group) * kCounterStride;
		auto global_ticket = [&](float local) { return (int)local * Q.groups + group; };
		/* known: highest frameset whose count has been OBSERVED complete; fenced: .. and acquired (fence after the
		 * observation). Framesets complete in order as far as this CTA can rely on it: a count is only consulted
		 * for the frameset an item needs, everything below it is implied by the ticket order. */
		int known_resolved = -1, known_splatted = -1, fenced_resolved = -1, fenced_splatted = -1;
		int pre_r_idx = -1, pre_s_idx = -1;
		unsigned pre_r_val = 0, pre_s_val = 0;
		int k = 0, retired_seen = 0; /* k: items published */
		unsigned long long idle = 0;
		long long t_stage = 0, t_dep = 0, t_depS = 0, t_pub = 0, t_fence = 0, t0 = prof_clock(), t1;
		int n_fence = 0, n_empty = 0, n_pollS = 0, n_pollR = 0;
		float ahead = claim_ticket(ticket_counter, 1.0f);
		while (true) {
			const Item it = decode_item(Q, global_ticket(ahead));
			if (it.kind == -1) break; /* past the end of the list */
			ahead = claim_ticket(ticket_counter, 1.0f); /* the next ticket: its round trip overlaps this item */
			if (it.kind == -2) {
				++n_empty;
				continue;
			}
			/* look-ahead polls: the counters of the NEXT framesets this CTA will depend on were loaded one item ago; their
			 * values are consumed here and the next loads issued, so a satisfied dependency never costs an L2 round trip */
			if (pre_r_idx >= 0 && pre_r_val >= (unsigned)Q.resolve_per_fs) known_resolved = pre_r_idx;
			if (pre_s_idx >= 0 && pre_s_val >= (unsigned)Q.splat_per_fs) known_splatted = pre_s_idx;
			pre_r_idx = known_resolved + 1 < Q.n_framesets ? known_resolved + 1 : -1;
			if (pre_r_idx >= 0) pre_r_val = ld_counter(done_resolve(pre_r_idx));
			pre_s_idx = known_splatted + 1 < Q.n_framesets ? known_splatted + 1 : -1;
			if (pre_s_idx >= 0) pre_s_val = ld_counter(done_splat(pre_s_idx));
			t1 = prof_clock();
			const int s = k % kStages;
			if (k >= kStages) mbar_wait(smem_u32(&done_bar[(k - kStages) % kRetireDepth]), (uint32_t)((k - kStages) / kRetireDepth) & 1u);
			while (k - retired_seen >= kRetireDepth) { /* rare: the retirer is a whole queue behind */
				retired_seen = ld_shared_volatile(&s_retired_count);
				if (k - retired_seen < kRetireDepth) break;
				__nanosleep(200);
				if (++idle > (1ull << 24)) __trap();
			}
			t_stage += prof_clock() - t1;
			t1 = prof_clock();
			/* the item's dependency. A newly observed count becomes an ACQUIRE through one fence.acq_rel.gpu (it pairs with
			 * the retirers' red.release.gpu on the counter), followed by a proxy fence: the other CTAs' REDs and re-arm stores, now
			 * visible to this thread, are ordered before the bulk copies (async proxy) it issues from here on */
			int need_r = -1, need_s = -1;
			if (it.kind == 0) need_r = it.f - Q.slots; /* key slot re-armed? */
			else need_s = it.f;                       /* every camera of the frameset splatted? */
			if (need_r > known_resolved) {
				++n_pollS;
				while (ld_counter(done_resolve(need_r)) < (unsigned)Q.resolve_per_fs) {
					__nanosleep(100);
					if (++idle > (1ull << 24)) __trap(); /* a protocol bug must be an error, not a hung GPU */
				}
				known_resolved = need_r;
			}
			if (need_s > known_splatted) {
				++n_pollR;
				while (ld_counter(done_splat(need_s)) < (unsigned)Q.splat_per_fs) {
					__nanosleep(100);
					if (++idle > (1ull << 24)) __trap();
				}
				known_splatted = need_s;
			}
			if (need_r > fenced_resolved || need_s > fenced_splatted) {
				const long long tf = prof_clock();
				asm volatile("fence.acq_rel.gpu;" ::: "memory");
				if (known_splatted > fenced_splatted) asm volatile("fence.proxy.async.global;" ::: "memory");
				fenced_resolved = known_resolved;
				fenced_splatted = known_splatted;
				t_fence += prof_clock() - tf;
				++n_fence;
			}
			idle = 0;
			t_dep += prof_clock() - t1;
			if (it.kind == 0) t_depS += prof_clock() - t1;
			t1 = prof_clock();
#if DFB_STREAM_INSTRUMENT
			pub_time[k % kRetireDepth] = t1;
#endif
			items[s] = it;
			retire_info[k % kRetireDepth] = 2 * it.f + it.kind;
			const uint32_t bar = smem_u32(&full_bar[s]);
			const uint32_t dst = base + (uint32_t)s * kStageStride;
			if (it.kind == 0) {
				const int ty = (int)Q.by_tiles_x.div((unsigned)it.b);
				const int u0 = (it.b - ty * Q.tiles_x) * kTileW, v0 = ty * kItemRows;
				const int image_d = it.f * Q.imgs_per_fs + Q.slot_d[it.a], image_c = it.f * Q.imgs_per_fs + Q.slot_c[it.a];
				mbar_expect_tx(bar, kStageTx);
				tma_load_3d(dst, &map_depth, u0, v0, image_d, bar, policy);
				tma_load_3d(dst + kItemDepthBytes, &map_color, u0 * 3, v0, image_c, bar, policy);
				tma_load_3d(dst + kItemDepthBytes + kItemColorBoxBytes, &map_color, u0 * 3 + 64, v0, image_c, bar, policy);
				tma_load_3d(dst + kItemDepthBytes + 2 * kItemColorBoxBytes, &map_color, u0 * 3 + 128, v0, image_c, bar, policy);
				bulk_load_1d(dst + kRayXOff, tile_rays + ((size_t)it.a * Q.tiles_per_image + it.b) * (kTileW + kItemRows), (kTileW + kItemRows) * 4, bar);
			} else { /* the key tile of the resolve item */
				const uint32_t first = (uint32_t)it.a * kResolveTile;
				const uint32_t count = Q.n_out_pixels - first < (uint32_t)kResolveTile ? Q.n_out_pixels - first : (uint32_t)kResolveTile;
				const char* src = reinterpret_cast<const char*>(keys) + (size_t)it.slot * Q.slot_bytes + (size_t)first * 8;
				mbar_expect_tx(bar, count * 8u);
				bulk_load_1d(dst, src, count * 8u, bar);
			}
			t_pub += prof_clock() - t1;
			++k;
		}
		/* k real items: tell the retirer, then publish the terminator the compute warps stop at */
		asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(&s_final_item)), "r"(k) : "memory");
		{
			const int s = k % kStages;
			if (k >= kStages) mbar_wait(smem_u32(&done_bar[(k - kStages) % kRetireDepth]), (uint32_t)((k - kStages) / kRetireDepth) & 1u);
			Item end;
			end.kind = -1;
			end.slot = end.epoch = 0;
			end.f = end.a = end.b = 0;
			items[s] = end;
			mbar_arrive(smem_u32(&full_bar[s]));
		}
		if (DFB_STREAM_INSTRUMENT && (Q.profile & 1) && (blockIdx.x % 37) == 0)
			printf("cta %4d producer: items %d (+%d empty tickets) total %lld wait-stage %lld dependency %lld (splat items %lld; polls S %d R %d; fences %d: %lld) publish %lld\n",
			       blockIdx.x, k, n_empty, prof_clock() - t0, t_stage, t_dep, t_depS, n_pollS, n_pollR, n_fence, t_fence, t_pub);
		return;
	}
	if (warp == kW + 1) {
		/* -------------------------------- retirer (one lane) --------------------------------- */
		if (lane != 0) return;
		int retired = 0;
		unsigned long long idle = 0;
		long long t_wait = 0, t_signal = 0, t0 = prof_clock(), t1;
		long long lat_seen = 0, lat_sig = 0, lat_max = 0, pend_pub = 0;
		int fences = 0, pend_items = 0;
		/* Finished items free their stage at once (the producer watches the same done barriers); their global signals are BATCHED: one gpu-scope fence
		 * covers every item that finished while the previous fence was draining. The delay is far below the distance of
		 * the dependencies (lag framesets). */
		constexpr int kPend = 4;
		int pend_idx[kPend], pend_cnt[kPend], n_pend = 0;
		auto flush = [&]() {
			/* RELEASE read-modify-writes at gpu scope: cumulative over everything the compute warps wrote before they arrived
			 * on done[] (observed by this thread through the barriers' acquire). MEMBAR.ALL.GPU + RED; unlike
			 * fence.acq_rel.gpu no CCTL.IVALL — the retirer acquires nothing, the SM's L1 stays */
#pragma unroll
			for (int i = 0; i < kPend; ++i)
				if (i < n_pend)
					asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(done_ctr + (size_t)pend_idx[i] * kCounterStride), "r"((unsigned)pend_cnt[i]) : "memory");
			n_pend = 0;
			++fences;
			if (DFB_STREAM_INSTRUMENT) { /* publish -> signalled, summed over the batch */
				const long long now = prof_clock();
				lat_sig += now * pend_items - pend_pub;
				pend_items = 0;
				pend_pub = 0;
			}
		};
		while (true) {
			t1 = prof_clock();
			const uint32_t bar = smem_u32(&done_bar[retired % kRetireDepth]), parity = (uint32_t)(retired / kRetireDepth) & 1u;
			bool got = mbar_test(bar, parity);
			if (!got && n_pend) { /* nothing more has finished: publish what is pending, then wait */
				t_wait += prof_clock() - t1;
				t1 = prof_clock();
				flush();
				t_signal += prof_clock() - t1;
				t1 = prof_clock();
			}
			bool end = false;
			while (!got) {
				if (retired >= ld_shared_volatile(&s_final_item)) {
					end = true;
					break;
				}
				got = mbar_try_wait_ns(bar, parity, 1000u);
				if (!got) {
					__nanosleep(200);
					if (++idle > (1ull << 24)) __trap();
				}
			}
			if (end) break;
			idle = 0;
			t_wait += prof_clock() - t1;
			t1 = prof_clock();
			const int idx = retire_info[retired % kRetireDepth];
#if DFB_STREAM_INSTRUMENT
			{
				const long long pt = pub_time[retired % kRetireDepth], d = t1 - pt;
				lat_seen += d;
				lat_max = d > lat_max ? d : lat_max;
				pend_pub += pt;
				++pend_items;
			}
#endif
			++retired;
			asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(smem_u32(&s_retired_count)), "r"(retired) : "memory");
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
			t_signal += prof_clock() - t1;
		}
		if (n_pend) flush();
		if (DFB_STREAM_INSTRUMENT && (Q.profile & 1) && (blockIdx.x % 37) == 0)
			printf("cta %4d retirer: items %d fences %d total %lld wait-done %lld fence+signal %lld | per item: publish->done seen %lld (max %lld), publish->signalled %lld\n",
			       blockIdx.x, retired, fences, prof_clock() - t0, t_wait, t_signal, lat_seen / (retired ? retired : 1), lat_max,
			       lat_sig / (retired ? retired : 1));
		return;
	}

	/* ---------------------------------------- compute warps ---------------------------------------- */
	const uint32_t dump_slot = ((uint32_t)blockIdx.x * (kW * 32) + (uint32_t)tid) & (kDumpWords - 1);
	const uint32_t lane_tbl = smem_u32(&s_lane_tbl[0][tid]);
	long long c_wait = 0, c_splat = 0, c_resolve = 0, c_arrive = 0, c0 = prof_clock();
	int n_splat = 0, n_resolve = 0;
#ifdef DFB_EXPERIMENTS
	const int reps = Q.repeat;
#else
	constexpr int reps = 1;
#endif
	/* item k of this CTA sits in stage k % kStages (parity (k / kStages) & 1 of its full barrier) and is reported on done
	 * barrier k % kRetireDepth: three small counters kept incrementally (a division per use cost ~15 instructions per item) */
	uint32_t s = 0, full_parity = 0, d = 0;
	for (;;) {
		long long tc = prof_clock();
		mbar_wait(smem_u32(&full_bar[0]) + 8u * s, full_parity);
		c_wait += prof_clock() - tc;
		tc = prof_clock();
		const Item it = items[s];
		if (it.kind < 0) break; /* terminator */
		const uint32_t stage = base + (uint32_t)s * kStageStride;
		/* a resolve item's key tile sits in the stage: its place in the ring is re-armed for the slot's next frameset NOW, with
		 * one bulk store (shared -> global) of the constant KEY_EMPTY tile — it completes while the item is being resolved.
		 * (Issued by the retirer after the item, as in round 1, the store's completion sat inside every flush: 6300 instead of
		 * 3000 cycles per flush, 157 k instead of 160 k framesets/s, and two more ring slots needed for the same rate.) */
		const bool rearms = it.kind == 1 && warp == 0 && lane == 0 && it.epoch == kKeyEpochs - 1;
		if (rearms) {
			const size_t tile0 = (size_t)it.a * kResolveTile;
			const size_t count = (size_t)Q.n_out_pixels - tile0 < (size_t)kResolveTile ? (size_t)Q.n_out_pixels - tile0 : (size_t)kResolveTile;
			char* kp = reinterpret_cast<char*>(keys) + (size_t)it.slot * Q.slot_bytes + tile0 * 8;
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(kp), "r"(base + kRearmOff), "r"((uint32_t)(count * 8)) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		}
		const int n_units = it.kind == 0 ? kSplatUnits : kResolveUnits;
#pragma unroll 1
		for (int unit = warp; unit < n_units; unit += kW) {
			if (it.kind == 0) {
#pragma unroll 1
				for (int rep = 0; rep < reps; ++rep)
					splat_unit<kRowMajor>(G, cams, P, Q, it, unit, stage, depth, color, keys, lane, dump_slot, lane_tbl);
			} else {
				resolve_unit(Q, it.f, it.a, unit, stage, out, lane, key_epoch_limit(it.epoch));
			}
		}
		if (it.kind == 0) {
			c_splat += prof_clock() - tc;
			++n_splat;
		} else {
			c_resolve += prof_clock() - tc;
			++n_resolve;
		}
		tc = prof_clock();
		if (rearms) { /* the store has landed (async proxy) before this warp reports the item (generic proxy) */
			asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
			asm volatile("fence.proxy.async.global;" ::: "memory");
		}
		__syncwarp();
		if (lane == 0) {
			/* ONE arrival: release of everything this warp wrote for the item (for the retirer) and "this warp no longer reads
			 * the stage" (for the producer, which waits for the done barrier of the item whose stage it wants to refill) */
			mbar_arrive(smem_u32(&done_bar[0]) + 8u * d);
		}
		c_arrive += prof_clock() - tc;
		if (++s == (uint32_t)kStages) {
			s = 0;
			full_parity ^= 1u;
		}
		if (++d == (uint32_t)kRetireDepth) d = 0;
	}
	if (DFB_STREAM_INSTRUMENT && (Q.profile & 1) && (blockIdx.x % 37) == 0 && lane == 0 && (warp == 1 || warp == 0))
		printf("cta %4d compute (warp %d): total %lld wait-full %lld splat %lld (%d items) resolve %lld (%d items) arrive %lld\n",
		       blockIdx.x, warp, prof_clock() - c0, c_wait, c_splat, n_splat, c_resolve, n_resolve, c_arrive);
}

}  // namespace

bool fuse_stream_supported(const FuseParams& P, const uint16_t* d_depth, const uint8_t* d_color, const uint8_t* d_out) {
	return pools_fit_tensor_maps(P.cam[0].width, d_depth, d_color) && ((size_t)P.n_out_pixels * 8 < (1ull << 32)) &&
	       (P.n_out_pixels % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15u) == 0) && (((size_t)P.n_out_pixels * 3) % 16 == 0);
}

size_t fuse_stream_counter_words(int n_framesets) { return (size_t)kCounterStride * (kFirstFramesetLine + 2 * (size_t)n_framesets); }

/* bytes the stream kernel needs BEHIND the allocated key slots (scratch for the branchless REDs) */
size_t fuse_stream_scratch_bytes() { return (size_t)kDumpWords * 8; }

/* tickets one frameset costs (the launcher's caller splits batches so that a launch stays below 2^23) */
long long fuse_stream_tickets_per_frameset(const FuseParams& P) {
	const int w = P.cam[0].width, h = P.cam[0].height;
	const long long tiles = (long long)((w + kTileW - 1) / kTileW) * ((h + kItemRows - 1) / kItemRows);
	return tiles * P.n_cameras + (P.n_out_pixels + kResolveTile - 1) / kResolveTile;
}

/* d_keys holds ring_slots >= slots frameset slots followed by fuse_stream_scratch_bytes(); d_counters =
 * fuse_stream_counter_words(n). *ring_phase (host, in/out) is the ring's state between launches: 0 with an all-ones ring;
 * after a launch the slots hold the stale keys of their last use and *ring_phase says which use comes next
 * (StreamParams::phase0). A ring that other code wants all-ones again must be cleared and its phase reset to 0. */
int launch_fuse_stream(const FuseParams& P, const RayTables& rays, const uint16_t* d_depth, const uint8_t* d_color, uint8_t* d_out,
                       unsigned long long* d_keys, int ring_slots, unsigned int* d_counters, int n_framesets, int slots, int lag,
                       bool row_major, const StreamLayout* layout, uint32_t* ring_phase, cudaStream_t s) {
	DFB_REQUIRE(ring_phase != nullptr, "NULL ring phase");
	const int w = P.cam[0].width, h = P.cam[0].height;
	StreamLayout packed;
	if (!layout) {
		packed.imgs_per_fs = P.n_cameras;
		packed.depth_stride_bytes = (size_t)w * h * 2;
		packed.color_stride_bytes = (size_t)w * h * 3;
		for (int c = 0; c < DFB_MAX_CAMERAS; ++c) packed.slot_d[c] = packed.slot_c[c] = (unsigned char)c;
		layout = &packed;
	}
	alignas(64) CUtensorMap md, mc;
	int rc = make_pool_tensor_maps(d_depth, d_color, w, h, (uint64_t)n_framesets * layout->imgs_per_fs, &md, &mc, kItemRows,
	                               layout->depth_stride_bytes, layout->color_stride_bytes);
	if (rc != DFB_OK) return rc;
	StreamParams Q;
	Q.imgs_per_fs = layout->imgs_per_fs;
	Q.depth_img_stride = layout->depth_stride_bytes / 2;
	Q.color_img_stride = layout->color_stride_bytes;
	for (int c = 0; c < DFB_MAX_CAMERAS; ++c) {
		Q.slot_d[c] = layout->slot_d[c];
		Q.slot_c[c] = layout->slot_c[c];
	}
	Q.n_framesets = n_framesets;
	Q.n_cameras = P.n_cameras;
	Q.tiles_x = (w + kTileW - 1) / kTileW;
	Q.tiles_per_image = Q.tiles_x * ((h + kItemRows - 1) / kItemRows);
	Q.splat_per_fs = Q.tiles_per_image * P.n_cameras;
	Q.resolve_per_fs = (P.n_out_pixels + kResolveTile - 1) / kResolveTile;
	Q.by_round.init((unsigned)(Q.splat_per_fs + Q.resolve_per_fs));
	Q.by_tiles.init((unsigned)Q.tiles_per_image);
	Q.by_tiles_x.init((unsigned)Q.tiles_x);
	if (slots > ring_slots) slots = ring_slots;
	if (slots < 1) slots = 1;
	if (lag > slots - 1) lag = slots - 1; /* R(f-lag) must be queued before S(f-lag+slots) */
	/* the list has n + lag rounds and the last `lag` of them hold resolve tickets only: a single frameset (the drop-in
	 * classes' call) runs S(0), R(0) with no empty tickets at all */
	if (lag > n_framesets - 1) lag = n_framesets - 1;
	if (lag < 0) lag = 0;
	Q.lag = lag;
	Q.slots = slots;
	Q.by_slots.init((unsigned)slots);
	const uint32_t period = (uint32_t)slots * (uint32_t)kKeyEpochs;
	Q.phase0 = *ring_phase % period; /* (a ring used with another slot count before must have been cleared: phase 0) */
	Q.n_out_pixels = (uint32_t)P.n_out_pixels;
	Q.slot_bytes = (uint32_t)P.n_out_pixels * 8u;
	DFB_REQUIRE((unsigned long long)P.n_out_pixels * 8ull * (unsigned long long)ring_slots + fuse_stream_scratch_bytes() < (1ull << 32),
	            "key ring too large for 32-bit offsets: use fewer slots");
	Q.dump_off = Q.slot_bytes * (uint32_t)ring_slots;
	const long long total = (long long)(n_framesets + lag) * (Q.splat_per_fs + Q.resolve_per_fs);
	DFB_REQUIRE(total + 65536 < (1ll << 24) * kTicketGroups && total < (1ll << 30),
	            "too many work items for one launch (float ticket counters): split the batch");
	Q.total_tickets = (int)total;
	Q.profile = 0;
	Q.repeat = 1;
#if DFB_STREAM_INSTRUMENT || defined(DFB_EXPERIMENTS)
	{
		const char* e = std::getenv("DFB_STREAM_PROFILE");
		Q.profile = e ? std::atoi(e) : 0;
		const char* r = std::getenv("DFB_STREAM_REPEAT");
		Q.repeat = r ? std::atoi(r) : 1;
		if (Q.repeat < 1) Q.repeat = 1;
	}
#endif
	StreamCams cams;
	for (int c = 0; c < DFB_MAX_CAMERAS; ++c) cams.K[c] = make_splat_cam(P, c < P.n_cameras ? c : 0, rays);
	SplatShared G = make_splat_shared(P);
	G.x_pitch = rays.x_pitch;
	G.y_base = (int)(rays.y - rays.x);
	G.y_pitch = rays.y_pitch;

	/* kCtasPerSm CTAs per SM (shared memory and registers are sized for exactly that). Per DEVICE, set up once: a
	 * single-frameset launch (the drop-in classes, 30 Hz) should not pay for attribute calls every time */
	int dev = 0;
	DFB_CUDA(cudaGetDevice(&dev));
	static std::mutex setup_mutex;
	static int sms_of[64] = {0};
	int sms = 0;
	{
		std::lock_guard<std::mutex> lk(setup_mutex);
		if (dev < 0 || dev >= 64 || sms_of[dev] == 0) {
			DFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
			DFB_CUDA(cudaFuncSetAttribute(fuse_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
			DFB_CUDA(cudaFuncSetAttribute(fuse_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
			if (dev >= 0 && dev < 64) sms_of[dev] = sms;
		} else {
			sms = sms_of[dev];
		}
	}
	long long grid = (long long)sms * kCtasPerSm;
	const long long real_items = (long long)n_framesets * (Q.splat_per_fs + Q.resolve_per_fs);
	if (grid > real_items) grid = real_items;
	Q.groups = grid < kTicketGroups ? (int)grid : kTicketGroups;
	DFB_CUDA(cudaMemsetAsync(d_counters, 0, fuse_stream_counter_words(n_framesets) * sizeof(unsigned int), s));
	/* (A persisting-L2 access-policy window over the key ring was measured and rejected: 126 k -> 107 k framesets/s,
	 * 2.6 GB of DRAM writes per 128 framesets. The ring stays in L2 on its own as long as it is <= ~46 MB.) */
	if (row_major)
		fuse_stream_kernel<true><<<(unsigned)grid, kThreads, kSmemBytes, s>>>(G, cams, P, Q, md, mc, rays.tiles, d_depth, d_color, d_keys, d_out,
		                                                                      d_counters);
	else
		fuse_stream_kernel<false><<<(unsigned)grid, kThreads, kSmemBytes, s>>>(G, cams, P, Q, md, mc, rays.tiles, d_depth, d_color, d_keys,
		                                                                       d_out, d_counters);
	count_launch();
	DFB_CUDA(cudaGetLastError());
	*ring_phase = (Q.phase0 + (uint32_t)n_framesets) % period;
	return DFB_OK;
}

}  // namespace dfb
