/*
 * post_ops.cu — the step right after the splat in FusionNode::processSyncedFrames
 * (fusion_node.cpp:789): Frameset::filterColor (frameset.cpp:194-208) =
 *   Kernels::filter_color_gauss_npp  (kernels.cu:632-653)  nppiFilterGauss_8u_C3R 3x3
 *   Kernels::filter_color_median_npp (kernels.cu:594-630)  nppiFilterMedian_8u_C3R 3x3
 * Both read a copy of the image (the reference does a D2D copy into the ColorFrame scratch buffer)
 * and write the interior (W-2)x(H-2); the one-pixel border keeps the splat output.
 * NPP is closed source: the reference pins nothing here; we pin against NPP 12.4 on the B200
 * (tests/test_ref_kernels_gpu.py).
 */
#include <cmath>

#include "dfb_internal.h"

namespace dfb {
namespace {

constexpr int kThreads = 256;

struct GaussTaps {
	float w[3];
};

__global__ void __launch_bounds__(kThreads)
gauss3x3_kernel(const GaussTaps G, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int width, int height) {
	const int row_bytes = width * 3;
	const int inner = (width - 2) * 3;
	const long long n = (long long)inner * (height - 2);
	const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
	if (t >= n) return;
	const int y = (int)(t / inner) + 1;
	const int xb = (int)(t % inner) + 3; /* byte offset inside the row (pixel 1 .. W-2, any channel) */
	float acc = 0.0f;
#pragma unroll
	for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
		for (int dx = -1; dx <= 1; ++dx)
			acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(G.w[dy + 1], G.w[dx + 1]),
			                               (float)src[(size_t)(y + dy) * row_bytes + xb + dx * 3]));
	int v = __float2int_rz(acc); /* NPP truncates (pinned against NPP 12.4 on a B200) */
	v = v > 255 ? 255 : v;
	dst[(size_t)y * row_bytes + xb] = (uint8_t)v;
}

__device__ __forceinline__ void sort2(int& a, int& b) {
	const int lo = min(a, b), hi = max(a, b);
	a = lo;
	b = hi;
}

__global__ void __launch_bounds__(kThreads)
median3x3_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int width, int height) {
	const int row_bytes = width * 3;
	const int inner = (width - 2) * 3;
	const long long n = (long long)inner * (height - 2);
	const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
	if (t >= n) return;
	const int y = (int)(t / inner) + 1;
	const int xb = (int)(t % inner) + 3;
	int p[9];
#pragma unroll
	for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
		for (int dx = -1; dx <= 1; ++dx) p[(dy + 1) * 3 + dx + 1] = src[(size_t)(y + dy) * row_bytes + xb + dx * 3];
	/* 19-exchange median-of-9 network */
	sort2(p[1], p[2]); sort2(p[4], p[5]); sort2(p[7], p[8]);
	sort2(p[0], p[1]); sort2(p[3], p[4]); sort2(p[6], p[7]);
	sort2(p[1], p[2]); sort2(p[4], p[5]); sort2(p[7], p[8]);
	sort2(p[0], p[3]); sort2(p[5], p[8]); sort2(p[4], p[7]);
	sort2(p[3], p[6]); sort2(p[1], p[4]); sort2(p[2], p[5]);
	sort2(p[4], p[7]); sort2(p[4], p[2]); sort2(p[6], p[4]);
	sort2(p[4], p[2]);
	/* reference quirk (kernels.cu:625 passes the un-offset destination): the result of the window centred on
	 * (y, x) is stored at (y-1, x-1); the last two rows / columns keep the splat output */
	dst[(size_t)(y - 1) * row_bytes + xb - 3] = (uint8_t)p[4];
}

/* uint8 -> float without the quarter-rate conversion pipe (16 lanes/clk/SM on B200, nine conversions per output byte
 * made the filter XU-bound): 0x4B000000 | b is the float 2^23 + b and the subtraction is exact */
__device__ __forceinline__ float byte_to_float(uint32_t b) { return __fadd_rn(__uint_as_float(0x4B000000u | b), -8388608.0f); }

/* Batched form for the fuser: frames src[f] -> dst[f] (distinct pools), EVERY byte of dst written — the filtered
 * value where the eager operator writes one, the splat output elsewhere — so no copy pass is needed. 4 bytes per thread.
 *   gauss : dst(y, b) filtered for 1 <= y <= H-2, 3 <= b < 3W-3                               (kernels.cu:632-653)
 *   median: dst(y, b) = median of the window centred on (y+1, b+3) for y <= H-3, b < 3W-6    (the shifted write) */
template <int kMedian>
__global__ void __launch_bounds__(kThreads)
post_filter_batch_kernel(const GaussTaps G, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int width, int height,
                         size_t frame_bytes) {
	const int row_bytes = width * 3;
	const size_t t = ((size_t)blockIdx.x * kThreads + threadIdx.x) * 4;
	if (t >= frame_bytes) return;
	const uint8_t* s = src + (size_t)blockIdx.y * frame_bytes;
	uint8_t* d = dst + (size_t)blockIdx.y * frame_bytes;
	const int y0 = (int)((uint32_t)t / (uint32_t)row_bytes), b0 = (int)((uint32_t)t - (uint32_t)y0 * (uint32_t)row_bytes);
	uint32_t packed = 0;
	float ww[9]; /* the nine tap products, rounded once like the reference's w[dy] * w[dx] */
#pragma unroll
	for (int k = 0; k < 9; ++k) ww[k] = __fmul_rn(G.w[k / 3], G.w[k % 3]);
#pragma unroll
	for (int e = 0; e < 4; ++e) {
		const size_t i = t + e;
		if (i >= frame_bytes) break;
		int y = y0, b = b0 + e; /* one 32-bit division per thread; a quad may straddle a row end when 3W % 4 != 0 */
		if (b >= row_bytes) {
			b -= row_bytes;
			++y;
		}
		const int cy = kMedian ? y + 1 : y, cb = kMedian ? b + 3 : b; /* centre of the window that produces this byte */
		int v;
		if (cy >= 1 && cy <= height - 2 && cb >= 3 && cb < row_bytes - 3) {
			if (kMedian) {
				int p[9];
#pragma unroll
				for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
					for (int dx = -1; dx <= 1; ++dx) p[(dy + 1) * 3 + dx + 1] = s[(size_t)(cy + dy) * row_bytes + cb + dx * 3];
				sort2(p[1], p[2]); sort2(p[4], p[5]); sort2(p[7], p[8]);
				sort2(p[0], p[1]); sort2(p[3], p[4]); sort2(p[6], p[7]);
				sort2(p[1], p[2]); sort2(p[4], p[5]); sort2(p[7], p[8]);
				sort2(p[0], p[3]); sort2(p[5], p[8]); sort2(p[4], p[7]);
				sort2(p[3], p[6]); sort2(p[1], p[4]); sort2(p[2], p[5]);
				sort2(p[4], p[7]); sort2(p[4], p[2]); sort2(p[6], p[4]);
				sort2(p[4], p[2]);
				v = p[4];
			} else {
				float acc = 0.0f;
#pragma unroll
				for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
					for (int dx = -1; dx <= 1; ++dx)
						acc = __fadd_rn(acc, __fmul_rn(ww[(dy + 1) * 3 + dx + 1], byte_to_float(s[(size_t)(cy + dy) * row_bytes + cb + dx * 3])));
				v = __float2int_rz(acc);
				v = v > 255 ? 255 : v;
			}
		} else {
			v = s[i];
		}
		packed |= (uint32_t)v << (8 * e);
	}
	if (t + 4 <= frame_bytes) {
		*reinterpret_cast<uint32_t*>(d + t) = packed;
	} else {
		for (int e = 0; t + e < frame_bytes; ++e) d[t + e] = (uint8_t)(packed >> (8 * e));
	}
}

/* Gauss, strip form: a thread owns 4 bytes of a row (one aligned word) over kStripRows output rows and slides down the
 * image. Each source row is loaded as three words, converted once, multiplied once per distinct tap (corner w0*w0, edge
 * w0*w1 = w1*w0, centre w1*w1: 22 products for the 10 source bytes) and added into the three output rows it belongs to —
 * in the reference's raster order (dy outer, dx inner), so the sums are bit-identical to the eager kernel.
 * Needs 3W % 4 == 0. ~24 instructions per output byte instead of ~50.
 * (A packed form — FMUL2 for the products, FADD2 for the sums, two output bytes per instruction — was tried and is NOT
 * usable: ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with -fmad=false, and NPP rounds every product;
 * the fused frames then differ from the reference, tests/test_fuse_gpu.py.) */
constexpr int kStripRows = 16;

__global__ void __launch_bounds__(kThreads)
gauss_strip_batch_kernel(const GaussTaps G, const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int width, int height,
                         size_t frame_bytes) {
	const int row_words = width * 3 / 4, row_bytes = width * 3;
	const int q = blockIdx.x * kThreads + threadIdx.x; /* word inside the row */
	if (q >= row_words) return;
	const int y_first = blockIdx.y * kStripRows;
	const uint32_t* s = reinterpret_cast<const uint32_t*>(src + (size_t)blockIdx.z * frame_bytes);
	uint32_t* d = reinterpret_cast<uint32_t*>(dst + (size_t)blockIdx.z * frame_bytes);
	const float wc = __fmul_rn(G.w[0], G.w[0]), we = __fmul_rn(G.w[0], G.w[1]), wm = __fmul_rn(G.w[1], G.w[1]);
	const int qm = q > 0 ? q - 1 : 0, qp = q + 1 < row_words ? q + 1 : q; /* clamped: out-of-row taps only feed border bytes */
	float accA[4] = {0, 0, 0, 0}, accB[4] = {0, 0, 0, 0};
	uint32_t prev_word = 0;
	const int y_last = min(y_first + kStripRows, height); /* one past the last output row of the strip */
	for (int r = y_first - 1; r <= y_last; ++r) {
		const int rr = r < 0 ? 0 : (r >= height ? height - 1 : r);
		const uint32_t* row = s + (size_t)rr * row_words;
		const uint32_t w0 = __ldg(row + qm), w1 = __ldg(row + q), w2 = __ldg(row + qp);
		float f[12]; /* bytes 4q-4 .. 4q+7; f[1..10] are used */
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			f[k] = byte_to_float((w0 >> (8 * k)) & 0xFFu);
			f[4 + k] = byte_to_float((w1 >> (8 * k)) & 0xFFu);
			f[8 + k] = byte_to_float((w2 >> (8 * k)) & 0xFFu);
		}
		float accC[4];
		uint32_t packed = 0;
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			const float l = f[e + 1], c = f[e + 4], rgt = f[e + 7];
			const float lc = __fmul_rn(wc, l), le = __fmul_rn(we, l), ce = __fmul_rn(we, c), cm = __fmul_rn(wm, c),
			            rc = __fmul_rn(wc, rgt), re = __fmul_rn(we, rgt);
			/* output row r-1: this source row is its dy = +1 */
			const float done = __fadd_rn(__fadd_rn(__fadd_rn(accA[e], lc), ce), rc);
			/* output row r: dy = 0 */
			accA[e] = __fadd_rn(__fadd_rn(__fadd_rn(accB[e], le), cm), re);
			/* output row r+1: dy = -1 (0 + lc == lc) */
			accC[e] = __fadd_rn(__fadd_rn(lc, ce), rc);
			int v = __float2int_rz(done); /* NPP truncates */
			v = v > 255 ? 255 : v;
			const int b = 4 * q + e, yo = r - 1;
			if (!(yo >= 1 && yo <= height - 2 && b >= 3 && b < row_bytes - 3)) v = (int)((prev_word >> (8 * e)) & 0xFFu); /* border: splat output */
			packed |= (uint32_t)v << (8 * e);
		}
#pragma unroll
		for (int e = 0; e < 4; ++e) accB[e] = accC[e];
		if (r - 1 >= y_first && r - 1 < y_last) d[(size_t)(r - 1) * row_words + q] = packed;
		prev_word = w1;
	}
}

GaussTaps gauss_taps() {
	/* sigma = 0.4 + (3/2)*0.6 = 1.0 (comment at kernels.cu:649); separable taps exp(-d^2/2)/sum */
	const float e = expf(-0.5f);
	const float sum = 1.0f + 2.0f * e;
	GaussTaps G;
	G.w[0] = e / sum;
	G.w[1] = 1.0f / sum;
	G.w[2] = e / sum;
	return G;
}

}  // namespace

/* Frameset::filterColor over n frames of a pool, src -> dst (used by the fuser's post-filter mode) */
int launch_post_filter_batch(const uint8_t* d_src, uint8_t* d_dst, int width, int height, int n_frames, int median, cudaStream_t s) {
	if (n_frames <= 0) return DFB_OK;
	const size_t frame_bytes = (size_t)width * height * 3;
	DFB_REQUIRE(frame_bytes % 4 == 0 && frame_bytes < (1ull << 32) && width >= 2 && (reinterpret_cast<uintptr_t>(d_dst) & 3u) == 0,
	            "post filter: frames must be 4-byte aligned and smaller than 4 GB");
	const dim3 grid((unsigned)((frame_bytes / 4 + kThreads - 1) / kThreads), (unsigned)n_frames);
	if (median) {
		post_filter_batch_kernel<1><<<grid, kThreads, 0, s>>>(gauss_taps(), d_src, d_dst, width, height, frame_bytes);
	} else if ((width * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(d_src) & 3u) == 0 && n_frames <= 65535) {
		const int row_words = width * 3 / 4;
		const dim3 sgrid((unsigned)((row_words + kThreads - 1) / kThreads), (unsigned)((height + kStripRows - 1) / kStripRows), (unsigned)n_frames);
		gauss_strip_batch_kernel<<<sgrid, kThreads, 0, s>>>(gauss_taps(), d_src, d_dst, width, height, frame_bytes);
	} else {
		post_filter_batch_kernel<0><<<grid, kThreads, 0, s>>>(gauss_taps(), d_src, d_dst, width, height, frame_bytes);
	}
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

}  // namespace dfb

using namespace dfb;

extern "C" int dfb_op_filter_color(uint8_t* d_color, uint8_t* d_scratch, int width, int height, int median, void* stream) {
	DFB_REQUIRE(d_color && d_scratch, "NULL argument");
	DFB_REQUIRE(width > 0 && height > 0, "bad image size");
	int rc = require_device();
	if (rc != DFB_OK) return rc;
	if (width < 3 || height < 3) return DFB_OK;
	cudaStream_t s = static_cast<cudaStream_t>(stream);
	const size_t bytes = (size_t)width * height * 3;
	DFB_CUDA(cudaMemcpyAsync(d_scratch, d_color, bytes, cudaMemcpyDeviceToDevice, s)); /* kernels.cu:616-620,644-647 */
	const long long n = (long long)(width - 2) * 3 * (height - 2);
	const unsigned blocks = (unsigned)((n + kThreads - 1) / kThreads);
	if (median) {
		median3x3_kernel<<<blocks, kThreads, 0, s>>>(d_scratch, d_color, width, height);
	} else {
		gauss3x3_kernel<<<blocks, kThreads, 0, s>>>(gauss_taps(), d_scratch, d_color, width, height);
	}
	count_launch();
	DFB_CUDA(cudaGetLastError());
	return DFB_OK;
}

