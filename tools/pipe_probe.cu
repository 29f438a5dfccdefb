// pipe_probe.cu — instruction-throughput probes for the splat body's instruction mix on sm_100a:
// scalar FFMA vs packed FFMA2, F2I / I2F (XU pipe), IMAD. One CTA of 1024 threads per SM, 8 independent
// chains per thread, cycles measured with clock64 on SM 0. Build: nvcc -gencode arch=compute_100a,code=sm_100a -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 2048;

template <int kMode>
__global__ void probe(float* out, long long* cycles, float seed, float seed2 = 1e-7f) {
	float a[8], b[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) {
		a[i] = seed + threadIdx.x * 1e-3f + i;
		b[i] = seed * 0.5f + i;
	}
	const float m = 1.0000001f, c = 1e-7f;
	unsigned q[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) q[i] = threadIdx.x * 2654435761u + i;
	__syncthreads();
	const long long t0 = clock64();
	for (int it = 0; it < kIters; ++it) {
		if (kMode == 0) { /* 16 scalar FFMA */
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				a[i] = __fmaf_rn(a[i], m, c);
				b[i] = __fmaf_rn(b[i], m, c);
			}
		} else if (kMode == 1) { /* 8 packed FFMA2 = the same 16 FMAs */
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				float2 v = __ffma2_rn(make_float2(a[i], b[i]), make_float2(m, m), make_float2(c, c));
				a[i] = v.x;
				b[i] = v.y;
			}
		} else if (kMode == 2) { /* 8 F2I + 8 I2F */
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				const int q = __float2int_rz(a[i]);
				a[i] = (float)(q & 0xFFFF) + b[i];
			}
		} else if (kMode == 3) { /* 16 IMAD */
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				int x = __float_as_int(a[i]), y = __float_as_int(b[i]);
				x = x * 3 + y;
				y = y * 5 + x;
				a[i] = __int_as_float(x);
				b[i] = __int_as_float(y);
			}
		} else if (kMode == 5 || kMode == 6 || kMode == 7) { /* FMA work + 16 independent integer (ALU pipe) ops */
			if (kMode == 5) {
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					float2 v = __ffma2_rn(make_float2(a[i], b[i]), make_float2(m, m), make_float2(c, c));
					a[i] = v.x;
					b[i] = v.y;
				}
			} else if (kMode == 6) {
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					a[i] = __fmaf_rn(a[i], m, c);
					b[i] = __fmaf_rn(b[i], m, c);
				}
			}
#pragma unroll
			for (int i = 0; i < 16; ++i) q[i] = (q[i] ^ (q[(i + 1) & 15] >> 3)) | 0x11u;
		} else if (kMode == 8) { /* 8 FFMA2 whose multiplier is a SCALAR from a uniform register (the splat body's form) */
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				float2 v = __ffma2_rn(make_float2(a[i], b[i]), make_float2(seed, seed), make_float2(a[(i + 1) & 7], b[(i + 1) & 7]));
				a[i] = v.x;
				b[i] = v.y;
			}
		} else if (kMode == 9) { /* 8 FMUL2 + 8 FADD2 with uniform scalars */
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				float2 v = __fmul2_rn(make_float2(a[i], b[i]), make_float2(seed, seed));
				v = __fadd2_rn(v, make_float2(seed2, seed2));
				a[i] = v.x;
				b[i] = v.y;
			}
		} else if (kMode == 4) { /* 8 FFMA2 + 8 scalar FFMA interleaved */
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				float2 v = __ffma2_rn(make_float2(a[i], b[i]), make_float2(m, m), make_float2(c, c));
				a[i] = v.x;
				b[i] = v.y;
				a[i + 4] = __fmaf_rn(a[i + 4], m, c);
				b[i + 4] = __fmaf_rn(b[i + 4], m, c);
			}
		}
	}
	const long long t1 = clock64();
	float s = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) s += a[i] + b[i];
	if (kMode >= 5) {
#pragma unroll
		for (int i = 0; i < 16; ++i) s += (float)q[i];
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int kMode>
void run(const char* name, double ops_per_iter_per_thread, int threads) {
	float* out;
	long long* cyc;
	cudaMalloc(&out, 148 * 1024 * sizeof(float));
	cudaMalloc(&cyc, 8);
	probe<kMode><<<148, threads>>>(out, cyc, 1.0f);
	probe<kMode><<<148, threads>>>(out, cyc, 1.0f);
	cudaDeviceSynchronize();
	long long h;
	cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
	const double warps = threads / 32.0;
	printf("%-28s threads %4d: %8lld cycles, %.2f warp-instr/cycle/SM, %.1f lane-ops/cycle/SM\n", name, threads, h,
	       warps * kIters * ops_per_iter_per_thread / h, 32.0 * warps * kIters * ops_per_iter_per_thread / h);
	cudaFree(out);
	cudaFree(cyc);
}

int main() {
	for (int threads : {512, 1024}) {
		run<0>("16 FFMA", 16, threads);
		run<1>("8 FFMA2 (=16 FMA)", 8, threads);
		run<2>("8 F2I + 8 I2F (+8 LOP,8 FADD)", 16, threads);
		run<3>("16 IMAD", 16, threads);
		run<4>("4 FFMA2 + 8 FFMA", 12, threads);
		run<5>("8 FFMA2 + 16 LOP3/SHF", 8 + 32, threads);
		run<6>("16 FFMA + 16 LOP3/SHF", 16 + 32, threads);
		run<7>("16 LOP3/SHF only", 32, threads);
		run<8>("8 FFMA2 with UR scalar", 8, threads);
		run<9>("8 FMUL2 + 8 FADD2 UR scalar", 16, threads);
	}
	return 0;
}
