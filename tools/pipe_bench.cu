/*
 * Integer-pipe micro-benchmark for the roofline denominators (SURVEY 8d asks
 * for measured IMAD / ALU peaks instead of Hopper-like assumptions).
 *
 * Each test runs ILP independent dependency chains per thread, 1024 threads per
 * SM on every SM, and reports thread-level operations per clock per SM from
 * clock64() deltas (so the result does not depend on the DVFS state), plus
 * wall-clock Gop/s from CUDA events.
 *
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/bin/pipe_bench tools/pipe_bench.cu
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef unsigned long long u64;
#define ILP 8
#define ITERS 4096

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
	fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct result { u64 cycles; };

template <class Body>
__global__ void __launch_bounds__(1024) bench_kernel(u64 *sink, u64 *cycles,
		u64 seed, Body body) {
	u64 v[ILP];
#pragma unroll
	for (int i = 0; i < ILP; i++) {
		v[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B97F4A7C15ull;
	}
	const u64 k1 = seed | 1, k2 = (seed >> 3) | 0x8000000000000001ull;
	__syncthreads();
	const u64 t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < ILP; i++) {
			body(v[i], k1, k2);
		}
	}
	const u64 t1 = clock64();
	u64 acc = 0;
#pragma unroll
	for (int i = 0; i < ILP; i++) acc ^= v[i];
	if (acc == 0x1234567) sink[0] = acc;
	__syncthreads();
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct op_imad { /* 32-bit mad.lo: 1 IMAD */
	static constexpr double ops = 1; static const char *name() { return "IMAD (mad.lo.u32)"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		unsigned x = (unsigned) v;
		asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"((unsigned) k1), "r"((unsigned) k2));
		v = x;
	}
};
struct op_imad_wide { /* mad.wide.u32 with 64-bit accumulate: 1 IMAD.WIDE.U32 */
	static constexpr double ops = 1; static const char *name() { return "IMAD.WIDE.U32 (mad.wide.u32)"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(v) : "r"((unsigned) v), "r"((unsigned) k1));
	}
};
struct op_iadd3 { /* 32-bit add: IADD3 */
	static constexpr double ops = 1; static const char *name() { return "IADD3 (add.u32)"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		unsigned x = (unsigned) v;
		asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"((unsigned) k1));
		v = x;
	}
};
struct op_lop3 {
	static constexpr double ops = 1; static const char *name() { return "LOP3 (xor.b32)"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		unsigned x = (unsigned) v;
		asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"((unsigned) k1), "r"((unsigned) k2));
		v = x;
	}
};
struct op_add64 { /* 64-bit add: IADD3 + IADD3.X */
	static constexpr double ops = 1; static const char *name() { return "add.u64 (IADD3+IADD3.X)"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		asm volatile("add.u64 %0, %0, %1;" : "+l"(v) : "l"(k1));
	}
};
struct op_csub64 { /* conditional subtract: x >= m ? x - m : x */
	static constexpr double ops = 1; static const char *name() { return "csub64 (x>=m?x-m:x) + add"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		v += k1;
		v = v >= k2 ? v - k2 : v;
	}
};
struct op_mulhi64 {
	static constexpr double ops = 1; static const char *name() { return "mul.hi.u64"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		v = __umul64hi(v, k2) + k1;
	}
};
struct op_mullo64 {
	static constexpr double ops = 1; static const char *name() { return "mul.lo.u64"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		v = v * k2 + k1;
	}
};
struct op_shoup { /* lazy Shoup modmul: y*w - hi(y*w')*q */
	static constexpr double ops = 1; static const char *name() { return "shoup_lazy modmul"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		const u64 q = k1 >> 4;
		v = v * k2 - __umul64hi(v, k1) * q;
	}
};
struct op_butterfly { /* Harvey CT butterfly on (v, v') pairs: counts as 1 butterfly */
	static constexpr double ops = 1; static const char *name() { return "Harvey CT butterfly (per chain: 1 bfly)"; }
	__device__ void operator()(u64 &v, u64 k1, u64 k2) const {
		const u64 q = k1 >> 4, twoq = 2 * q;
		u64 x = v, y = v ^ k2;
		const u64 xr = x >= twoq ? x - twoq : x;
		const u64 t = y * k2 - __umul64hi(y, k1) * q;
		x = xr + t;
		y = xr - t + twoq;
		v = x ^ (y >> 1);
	}
};

template <class Op>
static void run(int sms, double sm_mhz_hint) {
	u64 *sink, *cycles;
	CHECK(cudaMalloc(&sink, 8));
	CHECK(cudaMalloc(&cycles, sms * sizeof(u64)));
	cudaEvent_t e0, e1;
	CHECK(cudaEventCreate(&e0));
	CHECK(cudaEventCreate(&e1));
	for (int rep = 0; rep < 3; rep++) {
		CHECK(cudaEventRecord(e0));
		bench_kernel<<<sms, 1024>>>(sink, cycles, 0x123456789abcdefull + rep, Op());
		CHECK(cudaEventRecord(e1));
		CHECK(cudaEventSynchronize(e1));
	}
	float ms;
	CHECK(cudaEventElapsedTime(&ms, e0, e1));
	std::vector<u64> h(sms);
	CHECK(cudaMemcpy(h.data(), cycles, sms * sizeof(u64), cudaMemcpyDeviceToHost));
	double cyc = 0;
	for (int i = 0; i < sms; i++) cyc += (double) h[i];
	cyc /= sms;
	const double per_sm = 1024.0 * ILP * ITERS * Op::ops;
	printf("%-42s %8.2f ops/clk/SM   %9.1f Gop/s chip   (%.0f cyc, %.3f ms, eff clk %.0f MHz)\n",
			Op::name(), per_sm / cyc, per_sm * sms / (ms * 1e6), cyc, ms, cyc / (ms * 1e3));
	CHECK(cudaFree(sink));
	CHECK(cudaFree(cycles));
}

int main() {
	cudaDeviceProp prop;
	CHECK(cudaGetDeviceProperties(&prop, 0));
	printf("device: %s, %d SMs, clockRate %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
	const int sms = prop.multiProcessorCount;
	run<op_imad>(sms, 0);
	run<op_imad_wide>(sms, 0);
	run<op_iadd3>(sms, 0);
	run<op_lop3>(sms, 0);
	run<op_add64>(sms, 0);
	run<op_csub64>(sms, 0);
	run<op_mullo64>(sms, 0);
	run<op_mulhi64>(sms, 0);
	run<op_shoup>(sms, 0);
	run<op_butterfly>(sms, 0);
	return 0;
}
