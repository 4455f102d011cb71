/*
 * Register <-> lane transpose between two rounds of the radix-8 engine: through
 * a per-warp shared-memory buffer (what the row pass does: 8 STS.64 + 8 LDS.64,
 * __syncwarp) against warp shuffles (three butterfly steps of shfl.xor, two
 * 32-bit shuffles per 64-bit coefficient).  Both variants perform the same
 * 8 x 8 transpose of 64-bit values between the 8 registers of a thread and the
 * 8 lanes that differ in lane bits 4..2, which is the exchange between rounds
 * 0 and 1 of a 256-point tile (ntt_engine.cuh).  Reports cycles per exchange
 * per warp and exchanges per clock per SM; results are checked against each
 * other.
 *
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/bin/exchange_bench tools/exchange_bench.cu
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef unsigned long long u64;
#define ITERS 2048
#define THREADS 256

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
	fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

/* padded position, as xpad() in kernels_ntt.cu */
__device__ __forceinline__ int xpad(int i) { return i + ((i >> 5) << 2); }

/* new x[e] of lane l = old x[(l >> 2) & 7] of lane (l & 3) | (e << 2) */
__device__ __forceinline__ void exchange_smem(u64 (&x)[8], u64 *buf, int lane) {
	/* old layout: tile index i = (e << 5) | lane; new: i = ((lane >> 2) << 5) | (e << 2) | (lane & 3) */
#pragma unroll
	for (int e = 0; e < 8; e++) {
		buf[xpad((e << 5) | lane)] = x[e];
	}
	__syncwarp();
#pragma unroll
	for (int e = 0; e < 8; e++) {
		x[e] = buf[xpad(((lane >> 2) << 5) | (e << 2) | (lane & 3))];
	}
	__syncwarp();
}

__device__ __forceinline__ u64 shfl_xor64(u64 v, int mask) {
	unsigned lo = (unsigned) v, hi = (unsigned) (v >> 32);
	lo = __shfl_xor_sync(0xffffffffu, lo, mask);
	hi = __shfl_xor_sync(0xffffffffu, hi, mask);
	return ((u64) hi << 32) | lo;
}

__device__ __forceinline__ void exchange_shfl(u64 (&x)[8], int lane) {
	/* step k swaps register bit k with lane bit 2 + k */
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const bool upper = (lane >> (2 + k)) & 1;
#pragma unroll
		for (int e = 0; e < 8; e++) {
			if (e & (1 << k)) {
				continue;
			}
			/* the lower lane keeps x[e] and gives x[e | bit]; the upper lane
			 * keeps x[e | bit] and gives x[e] */
			const u64 give = upper ? x[e] : x[e | (1 << k)];
			const u64 got = shfl_xor64(give, 4 << k);
			if (upper) {
				x[e] = got;
			} else {
				x[e | (1 << k)] = got;
			}
		}
	}
}

template <bool SHFL>
__global__ void __launch_bounds__(THREADS) bench_kernel(u64 *out, u64 *cycles,
		u64 seed) {
	__shared__ u64 sm[(THREADS / 32) * (256 + 32 + 4)];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u64 *buf = sm + warp * (256 + 32 + 4);
	u64 x[8];
#pragma unroll
	for (int e = 0; e < 8; e++) {
		x[e] = seed + ((u64) blockIdx.x << 20) + (threadIdx.x << 3) + e;
	}
	__syncthreads();
	const u64 t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; it++) {
		if (SHFL) {
			exchange_shfl(x, lane);
		} else {
			exchange_smem(x, buf, lane);
		}
		x[it & 7] += it;   /* keep the iterations dependent and distinct */
	}
	const u64 t1 = clock64();
#pragma unroll
	for (int e = 0; e < 8; e++) {
		out[((size_t) blockIdx.x * THREADS + threadIdx.x) * 8 + e] = x[e];
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		cycles[blockIdx.x] = t1 - t0;
	}
}

int main() {
	cudaDeviceProp prop;
	CHECK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	const int ctas_per_sm = 2048 / THREADS;
	const int blocks = sms * ctas_per_sm;
	const size_t words = (size_t) blocks * THREADS * 8;
	u64 *out[2], *cycles;
	CHECK(cudaMalloc(&out[0], words * 8));
	CHECK(cudaMalloc(&out[1], words * 8));
	CHECK(cudaMalloc(&cycles, blocks * 8));
	std::vector<u64> cyc(blocks);
	double per_exchange[2];
	for (int v = 0; v < 2; v++) {
		for (int rep = 0; rep < 2; rep++) {
			if (v) {
				bench_kernel<true><<<blocks, THREADS>>>(out[v], cycles, 12345);
			} else {
				bench_kernel<false><<<blocks, THREADS>>>(out[v], cycles, 12345);
			}
			CHECK(cudaDeviceSynchronize());
		}
		CHECK(cudaMemcpy(cyc.data(), cycles, blocks * 8, cudaMemcpyDeviceToHost));
		double sum = 0;
		for (int b = 0; b < blocks; b++) {
			sum += (double) cyc[b];
		}
		const double cta_cycles = sum / blocks;
		/* ctas_per_sm CTAs of THREADS/32 warps run concurrently on an SM */
		const double warp_exchanges = (double) ITERS * (THREADS / 32) * ctas_per_sm;
		per_exchange[v] = cta_cycles / warp_exchanges;
		printf("%-28s %8.2f SM-cycles per warp exchange (256 coefficients), "
				"%6.2f coefficients/clk/SM\n",
				v ? "warp shuffles (shfl.xor)" : "shared memory + __syncwarp",
				per_exchange[v], 256.0 / per_exchange[v]);
	}
	std::vector<u64> a(words), b(words);
	CHECK(cudaMemcpy(a.data(), out[0], words * 8, cudaMemcpyDeviceToHost));
	CHECK(cudaMemcpy(b.data(), out[1], words * 8, cudaMemcpyDeviceToHost));
	size_t bad = 0;
	for (size_t i = 0; i < words; i++) {
		bad += a[i] != b[i];
	}
	printf("results identical: %s; shuffles / shared memory = %.2fx\n",
			bad ? "NO" : "yes", per_exchange[1] / per_exchange[0]);
	return bad != 0;
}
