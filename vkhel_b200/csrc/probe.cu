/*
 * Integer-pipe probes: the denominators of the transform's binding roofline,
 * measured on the device the context runs on, in the run that reports them
 * (bench.py's `issue_roofline`), instead of constants from an earlier session.
 *
 * The NTT butterfly is bound by the "fmaheavy" pipe, which executes IMAD,
 * IMAD.WIDE and IMAD.HI.  Every probe is a kernel of 1024 threads per SM on
 * every SM, each thread running CHAINS independent dependency chains of one
 * instruction for ITERS iterations (written in PTX so that ptxas can neither
 * remove nor re-balance them); the rate is thread-instructions per SM clock,
 * with the SM clock taken from clock64() and the elapsed time from CUDA events
 * (their ratio is the clock the probe ran at).  The butterfly probes run the
 * library's own butterflies (modarith.cuh: ct_lazy3 / gs_lazy3, the code the
 * transform kernels inline) on register operands only -- no loads, exchanges
 * or stores -- which is the rate the kernels would reach if those were free.
 *
 * Nothing here is on the transform path; reference: none (the reference has no
 * measurement code at all).
 */
#include "common.cuh"
#include "vkhel_ext.h"

#define PROBE_THREADS 1024
#define PROBE_CHAINS 8
#define PROBE_ITERS 4096

enum { OP_IMAD = 0, OP_IMAD_WIDE, OP_IMAD_HI, OP_LOP3, OP_BFLY_CT, OP_BFLY_GS,
	OP_SHF, OP_IADD3, OP_ADD64_3, OP_CSUB64, OP_COUNT };

struct probe_args {
	u64 q, w, wp;
	u64 zero;        /* opaque zero, as in the transform kernels */
	u64 *sink;
	u64 *cycles;     /* per block */
};

template <int OP>
__global__ void __launch_bounds__(PROBE_THREADS)
probe_kernel(const __grid_constant__ probe_args a) {
	unsigned r[PROBE_CHAINS];
	u64 acc[PROBE_CHAINS];
	u64 x[PROBE_CHAINS / 2], y[PROBE_CHAINS / 2];
	const unsigned b = (unsigned) a.w | 1u, c = (unsigned) a.wp | 3u;
#pragma unroll
	for (int i = 0; i < PROBE_CHAINS; i++) {
		r[i] = threadIdx.x * 2654435761u + i;
		acc[i] = (u64) r[i] * 0x9E3779B97F4A7C15ull;
	}
#pragma unroll
	for (int i = 0; i < PROBE_CHAINS / 2; i++) {
		x[i] = (a.q >> 1) + threadIdx.x * 977 + i;
		y[i] = (a.q >> 2) + threadIdx.x * 131 + 7 * i;
	}
	const u64 q = a.q, bq = 3 * a.q;
	__syncthreads();
	const u64 t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < PROBE_ITERS; it++) {
#pragma unroll
		for (int i = 0; i < PROBE_CHAINS; i++) {
			if (OP == OP_IMAD) {
				asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
			} else if (OP == OP_IMAD_WIDE) {
				/* 32 x 32 -> 64 product; both result words feed the next
				 * multiplicand through one ALU-pipe instruction, so that the
				 * chain is a true dependency and nothing can be hoisted */
				asm volatile("{\n\t.reg .u32 lo, hi;\n\t"
						"mul.wide.u32 %0, %1, %2;\n\t"
						"mov.b64 {lo, hi}, %0;\n\t"
						"lop3.b32 %1, lo, hi, %3, 0x96;\n\t}"
						: "+l"(acc[i]), "+r"(r[i]) : "r"(c), "r"(b));
			} else if (OP == OP_IMAD_HI) {
				asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));
			} else if (OP == OP_LOP3) {
				asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(c));
			} else if (OP == OP_SHF) {
				/* funnel shift, the form a sparse-modulus product would use */
				asm volatile("shf.l.clamp.b32 %0, %1, %0, 28;" : "+r"(r[i]) : "r"(b));
			} else if (OP == OP_IADD3) {
				/* three-input 32-bit addition */
				asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;"
						: "+r"(r[i]) : "r"(b), "r"((unsigned) a.zero));
			} else if (OP == OP_ADD64_3) {
				/* three-input 64-bit addition: IADD3 with two carry-outs +
				 * IADD3.X with two carry-ins */
				asm volatile("add.u64 %0, %0, %1;\n\tadd.u64 %0, %0, %2;"
						: "+l"(acc[i]) : "l"(a.wp), "l"(a.zero));
			} else if (OP == OP_CSUB64) {
				/* the butterflies' conditional subtraction (+ an addition that
				 * keeps the value moving) */
				acc[i] = csub(acc[i] + a.wp + a.zero, bq);
			}
		}
		if (OP == OP_BFLY_CT || OP == OP_BFLY_GS) {
#pragma unroll
			for (int i = 0; i < PROBE_CHAINS / 2; i++) {
				if (OP == OP_BFLY_CT) {
					ct_lazy3(x[i], y[i], a.w, a.wp, q, bq, a.zero);
				} else {
					gs_lazy3(x[i], y[i], a.w, a.wp, q, bq, a.zero);
				}
			}
			/* rotate the pairs so that the x and y roles mix as in a transform */
			const u64 tmp = y[0];
#pragma unroll
			for (int i = 0; i < PROBE_CHAINS / 2 - 1; i++) {
				y[i] = y[i + 1];
			}
			y[PROBE_CHAINS / 2 - 1] = tmp;
		}
	}
	const u64 t1 = clock64();
	u64 out = 0;
#pragma unroll
	for (int i = 0; i < PROBE_CHAINS; i++) {
		out ^= r[i] ^ acc[i];
	}
#pragma unroll
	for (int i = 0; i < PROBE_CHAINS / 2; i++) {
		out ^= x[i] ^ y[i];
	}
	if (out == 0x123456789abcdefull) {
		a.sink[0] = out;   /* never true in practice; keeps the chains alive */
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		a.cycles[blockIdx.x] = t1 - t0;
	}
}

template <int OP>
static void run_probe(struct vkhel_ctx *ctx, const probe_args &a, u64 *host_cycles,
		double *per_clk_sm, double *clock_mhz) {
	const int blocks = ctx->dev.sm_count;
	cudaStream_t stream = ctx_stream(ctx);
	cudaEvent_t e0, e1;
	CUDA_CHECK(cudaEventCreate(&e0));
	CUDA_CHECK(cudaEventCreate(&e1));
	double best_rate = 0, best_clock = 0;
	for (int rep = 0; rep < 4; rep++) {   /* the first repetition warms up */
		CUDA_CHECK(cudaEventRecord(e0, stream));
		probe_kernel<OP><<<blocks, PROBE_THREADS, 0, stream>>>(a);
		CUDA_CHECK(cudaGetLastError());
		CUDA_CHECK(cudaEventRecord(e1, stream));
		CUDA_CHECK(cudaMemcpyAsync(host_cycles, a.cycles, blocks * sizeof(u64),
					cudaMemcpyDeviceToHost, stream));
		CUDA_CHECK(cudaStreamSynchronize(stream));
		ctx->dev.launches++;
		float ms = 0;
		CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
		u64 worst = 0;
		for (int b = 0; b < blocks; b++) {
			worst = host_cycles[b] > worst ? host_cycles[b] : worst;
		}
		const bool bfly = OP == OP_BFLY_CT || OP == OP_BFLY_GS;
		const double ops = (double) PROBE_THREADS * PROBE_ITERS
			* (bfly ? PROBE_CHAINS / 2 : PROBE_CHAINS);
		const double rate = ops / (double) worst;
		if (rep > 0 && rate > best_rate) {
			best_rate = rate;
			best_clock = (double) worst / (ms * 1e-3) / 1e6;
		}
	}
	CUDA_CHECK(cudaEventDestroy(e0));
	CUDA_CHECK(cudaEventDestroy(e1));
	*per_clk_sm = best_rate;
	*clock_mhz = best_clock;
}

/* out[0..count): see vkhel_ext.h for the order.  Returns the number of values
 * written. */
extern "C" int vkhel_ctx_probe_int_peaks(struct vkhel_ctx *ctx, double *out,
		int count) {
	CUDA_CHECK(cudaSetDevice(ctx->dev.device));
	defer_flush(ctx);
	const int blocks = ctx->dev.sm_count;
	probe_args a;
	/* a 60-bit NTT prime (SURVEY App. C, P[0]) and an arbitrary twiddle */
	a.q = 1152921504606584833ull;
	a.w = 987813353222176621ull;
	a.wp = nt_compute_barrett_factor(a.w, a.q, 64);
	a.zero = 0;
	a.sink = (u64 *) device_alloc(ctx, sizeof(u64));
	a.cycles = (u64 *) device_alloc(ctx, blocks * sizeof(u64));
	u64 *host_cycles = (u64 *) malloc(blocks * sizeof(u64));
	VK_REQUIRE(host_cycles, "out of host memory");
	double rate[OP_COUNT], clk[OP_COUNT];
	run_probe<OP_IMAD>(ctx, a, host_cycles, &rate[OP_IMAD], &clk[OP_IMAD]);
	run_probe<OP_IMAD_WIDE>(ctx, a, host_cycles, &rate[OP_IMAD_WIDE], &clk[OP_IMAD_WIDE]);
	run_probe<OP_IMAD_HI>(ctx, a, host_cycles, &rate[OP_IMAD_HI], &clk[OP_IMAD_HI]);
	run_probe<OP_LOP3>(ctx, a, host_cycles, &rate[OP_LOP3], &clk[OP_LOP3]);
	run_probe<OP_BFLY_CT>(ctx, a, host_cycles, &rate[OP_BFLY_CT], &clk[OP_BFLY_CT]);
	run_probe<OP_BFLY_GS>(ctx, a, host_cycles, &rate[OP_BFLY_GS], &clk[OP_BFLY_GS]);
	run_probe<OP_SHF>(ctx, a, host_cycles, &rate[OP_SHF], &clk[OP_SHF]);
	run_probe<OP_IADD3>(ctx, a, host_cycles, &rate[OP_IADD3], &clk[OP_IADD3]);
	run_probe<OP_ADD64_3>(ctx, a, host_cycles, &rate[OP_ADD64_3], &clk[OP_ADD64_3]);
	run_probe<OP_CSUB64>(ctx, a, host_cycles, &rate[OP_CSUB64], &clk[OP_CSUB64]);
	free(host_cycles);
	device_free(ctx, a.sink);
	device_free(ctx, a.cycles);
	const double vals[] = {
		(double) blocks, clk[OP_BFLY_CT],
		rate[OP_IMAD], rate[OP_IMAD_WIDE], rate[OP_IMAD_HI], rate[OP_LOP3],
		rate[OP_BFLY_CT], rate[OP_BFLY_GS],
		rate[OP_SHF], rate[OP_IADD3], rate[OP_ADD64_3], rate[OP_CSUB64],
	};
	int n = (int) (sizeof(vals) / sizeof(vals[0]));
	if (n > count) {
		n = count;
	}
	for (int i = 0; i < n; i++) {
		out[i] = vals[i];
	}
	return n;
}
