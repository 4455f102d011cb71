/*
 * Butterfly formulation micro-benchmark: which way of writing the lazy
 * Cooley-Tukey / Gentleman-Sande butterfly costs the fewest FMA-pipe issue
 * slots on sm_100a.  Each thread keeps 4 (x,y) pairs in registers (the ILP of
 * one radix-8 stage) and applies the butterfly ITERS times; 1024 threads per
 * SM on every SM.  Output: butterflies per clock per SM (clock64 based).
 *
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/bin/bfly_bench tools/bfly_bench.cu
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef unsigned long long u64;
typedef unsigned int u32;
#define PAIRS 4
#define ITERS 2048

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
	fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct consts { u64 q, twoq, w, wp; };
struct fp_consts { double dp0, c0, dp1, c1; };
struct tw48 { unsigned long long w, wp; fp_consts f; };
__device__ __forceinline__ fp_consts make_fp(unsigned long long wp);
__shared__ ulonglong2 sm_tw[256];

template <class B>
__global__ void __launch_bounds__(1024) bench(u64 *sink, u64 *cycles, consts c, B b) {
	u64 x[PAIRS], y[PAIRS];
#pragma unroll
	for (int i = 0; i < PAIRS; i++) {
		x[i] = (c.q >> 1) + threadIdx.x * 977 + i;
		y[i] = (c.q >> 2) + threadIdx.x * 131 + 7 * i;
	}
	if (threadIdx.x < 256) sm_tw[threadIdx.x] = make_ulonglong2(c.w + threadIdx.x, c.wp - threadIdx.x);
	{
		extern __shared__ tw48 sm48[];
		if (threadIdx.x < 128) {
			tw48 tw; tw.w = c.w + threadIdx.x; tw.wp = c.wp - threadIdx.x; tw.f = make_fp(tw.wp);
			sm48[threadIdx.x] = tw;
		}
	}
	__syncthreads();
	const u64 t0 = clock64();
#pragma unroll 1
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < PAIRS; i++) {
			b(x[i], y[i], c);
		}
		/* rotate the pairs so that x and y roles mix as in a real transform */
		const u64 tmp = y[0];
#pragma unroll
		for (int i = 0; i < PAIRS - 1; i++) y[i] = y[i + 1];
		y[PAIRS - 1] = tmp;
	}
	const u64 t1 = clock64();
	u64 acc = 0;
#pragma unroll
	for (int i = 0; i < PAIRS; i++) acc ^= x[i] ^ y[i];
	if (acc == 0x1234567) sink[0] = acc;
	__syncthreads();
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ u64 csub(u64 x, u64 m) { return x >= m ? x - m : x; }

/* ---- V0: Harvey CT as shipped (modarith.cuh ct_lazy) ---- */
struct ct_v0 { static const char *name() { return "CT v0 harvey (csub x, exact mulhi)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 xr = csub(x, c.twoq);
		const u64 t = y * c.w - __umul64hi(y, c.wp) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };

/* ---- V1: no range correction (cost of the csub) ---- */
struct ct_v1 { static const char *name() { return "CT v1 no csub (range grows 2q/stage)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 t = y * c.w - __umul64hi(y, c.wp) * c.q;
		const u64 xo = x;
		x = xo + t; y = xo - t + c.twoq;
	} };

/* approximate high product: drops the low x low partial product and the
 * carries of the middle column: result in [hi-2, hi] */
__device__ __forceinline__ u64 mulhi_approx(u64 a, u64 b) {
	const u32 a0 = (u32) a, a1 = (u32) (a >> 32), b0 = (u32) b, b1 = (u32) (b >> 32);
	return (u64) a1 * b1 + __umulhi(a1, b0) + __umulhi(a0, b1);
}

/* ---- V2: approximate mulhi, t in [0,4q), no csub ---- */
struct ct_v2 { static const char *name() { return "CT v2 approx mulhi (3 hi products), no csub"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 t = y * c.w - mulhi_approx(y, c.wp) * c.q;
		const u64 xo = x;
		x = xo + t; y = xo - t + 2 * c.twoq;
	} };

/* ---- V3: approximate mulhi + csub ---- */
struct ct_v3 { static const char *name() { return "CT v3 approx mulhi + csub(x,4q)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 xr = csub(x, 2 * c.twoq);
		const u64 t = y * c.w - mulhi_approx(y, c.wp) * c.q;
		x = xr + t; y = xr - t + 2 * c.twoq;
	} };

/* ---- V4: csub written with the sign of the difference ---- */
struct ct_v4 { static const char *name() { return "CT v4 harvey, csub via sign mask"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 d = x - c.twoq;
		const u64 xr = ((long long) d < 0) ? x : d;   /* valid while x < 2^63 */
		const u64 t = y * c.w - __umul64hi(y, c.wp) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };

/* ---- V5: fold the correction into 3-input adds (keeps adds on the ALU pipe) ---- */
struct ct_v5 { static const char *name() { return "CT v5 harvey, correction folded into 3-input adds"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 corr = x >= c.twoq ? c.twoq : 0;     /* 0 or 2q */
		const u64 t = y * c.w - __umul64hi(y, c.wp) * c.q;
		const u64 xo = x;
		x = xo - corr + t;
		y = xo + (c.twoq - corr) - t;
	} };

/* ---- V6: min-based csub: min(x, x - 2q) as unsigned ---- */
struct ct_v6 { static const char *name() { return "CT v6 harvey, csub = umin(x, x-2q)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 xr = min(x, x - c.twoq);
		const u64 t = y * c.w - __umul64hi(y, c.wp) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };

/* ---- mulhi through explicit 32-bit partial products, carries kept minimal ---- */
__device__ __forceinline__ u64 mulhi_explicit(u64 a, u64 b) {
	const u32 a0 = (u32) a, a1 = (u32) (a >> 32), b0 = (u32) b, b1 = (u32) (b >> 32);
	const u64 p00 = (u64) a0 * b0;
	const u64 p01 = (u64) a0 * b1 + (p00 >> 32);        /* cannot overflow */
	const u64 p10 = (u64) a1 * b0 + (u32) p01;          /* cannot overflow */
	return (u64) a1 * b1 + (p01 >> 32) + (p10 >> 32);
}
struct ct_v7 { static const char *name() { return "CT v7 harvey, mulhi as 4 mad.wide chain"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 xr = csub(x, c.twoq);
		const u64 t = y * c.w - mulhi_explicit(y, c.wp) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };

/* ---- V8: as v0 but w, w' are per-thread values in ordinary registers ---- */
struct ct_v8 { static const char *name() { return "CT v8 harvey, per-thread w/w' (vector regs)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		/* derive thread-dependent twiddles once; the compiler cannot keep
		 * them in uniform registers */
		const u64 w = c.w ^ (threadIdx.x & 1), wp = c.wp ^ (threadIdx.x & 2);
		const u64 xr = csub(x, c.twoq);
		const u64 t = y * w - __umul64hi(y, wp) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };

/* ---- V9: twiddle pair fetched from shared memory for every butterfly ---- */
struct ct_v9 { static const char *name() { return "CT v9 harvey, w/w' LDS.128 per butterfly"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const ulonglong2 w = sm_tw[(threadIdx.x + (unsigned) x) & 255];
		const u64 xr = csub(x, c.twoq);
		const u64 t = y * w.x - __umul64hi(y, w.y) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };

/* ---- V10: quotient estimate with the two cross products on the FP64 pipe ----
 * hi32(a*b) for 32-bit a, b is the low mantissa word of fma_rz(a, b, 2^84).
 * With A = 2^52 + a (bit pattern 0x43300000:a) and c = 2^84 - 2^52*b
 * precomputed per twiddle, fma_rz(A, b, c) = a*b + 2^84 exactly, so neither
 * operand needs a conversion.  h' = y1*p1 + hi32(y1*p0) + hi32(y0*p1) is in
 * [hi-2, hi]  =>  t in [0,4q), values in [0,8q), one csub(x, 4q). */
__device__ __forceinline__ u64 mulhi_fp64(u64 y, u64 wp, const fp_consts &f) {
	const u32 y0 = (u32) y, y1 = (u32) (y >> 32), p1 = (u32) (wp >> 32);
	const double A0 = __hiloint2double(0x43300000, (int) y0);
	const double A1 = __hiloint2double(0x43300000, (int) y1);
	const u32 k0 = (u32) __double2loint(__fma_rz(A1, f.dp0, f.c0)); /* hi32(y1*p0) */
	const u32 k1 = (u32) __double2loint(__fma_rz(A0, f.dp1, f.c1)); /* hi32(y0*p1) */
	return (u64) y1 * p1 + k0 + k1;
}
__device__ __forceinline__ fp_consts make_fp(u64 wp) {
	fp_consts f;
	const u32 p0 = (u32) wp, p1 = (u32) (wp >> 32);
	f.dp0 = (double) p0; f.c0 = 0x1p84 - 0x1p52 * (double) p0;
	f.dp1 = (double) p1; f.c1 = 0x1p84 - 0x1p52 * (double) p1;
	return f;
}
struct ct_v10 { static const char *name() { return "CT v10 fp64 cross products, csub(x,4q), uniform tw"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const fp_consts f = make_fp(c.wp);
		const u64 fourq = 2 * c.twoq;
		const u64 xr = csub(x, fourq);
		const u64 t = y * c.w - mulhi_fp64(y, c.wp, f) * c.q;
		x = xr + t; y = xr - t + fourq;
	} };
struct ct_v11 { static const char *name() { return "CT v11 fp64 cross products, 48-byte twiddle from smem"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		extern __shared__ tw48 sm48[];
		const tw48 tw = sm48[(threadIdx.x + (unsigned) x) & 127];
		const u64 fourq = 2 * c.twoq;
		const u64 xr = csub(x, fourq);
		const u64 t = y * tw.w - mulhi_fp64(y, tw.wp, tw.f) * c.q;
		x = xr + t; y = xr - t + fourq;
	} };
struct only_dfma { static const char *name() { return "2x DFMA only (fp64 pipe rate; per pair)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		double a = __longlong_as_double((long long) x), b = __longlong_as_double((long long) y);
		a = __fma_rz(a, 1.0000001, b); b = __fma_rz(b, 0.9999999, a);
		x = (u64) __double_as_longlong(a); y = (u64) __double_as_longlong(b);
	} };

/* ---- V12: csub through the borrow of a 64-bit subtraction (PTX carry chain) ---- */
__device__ __forceinline__ u64 csub_borrow(u64 x, u64 m) {
	const u32 xl = (u32) x, xh = (u32) (x >> 32), ml = (u32) m, mh = (u32) (m >> 32);
	u32 tl, th, b;
	asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;"
			: "=r"(tl), "=r"(th), "=r"(b) : "r"(xl), "r"(xh), "r"(ml), "r"(mh));
	const u32 rl = b ? xl : tl, rh = b ? xh : th;
	return ((u64) rh << 32) | rl;
}
struct ct_v12 { static const char *name() { return "CT v12 harvey, csub via borrow chain"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 xr = csub_borrow(x, c.twoq);
		const u64 t = y * c.w - __umul64hi(y, c.wp) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };
struct ct_v13 { static const char *name() { return "CT v13 = v12 with twiddle LDS.128 per butterfly"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const ulonglong2 w = sm_tw[(threadIdx.x + (unsigned) x) & 255];
		const u64 xr = csub_borrow(x, c.twoq);
		const u64 t = y * w.x - __umul64hi(y, w.y) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };
struct gs_v12 { static const char *name() { return "GS v12 csub via borrow chain"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 s = x + y, d = x - y + c.twoq;
		x = csub_borrow(s, c.twoq);
		y = d * c.w - __umul64hi(d, c.wp) * c.q;
	} };

/* ---- V14: v12 + mulhi as an explicit mad/madc carry chain ---- */
__device__ __forceinline__ u64 mulhi_chain(u64 a, u64 b) {
	const u32 a0 = (u32) a, a1 = (u32) (a >> 32), b0 = (u32) b, b1 = (u32) (b >> 32);
	u32 r0, r1, r2;
	asm("{\n\t"
		"mul.hi.u32 %0, %3, %5;\n\t"
		"mad.lo.cc.u32 %0, %3, %6, %0;\n\t"
		"madc.hi.u32 %1, %3, %6, 0;\n\t"
		"mad.lo.cc.u32 %0, %4, %5, %0;\n\t"
		"madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
		"addc.u32 %2, 0, 0;\n\t"
		"mad.lo.cc.u32 %1, %4, %6, %1;\n\t"
		"madc.hi.u32 %2, %4, %6, %2;\n\t"
		"}" : "=&r"(r0), "=&r"(r1), "=&r"(r2) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
	return ((u64) r2 << 32) | r1;
}
struct ct_v14 { static const char *name() { return "CT v14 = v12 + mulhi mad/madc chain"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 xr = csub_borrow(x, c.twoq);
		const u64 t = y * c.w - mulhi_chain(y, c.wp) * c.q;
		x = xr + t; y = xr - t + c.twoq;
	} };
/* ---- V15: v12 with the two low products fused: t = lo64(y*w - hi*q) by hand ---- */
struct ct_v15 { static const char *name() { return "CT v15 = v12, low products with mad.wide + mad.lo"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 xr = csub_borrow(x, c.twoq);
		const u64 hi = __umul64hi(y, c.wp);
		const u32 y0 = (u32) y, y1 = (u32) (y >> 32), w0 = (u32) c.w, w1 = (u32) (c.w >> 32);
		const u32 h0 = (u32) hi, h1 = (u32) (hi >> 32), q0 = (u32) c.q, q1 = (u32) (c.q >> 32);
		u64 a = (u64) y0 * w0;                       /* mul.wide */
		u32 ah = (u32) (a >> 32) + y0 * w1 + y1 * w0;  /* 2 mad.lo */
		u64 b = (u64) h0 * q0;
		u32 bh = (u32) (b >> 32) + h0 * q1 + h1 * q0;
		const u64 t = (((u64) ah << 32) | (u32) a) - (((u64) bh << 32) | (u32) b);
		x = xr + t; y = xr - t + c.twoq;
	} };

/* ---- V16: v12 with t = y*w + hi*(-q): no negation of the second product ---- */
struct ct_v16 { static const char *name() { return "CT v16 = v12, t = y*w + hi*(2^64-q)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 nq = 0 - c.q;
		const u64 xr = csub_borrow(x, c.twoq);
		const u64 t = y * c.w + __umul64hi(y, c.wp) * nq;
		x = xr + t; y = xr - t + c.twoq;
	} };
/* ---- V17: v16 + x' folded into the multiply-accumulate chain ---- */
struct ct_v17 { static const char *name() { return "CT v17 = v16, x' = (xr + y*w) + hi*nq, y' = 2xr+2q-x'"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 nq = 0 - c.q;
		const u64 xr = csub_borrow(x, c.twoq);
		const u64 hi = __umul64hi(y, c.wp);
		const u64 xn = hi * nq + (y * c.w + xr);
		y = (xr + xr + c.twoq) - xn;
		x = xn;
	} };
struct gs_v16 { static const char *name() { return "GS v16 = v12, t = d*w + hi*(2^64-q)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 nq = 0 - c.q;
		const u64 s = x + y, d = x - y + c.twoq;
		x = csub_borrow(s, c.twoq);
		y = d * c.w + __umul64hi(d, c.wp) * nq;
	} };

/* ---- V18: v14 without the a0*b0 partial product: h' in {hi-1, hi}, t in [0,3q),
 * values in [0,6q), one csub(x, 3q); needs 6q < 2^64 ---- */
__device__ __forceinline__ u64 mulhi_chain_approx(u64 a, u64 b) {
	const u32 a0 = (u32) a, a1 = (u32) (a >> 32), b0 = (u32) b, b1 = (u32) (b >> 32);
	u32 r1, r2;
	asm("{\n\t"
		".reg .u32 r0;\n\t"
		"mul.lo.u32 r0, %2, %5;\n\t"
		"mul.hi.u32 %0, %2, %5;\n\t"
		"mad.lo.cc.u32 r0, %3, %4, r0;\n\t"
		"madc.hi.cc.u32 %0, %3, %4, %0;\n\t"
		"addc.u32 %1, 0, 0;\n\t"
		"mad.lo.cc.u32 %0, %3, %5, %0;\n\t"
		"madc.hi.u32 %1, %3, %5, %1;\n\t"
		"}" : "=&r"(r1), "=&r"(r2) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
	return ((u64) r2 << 32) | r1;
}
struct ct_v18 { static const char *name() { return "CT v18 = v14 minus lo*lo product, csub(x,3q)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q;
		const u64 xr = csub_borrow(x, threeq);
		const u64 t = y * c.w - mulhi_chain_approx(y, c.wp) * c.q;
		x = xr + t; y = xr - t + threeq;
	} };
struct gs_v18 { static const char *name() { return "GS v18 approx chain, range [0,3q)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q;
		const u64 s = x + y, d = x - y + threeq;
		x = csub_borrow(s, threeq);
		y = d * c.w - mulhi_chain_approx(d, c.wp) * c.q;
	} };
struct gs_v14 { static const char *name() { return "GS v14 exact chain"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 s = x + y, d = x - y + c.twoq;
		x = csub_borrow(s, c.twoq);
		y = d * c.w - mulhi_chain(d, c.wp) * c.q;
	} };

/* ---- V19: v18 with the low 64 bits y*w + h*(2^64-q) as one explicit mad chain ---- */
__device__ __forceinline__ u64 shoup3_ptx(u64 y, u64 w, u64 wp, u64 nq) {
	const u32 y0 = (u32) y, y1 = (u32) (y >> 32), w0 = (u32) w, w1 = (u32) (w >> 32);
	const u32 p0 = (u32) wp, p1 = (u32) (wp >> 32), n0 = (u32) nq, n1 = (u32) (nq >> 32);
	u32 t0, t1;
	asm("{\n\t"
		".reg .u32 r0, h0, h1;\n\t"
		/* h = approximate high 64 bits of y * wp */
		"mul.lo.u32 r0, %2, %7;\n\t"
		"mul.hi.u32 h0, %2, %7;\n\t"
		"mad.lo.cc.u32 r0, %3, %6, r0;\n\t"
		"madc.hi.cc.u32 h0, %3, %6, h0;\n\t"
		"addc.u32 h1, 0, 0;\n\t"
		"mad.lo.cc.u32 h0, %3, %7, h0;\n\t"
		"madc.hi.u32 h1, %3, %7, h1;\n\t"
		/* t = lo64(y*w) + lo64(h*nq) */
		"mul.lo.u32 %0, %2, %4;\n\t"
		"mul.hi.u32 %1, %2, %4;\n\t"
		"mad.lo.cc.u32 %0, h0, %8, %0;\n\t"
		"madc.hi.u32 %1, h0, %8, %1;\n\t"
		"mad.lo.u32 %1, %2, %5, %1;\n\t"
		"mad.lo.u32 %1, %3, %4, %1;\n\t"
		"mad.lo.u32 %1, h0, %9, %1;\n\t"
		"mad.lo.u32 %1, h1, %8, %1;\n\t"
		"}" : "=&r"(t0), "=&r"(t1)
		: "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(p0), "r"(p1), "r"(n0), "r"(n1));
	return ((u64) t1 << 32) | t0;
}
struct ct_v19 { static const char *name() { return "CT v19 = v18, whole Shoup product as one PTX chain"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q, nq = 0 - c.q;
		const u64 xr = csub_borrow(x, threeq);
		const u64 t = shoup3_ptx(y, c.w, c.wp, nq);
		x = xr + t; y = xr - t + threeq;
	} };
struct gs_v19 { static const char *name() { return "GS v19 whole Shoup product as one PTX chain"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q, nq = 0 - c.q;
		const u64 s = x + y, d = x - y + threeq;
		x = csub_borrow(s, threeq);
		y = shoup3_ptx(d, c.w, c.wp, nq);
	} };


/* ---- V20: x' comes out of the multiply chain (accumulator starts at xr), y' = (2xr+3q) - x' ---- */
__device__ __forceinline__ u64 shoup3_acc_ptx(u64 acc, u64 y, u64 w, u64 wp, u64 nq) {
	const u32 y0 = (u32) y, y1 = (u32) (y >> 32), w0 = (u32) w, w1 = (u32) (w >> 32);
	const u32 p0 = (u32) wp, p1 = (u32) (wp >> 32), n0 = (u32) nq, n1 = (u32) (nq >> 32);
	const u32 a0 = (u32) acc, a1 = (u32) (acc >> 32);
	u32 t0, t1;
	asm("{\n\t"
		".reg .u32 r0, h0, h1;\n\t"
		"mul.lo.u32 r0, %2, %7;\n\t"
		"mul.hi.u32 h0, %2, %7;\n\t"
		"mad.lo.cc.u32 r0, %3, %6, r0;\n\t"
		"madc.hi.cc.u32 h0, %3, %6, h0;\n\t"
		"addc.u32 h1, 0, 0;\n\t"
		"mad.lo.cc.u32 h0, %3, %7, h0;\n\t"
		"madc.hi.u32 h1, %3, %7, h1;\n\t"
		/* t = acc + lo64(y*w) + lo64(h*nq) */
		"mad.lo.cc.u32 %0, %2, %4, %10;\n\t"
		"madc.hi.u32 %1, %2, %4, %11;\n\t"
		"mad.lo.cc.u32 %0, h0, %8, %0;\n\t"
		"madc.hi.u32 %1, h0, %8, %1;\n\t"
		"mad.lo.u32 %1, %2, %5, %1;\n\t"
		"mad.lo.u32 %1, %3, %4, %1;\n\t"
		"mad.lo.u32 %1, h0, %9, %1;\n\t"
		"mad.lo.u32 %1, h1, %8, %1;\n\t"
		"}" : "=&r"(t0), "=&r"(t1)
		: "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(p0), "r"(p1), "r"(n0), "r"(n1),
		  "r"(a0), "r"(a1));
	return ((u64) t1 << 32) | t0;
}
struct ct_v20 { static const char *name() { return "CT v20 = v19, x' out of the chain (acc = xr), y' = 2xr+3q-x'"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q, nq = 0 - c.q;
		const u64 xr = csub_borrow(x, threeq);
		const u64 xn = shoup3_acc_ptx(xr, y, c.w, c.wp, nq);
		y = xr + xr + threeq - xn; x = xn;
	} };
struct ct_v24 { static const char *name() { return "CT v24 = v19 without csub (bound +3q per stage)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q, nq = 0 - c.q;
		const u64 xr = x;
		const u64 t = shoup3_ptx(y, c.w, c.wp, nq);
		x = xr + t; y = xr - t + threeq;
	} };
struct ct_v25 { static const char *name() { return "CT v25 = v20 without csub"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q, nq = 0 - c.q;
		const u64 xr = x;
		const u64 xn = shoup3_acc_ptx(xr, y, c.w, c.wp, nq);
		y = xr + xr + threeq - xn; x = xn;
	} };
struct gs_v24 { static const char *name() { return "GS v24 = v19 without csub"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q, nq = 0 - c.q;
		const u64 s = x + y, d = x - y + threeq;
		x = s;
		y = shoup3_ptx(d, c.w, c.wp, nq);
	} };
/* pipe-cost probes: N independent multiply forms per pair */
struct only_wide_rz { static const char *name() { return "probe: 4x mul.wide.u32 (C = RZ)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u32 a = (u32) x, b = (u32) (x >> 32), d = (u32) y, e = (u32) (y >> 32);
		u64 r0, r1, r2, r3;
		asm("mul.wide.u32 %0, %1, %2;" : "=l"(r0) : "r"(a), "r"((u32) c.w));
		asm("mul.wide.u32 %0, %1, %2;" : "=l"(r1) : "r"(b), "r"((u32) c.wp));
		asm("mul.wide.u32 %0, %1, %2;" : "=l"(r2) : "r"(d), "r"((u32) c.q));
		asm("mul.wide.u32 %0, %1, %2;" : "=l"(r3) : "r"(e), "r"((u32) (c.w >> 32)));
		x = r0 ^ r1; y = r2 ^ r3;
	} };
struct only_wide_acc { static const char *name() { return "probe: 4x mad.wide.u32 (64-bit C)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u32 a = (u32) x, b = (u32) (x >> 32), d = (u32) y, e = (u32) (y >> 32);
		u64 r0, r1;
		asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r0) : "r"(a), "r"((u32) c.w), "l"(y));
		asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r0) : "r"(b), "r"((u32) c.wp), "l"(r0));
		asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r1) : "r"(d), "r"((u32) c.q), "l"(x));
		asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r1) : "r"(e), "r"((u32) (c.w >> 32)), "l"(r1));
		x = r0; y = r1;
	} };
struct only_hi_acc { static const char *name() { return "probe: 4x mad.hi.u32"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		u32 a = (u32) x, b = (u32) (x >> 32), d = (u32) y, e = (u32) (y >> 32);
		asm("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"((u32) c.w), "r"(b));
		asm("mad.hi.u32 %0, %0, %1, %2;" : "+r"(b) : "r"((u32) c.wp), "r"(d));
		asm("mad.hi.u32 %0, %0, %1, %2;" : "+r"(d) : "r"((u32) c.q), "r"(e));
		asm("mad.hi.u32 %0, %0, %1, %2;" : "+r"(e) : "r"((u32) (c.w >> 32)), "r"(a));
		x = ((u64) b << 32) | a; y = ((u64) e << 32) | d;
	} };
struct only_lo { static const char *name() { return "probe: 4x mad.lo.u32"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		u32 a = (u32) x, b = (u32) (x >> 32), d = (u32) y, e = (u32) (y >> 32);
		asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"((u32) c.w), "r"(b));
		asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b) : "r"((u32) c.wp), "r"(d));
		asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"((u32) c.q), "r"(e));
		asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(e) : "r"((u32) (c.w >> 32)), "r"(a));
		x = ((u64) b << 32) | a; y = ((u64) e << 32) | d;
	} };


/* ---- V26: no IMAD.HI -- three wide products, word sums on the ALU ---- */
__device__ __forceinline__ u64 shoup3_nohi_ptx(u64 y, u64 w, u64 wp, u64 nq) {
	const u32 y0 = (u32) y, y1 = (u32) (y >> 32), w0 = (u32) w, w1 = (u32) (w >> 32);
	const u32 p0 = (u32) wp, p1 = (u32) (wp >> 32), n0 = (u32) nq, n1 = (u32) (nq >> 32);
	u32 t0, t1;
	asm("{\n\t"
		".reg .u32 a0, a1, b0, b1, c0, c1, j, h0, h1;\n\t"
		".reg .u64 A, B, C;\n\t"
		"mul.wide.u32 A, %2, %7;\n\t"          /* y0*p1 */
		"mul.wide.u32 B, %3, %6;\n\t"          /* y1*p0 */
		"mul.wide.u32 C, %3, %7;\n\t"          /* y1*p1 */
		"mov.b64 {a0, a1}, A;\n\t"
		"mov.b64 {b0, b1}, B;\n\t"
		"mov.b64 {c0, c1}, C;\n\t"
		"add.cc.u32 j, a0, b0;\n\t"            /* word 1: only its carry */
		"addc.cc.u32 h0, a1, b1;\n\t"          /* word 2 */
		"addc.u32 h1, c1, 0;\n\t"              /* word 3 */
		"add.cc.u32 h0, h0, c0;\n\t"
		"addc.u32 h1, h1, 0;\n\t"
		"mul.lo.u32 %0, %2, %4;\n\t"
		"mul.hi.u32 %1, %2, %4;\n\t"
		"mad.lo.cc.u32 %0, h0, %8, %0;\n\t"
		"madc.hi.u32 %1, h0, %8, %1;\n\t"
		"mad.lo.u32 %1, %2, %5, %1;\n\t"
		"mad.lo.u32 %1, %3, %4, %1;\n\t"
		"mad.lo.u32 %1, h0, %9, %1;\n\t"
		"mad.lo.u32 %1, h1, %8, %1;\n\t"
		"}" : "=&r"(t0), "=&r"(t1)
		: "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(p0), "r"(p1), "r"(n0), "r"(n1));
	return ((u64) t1 << 32) | t0;
}
struct ct_v26 { static const char *name() { return "CT v26 = v19 with three wide products, no IMAD.HI"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q, nq = 0 - c.q;
		const u64 xr = csub_borrow(x, threeq);
		const u64 t = shoup3_nohi_ptx(y, c.w, c.wp, nq);
		x = xr + t; y = xr - t + threeq;
	} };
/* ---- V27: v26 dropping the word-1 sum as well: h in [H-2, H], product in [0,4q) ---- */
__device__ __forceinline__ u64 shoup4_ptx(u64 y, u64 w, u64 wp, u64 nq) {
	const u32 y0 = (u32) y, y1 = (u32) (y >> 32), w0 = (u32) w, w1 = (u32) (w >> 32);
	const u32 p0 = (u32) wp, p1 = (u32) (wp >> 32), n0 = (u32) nq, n1 = (u32) (nq >> 32);
	u32 t0, t1;
	asm("{\n\t"
		".reg .u32 a0, a1, b0, b1, c0, c1, h0, h1;\n\t"
		".reg .u64 A, B, C;\n\t"
		"mul.wide.u32 A, %2, %7;\n\t"
		"mul.wide.u32 B, %3, %6;\n\t"
		"mul.wide.u32 C, %3, %7;\n\t"
		"mov.b64 {a0, a1}, A;\n\t"
		"mov.b64 {b0, b1}, B;\n\t"
		"mov.b64 {c0, c1}, C;\n\t"
		"add.cc.u32 h0, a1, b1;\n\t"
		"addc.u32 h1, c1, 0;\n\t"
		"add.cc.u32 h0, h0, c0;\n\t"
		"addc.u32 h1, h1, 0;\n\t"
		"mul.lo.u32 %0, %2, %4;\n\t"
		"mul.hi.u32 %1, %2, %4;\n\t"
		"mad.lo.cc.u32 %0, h0, %8, %0;\n\t"
		"madc.hi.u32 %1, h0, %8, %1;\n\t"
		"mad.lo.u32 %1, %2, %5, %1;\n\t"
		"mad.lo.u32 %1, %3, %4, %1;\n\t"
		"mad.lo.u32 %1, h0, %9, %1;\n\t"
		"mad.lo.u32 %1, h1, %8, %1;\n\t"
		"}" : "=&r"(t0), "=&r"(t1)
		: "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(p0), "r"(p1), "r"(n0), "r"(n1));
	return ((u64) t1 << 32) | t0;
}
struct ct_v27 { static const char *name() { return "CT v27 = v26 without the word-1 sum: [0,4q) product, csub(x,4q)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 fourq = 2 * c.twoq, nq = 0 - c.q;
		const u64 xr = csub_borrow(x, fourq);
		const u64 t = shoup4_ptx(y, c.w, c.wp, nq);
		x = xr + t; y = xr - t + fourq;
	} };
/* ---- V28: v19 with the four narrow products summed apart and the high words of
 * x' and y' formed by three-input adds (keeps them off the FMA pipe) ---- */
__device__ __forceinline__ void ct_split_ptx(u64 &x, u64 &y, u64 xr, u64 w, u64 wp, u64 nq, u64 threeq) {
	const u32 y0 = (u32) y, y1 = (u32) (y >> 32), w0 = (u32) w, w1 = (u32) (w >> 32);
	const u32 p0 = (u32) wp, p1 = (u32) (wp >> 32), n0 = (u32) nq, n1 = (u32) (nq >> 32);
	const u32 x0 = (u32) xr, x1 = (u32) (xr >> 32);
	const u32 m0 = (u32) threeq, m1 = (u32) (threeq >> 32);
	u32 o0, o1, z0, z1;
	asm("{\n\t"
		".reg .u32 r0, h0, h1, t0, t1, s, e0, e1;\n\t"
		"mul.lo.u32 r0, %4, %9;\n\t"
		"mul.hi.u32 h0, %4, %9;\n\t"
		"mad.lo.cc.u32 r0, %5, %8, r0;\n\t"
		"madc.hi.cc.u32 h0, %5, %8, h0;\n\t"
		"addc.u32 h1, 0, 0;\n\t"
		"mad.lo.cc.u32 h0, %5, %9, h0;\n\t"
		"madc.hi.u32 h1, %5, %9, h1;\n\t"
		"mul.lo.u32 t0, %4, %6;\n\t"
		"mul.hi.u32 t1, %4, %6;\n\t"
		"mad.lo.cc.u32 t0, h0, %10, t0;\n\t"
		"madc.hi.u32 t1, h0, %10, t1;\n\t"
		"mul.lo.u32 s, %4, %7;\n\t"            /* narrow products apart */
		"mad.lo.u32 s, %5, %6, s;\n\t"
		"mad.lo.u32 s, h0, %11, s;\n\t"
		"mad.lo.u32 s, h1, %10, s;\n\t"
		/* x' = xr + t + (s << 32) */
		"add.cc.u32 %0, %12, t0;\n\t"
		"addc.u32 %1, %13, t1;\n\t"
		"add.u32 %1, %1, s;\n\t"
		/* y' = xr + 3q - t - (s << 32) */
		"add.cc.u32 e0, %12, %14;\n\t"
		"addc.u32 e1, %13, %15;\n\t"
		"sub.cc.u32 %2, e0, t0;\n\t"
		"subc.u32 %3, e1, t1;\n\t"
		"sub.u32 %3, %3, s;\n\t"
		"}" : "=&r"(o0), "=&r"(o1), "=&r"(z0), "=&r"(z1)
		: "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(p0), "r"(p1), "r"(n0), "r"(n1),
		  "r"(x0), "r"(x1), "r"(m0), "r"(m1));
	x = ((u64) o1 << 32) | o0;
	y = ((u64) z1 << 32) | z0;
}
struct ct_v28 { static const char *name() { return "CT v28 = v19, narrow products apart, 3-input high-word adds"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 threeq = c.twoq + c.q, nq = 0 - c.q;
		const u64 xr = csub_borrow(x, threeq);
		ct_split_ptx(x, y, xr, c.w, c.wp, nq, threeq);
	} };

/* ---- GS variants ---- */
struct gs_v0 { static const char *name() { return "GS v0 harvey (as shipped)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 s = x + y, d = x - y + c.twoq;
		x = csub(s, c.twoq);
		y = d * c.w - __umul64hi(d, c.wp) * c.q;
	} };
struct gs_v1 { static const char *name() { return "GS v1 no csub"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 s = x + y, d = x - y + c.twoq;
		x = s;
		y = d * c.w - __umul64hi(d, c.wp) * c.q;
	} };
struct gs_v2 { static const char *name() { return "GS v2 approx mulhi, no csub"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 s = x + y, d = x - y + c.twoq;
		x = s;
		y = d * c.w - mulhi_approx(d, c.wp) * c.q;
	} };
struct gs_v6 { static const char *name() { return "GS v6 csub = umin"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u64 s = x + y, d = x - y + c.twoq;
		x = min(s, s - c.twoq);
		y = d * c.w - __umul64hi(d, c.wp) * c.q;
	} };

/* ---- pieces ---- */
struct only_shoup { static const char *name() { return "shoup_lazy only (1 per pair)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		y = y * c.w - __umul64hi(y, c.wp) * c.q; x ^= y;
	} };
struct only_shoup_approx { static const char *name() { return "shoup approx only (1 per pair)"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		y = y * c.w - mulhi_approx(y, c.wp) * c.q; x ^= y;
	} };
struct only_mulhi32 { static const char *name() { return "2x mul.hi.u32 only"; }
	__device__ __forceinline__ void operator()(u64 &x, u64 &y, const consts &c) const {
		const u32 r = __umulhi((u32) y, (u32) c.wp) + __umulhi((u32) (y >> 32), (u32) (c.wp >> 32));
		y = ((u64) r << 32) | (u32) x; x += y;
	} };

static int g_threads = 1024;

template <class B>
static void run(int sms, const consts &c) {
	u64 *sink, *cycles;
	CHECK(cudaMalloc(&sink, 8));
	CHECK(cudaMalloc(&cycles, sms * sizeof(u64)));
	for (int rep = 0; rep < 2; rep++) {
		bench<<<sms, g_threads, 128 * 48>>>(sink, cycles, c, B());
		CHECK(cudaDeviceSynchronize());
	}
	std::vector<u64> h(sms);
	CHECK(cudaMemcpy(h.data(), cycles, sms * sizeof(u64), cudaMemcpyDeviceToHost));
	double cyc = 0;
	for (int i = 0; i < sms; i++) cyc += (double) h[i];
	cyc /= sms;
	const double per_sm = (double) g_threads * PAIRS * ITERS;
	printf("%-52s %6.3f bfly/clk/SM  -> %7.1f Gbfly/s @1.965GHz x148\n", B::name(),
			per_sm / cyc, per_sm / cyc * 148 * 1.965);
	CHECK(cudaFree(sink));
	CHECK(cudaFree(cycles));
}

int main() {
	cudaDeviceProp prop;
	CHECK(cudaGetDeviceProperties(&prop, 0));
	const int sms = prop.multiProcessorCount;
	consts c;
	c.q = 1152921504606584833ull; c.twoq = 2 * c.q;
	c.w = 987813353222176621ull;
	c.wp = (u64) ((((unsigned __int128) c.w) << 64) / c.q);
	/* occupancy sweep: how many warps per SM does the butterfly need */
	for (int th = 128; th <= 1024; th += 128) {
		g_threads = th;
		printf("[%4d threads/SM] ", th);
		run<ct_v0>(sms, c);
	}
	g_threads = 1024;
	run<ct_v0>(sms, c); run<ct_v8>(sms, c); run<ct_v9>(sms, c); run<ct_v12>(sms, c); run<ct_v14>(sms, c); run<ct_v18>(sms, c); run<ct_v19>(sms, c); run<gs_v19>(sms, c); run<ct_v16>(sms, c); run<ct_v17>(sms, c); run<gs_v12>(sms, c); run<gs_v14>(sms, c); run<gs_v18>(sms, c); run<gs_v16>(sms, c); run<ct_v1>(sms, c); run<ct_v2>(sms, c); run<ct_v3>(sms, c);
	run<ct_v4>(sms, c); run<ct_v5>(sms, c); run<ct_v6>(sms, c); run<ct_v7>(sms, c);
	run<gs_v0>(sms, c); run<gs_v1>(sms, c); run<gs_v2>(sms, c); run<gs_v6>(sms, c);
	run<only_shoup>(sms, c); run<only_shoup_approx>(sms, c); run<only_mulhi32>(sms, c);
	run<ct_v26>(sms, c); run<ct_v27>(sms, c); run<ct_v28>(sms, c); run<ct_v20>(sms, c); run<ct_v24>(sms, c); run<ct_v25>(sms, c); run<gs_v24>(sms, c);
	run<only_wide_rz>(sms, c); run<only_wide_acc>(sms, c); run<only_hi_acc>(sms, c); run<only_lo>(sms, c);
	return 0;
}
