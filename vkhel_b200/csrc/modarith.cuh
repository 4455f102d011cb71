/*
 * Device-side 64-bit modular arithmetic for sm_100a.
 *
 * Everything here produces the same residues as the reference's GLSL
 * arithmetic (src/kernels/shaders/X.comp), which emulates 64x64->128 products
 * with four 32x32 multiplies (e.g. nttfwdbutterfly.comp:21-29); on sm_100a the
 * high half comes from mul.hi.u64 (IMAD.WIDE.U32 chains).  Because every
 * reference kernel stores the canonical residue in [0,q), any exact algorithm
 * followed by a canonical reduction is bit-identical to it.
 */
#ifndef VKHEL_MODARITH_CUH
#define VKHEL_MODARITH_CUH

#include <stdint.h>

typedef unsigned long long u64;

/* High 64 bits of a 64x64 product as an explicit mad/madc carry chain.  Same
 * value as __umul64hi; ptxas schedules this form with fewer carry fix-ups
 * (tools/bfly_bench.cu v14: +2.7 % butterfly rate over mul.hi.u64). */
__device__ __forceinline__ u64 mulhi64(u64 a, u64 b) {
	const unsigned a0 = (unsigned) a, a1 = (unsigned) (a >> 32);
	const unsigned b0 = (unsigned) b, b1 = (unsigned) (b >> 32);
	unsigned r1, r2;
	asm("{\n\t"
	    ".reg .u32 r0;\n\t"
	    "mul.hi.u32 r0, %2, %4;\n\t"           /* word 1 so far: hi(a0*b0) */
	    "mad.lo.cc.u32 r0, %2, %5, r0;\n\t"    /* + lo(a0*b1) */
	    "madc.hi.u32 %0, %2, %5, 0;\n\t"       /* word 2: hi(a0*b1) + carry */
	    "mad.lo.cc.u32 r0, %3, %4, r0;\n\t"    /* word 1 += lo(a1*b0) */
	    "madc.hi.cc.u32 %0, %3, %4, %0;\n\t"   /* word 2 += hi(a1*b0) + carry */
	    "addc.u32 %1, 0, 0;\n\t"               /* word 3: carry */
	    "mad.lo.cc.u32 %0, %3, %5, %0;\n\t"    /* word 2 += lo(a1*b1) */
	    "madc.hi.u32 %1, %3, %5, %1;\n\t"      /* word 3 += hi(a1*b1) + carry */
	    "}"
	    : "=&r"(r1), "=&r"(r2)
	    : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
	return ((u64) r2 << 32) | r1;
}

/* The whole product as one PTX chain: h = high 64 bits of y*wp (exact, or
 * without the y0*wp0 partial product when APPROX), then
 * t = lo64(y*w) + lo64(h*nq) with nq = 2^64 - q, i.e. y*w - h*q.  Written out
 * so that ptxas emits the ten (nine) multiplies as IMAD / IMAD.WIDE / IMAD.HI
 * with their accumulators chained and no separate add, negate or move
 * (tools/bfly_bench.cu v19: +7 % butterfly rate over the C expression). */
/* (Measured and dropped, profiles/r02_variants_a.txt: for moduli just below a
 * power of two the high word of 2^64 - q is 2^32 - 2^s, so h0*n1 is a shift and
 * a subtraction instead of a multiply -- one fmaheavy slot of 16 fewer per
 * butterfly, and the transform 5-7 % SLOWER: the funnel shift runs at 49.5
 * per clock per SM and three-input additions with two carries at a third of
 * the IADD3 rate, vkhel_ctx_probe_int_peaks.) */
template <bool APPROX>
__device__ __forceinline__ u64 shoup_chain(u64 y, u64 w, u64 wp, u64 nq) {
	const unsigned y0 = (unsigned) y, y1 = (unsigned) (y >> 32);
	const unsigned w0 = (unsigned) w, w1 = (unsigned) (w >> 32);
	const unsigned p0 = (unsigned) wp, p1 = (unsigned) (wp >> 32);
	const unsigned n0 = (unsigned) nq, n1 = (unsigned) (nq >> 32);
	unsigned t0, t1;
	if (APPROX) {
		asm("{\n\t"
		    ".reg .u32 r0, h0, h1;\n\t"
		    "mul.lo.u32 r0, %2, %7;\n\t"           /* word 1: lo(y0*p1) */
		    "mul.hi.u32 h0, %2, %7;\n\t"           /* word 2: hi(y0*p1) */
		    "mad.lo.cc.u32 r0, %3, %6, r0;\n\t"    /* word 1 += lo(y1*p0) */
		    "madc.hi.cc.u32 h0, %3, %6, h0;\n\t"   /* word 2 += hi(y1*p0) + c */
		    "addc.u32 h1, 0, 0;\n\t"
		    "mad.lo.cc.u32 h0, %3, %7, h0;\n\t"    /* word 2 += lo(y1*p1) */
		    "madc.hi.u32 h1, %3, %7, h1;\n\t"      /* word 3 += hi(y1*p1) + c */
		    "mul.lo.u32 %0, %2, %4;\n\t"           /* t = y0*w0 */
		    "mul.hi.u32 %1, %2, %4;\n\t"
		    "mad.lo.cc.u32 %0, h0, %8, %0;\n\t"    /* t += h0*n0 */
		    "madc.hi.u32 %1, h0, %8, %1;\n\t"
		    "mad.lo.u32 %1, %2, %5, %1;\n\t"       /* t.hi += y0*w1 */
		    "mad.lo.u32 %1, %3, %4, %1;\n\t"       /* t.hi += y1*w0 */
		    "mad.lo.u32 %1, h0, %9, %1;\n\t"       /* t.hi += h0*n1 */
		    "mad.lo.u32 %1, h1, %8, %1;\n\t"       /* t.hi += h1*n0 */
		    "}"
		    : "=&r"(t0), "=&r"(t1)
		    : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(p0), "r"(p1), "r"(n0),
		      "r"(n1));
	} else {
		asm("{\n\t"
		    ".reg .u32 r0, h0, h1;\n\t"
		    "mul.hi.u32 r0, %2, %6;\n\t"           /* word 1: hi(y0*p0) */
		    "mad.lo.cc.u32 r0, %2, %7, r0;\n\t"    /* + lo(y0*p1) */
		    "madc.hi.u32 h0, %2, %7, 0;\n\t"       /* word 2: hi(y0*p1) + c */
		    "mad.lo.cc.u32 r0, %3, %6, r0;\n\t"    /* word 1 += lo(y1*p0) */
		    "madc.hi.cc.u32 h0, %3, %6, h0;\n\t"   /* word 2 += hi(y1*p0) + c */
		    "addc.u32 h1, 0, 0;\n\t"
		    "mad.lo.cc.u32 h0, %3, %7, h0;\n\t"    /* word 2 += lo(y1*p1) */
		    "madc.hi.u32 h1, %3, %7, h1;\n\t"      /* word 3 += hi(y1*p1) + c */
		    "mul.lo.u32 %0, %2, %4;\n\t"
		    "mul.hi.u32 %1, %2, %4;\n\t"
		    "mad.lo.cc.u32 %0, h0, %8, %0;\n\t"
		    "madc.hi.u32 %1, h0, %8, %1;\n\t"
		    "mad.lo.u32 %1, %2, %5, %1;\n\t"
		    "mad.lo.u32 %1, %3, %4, %1;\n\t"
		    "mad.lo.u32 %1, h0, %9, %1;\n\t"
		    "mad.lo.u32 %1, h1, %8, %1;\n\t"
		    "}"
		    : "=&r"(t0), "=&r"(t1)
		    : "r"(y0), "r"(y1), "r"(w0), "r"(w1), "r"(p0), "r"(p1), "r"(n0),
		      "r"(n1));
	}
	return ((u64) t1 << 32) | t0;
}

/* ---- Shoup multiplication by a fixed factor --------------------------------
 * w < q, wp = floor(w * 2^64 / q).  For ANY 64-bit y the lazy result is
 * y*w mod q + {0,q}, i.e. in [0,2q) (reference: nttfwdbutterfly.comp:44-48,
 * elemmulconst.comp:42-46, which then subtract q once). Needs q < 2^63. */
__device__ __forceinline__ u64 shoup_lazy(u64 y, u64 w, u64 wp, u64 q) {
	return shoup_chain<false>(y, w, wp, 0 - q);
}

/* The same chain without the a0*b0 partial product: the word-1 column loses a
 * summand below 2^32, so its carry into the result drops by at most one and
 * the value is mulhi64(a, b) or one less.  Saves one of the six wide
 * multiplies of a Shoup product (tools/bfly_bench.cu v18: +11 % forward,
 * +14 % inverse butterfly rate). */
__device__ __forceinline__ u64 mulhi64_approx(u64 a, u64 b) {
	const unsigned a0 = (unsigned) a, a1 = (unsigned) (a >> 32);
	const unsigned b0 = (unsigned) b, b1 = (unsigned) (b >> 32);
	unsigned r1, r2;
	asm("{\n\t"
	    ".reg .u32 r0;\n\t"
	    "mul.lo.u32 r0, %2, %5;\n\t"           /* word 1: lo(a0*b1) */
	    "mul.hi.u32 %0, %2, %5;\n\t"           /* word 2: hi(a0*b1) */
	    "mad.lo.cc.u32 r0, %3, %4, r0;\n\t"    /* word 1 += lo(a1*b0) */
	    "madc.hi.cc.u32 %0, %3, %4, %0;\n\t"   /* word 2 += hi(a1*b0) + carry */
	    "addc.u32 %1, 0, 0;\n\t"               /* word 3: carry */
	    "mad.lo.cc.u32 %0, %3, %5, %0;\n\t"    /* word 2 += lo(a1*b1) */
	    "madc.hi.u32 %1, %3, %5, %1;\n\t"      /* word 3 += hi(a1*b1) + carry */
	    "}"
	    : "=&r"(r1), "=&r"(r2)
	    : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
	return ((u64) r2 << 32) | r1;
}

/* Shoup product with the approximate quotient: the quotient estimate is at
 * most one further below the true one, so the result is y*w mod q plus
 * {0, q, 2q}: in [0,3q) for ANY 64-bit y.  Needs 3q < 2^64. */
__device__ __forceinline__ u64 shoup_lazy3(u64 y, u64 w, u64 wp, u64 q) {
	return shoup_chain<true>(y, w, wp, 0 - q);
}

__device__ __forceinline__ u64 shoup_canon(u64 y, u64 w, u64 wp, u64 q) {
	const u64 r = shoup_lazy(y, w, wp, q);
	return r >= q ? r - q : r;
}

/* x - (x >= m ? m : 0).
 * Written as a 64-bit subtraction whose borrow selects the result: ptxas turns
 * the C expression `x >= m ? x - m : x` into 2 ISETP + 2 SEL + 2 subtract
 * instructions, this form into 2 subtracts + 1 borrow word + 2 SEL
 * (tools/bfly_bench.cu v12: +4 % forward, +7 % inverse butterfly rate). */
__device__ __forceinline__ u64 csub(u64 x, u64 m) {
	const unsigned xl = (unsigned) x, xh = (unsigned) (x >> 32);
	const unsigned ml = (unsigned) m, mh = (unsigned) (m >> 32);
	unsigned tl, th, borrow;
	asm("sub.cc.u32 %0, %3, %5;\n\t"
	    "subc.cc.u32 %1, %4, %6;\n\t"
	    "subc.u32 %2, 0, 0;"
	    : "=r"(tl), "=r"(th), "=r"(borrow)
	    : "r"(xl), "r"(xh), "r"(ml), "r"(mh));
	const unsigned rl = borrow ? xl : tl, rh = borrow ? xh : th;
	return ((u64) rh << 32) | rl;
}

/* ---- Harvey butterflies, lazy ranges (q < 2^62) ------------------------------
 * forward (Cooley-Tukey, reference nttfwdbutterfly.comp:41-57):
 *   in : x, y in [0,4q)      out: x' = x + w*y, y' = x - w*y, both in [0,4q)
 * inverse (Gentleman-Sande, reference nttrevbutterfly.comp:41-57):
 *   in : x, y in [0,2q)      out: x' = x + y, y' = (x - y)*w, both in [0,2q)
 */
/* `zr` is a zero the compiler cannot see through (a kernel parameter).  The
 * two-input sums x' = xr + t and s = x + y are written as three-input sums with
 * it: ptxas then keeps both words of the addition on the ALU pipe (IADD3 /
 * IADD3.X with three addends), where it otherwise moves the high word of about
 * two in three such additions to IMAD.X "to balance the pipes" -- 1.1 issue
 * slots each on the fmaheavy pipe, the one that binds these kernels
 * (profiles/r02_sass_census.txt). */
__device__ __forceinline__ void ct_lazy(u64 &x, u64 &y, u64 w, u64 wp,
		u64 q, u64 twoq, u64 zr = 0) {
	const u64 xr = csub(x, twoq);
	const u64 t = shoup_lazy(y, w, wp, q);
	x = xr + t + zr;
	y = xr - t + twoq;
}

__device__ __forceinline__ void gs_lazy(u64 &x, u64 &y, u64 w, u64 wp,
		u64 q, u64 twoq, u64 zr = 0) {
	const u64 s = x + y + zr;
	const u64 d = x - y + twoq;
	x = csub(s, twoq);
	y = shoup_lazy(d, w, wp, q);
}

/* ---- the same butterflies around shoup_lazy3 (6q < 2^64, i.e. q < 2^61.4) -------
 * forward: x, y in [0,6q) -> [0,6q); inverse: x, y in [0,3q) -> [0,3q).
 * threeq = 3q.  One conditional subtraction per butterfly, as above. */
__device__ __forceinline__ void ct_lazy3(u64 &x, u64 &y, u64 w, u64 wp,
		u64 q, u64 threeq, u64 zr = 0) {
	const u64 xr = csub(x, threeq);
	const u64 t = shoup_lazy3(y, w, wp, q);
	x = xr + t + zr;
	y = xr - t + threeq;
}

__device__ __forceinline__ void gs_lazy3(u64 &x, u64 &y, u64 w, u64 wp,
		u64 q, u64 threeq, u64 zr = 0) {
	const u64 s = x + y + zr;
	const u64 d = x - y + threeq;
	x = csub(s, threeq);
	y = shoup_lazy3(d, w, wp, q);
}

/* ---- strict butterflies: every value canonical (2^62 <= q < 2^63) ---------- */
__device__ __forceinline__ void ct_strict(u64 &x, u64 &y, u64 w, u64 wp,
		u64 q) {
	const u64 t = shoup_canon(y, w, wp, q);
	const u64 s = x + t;           /* < 2q < 2^64 */
	const u64 d = x - t;
	y = x < t ? d + q : d;
	x = csub(s, q);
}

__device__ __forceinline__ void gs_strict(u64 &x, u64 &y, u64 w, u64 wp,
		u64 q) {
	const u64 s = x + y;
	const u64 d = x - y;
	const u64 dd = x < y ? d + q : d;
	x = csub(s, q);
	y = shoup_canon(dd, w, wp, q);
}

/* ---- general reduction for the element-wise kernels -------------------------
 * Exact for every modulus 2 <= q < 2^64 and every input, which covers the
 * reference tests' tiny moduli (2, 3, 5, 10, 17: test/vector.c:167-241,
 * examples/example.c:47) where the reference's own Barrett shift is
 * ill-defined.
 *
 * struct modulus is built on the host (make_modulus in vector.cu):
 *   mu   = floor(2^64 / q)                       one-word reciprocal
 *   s    = clz(q), d = q << s                    normalised divisor
 *   v    = floor((2^128 - 1) / d) - 2^64         Moller-Granlund reciprocal
 */
struct modulus {
	u64 q;
	u64 mu;
	u64 d;
	u64 v;
	unsigned s;
};

/* x mod q, any 64-bit x */
__device__ __forceinline__ u64 reduce64(u64 x, const modulus &m) {
	/* floor(x*mu/2^64) is floor(x/q) or one less */
	const u64 r = x - __umul64hi(x, m.mu) * m.q;
	return r >= m.q ? r - m.q : r;
}

/* (hi*2^64 + lo) mod q for hi < q: division of a two-word numerator by a
 * normalised one-word divisor (Moller & Granlund 2011, Alg. 4) */
__device__ __forceinline__ u64 reduce128(u64 hi, u64 lo, const modulus &m) {
	const u64 u1 = m.s ? (hi << m.s) | (lo >> (64 - m.s)) : hi;
	const u64 u0 = lo << m.s;
	/* (q1,q0) = v*u1 + (u1,u0) */
	u64 q0 = m.v * u1;
	u64 q1 = __umul64hi(m.v, u1);
	q0 += u0;
	q1 += u1 + (q0 < u0) + 1;
	u64 r = u0 - q1 * m.d;
	if (r > q0) {
		r += m.d;
	}
	if (r >= m.d) {
		r -= m.d;
	}
	return r >> m.s;
}

/* a*b mod q for a, b < q */
__device__ __forceinline__ u64 mulmod(u64 a, u64 b, const modulus &m) {
	return reduce128(__umul64hi(a, b), a * b, m);
}

#endif
