/*
 * Element-wise kernels: the sm_100a replacements of the reference's
 * elemmul / elemfma / elemmulconst / elemgtadd / elemgtsub / elemmodbytwo
 * compute shaders (src/kernels/shaders/X.comp) and of their record functions
 * (src/kernels/X.c), which computed the Barrett constants per call.
 *
 * All of them are HBM-bound streaming kernels (24 or 16 bytes per element,
 * a few dozen integer instructions): 128-bit coalesced loads and stores, four
 * vectors in flight per thread, one CTA per contiguous chunk.
 * Each output is the canonical residue, so results are bit-identical to the
 * shaders wherever those are defined.
 */
#include "common.cuh"

typedef unsigned __int128 u128;

struct modulus make_modulus(uint64_t q) {
	VK_REQUIRE(q >= 2, "modulus must be at least 2");
	struct modulus m;
	m.q = q;
	m.mu = (u64) ((((u128) 1) << 64) / q);
	m.s = (unsigned) __builtin_clzll(q);
	m.d = q << m.s;
	m.v = (u64) ((~(u128) 0) / m.d - (((u128) 1) << 64));
	return m;
}

/* ---- per-element operations ------------------------------------------------------ */
/* (a mod q)(b mod q) mod q for arbitrary 64-bit a, b (reference
 * elemmul.comp:62-73 reduces both operands first).  Here the full product is
 * formed once: a*b = hi*2^64 + lo = (hi mod q)*2^64 + lo (mod q), and the
 * two-word division needs hi < q -- true whenever the operands are canonical,
 * so the extra reduction of hi sits behind a branch that canonical data never
 * takes. */
__device__ __forceinline__ u64 mulmod_any(u64 a, u64 b, const modulus &m) {
	u64 hi = __umul64hi(a, b);
	const u64 lo = a * b;
	if (hi >= m.q) {
		hi = reduce64(hi, m);
	}
	return reduce128(hi, lo, m);
}

struct op_mul {
	modulus m;
	__device__ __forceinline__ u64 operator()(u64 a, u64 b) const {
		return mulmod_any(a, b, m);
	}
};

struct op_fma { /* contract of elemfma (shader defect Q2) */
	modulus m;
	u64 mult; /* reduced on the host: mult <= q - 1 */
	__device__ __forceinline__ u64 operator()(u64 a, u64 b) const {
		/* a*mult + b for arbitrary a, b: the high word of a*mult is at most
		 * mult - 1 <= q - 2, the carry of the addition raises it to at most
		 * q - 1, so the two-word division applies directly */
		const u64 lo = a * mult;
		const u64 sum = lo + b;
		const u64 hi = __umul64hi(a, mult) + (sum < lo);
		return reduce128(hi, sum, m);
	}
};

struct op_mulconst { /* reference elemmulconst.comp:35-49 */
	u64 q, b, bp;
	__device__ __forceinline__ u64 operator()(u64 a, u64) const {
		return shoup_canon(a, b, bp, q);
	}
};

struct op_mulconst_wide { /* same for q >= 2^63 where Shoup's range breaks */
	modulus m;
	u64 b;
	__device__ __forceinline__ u64 operator()(u64 a, u64) const {
		return mulmod(reduce64(a, m), b, m);
	}
};

struct op_gtadd { /* reference elemgtadd.comp:20-30 */
	u64 bound, diff;
	__device__ __forceinline__ u64 operator()(u64 a, u64) const {
		return a > bound ? a + diff : a;
	}
};

struct op_gtsub { /* reference elemgtsub.comp:49-67 */
	modulus m;
	u64 bound, diff; /* diff already reduced */
	__device__ __forceinline__ u64 operator()(u64 a, u64) const {
		const u64 r = reduce64(a, m);
		if (a > bound) { /* the compare uses the unreduced input */
			return r >= diff ? r - diff : r + (m.q - diff);
		}
		return r;
	}
};

struct op_modbytwo { /* reference elemmodbytwo.comp:19-27 */
	u64 signed_bound;
	__device__ __forceinline__ u64 operator()(u64 a, u64) const {
		const u64 u = a & 1;
		return a > signed_bound ? 1 - u : u;
	}
};

/* ---- streaming skeleton -------------------------------------------------------------
 * result may alias an operand (in-place ops), hence no __restrict__.
 * Vector body: a CTA takes one contiguous chunk of ELEM_THREADS * elem_unroll
 * 16-byte vectors, each thread elem_unroll of them with all loads issued
 * before the first use, and exits.  No grid-stride loop: with as many CTAs as
 * chunks the hardware block scheduler balances the SMs (the two dies do not
 * stream at the same rate); a grid of one resident wave that loops was
 * measured 4-20 % slower at 2^27 elements (elemfma 6104 against 6722 GB/s,
 * elemgtsub 5482 against 6843).  `len` elements; pointers 16 B aligned (the
 * scalar kernel below covers odd tails and unaligned sub-ranges). */
#define ELEM_THREADS 256
/* vectors per thread: 4, except for the full product (64 registers at 4, i.e.
 * half occupancy; measured 7094 GB/s with 2 against 6652 with 4, while
 * elemfma loses with 2: 6521 against 7031) */
template <class Op> struct elem_unroll { static constexpr int value = 4; };
template <> struct elem_unroll<op_mul> { static constexpr int value = 2; };

template <bool TWO_INPUTS, class Op>
__global__ void __launch_bounds__(ELEM_THREADS)
elem_vec_kernel(const ulonglong2 *a,
		const ulonglong2 *b, ulonglong2 *out,
		u64 nvec, const Op op) {
	constexpr int ELEM_UNROLL = elem_unroll<Op>::value;
	const u64 i = (u64) blockIdx.x * (ELEM_THREADS * ELEM_UNROLL) + threadIdx.x;
	if (i + (ELEM_UNROLL - 1) * ELEM_THREADS < nvec) {
		ulonglong2 va[ELEM_UNROLL], vb[ELEM_UNROLL];
#pragma unroll
		for (int u = 0; u < ELEM_UNROLL; u++) {
			va[u] = a[i + u * ELEM_THREADS];
			if (TWO_INPUTS) {
				vb[u] = b[i + u * ELEM_THREADS];
			} else {
				vb[u] = make_ulonglong2(0, 0);
			}
		}
#pragma unroll
		for (int u = 0; u < ELEM_UNROLL; u++) {
			out[i + u * ELEM_THREADS] = make_ulonglong2(op(va[u].x, vb[u].x),
					op(va[u].y, vb[u].y));
		}
		return;
	}
	/* the last, partial chunk */
	for (u64 j = i; j < nvec; j += ELEM_THREADS) {
		const ulonglong2 va = a[j];
		const ulonglong2 vb = TWO_INPUTS ? b[j] : make_ulonglong2(0, 0);
		out[j] = make_ulonglong2(op(va.x, vb.x), op(va.y, vb.y));
	}
}

template <bool TWO_INPUTS, class Op>
__global__ void __launch_bounds__(ELEM_THREADS)
elem_scalar_kernel(const u64 *a, const u64 *b, u64 *out, u64 len,
		const Op op) {
	const u64 i = (u64) blockIdx.x * ELEM_THREADS + threadIdx.x;
	if (i < len) {
		out[i] = op(a[i], TWO_INPUTS ? b[i] : 0);
	}
}

/* one CTA per chunk of `per_block` work items */
static unsigned grid_for(u64 work_items, unsigned per_block) {
	const u64 blocks = (work_items + per_block - 1) / per_block;
	VK_REQUIRE(blocks <= 0x7fffffffull, "vector too long for one launch");
	return (unsigned) (blocks ? blocks : 1);
}

template <bool TWO_INPUTS, class Op>
static void launch_elem(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *out, uint64_t len, const Op &op) {
	if (len == 0) {
		return;
	}
	cudaStream_t stream = ctx_stream(ctx);
	const uintptr_t align = (uintptr_t) a | (uintptr_t) out
		| (TWO_INPUTS ? (uintptr_t) b : 0);
	uint64_t done = 0;
	if ((align & 15) == 0 && len >= 2) {
		const u64 nvec = len / 2;
		elem_vec_kernel<TWO_INPUTS, Op>
			<<<grid_for(nvec, ELEM_THREADS * elem_unroll<Op>::value),
				ELEM_THREADS, 0, stream>>>(
				(const ulonglong2 *) a, (const ulonglong2 *) b,
				(ulonglong2 *) out, nvec, op);
		CUDA_CHECK(cudaGetLastError());
		ctx->dev.launches++;
		done = nvec * 2;
	}
	if (done < len) {
		const u64 rest = len - done;
		elem_scalar_kernel<TWO_INPUTS, Op>
			<<<grid_for(rest, ELEM_THREADS), ELEM_THREADS, 0, stream>>>(
				a + done, TWO_INPUTS ? b + done : NULL, out + done, rest, op);
		CUDA_CHECK(cudaGetLastError());
		ctx->dev.launches++;
	}
}

void launch_elemmul(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *out, uint64_t len, uint64_t q) {
	op_mul op = { make_modulus(q) };
	launch_elem<true>(ctx, a, b, out, len, op);
}

void launch_elemfma(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *out, uint64_t len, uint64_t mult, uint64_t q) {
	op_fma op = { make_modulus(q), mult % q };
	launch_elem<true>(ctx, a, b, out, len, op);
}

void launch_elemmulconst(struct vkhel_ctx *ctx, const u64 *in, u64 *out,
		uint64_t len, uint64_t b, uint64_t q) {
	if (q >> 63) {
		op_mulconst_wide op = { make_modulus(q), b % q };
		launch_elem<false>(ctx, in, NULL, out, len, op);
		return;
	}
	b %= q;
	op_mulconst op = { q, b, nt_compute_barrett_factor(b, q, 64) };
	launch_elem<false>(ctx, in, NULL, out, len, op);
}

void launch_elemgtadd(struct vkhel_ctx *ctx, const u64 *in, u64 *out,
		uint64_t len, uint64_t bound, uint64_t diff) {
	op_gtadd op = { bound, diff };
	launch_elem<false>(ctx, in, NULL, out, len, op);
}

void launch_elemgtsub(struct vkhel_ctx *ctx, const u64 *in, u64 *out,
		uint64_t len, uint64_t bound, uint64_t diff, uint64_t q) {
	op_gtsub op = { make_modulus(q), bound, diff % q };
	launch_elem<false>(ctx, in, NULL, out, len, op);
}

void launch_elemmodbytwo(struct vkhel_ctx *ctx, const u64 *in, u64 *out,
		uint64_t len, uint64_t signed_bound) {
	op_modbytwo op = { signed_bound };
	launch_elem<false>(ctx, in, NULL, out, len, op);
}

/* ---- RNS element-wise product: [batch][limbs][n], one modulus per limb ------------- */
#define RNS_MAX_LIMBS 64

struct rns_moduli {
	modulus m[RNS_MAX_LIMBS];
};

__global__ void __launch_bounds__(ELEM_THREADS)
elemmul_rns_kernel(const ulonglong2 *a,
		const ulonglong2 *b, ulonglong2 *out,
		u64 nvec, unsigned log2_vec_per_poly, unsigned limbs,
		const __grid_constant__ rns_moduli mods) {
	/* one chunk of 2 * ELEM_THREADS vectors per CTA, as above */
	const u64 i = (u64) blockIdx.x * (2 * ELEM_THREADS) + threadIdx.x;
	const u64 j = i + ELEM_THREADS;
	if (i >= nvec) {
		return;
	}
	const bool second = j < nvec;
	const ulonglong2 a0 = a[i], b0 = b[i];
	ulonglong2 a1 = make_ulonglong2(0, 0), b1 = a1;
	if (second) {
		a1 = a[j];
		b1 = b[j];
	}
	const modulus &m0 = mods.m[(i >> log2_vec_per_poly) % limbs];
	out[i] = make_ulonglong2(mulmod_any(a0.x, b0.x, m0),
			mulmod_any(a0.y, b0.y, m0));
	if (second) {
		const modulus &m1 = mods.m[(j >> log2_vec_per_poly) % limbs];
		out[j] = make_ulonglong2(mulmod_any(a1.x, b1.x, m1),
				mulmod_any(a1.y, b1.y, m1));
	}
}

void launch_elemmul_rns(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *out, const uint64_t *mods, uint64_t limbs, uint64_t n,
		uint64_t batch) {
	VK_REQUIRE(limbs >= 1 && limbs <= RNS_MAX_LIMBS,
			"elemmul_rns: 1..%d limbs", RNS_MAX_LIMBS);
	VK_REQUIRE(n >= 2 && (n & (n - 1)) == 0,
			"elemmul_rns: n must be a power of two >= 2");
	if (batch == 0) {
		return;
	}
	rns_moduli params;
	for (uint64_t l = 0; l < limbs; l++) {
		params.m[l] = make_modulus(mods[l]);
	}
	for (uint64_t l = limbs; l < RNS_MAX_LIMBS; l++) {
		params.m[l] = params.m[0];
	}
	const u64 nvec = limbs * n * batch / 2;
	const unsigned log2_vec_per_poly = (unsigned) nt_ceil_log2(n) - 2;
	elemmul_rns_kernel<<<grid_for(nvec, ELEM_THREADS * 2), ELEM_THREADS, 0,
		ctx_stream(ctx)>>>((const ulonglong2 *) a, (const ulonglong2 *) b,
				(ulonglong2 *) out, nvec, log2_vec_per_poly, (unsigned) limbs,
				params);
	CUDA_CHECK(cudaGetLastError());
	ctx->dev.launches++;
}
