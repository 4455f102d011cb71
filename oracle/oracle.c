/*
 * CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the reference's algorithm for the hot path, used
 * only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker and as the timed CPU baseline.  The
 * library (vkhel_b200/) never links, loads or calls anything in this
 * directory.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against the golden vectors of the reference's own tests (test/vector.c,
 * test/ntt.c, test/numbers.c, copied as data into tests/golden/) and, when
 * oracle/_ref is built, against the reference's own numbers.c/ntt_tables.c
 * compiled from /root/reference.  The transform and the element-wise
 * arithmetic exist in the reference only as GLSL shaders (no Vulkan toolchain
 * in this image), so for those the golden vectors are the anchor.
 *
 * Each function cites the reference lines it follows.  "Literal" functions
 * reproduce the shader arithmetic step by step (including the four-way 32-bit
 * split of the 64x64 product); "canonical" functions state the mathematical
 * contract with 128-bit integers.  tests assert literal == canonical wherever
 * the shader arithmetic is well defined.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---- 64x64 -> 128 product the way every shader does it ----------------------------
 * reference: mul64() in nttfwdbutterfly.comp:21-29, elemmul.comp:24-34:
 * four 32x32 partial products; the middle column's carries are collected in
 * `cross` before being folded into the high word. */
static void shader_mul64(uint64_t a, uint64_t b, uint64_t *hi, uint64_t *lo) {
	const uint64_t a_lo = a & 0xffffffffu, a_hi = a >> 32;
	const uint64_t b_lo = b & 0xffffffffu, b_hi = b >> 32;
	const uint64_t p_ll = a_lo * b_lo;
	const uint64_t p_hl = a_hi * b_lo;
	const uint64_t p_lh = a_lo * b_hi;
	const uint64_t p_hh = a_hi * b_hi;
	const uint64_t cross = (p_ll >> 32) + (p_hl & 0xffffffffu) + p_lh;
	*hi = (p_hl >> 32) + (cross >> 32) + p_hh;
	*lo = (cross << 32) | (p_ll & 0xffffffffu);
}

static uint64_t shader_mulhi(uint64_t a, uint64_t b) {
	uint64_t hi, lo;
	shader_mul64(a, b, &hi, &lo);
	return hi;
}

static unsigned bit_length(uint64_t v) { /* nt_ceil_log2, numbers.h:11-13 */
	return 64 - (unsigned) __builtin_clzll(v);
}

/* ---- host number theory (reference src/numbers.c) ------------------------------- */
/* floor(f * 2^64 / q): numbers.c:30-34 */
uint64_t oracle_shoup_factor(uint64_t f, uint64_t q) {
	return (uint64_t) (((u128) f << 64) / q);
}

/* numbers.c:5-28,36-40: Barrett with mu = floor(2^(bits+62) / q) recomputed
 * per call; the barrett_factor argument is ignored there as well */
uint64_t oracle_multiply_mod(uint64_t a, uint64_t b, uint64_t q) {
	const unsigned bits = bit_length(q);
	const u128 prod = (u128) a * b;
	const uint64_t hi = (uint64_t) (prod >> 64), lo = (uint64_t) prod;
	const uint64_t mu = (uint64_t) ((((u128) 1 << (bits + 62 - 64)) << 64) / q);
	const unsigned shift = bits - 2;
	const uint64_t top = (shift == 0) ? 0 : (hi << (64 - shift));
	const uint64_t num_c = top + (lo >> shift);
	const uint64_t q_hat = (uint64_t) (((u128) num_c * mu) >> 64);
	uint64_t z = lo - q_hat * q;
	if (z >= q) {
		z -= q;
	}
	return z;
}

/* numbers.c:42-59 */
uint64_t oracle_power_mod(uint64_t base, uint64_t exp, uint64_t q) {
	uint64_t result = 1;
	base %= q;
	while (exp > 0) {
		if (exp & 1) {
			result = oracle_multiply_mod(result, base, q);
		}
		base = oracle_multiply_mod(base, base, q);
		exp >>= 1;
	}
	return result;
}

/* numbers.c:71-99: extended Euclid in signed 64-bit */
uint64_t oracle_inverse_mod(uint64_t input, uint64_t q) {
	if (q == 1) {
		return 0;
	}
	int64_t a = (int64_t) input, b = (int64_t) q;
	int64_t y = 0, x = 1;
	while (a > 1) {
		const int64_t quot = a / b, rem = a % b;
		a = b;
		b = rem;
		const int64_t prev_y = y;
		y = x - quot * y;
		x = prev_y;
	}
	if (x < 0) {
		x += (int64_t) q;
	}
	return (uint64_t) x;
}

/* ---- tables (reference src/ntt_tables.c:8-44) ------------------------------------- */
static uint64_t reverse_bits(uint64_t v, unsigned width) {
	uint64_t r = 0;
	for (unsigned i = 0; i < width; i++) {
		if (v & ((uint64_t) 1 << i)) {
			r |= (uint64_t) 1 << (width - 1 - i);
		}
	}
	return r;
}

void oracle_tables(uint64_t n, uint64_t q, uint64_t w,
		uint64_t *roots, uint64_t *inv_roots,
		uint64_t *roots_shoup, uint64_t *inv_roots_shoup) {
	const unsigned width = bit_length(n) - 1;
	roots[0] = 1;
	inv_roots[0] = oracle_inverse_mod(1, q);
	uint64_t prev = 1;
	for (uint64_t i = 1; i < n; i++) {
		const uint64_t idx = reverse_bits(i, width);
		roots[idx] = oracle_multiply_mod(prev, w, q);      /* :27-28 */
		inv_roots[idx] = oracle_inverse_mod(roots[idx], q); /* :29-30 */
		prev = roots[idx];
	}
	for (uint64_t i = 0; i < n; i++) {                      /* :35-43 */
		roots_shoup[i] = oracle_shoup_factor(roots[i], q);
		inv_roots_shoup[i] = oracle_shoup_factor(inv_roots[i], q);
	}
}

/* ---- butterflies (shaders) ---------------------------------------------------------- */
/* Shoup product reduced to [0,q): nttfwdbutterfly.comp:44-48,
 * nttrevbutterfly.comp:50-54, elemmulconst.comp:42-46 */
static uint64_t shader_shoup(uint64_t y, uint64_t w, uint64_t w_shoup,
		uint64_t q) {
	uint64_t r = y * w - shader_mulhi(y, w_shoup) * q;
	if (r >= q) {
		r -= q;
	}
	return r;
}

/* (x - y) mod q the shader way: add q until x >= y (nttfwdbutterfly.comp:50-54) */
static uint64_t shader_sub_mod(uint64_t x, uint64_t y, uint64_t q) {
	while (x < y) {
		x += q;
	}
	return (x - y) % q;
}

/* forward transform: stage loop src/vector.c:536-566, butterfly
 * nttfwdbutterfly.comp:41-57.  The first stage reads `in`, every later stage
 * works in place on `out` (vector.c:565 `input = result`). */
void oracle_forward(const uint64_t *in, uint64_t *out, uint64_t n, uint64_t q,
		const uint64_t *roots, const uint64_t *roots_shoup) {
	const uint64_t *src = in;
	uint64_t t = n / 2;
	for (uint64_t m = 1; m < n; m *= 2) {
		uint64_t offset = 0;
		for (uint64_t i = 0; i < m; i++) {
			const uint64_t w = roots[m + i], ws = roots_shoup[m + i];
			for (uint64_t pos = 0; pos < t; pos++) {
				const uint64_t X = src[offset + pos];
				const uint64_t Y = src[offset + pos + t];
				const uint64_t WY = shader_shoup(Y, w, ws, q);
				out[offset + pos] = (X + WY) % q;
				out[offset + pos + t] = shader_sub_mod(X, WY, q);
			}
			offset += 2 * t;
		}
		t /= 2;
		src = out;
	}
}

/* inverse transform: stage loop src/vector.c:599-630, butterfly
 * nttrevbutterfly.comp:41-57, then every element of the result vector
 * (out_len of them, not n: vector.c:633-639 + elemmulconst.c:157-169) is
 * multiplied by n^-1 with elemmulconst.comp:42-48. */
void oracle_inverse(const uint64_t *in, uint64_t *out, uint64_t n,
		uint64_t out_len, uint64_t q,
		const uint64_t *inv_roots, const uint64_t *inv_roots_shoup) {
	const uint64_t *src = in;
	uint64_t t = 1;
	for (uint64_t m = n / 2; m >= 1; m /= 2) {
		uint64_t offset = 0;
		for (uint64_t i = 0; i < m; i++) {
			const uint64_t w = inv_roots[m + i], ws = inv_roots_shoup[m + i];
			for (uint64_t pos = 0; pos < t; pos++) {
				const uint64_t X = src[offset + pos];
				const uint64_t Y = src[offset + pos + t];
				const uint64_t diff = shader_sub_mod(X, Y, q);
				/* the shader multiplies the un-reduced x-y+kq; reducing it
				 * first does not change the canonical Shoup result */
				out[offset + pos] = (X + Y) % q;
				out[offset + pos + t] = shader_shoup(diff, w, ws, q);
			}
			offset += 2 * t;
		}
		t *= 2;
		src = out;
	}
	const uint64_t inv_n = oracle_inverse_mod(n % q, q);
	const uint64_t inv_n_shoup = oracle_shoup_factor(inv_n, q);
	for (uint64_t k = 0; k < out_len; k++) {
		out[k] = shader_shoup(out[k], inv_n, inv_n_shoup, q);
	}
}

/* ---- element-wise shaders ------------------------------------------------------------- */
/* elemmul.comp:36-47 / elemgtsub.comp:38-49 with the host constants of
 * elemmul.c:159-168: mu = floor(2^(bits+62)/q), shift bits-2 */
static uint64_t shader_reduce64(uint64_t lo, uint64_t q, uint64_t mu,
		unsigned bits) {
	const uint64_t num_c = lo >> (bits - 2);
	uint64_t z = lo - shader_mulhi(num_c, mu) * q;
	if (z >= q) {
		z -= q;
	}
	return z;
}

/* elemmul.comp:49-60 */
static uint64_t shader_reduce128(uint64_t hi, uint64_t lo, uint64_t q,
		uint64_t mu, unsigned bits) {
	const uint64_t num_c = (hi << (64 - (bits - 2))) + (lo >> (bits - 2));
	uint64_t z = lo - shader_mulhi(num_c, mu) * q;
	if (z >= q) {
		z -= q;
	}
	return z;
}

static uint64_t barrett_mu(uint64_t q) { /* elemmul.c:159-162 */
	const unsigned bits = bit_length(q);
	return (uint64_t) ((((u128) 1 << (bits + 62 - 64)) << 64) / q);
}

/* 1 when the shader's Barrett arithmetic is well defined for q: the shifts
 * by bits-2 and 66-bits must lie in [0,63] and q must stay clear of 2^63
 *  */
int oracle_barrett_defined(uint64_t q) {
	const unsigned bits = bit_length(q);
	return bits >= 3 && bits <= 62;
}

/* elemmul.comp:62-73, literal */
void oracle_elemmul_literal(const uint64_t *a, const uint64_t *b,
		uint64_t *out, uint64_t len, uint64_t q) {
	const unsigned bits = bit_length(q);
	const uint64_t mu = barrett_mu(q);
	for (uint64_t k = 0; k < len; k++) {
		uint64_t hi, lo;
		shader_mul64(shader_reduce64(a[k], q, mu, bits),
				shader_reduce64(b[k], q, mu, bits), &hi, &lo);
		out[k] = shader_reduce128(hi, lo, q, mu, bits);
	}
}

/* contract of elemmul: (a mod q)(b mod q) mod q */
void oracle_elemmul(const uint64_t *a, const uint64_t *b, uint64_t *out,
		uint64_t len, uint64_t q) {
	for (uint64_t k = 0; k < len; k++) {
		out[k] = (uint64_t) (((u128) (a[k] % q) * (b[k] % q)) % q);
	}
}

/* elemfma.comp:36-55 literal, INCLUDING its sign defects (`mod - product`,
 * `mod - sum`) -- kept only to document where the shader
 * and the contract part ways; the library implements the contract below. */
void oracle_elemfma_literal(const uint64_t *a, const uint64_t *b,
		uint64_t *out, uint64_t len, uint64_t mult, uint64_t q) {
	const uint64_t mult_shoup = oracle_shoup_factor(mult, q); /* elemfma.c:159-161 */
	for (uint64_t k = 0; k < len; k++) {
		uint64_t product = a[k] * mult - shader_mulhi(a[k], mult_shoup) * q;
		if (product >= q) {
			product = q - product;
		}
		const uint64_t sum = product + b[k];
		out[k] = sum >= q ? q - sum : sum;
	}
}

/* contract of elemfma pinned by test/vector.c:39-83: (a*mult + b) mod q */
void oracle_elemfma(const uint64_t *a, const uint64_t *b, uint64_t *out,
		uint64_t len, uint64_t mult, uint64_t q) {
	for (uint64_t k = 0; k < len; k++) {
		const u128 v = (u128) (a[k] % q) * (mult % q) + (b[k] % q);
		out[k] = (uint64_t) (v % q);
	}
}

/* elemmulconst.comp:35-49 with elemmulconst.c:154-156 */
void oracle_elemmulconst(const uint64_t *in, uint64_t *out, uint64_t len,
		uint64_t b, uint64_t q) {
	const uint64_t b_shoup = oracle_shoup_factor(b, q);
	for (uint64_t k = 0; k < len; k++) {
		out[k] = shader_shoup(in[k], b, b_shoup, q);
	}
}

/* elemgtadd.comp:20-30 */
void oracle_elemgtadd(const uint64_t *in, uint64_t *out, uint64_t len,
		uint64_t bound, uint64_t diff) {
	for (uint64_t k = 0; k < len; k++) {
		out[k] = in[k] > bound ? in[k] + diff : in[k];
	}
}

/* elemgtsub.comp:49-67.  The shader reduces with its Barrett reduce64, which
 * is only defined for bits(q) >= 2 ... ; `literal` selects it, otherwise the
 * plain remainder states the same contract. */
void oracle_elemgtsub(const uint64_t *in, uint64_t *out, uint64_t len,
		uint64_t bound, uint64_t diff, uint64_t q, int literal) {
	const unsigned bits = bit_length(q);
	const uint64_t mu = literal ? barrett_mu(q) : 0;
	for (uint64_t k = 0; k < len; k++) {
		const uint64_t reduced = literal
			? shader_reduce64(in[k], q, mu, bits) : in[k] % q;
		const uint64_t diff_reduced = literal
			? shader_reduce64(diff, q, mu, bits) : diff % q;
		if (in[k] > bound) {
			if (literal) {
				/* 64-bit wrap-around as in the shader (needs q < 2^63) */
				uint64_t z = reduced + q - diff_reduced;
				if (z >= q) {
					z -= q;
				}
				out[k] = z;
			} else {
				out[k] = (uint64_t) (((u128) reduced + q - diff_reduced) % q);
			}
		} else {
			out[k] = reduced;
		}
	}
}

/* elemmodbytwo.comp:19-27 */
void oracle_elemmodbytwo(const uint64_t *in, uint64_t *out, uint64_t len,
		uint64_t signed_bound) {
	for (uint64_t k = 0; k < len; k++) {
		const uint64_t u = in[k] & 1;
		out[k] = in[k] > signed_bound ? 1 - u : u;
	}
}

/* dispatch rule of vkhel_vector_elemmod, src/vector.c:360-368 */
void oracle_elemmod(const uint64_t *in, uint64_t *out, uint64_t len,
		uint64_t mod, uint64_t q) {
	if (mod == 2) {
		oracle_elemmodbytwo(in, out, len, q / 2);
	} else {
		oracle_elemgtsub(in, out, len, q / 2, q, mod, 0);
	}
}

/* ---- independent cross-check: schoolbook negacyclic product ------------------------ */
/* c = a * b mod (x^n + 1, q); O(n^2), for small n only */
void oracle_negacyclic_schoolbook(const uint64_t *a, const uint64_t *b,
		uint64_t *c, uint64_t n, uint64_t q) {
	for (uint64_t k = 0; k < n; k++) {
		c[k] = 0;
	}
	for (uint64_t i = 0; i < n; i++) {
		for (uint64_t j = 0; j < n; j++) {
			const uint64_t p = (uint64_t) (((u128) a[i] * b[j]) % q);
			const uint64_t k = (i + j) % n;
			if (i + j < n) {
				c[k] = (uint64_t) (((u128) c[k] + p) % q);
			} else {
				c[k] = (uint64_t) (((u128) c[k] + q - p) % q);
			}
		}
	}
}

/* ---- batched drivers (OpenMP over polynomials): the timed CPU baseline -------------- */
/* layout [polys][n]; polynomial p uses table p % limbs; tables[l] points at
 * 4 arrays of n: roots, roots_shoup, inv_roots, inv_roots_shoup */
void oracle_forward_batch(const uint64_t *in, uint64_t *out, uint64_t n,
		uint64_t polys, uint64_t limbs, const uint64_t *moduli,
		const uint64_t *const *roots, const uint64_t *const *roots_shoup,
		int threads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
	for (int64_t p = 0; p < (int64_t) polys; p++) {
		const uint64_t l = (uint64_t) p % limbs;
		oracle_forward(in + p * n, out + p * n, n, moduli[l], roots[l],
				roots_shoup[l]);
	}
}

void oracle_inverse_batch(const uint64_t *in, uint64_t *out, uint64_t n,
		uint64_t polys, uint64_t limbs, const uint64_t *moduli,
		const uint64_t *const *inv_roots,
		const uint64_t *const *inv_roots_shoup, int threads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
	for (int64_t p = 0; p < (int64_t) polys; p++) {
		const uint64_t l = (uint64_t) p % limbs;
		oracle_inverse(in + p * n, out + p * n, n, n, moduli[l], inv_roots[l],
				inv_roots_shoup[l]);
	}
}
