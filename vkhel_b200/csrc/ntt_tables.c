/*
 * NTT twiddle tables (host).  Contract of the reference's
 * src/ntt_tables.c:17-44,65-87:
 *   roots_of_unity[brv(i)]      = w^i mod q          (brv over log2 n bits)
 *   inv_roots_of_unity[k]       = roots_of_unity[k]^-1 mod q
 *   roots_barrett_factors[k]    = floor(roots_of_unity[k] * 2^64 / q)
 *   inv_roots_barrett_factors[k]= floor(inv_roots_of_unity[k] * 2^64 / q)
 * The reference obtains every inverse with its own extended-Euclid run; here
 * the inverses are the powers of w^-1 (one inversion in total), and the power
 * chain is advanced with a Shoup multiplication by the fixed factor.  Values
 * are identical; n = 2^16 takes ~3 ms instead of ~21 ms.
 *
 * The device mirror (B200 addition) is owned by the device layer
 * (device.cu: ntt_tables_device_pairs / ntt_tables_release_device).
 */
#include <assert.h>
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include "priv/ntt_tables.h"
#include "priv/numbers.h"

typedef unsigned __int128 u128;

void ntt_tables_release_device(struct vkhel_ntt_tables *ntt); /* device.cu */

static uint64_t next_serial = 1;

static uint64_t bit_reverse(uint64_t v, unsigned width) {
	uint64_t out = 0;
	for (unsigned b = 0; b < width; b++) {
		out = (out << 1) | ((v >> b) & 1);
	}
	return out;
}

/* x * f mod q for fixed f < q with f_shoup = floor(f * 2^64 / q); exact for
 * any 64-bit x when q < 2^63 */
static inline uint64_t mul_fixed(uint64_t x, uint64_t f, uint64_t f_shoup,
		uint64_t q) {
	const uint64_t hi = (uint64_t) (((u128) x * f_shoup) >> 64);
	uint64_t r = x * f - hi * q;
	return r >= q ? r - q : r;
}

static void fill_power_chain(uint64_t *out, uint64_t n, unsigned log2n,
		uint64_t base, uint64_t q) {
	const uint64_t base_shoup = nt_compute_barrett_factor(base, q, 64);
	uint64_t p = 1 % q;
	out[0] = p;
	for (uint64_t i = 1; i < n; i++) {
		if (q >> 63) {
			p = nt_multiply_mod(p, base, q, 0);
		} else {
			p = mul_fixed(p, base, base_shoup, q);
		}
		out[bit_reverse(i, log2n)] = p;
	}
}

/* allocate the struct and its four arrays and fill in the scalars; the arrays
 * are filled by the caller (host chain below, or the device generator of
 * tables_device.cu) */
struct vkhel_ntt_tables *ntt_tables_alloc(uint64_t n, uint64_t q, uint64_t w) {
	assert(n >= 1 && (n & (n - 1)) == 0 && "n must be a power of two");
	assert(q >= 2);

	struct vkhel_ntt_tables *ntt = calloc(1, sizeof(*ntt));
	assert(ntt);
	ntt->n = n;
	ntt->q = q;
	ntt->w = w;
	ntt->serial = __atomic_fetch_add(&next_serial, 1, __ATOMIC_RELAXED);
	ntt->log2n = nt_ceil_log2(n) - 1;

	ntt->roots_of_unity = malloc(sizeof(uint64_t) * n);
	ntt->inv_roots_of_unity = malloc(sizeof(uint64_t) * n);
	ntt->roots_barrett_factors = malloc(sizeof(uint64_t) * n);
	ntt->inv_roots_barrett_factors = malloc(sizeof(uint64_t) * n);
	assert(ntt->roots_of_unity && ntt->inv_roots_of_unity
			&& ntt->roots_barrett_factors
			&& ntt->inv_roots_barrett_factors);

	/* n^-1 for the inverse transform's scaling (reference vector.c:633) */
	ntt->inv_n = nt_inverse_mod(n % q, q);
	ntt->inv_n_shoup = nt_compute_barrett_factor(ntt->inv_n, q, 64);
	return ntt;
}

struct vkhel_ntt_tables *vkhel_ntt_tables_create(uint64_t n,
		uint64_t q, uint64_t w) {
	struct vkhel_ntt_tables *ntt = ntt_tables_alloc(n, q, w);

	const uint64_t w_red = w % q;
	fill_power_chain(ntt->roots_of_unity, n, ntt->log2n, w_red, q);
	/* w = 0 has no inverse; the reference would fault in that case too */
	const uint64_t w_inv = (n > 1) ? nt_inverse_mod(w_red, q) : 1 % q;
	fill_power_chain(ntt->inv_roots_of_unity, n, ntt->log2n, w_inv, q);

	for (uint64_t i = 0; i < n; i++) {
		ntt->roots_barrett_factors[i] = nt_compute_barrett_factor(
				ntt->roots_of_unity[i], q, 64);
		ntt->inv_roots_barrett_factors[i] = nt_compute_barrett_factor(
				ntt->inv_roots_of_unity[i], q, 64);
	}
	return ntt;
}

void vkhel_ntt_tables_destroy(struct vkhel_ntt_tables *ntt) {
	if (!ntt) {
		return;
	}
	ntt_tables_release_device(ntt);
	free(ntt->roots_of_unity);
	free(ntt->inv_roots_of_unity);
	free(ntt->roots_barrett_factors);
	free(ntt->inv_roots_barrett_factors);
	free(ntt); /* the reference leaks the struct (ntt_tables.c:82-87) */
}

void vkhel_ntt_tables_dbgprint(struct vkhel_ntt_tables *ntt) {
	printf("ntt_tables: (n=%" PRIu64 " q=%" PRIu64 " w=%" PRIu64 ")\n",
			ntt->n, ntt->q, ntt->w);
	const uint64_t *rows[2] = {
		ntt->roots_of_unity, ntt->inv_roots_of_unity };
	const char *names[2] = { "roots_of_unity", "inv_roots_of_unity" };
	for (int r = 0; r < 2; r++) {
		printf("\t%s: ", names[r]);
		for (uint64_t i = 0; i < ntt->n; i++) {
			printf(i + 1 == ntt->n ? "%" PRIu64 : "%" PRIu64 ", ",
					rows[r][i]);
		}
		printf("\n");
	}
}
