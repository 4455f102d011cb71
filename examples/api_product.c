/*
 * The reference's polynomial product from C, exactly as a reference user
 * writes it (examples/example.c:18-60 of the reference): per pair of
 * polynomials
 *     vkhel_vector_forward_transform(a, a, ntt);
 *     vkhel_vector_forward_transform(b, b, ntt);
 *     vkhel_vector_elemmul(a, b, c, q);
 *     vkhel_vector_inverse_transform(c, c, ntt);
 * for `count` pairs, then one synchronisation.  Prints host microseconds per
 * product and checks the first product against the schoolbook negacyclic
 * product.
 *
 * The library records the two forward transforms (one batched launch), and
 * the element-wise product goes out inside the first pass of the inverse
 * transform.  VKHEL_NO_FUSED_PRODUCT=1 launches the product on its own,
 * VKHEL_NO_DEFER=1 launches every call on its own (7 kernels per product).
 *
 *   build/bin/api_product [log2n] [count]
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <vkhel.h>
#include <vkhel_ext.h>

static uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) {
	return (uint64_t) ((unsigned __int128) a * b % q);
}

static uint64_t powmod(uint64_t b, uint64_t e, uint64_t q) {
	uint64_t r = 1, x = b % q;
	while (e) {
		if (e & 1) {
			r = mulmod(r, x, q);
		}
		x = mulmod(x, x, q);
		e >>= 1;
	}
	return r;
}

static double now_us(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(int argc, char **argv) {
	const unsigned log2n = argc > 1 ? (unsigned) atoi(argv[1]) : 12;
	const size_t count = argc > 2 ? (size_t) atol(argv[2]) : 256;
	const uint64_t n = 1ull << log2n;
	const uint64_t q = 1152921504606584833ull; /* largest prime < 2^60, 1 mod 2^18 */
	uint64_t psi = 0;
	for (uint64_t x = 2; !psi; x++) {
		const uint64_t c = powmod(x, (q - 1) / (2 * n), q);
		if (powmod(c, n, q) == q - 1) {
			psi = c;
		}
	}
	struct vkhel_ctx *ctx = vkhel_ctx_create();
	struct vkhel_ntt_tables *ntt = vkhel_ntt_tables_create(n, q, psi);
	struct vkhel_vector **a = calloc(count, sizeof(*a));
	struct vkhel_vector **b = calloc(count, sizeof(*b));
	struct vkhel_vector **c = calloc(count, sizeof(*c));
	uint64_t *ha = malloc(n * sizeof(*ha)), *hb = malloc(n * sizeof(*hb));
	for (uint64_t i = 0; i < n; i++) {
		ha[i] = (i * 0x9E3779B97F4A7C15ull + 1) % q;
		hb[i] = (i * 0xC2B2AE3D27D4EB4Full + 7) % q;
	}
	for (size_t v = 0; v < count; v++) {
		a[v] = vkhel_vector_create(ctx, n);
		b[v] = vkhel_vector_create(ctx, n);
		c[v] = vkhel_vector_create(ctx, n);
	}
	const int reps = 6;
	double best = 1e30;
	for (int r = 0; r < reps; r++) {
		for (size_t v = 0; v < count; v++) {
			vkhel_vector_copy_from_host(a[v], ha);
			vkhel_vector_copy_from_host(b[v], hb);
		}
		vkhel_ctx_sync(ctx);
		const double t0 = now_us();
		for (size_t v = 0; v < count; v++) {
			vkhel_vector_forward_transform(a[v], a[v], ntt);
			vkhel_vector_forward_transform(b[v], b[v], ntt);
			vkhel_vector_elemmul(a[v], b[v], c[v], q);
			vkhel_vector_inverse_transform(c[v], c[v], ntt);
		}
		vkhel_ctx_sync(ctx);
		const double dt = now_us() - t0;
		if (r > 0 && dt < best) {
			best = dt;
		}
	}
	/* first and last product against the schoolbook negacyclic product
	 * (a few coefficients when n is large) */
	uint64_t *m0 = NULL, *m1 = NULL;
	vkhel_vector_map(c[0], (void **) &m0, n * sizeof(uint64_t));
	vkhel_vector_map(c[count - 1], (void **) &m1, n * sizeof(uint64_t));
	int ok = 1;
	const uint64_t step = n > 1024 ? n / 64 : 1;
	for (uint64_t k = 0; k < n; k += step) {
		uint64_t acc = 0;
		for (uint64_t i = 0; i < n; i++) {
			/* x^n = -1: a[i] * b[k - i], negated when the exponent wraps */
			const uint64_t j = (k + n - i) % n;
			const uint64_t t = mulmod(ha[i], hb[j], q);
			acc = i <= k ? (acc + t) % q : (acc + q - t) % q;
		}
		ok &= m0[k] == acc && m1[k] == acc;
	}
	vkhel_vector_unmap(c[0]);
	vkhel_vector_unmap(c[count - 1]);
	printf("{\"config\": \"reference product sequence, C\", \"log2n\": %u, "
			"\"products\": %zu, \"us_per_product\": %.3f, \"products_per_s\": %.0f, "
			"\"fused_products\": %" PRIu64 ", \"kernel_launches\": %" PRIu64
			", \"matches_schoolbook\": %s}\n",
			log2n, count, best / count, count / best * 1e6,
			vkhel_ctx_fused_products(ctx), vkhel_ctx_launch_count(ctx),
			ok ? "true" : "false");
	for (size_t v = 0; v < count; v++) {
		vkhel_vector_destroy(a[v]);
		vkhel_vector_destroy(b[v]);
		vkhel_vector_destroy(c[v]);
	}
	vkhel_ntt_tables_destroy(ntt);
	vkhel_ctx_destroy(ctx);
	free(a);
	free(b);
	free(c);
	free(ha);
	free(hb);
	return ok ? 0 : 1;
}
