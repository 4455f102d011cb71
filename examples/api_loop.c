/*
 * The reference API in a loop, from C: `count` separate vectors of n
 * coefficients, vkhel_vector_forward_transform on each (one call per vector,
 * exactly as a reference user writes it, src/vector.c:513-574), then one
 * synchronisation.  Prints host microseconds per transform.
 *
 * The calls are only recorded by the library and go out as one indirect
 * batched launch at the synchronisation; run with VKHEL_NO_DEFER=1 to see one
 * pair of launches per call instead.
 *
 *   build/bin/api_loop [log2n] [count]
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <vkhel.h>
#include <vkhel_ext.h>

static uint64_t powmod(uint64_t b, uint64_t e, uint64_t q) {
	unsigned __int128 r = 1, x = b % q;
	while (e) {
		if (e & 1) {
			r = r * x % q;
		}
		x = x * x % q;
		e >>= 1;
	}
	return (uint64_t) r;
}

static double now_us(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(int argc, char **argv) {
	const unsigned log2n = argc > 1 ? (unsigned) atoi(argv[1]) : 12;
	const size_t count = argc > 2 ? (size_t) atol(argv[2]) : 1024;
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
	struct vkhel_vector **vecs = calloc(count, sizeof(*vecs));
	uint64_t *host = malloc(n * sizeof(*host));
	for (uint64_t i = 0; i < n; i++) {
		host[i] = (i * 0x9E3779B97F4A7C15ull) % q;
	}
	for (size_t v = 0; v < count; v++) {
		vecs[v] = vkhel_vector_create(ctx, n);
		vkhel_vector_copy_from_host(vecs[v], host);
	}
	const int reps = 6;
	double best = 1e30;
	for (int r = 0; r < reps; r++) {
		const double t0 = now_us();
		for (size_t v = 0; v < count; v++) {
			vkhel_vector_forward_transform(vecs[v], vecs[v], ntt);
		}
		vkhel_ctx_sync(ctx);
		const double dt = now_us() - t0;
		if (r > 0 && dt < best) {
			best = dt;
		}
	}
	/* all vectors started equal and saw the same transforms */
	uint64_t *m0 = NULL, *m1 = NULL;
	vkhel_vector_map(vecs[0], (void **) &m0, n * sizeof(uint64_t));
	vkhel_vector_map(vecs[count - 1], (void **) &m1, n * sizeof(uint64_t));
	int same = 1;
	for (uint64_t i = 0; i < n; i++) {
		same &= m0[i] == m1[i];
	}
	vkhel_vector_unmap(vecs[0]);
	vkhel_vector_unmap(vecs[count - 1]);
	uint64_t batches = 0, carried = 0;
	vkhel_ctx_deferred_stats(ctx, &batches, &carried);
	printf("{\"config\": \"reference API loop, C\", \"log2n\": %u, \"vectors\": %zu, "
			"\"us_per_transform\": %.3f, \"transforms_per_s\": %.0f, "
			"\"batched_launches\": %" PRIu64 ", \"transforms_in_batches\": %" PRIu64
			", \"consistent\": %s}\n",
			log2n, count, best / count, count / best * 1e6, batches, carried,
			same ? "true" : "false");
	for (size_t v = 0; v < count; v++) {
		vkhel_vector_destroy(vecs[v]);
	}
	vkhel_ntt_tables_destroy(ntt);
	vkhel_ctx_destroy(ctx);
	free(vecs);
	free(host);
	return same ? 0 : 1;
}
