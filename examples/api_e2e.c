/*
 * End to end through the reference's 18 entry points only, from C, the way a
 * reference user moves data (src/vector.c:262-296, 513-657): for `count`
 * polynomials held as one vector each, from pageable host memory,
 *
 *     vkhel_vector_copy_from_host   (each vector)
 *     vkhel_vector_forward_transform (each vector)
 *     vkhel_vector_inverse_transform (each vector)
 *     vkhel_vector_map, read the result out, vkhel_vector_unmap (each vector)
 *
 * Host wall clock per step (a step ends when the last result has been read:
 * map waits for it).  The round trip is checked against the input.  Prints one
 * JSON line; bench.py reports it as `e2e_reference_api`.
 *
 *   build/bin/api_e2e [log2n] [count] [steps]
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
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
	const unsigned log2n = argc > 1 ? (unsigned) atoi(argv[1]) : 16;
	const size_t count = argc > 2 ? (size_t) atol(argv[2]) : 64;
	const int steps = argc > 3 ? atoi(argv[3]) : 5;
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
	/* pageable host memory, as a reference user's arrays are */
	uint64_t *in = malloc(count * n * sizeof(*in));
	uint64_t *out = malloc(count * n * sizeof(*out));
	uint64_t s = 0x9E3779B97F4A7C15ull;
	for (size_t i = 0; i < count * n; i++) {
		s ^= s << 13;
		s ^= s >> 7;
		s ^= s << 17;
		in[i] = s % q;
	}
	memset(out, 0, count * n * sizeof(*out));
	for (size_t v = 0; v < count; v++) {
		vecs[v] = vkhel_vector_create(ctx, n);
	}
	double best = 1e30, total = 0;
	/* host time of the phases: upload calls, transform calls, first map (which
	 * waits for everything enqueued so far), the other maps, the caller's own
	 * copy out of the mapped buffers */
	double t_in = 0, t_ntt = 0, t_first = 0, t_maps = 0, t_read = 0;
	for (int r = 0; r <= steps; r++) {   /* step 0 warms up */
		const double t0 = now_us();
		for (size_t v = 0; v < count; v++) {
			vkhel_vector_copy_from_host(vecs[v], in + v * n);
		}
		const double t1 = now_us();
		for (size_t v = 0; v < count; v++) {
			vkhel_vector_forward_transform(vecs[v], vecs[v], ntt);
		}
		for (size_t v = 0; v < count; v++) {
			vkhel_vector_inverse_transform(vecs[v], vecs[v], ntt);
		}
		const double t2 = now_us();
		double first = 0, maps = 0, read = 0;
		for (size_t v = 0; v < count; v++) {
			uint64_t *mapped = NULL;
			const double a = now_us();
			vkhel_vector_map(vecs[v], (void **) &mapped, n * sizeof(uint64_t));
			const double b = now_us();
			memcpy(out + v * n, mapped, n * sizeof(uint64_t));
			const double c = now_us();
			vkhel_vector_unmap(vecs[v]);
			const double d = now_us();
			if (v == 0) {
				first += b - a;
			} else {
				maps += b - a;
			}
			read += c - b;
			maps += d - c;
		}
		const double dt = now_us() - t0;
		if (r > 0) {
			total += dt;
			t_in += t1 - t0;
			t_ntt += t2 - t1;
			t_first += first;
			t_maps += maps;
			t_read += read;
			if (dt < best) {
				best = dt;
			}
		}
	}
	vkhel_ctx_sync(ctx);
	const int same = memcmp(in, out, count * n * sizeof(*in)) == 0;
	const double mean = total / steps;
	printf("{\"config\": \"reference API end to end, C\", \"log2n\": %u, "
			"\"vectors\": %zu, \"steps\": %d, \"us_per_step_mean\": %.1f, "
			"\"us_per_step_best\": %.1f, \"ntt_per_s\": %.0f, "
			"\"us_copy_from_host\": %.1f, \"us_transform_calls\": %.1f, "
			"\"us_first_map\": %.1f, \"us_other_maps_and_unmaps\": %.1f, "
			"\"us_caller_reads\": %.1f, "
			"\"readahead_hits\": %" PRIu64 ", \"round_trip_exact\": %s}\n",
			log2n, count, steps, mean, best, 2.0 * count / mean * 1e6,
			t_in / steps, t_ntt / steps, t_first / steps, t_maps / steps,
			t_read / steps,
			vkhel_ctx_readahead_hits(ctx), same ? "true" : "false");
	for (size_t v = 0; v < count; v++) {
		vkhel_vector_destroy(vecs[v]);
	}
	vkhel_ntt_tables_destroy(ntt);
	vkhel_ctx_destroy(ctx);
	free(vecs);
	free(in);
	free(out);
	return same ? 0 : 1;
}
