/*
 * Dump program for tools/vulkan_parity.sh: runs seeded random vectors through a
 * vkhel library using ONLY the reference's public API (include/vkhel/vkhel.h:
 * the 18 entry points) and prints every result, one line per case:
 *
 *     <op> n=<n> q=<q> w=<w> seed=<s> [k=v ...] : v0 v1 v2 ...
 *
 * Linked against the REAL reference (built by meson over a Vulkan ICD) it shows
 * what the GLSL shaders compute; linked against this repository's libvkhel.so
 * it shows the CUDA path.  tools/vulkan_parity_check.py recomputes the inputs
 * from the seeds and compares the outputs with the CPU oracle.
 *
 *   cc -Iinclude/vkhel tools/vulkan_dump.c -L<libdir> -lvkhel -o vulkan_dump
 */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>

#include <vkhel.h>

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

static uint64_t find_psi(uint64_t n, uint64_t q) {
	for (uint64_t x = 2;; x++) {
		const uint64_t c = powmod(x, (q - 1) / (2 * n), q);
		if (powmod(c, n, q) == q - 1) {
			return c;
		}
	}
}

/* xorshift64, one stream per seed */
static void fill(uint64_t *out, uint64_t count, uint64_t seed, uint64_t mod) {
	uint64_t s = seed | 1;
	for (uint64_t i = 0; i < count; i++) {
		s ^= s << 13;
		s ^= s >> 7;
		s ^= s << 17;
		out[i] = mod ? s % mod : s;
	}
}

static void dump(const char *head, struct vkhel_vector *v, uint64_t len) {
	uint64_t *m = NULL;
	vkhel_vector_map(v, (void **) &m, len * sizeof(uint64_t));
	printf("%s :", head);
	for (uint64_t i = 0; i < len; i++) {
		printf(" %" PRIu64, m[i]);
	}
	printf("\n");
	vkhel_vector_unmap(v);
}

int main(void) {
	struct vkhel_ctx *ctx = vkhel_ctx_create();
	/* 60-bit and 61-bit NTT primes, the reference's own 50-bit test modulus */
	const uint64_t qs[] = { 1152921504606584833ull, 2305843009211596801ull,
		1125891450734593ull };
	char head[256];
	uint64_t seed = 1000;
	for (int qi = 0; qi < 3; qi++) {
		const uint64_t q = qs[qi];
		for (uint64_t n = 4; n <= 4096; n *= 4) {
			if ((q - 1) % (2 * n)) {
				continue;
			}
			const uint64_t w = find_psi(n, q);
			struct vkhel_ntt_tables *ntt = vkhel_ntt_tables_create(n, q, w);
			uint64_t *a = malloc(n * sizeof(*a)), *b = malloc(n * sizeof(*b));
			struct vkhel_vector *va = vkhel_vector_create(ctx, n);
			struct vkhel_vector *vb = vkhel_vector_create(ctx, n);
			struct vkhel_vector *vc = vkhel_vector_create(ctx, n);

			fill(a, n, ++seed, q);
			vkhel_vector_copy_from_host(va, a);
			vkhel_vector_forward_transform(va, vc, ntt);
			snprintf(head, sizeof(head), "forward n=%" PRIu64 " q=%" PRIu64
					" w=%" PRIu64 " seed=%" PRIu64, n, q, w, seed);
			dump(head, vc, n);
			vkhel_vector_inverse_transform(va, vc, ntt);
			snprintf(head, sizeof(head), "inverse n=%" PRIu64 " q=%" PRIu64
					" w=%" PRIu64 " seed=%" PRIu64, n, q, w, seed);
			dump(head, vc, n);

			/* arbitrary 64-bit operands for the product */
			fill(a, n, ++seed, 0);
			fill(b, n, seed + 7777, 0);
			vkhel_vector_copy_from_host(va, a);
			vkhel_vector_copy_from_host(vb, b);
			vkhel_vector_elemmul(va, vb, vc, q);
			snprintf(head, sizeof(head), "elemmul n=%" PRIu64 " q=%" PRIu64
					" w=0 seed=%" PRIu64, n, q, seed);
			dump(head, vc, n);

			/* elemfma within its contract (multiplier < q, b < q); half of the
			 * sums wrap, which is where the shader's defect shows */
			fill(a, n, ++seed, q);
			fill(b, n, seed + 7777, q);
			vkhel_vector_copy_from_host(va, a);
			vkhel_vector_copy_from_host(vb, b);
			const uint64_t mult = (seed * 0x9E3779B97F4A7C15ull) % q;
			vkhel_vector_elemfma(va, vb, vc, mult, q);
			snprintf(head, sizeof(head), "elemfma n=%" PRIu64 " q=%" PRIu64
					" w=0 seed=%" PRIu64 " mult=%" PRIu64, n, q, seed, mult);
			dump(head, vc, n);
			vkhel_vector_elemfma(va, vb, vc, 1, q);
			snprintf(head, sizeof(head), "elemfma n=%" PRIu64 " q=%" PRIu64
					" w=0 seed=%" PRIu64 " mult=1", n, q, seed);
			dump(head, vc, n);

			fill(a, n, ++seed, q);
			vkhel_vector_copy_from_host(va, a);
			vkhel_vector_elemmod(va, vc, 2, q);
			snprintf(head, sizeof(head), "elemmod n=%" PRIu64 " q=%" PRIu64
					" w=0 seed=%" PRIu64 " mod=2", n, q, seed);
			dump(head, vc, n);
			vkhel_vector_elemmod(va, vc, 65537, q);
			snprintf(head, sizeof(head), "elemmod n=%" PRIu64 " q=%" PRIu64
					" w=0 seed=%" PRIu64 " mod=65537", n, q, seed);
			dump(head, vc, n);
			vkhel_vector_elemgtadd(va, vc, q / 3, 12345);
			snprintf(head, sizeof(head), "elemgtadd n=%" PRIu64 " q=%" PRIu64
					" w=0 seed=%" PRIu64 " bound=%" PRIu64 " diff=12345", n, q,
					seed, q / 3);
			dump(head, vc, n);
			vkhel_vector_elemgtsub(va, vc, q / 3, 999, 1000003);
			snprintf(head, sizeof(head), "elemgtsub n=%" PRIu64 " q=%" PRIu64
					" w=0 seed=%" PRIu64 " bound=%" PRIu64 " diff=999 mod=1000003",
					n, q, seed, q / 3);
			dump(head, vc, n);

			vkhel_vector_destroy(va);
			vkhel_vector_destroy(vb);
			vkhel_vector_destroy(vc);
			vkhel_ntt_tables_destroy(ntt);
			free(a);
			free(b);
		}
	}
	vkhel_ctx_destroy(ctx);
	return 0;
}
