/*
 * Multi-GPU use of the C API: an RNS batch [batch][limbs][n] is sharded by
 * limb over all visible GPUs (one vkhel context and one host thread per GPU,
 * SURVEY 8e), transformed with no exchange between the GPUs, and gathered on
 * GPU 0 over NVLink with vkhel_vector_copy_peer only because the caller asks
 * for it.  Each limb is checked against a single-GPU transform.
 *
 *   build/bin/multi_gpu [log2n] [limbs] [batch]
 */
#include <inttypes.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vkhel.h>
#include <vkhel_ext.h>
#include "priv/numbers.h"

#define MAX_GPUS 16

struct shard {
	int device, gpus;
	uint64_t n, limbs, batch;
	const uint64_t *primes, *psis;
	const uint64_t *input;          /* [batch][limbs][n] on the host */
	struct vkhel_ctx *ctx;
	struct vkhel_vector *vec;       /* [batch][own limbs][n] */
	uint64_t lo, hi;                /* limb range */
};

static void *run_shard(void *arg) {
	struct shard *s = arg;
	const uint64_t own = s->hi - s->lo;
	s->ctx = vkhel_ctx_create_device(s->device);
	s->vec = vkhel_vector_create2(s->ctx, s->batch * own * s->n, false);
	struct vkhel_ntt_tables **tables = calloc(own ? own : 1, sizeof(*tables));
	for (uint64_t l = 0; l < own; l++) {
		tables[l] = vkhel_ntt_tables_create(s->n, s->primes[s->lo + l],
				s->psis[s->lo + l]);
	}
	/* upload this GPU's limbs of every batch entry */
	for (uint64_t b = 0; b < s->batch && own; b++) {
		vkhel_vector_upload(s->vec,
				s->input + (b * s->limbs + s->lo) * s->n,
				b * own * s->n, own * s->n);
	}
	if (own) {
		vkhel_vector_forward_transform_rns(s->vec, s->vec, tables, own,
				s->batch);
	}
	vkhel_ctx_sync(s->ctx);
	for (uint64_t l = 0; l < own; l++) {
		vkhel_ntt_tables_destroy(tables[l]);
	}
	free(tables);
	return NULL;
}

static uint64_t next_prime_below(uint64_t start, uint64_t step) {
	for (uint64_t c = start; ; c -= step) {
		int prime = c % 2 == 1;
		/* Miller-Rabin to the deterministic 64-bit bases, with the
		 * library's own host arithmetic */
		static const uint64_t bases[] = { 2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37 };
		uint64_t d = c - 1;
		int r = 0;
		while (d % 2 == 0) { d /= 2; r++; }
		for (unsigned i = 0; prime && i < sizeof(bases) / sizeof(bases[0]); i++) {
			uint64_t x = nt_power_mod(bases[i], d, c);
			if (x == 1 || x == c - 1) continue;
			int composite = 1;
			for (int j = 1; j < r; j++) {
				x = nt_multiply_mod(x, x, c, 0);
				if (x == c - 1) { composite = 0; break; }
			}
			if (composite) prime = 0;
		}
		if (prime) return c;
	}
}

int main(int argc, char **argv) {
	const unsigned log2n = argc > 1 ? atoi(argv[1]) : 12;
	const uint64_t limbs = argc > 2 ? atoi(argv[2]) : 8;
	const uint64_t batch = argc > 3 ? atoi(argv[3]) : 4;
	const uint64_t n = 1ull << log2n;
	int gpus = vkhel_device_count();
	if (gpus < 1) {
		fprintf(stderr, "multi_gpu: no CUDA device\n");
		return 1;
	}
	if (gpus > MAX_GPUS) gpus = MAX_GPUS;

	/* NTT primes below 2^60 with q = 1 (mod 2^18), and 2n-th roots */
	uint64_t *primes = malloc(limbs * sizeof(uint64_t));
	uint64_t *psis = malloc(limbs * sizeof(uint64_t));
	uint64_t cand = (1ull << 60) + 1;
	for (uint64_t l = 0; l < limbs; l++) {
		cand = next_prime_below(cand - (1ull << 18), 1ull << 18);
		primes[l] = cand;
		for (uint64_t x = 2; ; x++) {
			const uint64_t psi = nt_power_mod(x, (cand - 1) / (2 * n), cand);
			if (nt_is_primitive_root(psi, 2 * n, cand)) { psis[l] = psi; break; }
		}
	}

	const uint64_t total = batch * limbs * n;
	uint64_t *input = vkhel_host_alloc(total * sizeof(uint64_t));
	uint64_t state = 0x9E3779B97F4A7C15ull;
	for (uint64_t b = 0; b < batch; b++)
		for (uint64_t l = 0; l < limbs; l++)
			for (uint64_t k = 0; k < n; k++) {
				state ^= state << 13; state ^= state >> 7; state ^= state << 17;
				input[(b * limbs + l) * n + k] = state % primes[l];
			}

	/* one host thread + context per GPU, contiguous limb ranges */
	struct shard shards[MAX_GPUS];
	pthread_t threads[MAX_GPUS];
	for (int g = 0; g < gpus; g++) {
		struct shard *s = &shards[g];
		memset(s, 0, sizeof(*s));
		s->device = g; s->gpus = gpus; s->n = n; s->limbs = limbs; s->batch = batch;
		s->primes = primes; s->psis = psis; s->input = input;
		const uint64_t base = limbs / gpus, extra = limbs % gpus;
		s->lo = g * base + ((uint64_t) g < extra ? (uint64_t) g : extra);
		s->hi = s->lo + base + ((uint64_t) g < extra ? 1 : 0);
		pthread_create(&threads[g], NULL, run_shard, s);
	}
	for (int g = 0; g < gpus; g++) pthread_join(threads[g], NULL);

	/* optional gather on GPU 0: [batch][limbs][n] */
	struct vkhel_vector *all = vkhel_vector_create2(shards[0].ctx, total, false);
	for (int g = 0; g < gpus; g++) {
		const uint64_t own = shards[g].hi - shards[g].lo;
		for (uint64_t b = 0; b < batch && own; b++) {
			vkhel_vector_copy_peer(all, (b * limbs + shards[g].lo) * n,
					shards[g].vec, b * own * n, own * n);
		}
	}
	uint64_t *gathered = vkhel_host_alloc(total * sizeof(uint64_t));
	vkhel_vector_download(all, gathered, 0, total);
	vkhel_ctx_sync(shards[0].ctx);

	/* reference result: everything on GPU 0 in one call */
	struct vkhel_ntt_tables **tables = calloc(limbs, sizeof(*tables));
	for (uint64_t l = 0; l < limbs; l++)
		tables[l] = vkhel_ntt_tables_create(n, primes[l], psis[l]);
	struct vkhel_vector *single = vkhel_vector_create2(shards[0].ctx, total, false);
	vkhel_vector_upload(single, input, 0, total);
	vkhel_vector_forward_transform_rns(single, single, tables, limbs, batch);
	uint64_t *expect = vkhel_host_alloc(total * sizeof(uint64_t));
	vkhel_vector_download(single, expect, 0, total);
	vkhel_ctx_sync(shards[0].ctx);

	const int ok = memcmp(gathered, expect, total * sizeof(uint64_t)) == 0;
	printf("multi_gpu: %d GPU(s), n=2^%u, %" PRIu64 " limbs x batch %" PRIu64
			": sharded + gathered %s single-GPU result\n", gpus, log2n, limbs,
			batch, ok ? "==" : "!=");

	vkhel_vector_destroy(single);
	vkhel_vector_destroy(all);
	for (uint64_t l = 0; l < limbs; l++) vkhel_ntt_tables_destroy(tables[l]);
	free(tables);
	for (int g = 0; g < gpus; g++) {
		vkhel_vector_destroy(shards[g].vec);
		if (g) vkhel_ctx_destroy(shards[g].ctx);
	}
	vkhel_ctx_destroy(shards[0].ctx);
	vkhel_host_free(input); vkhel_host_free(gathered); vkhel_host_free(expect);
	free(primes); free(psis);
	return ok ? 0 : 1;
}
