/*
 * NTT twiddle tables.  The first seven fields have the names, types and
 * order of the reference's struct (include/priv/ntt_tables.h:6-15; read
 * directly by test/ntt.c:19-23).  The fields after them are the B200 device
 * mirror: interleaved (w, floor(w*2^64/q)) pairs in the reference's
 * bit-reversed order, uploaded once per (table, device) on first use.
 */
#ifndef PRIV_NTT_TABLES_H
#define PRIV_NTT_TABLES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKHEL_MAX_DEVICES 16

struct vkhel_ntt_tables {
	uint64_t n; /* degree */
	uint64_t q; /* modulus */
	uint64_t w; /* root of unity (primitive 2n-th) */

	uint64_t *roots_of_unity;            /* [brv(i)] = w^i */
	uint64_t *inv_roots_of_unity;        /* element-wise inverses */
	uint64_t *roots_barrett_factors;     /* floor(root * 2^64 / q) */
	uint64_t *inv_roots_barrett_factors;

	/* ---- B200 additions (not in the reference) ---- */
	uint64_t serial;      /* unique id, keys the per-context plan cache */
	uint64_t log2n;
	uint64_t inv_n;       /* n^-1 mod q */
	uint64_t inv_n_shoup; /* floor(inv_n * 2^64 / q) */
	/* device mirrors, indexed by CUDA device ordinal; each is 2n pairs:
	 * [0,n) forward (w,w'), [n,2n) inverse (w^-1, w^-1'), 16 B per pair */
	void *dev_pairs[VKHEL_MAX_DEVICES];
};

void vkhel_ntt_tables_dbgprint(struct vkhel_ntt_tables *);
/* struct + arrays + scalars, arrays not yet filled (ntt_tables.c) */
struct vkhel_ntt_tables *ntt_tables_alloc(uint64_t n, uint64_t q, uint64_t w);

#ifdef __cplusplus
}
#endif

#endif
