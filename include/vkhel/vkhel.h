/*
 * vkhel public C API -- B200-native implementation.
 *
 * This header declares the same 18 entry points, with the same argument
 * meaning, as the reference's include/vkhel/vkhel.h:8-53, so programs written
 * against the reference (examples/example.c, test/vector.c) compile and link
 * unchanged.  Everything underneath is a CUDA device layer and hand-written
 * sm_100a kernels; there is no CPU fallback: every call aborts with a message
 * if no CUDA device is usable.
 *
 * Error convention (reference: assert()/abort, src/vector.c:221,302,517 ...):
 * no error codes.  A failed CUDA call or a violated precondition prints
 * "vkhel: <file>:<line>: <what>" to stderr and calls abort().
 *
 * Threading (reference: one VkQueue, one command pool): a context
 * and its vectors must be driven from one host thread at a time.  Operations
 * are enqueued asynchronously on the context's stream; map / dbgprint /
 * destroy observe completed work.
 *
 * Additive B200 extensions (batched / RNS transforms, fused polynomial
 * multiply, pinned host memory, timing) live in <vkhel_ext.h>.
 */
#ifndef VKHEL_H
#define VKHEL_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- context (reference: src/vkhel.c:4-13) ------------------------------ */
struct vkhel_ctx;
struct vkhel_ctx *vkhel_ctx_create(void); /* reference: empty parens */
void vkhel_ctx_destroy(struct vkhel_ctx *);

/* ---- NTT tables (reference: src/ntt_tables.c:65-87) ---------------------
 * n: transform size (power of two), q: modulus, w: primitive 2n-th root of
 * unity mod q.  Host-only object, independent of any context; the device
 * mirror is uploaded lazily on first use by a transform. */
struct vkhel_ntt_tables;
struct vkhel_ntt_tables *vkhel_ntt_tables_create(
		uint64_t n, uint64_t q, uint64_t w);
void vkhel_ntt_tables_destroy(struct vkhel_ntt_tables *);

/* ---- vectors (reference: src/vector.c:206-296) --------------------------- */
struct vkhel_vector;
struct vkhel_vector *vkhel_vector_create(struct vkhel_ctx *, uint64_t length);
struct vkhel_vector *vkhel_vector_create2(struct vkhel_ctx *, uint64_t length,
		bool zero);
void vkhel_vector_destroy(struct vkhel_vector *);
struct vkhel_vector *vkhel_vector_dup(struct vkhel_vector *);
void vkhel_vector_copy_from_host(struct vkhel_vector *, const uint64_t *);
/* map: read-write host view of the whole vector (the size argument is
 * accepted for compatibility; length*8 bytes are always staged, which is a
 * superset of both reference call conventions). */
void vkhel_vector_map(struct vkhel_vector *, void **, size_t);
void vkhel_vector_unmap(struct vkhel_vector *);

/* ---- element-wise modular ops (reference: src/vector.c:298-511) ---------- */
/* result[k] = (a[k] * multiplier + b[k]) mod `mod` */
void vkhel_vector_elemfma(
		const struct vkhel_vector *a,
		const struct vkhel_vector *b,
		struct vkhel_vector *result,
		uint64_t multiplier, uint64_t mod);
/* centred reduction of a value in [0,q) to modulus `mod` */
void vkhel_vector_elemmod(
		const struct vkhel_vector *a,
		struct vkhel_vector *result, uint64_t mod, uint64_t q);
/* result[k] = (a[k] mod `mod`) * (b[k] mod `mod`) mod `mod` */
void vkhel_vector_elemmul(
		const struct vkhel_vector *a,
		const struct vkhel_vector *b,
		struct vkhel_vector *result, uint64_t mod);
/* result[k] = operand[k] > bound ? operand[k] + diff : operand[k] */
void vkhel_vector_elemgtadd(const struct vkhel_vector *operand,
		struct vkhel_vector *result,
		uint64_t bound, uint64_t diff);
/* result[k] = operand[k] > bound ? (operand[k] - diff) mod `mod`
 *                                : operand[k] mod `mod` */
void vkhel_vector_elemgtsub(
		const struct vkhel_vector *operand,
		struct vkhel_vector *result,
		uint64_t bound, uint64_t diff, uint64_t mod);

/* ---- negacyclic NTT (reference: src/vector.c:513-657) ---------------------
 * forward: Cooley-Tukey, natural order in, bit-reversed order out.
 * inverse: Gentleman-Sande, bit-reversed in, natural out, scaled by n^-1.
 * Only the first ntt->n elements are transformed; result may alias operand. */
void vkhel_vector_forward_transform(
		const struct vkhel_vector *operand,
		struct vkhel_vector *result,
		struct vkhel_ntt_tables *ntt);
void vkhel_vector_inverse_transform(
		const struct vkhel_vector *operand,
		struct vkhel_vector *result,
		struct vkhel_ntt_tables *ntt);

#ifdef __cplusplus
}
#endif

#endif
