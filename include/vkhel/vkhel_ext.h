/*
 * vkhel B200 extensions -- additive symbols (they match the vkhel_* export
 * glob of vkhel.syms, so the shared library exports them with no change to
 * the version script).  The reference API has no batch dimension, no device
 * argument and no asynchronous transfer; these calls add them
 * without touching the meaning of the 18 reference entry points.
 *
 * Layouts: a batched vector holds `batch` polynomials of ntt->n coefficients
 * back to back ([batch][n]); an RNS vector holds [batch][limbs][n] with limb l
 * reduced modulo ntt[l]->q.
 */
#ifndef VKHEL_EXT_H
#define VKHEL_EXT_H

#include <vkhel.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- device layer (replaces the reference's src/vulkan.c:153-225) -------- */
int vkhel_device_count(void);
/* like vkhel_ctx_create() (reference src/vkhel.c:4-8) but on CUDA device
 * `device`; vkhel_ctx_create() uses $VKHEL_DEVICE or device 0 (the reference
 * always takes physical device 0, src/vulkan.c:159-166). */
struct vkhel_ctx *vkhel_ctx_create_device(int device);
int vkhel_ctx_device(const struct vkhel_ctx *);
/* block until everything enqueued on the context has completed (the reference
 * does this after every op, src/vector.c:327) */
void vkhel_ctx_sync(struct vkhel_ctx *);
/* the context's cudaStream_t, for callers that interleave their own work.
 * What is recorded so far (see vkhel_ctx_deferred_stats) is launched, and from
 * this call on the context records nothing: a caller holding the stream -- or
 * a raw pointer from vkhel_vector_device_ptr, for operations on that vector --
 * finds every later call already enqueued when it returns. */
void *vkhel_ctx_stream(struct vkhel_ctx *);

/* vkhel_ntt_tables_create (reference src/ntt_tables.c:65-80) with the four
 * arrays computed on the context's GPU, one thread per power of w, and the
 * device mirror of the tables left in place for the transforms; same values,
 * same struct, destroyed with vkhel_ntt_tables_destroy.  The plain
 * vkhel_ntt_tables_create stays a host function (it takes no context). */
struct vkhel_ntt_tables *vkhel_ntt_tables_create_on(struct vkhel_ctx *,
		uint64_t n, uint64_t q, uint64_t w);

/* ---- pinned host memory + asynchronous transfers --------------------------
 * (reference: map/unmap staging, src/vector.c:262-296) */
void *vkhel_host_alloc(size_t bytes);
void vkhel_host_free(void *);
uint64_t vkhel_vector_length(const struct vkhel_vector *);
void *vkhel_vector_device_ptr(struct vkhel_vector *);
/* vkhel_vector_map (reference src/vector.c:270-283) for the element range
 * [offset, offset + count) only: stages and copies count*8 bytes instead of the
 * whole vector; vkhel_vector_unmap writes that range back.  The reference's
 * own map always moves the whole vector (its size argument is not usable, see
 * vkhel.h). */
void vkhel_vector_map_range(struct vkhel_vector *, void **mem,
		uint64_t offset, uint64_t count);
/* maps whose device -> host copy had been started before they were called: a
 * map of a vector that a recorded batch of transforms produced also starts the
 * copies of the next vectors of that batch (vector.cu, read-ahead) */
uint64_t vkhel_ctx_readahead_hits(const struct vkhel_ctx *);
/* batched forward transforms that were followed at once by the in-place
 * inverse transform of their result and therefore stored lazy residues the
 * inverse overwrote (vector.cu, held forward transform); does not flush */
uint64_t vkhel_ctx_lazy_forwards(const struct vkhel_ctx *);
/* enqueue copies of `count` elements starting at element `offset`; the host
 * buffer must stay valid until vkhel_ctx_sync / map / destroy */
void vkhel_vector_upload(struct vkhel_vector *, const uint64_t *src,
		uint64_t offset, uint64_t count);
void vkhel_vector_download(const struct vkhel_vector *, uint64_t *dst,
		uint64_t offset, uint64_t count);

/* copy `count` elements between vectors that may live in different contexts
 * (different GPUs: the copy goes over NVLink peer-to-peer).  Ordered after the
 * work enqueued so far on the source context and before later work on the
 * destination context.  This is the "gather the shards on request" step of a
 * multi-GPU run; the transforms themselves need no exchange. */
void vkhel_vector_copy_peer(struct vkhel_vector *dst, uint64_t dst_offset,
		const struct vkhel_vector *src, uint64_t src_offset, uint64_t count);

/* ---- batched / RNS transforms --------------------------------------------
 * Same arithmetic per polynomial as vkhel_vector_forward_transform /
 * vkhel_vector_inverse_transform (reference src/vector.c:513-657); exactly
 * batch*limbs*n elements are read and written.  result may alias operand. */
void vkhel_vector_forward_transform_batch(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *ntt, uint64_t batch);
void vkhel_vector_inverse_transform_batch(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *ntt, uint64_t batch);
void vkhel_vector_forward_transform_rns(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs, uint64_t batch);
void vkhel_vector_inverse_transform_rns(
		const struct vkhel_vector *operand, struct vkhel_vector *result,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs, uint64_t batch);

/* element-wise product / fma with one modulus per limb, [batch][limbs][n]
 * (per element: reference src/kernels/shaders/elemmul.comp:62-73 and the
 * canonical elemfma contract) */
void vkhel_vector_elemmul_rns(
		const struct vkhel_vector *a, const struct vkhel_vector *b,
		struct vkhel_vector *result, const uint64_t *mods,
		uint64_t limbs, uint64_t n, uint64_t batch);

/* negacyclic polynomial product c = INTT(NTT(a) (*) NTT(b)) per limb, the
 * reference call sequence forward,forward,elemmul,inverse
 * (src/vector.c:388-427,513-657) in one call.  a and b are left untouched. */
void vkhel_vector_polymul_rns(
		const struct vkhel_vector *a, const struct vkhel_vector *b,
		struct vkhel_vector *result,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs, uint64_t batch);

/* ---- device timing (CUDA events on the context's stream) ----------------- */
struct vkhel_timer;
struct vkhel_timer *vkhel_timer_create(struct vkhel_ctx *);
void vkhel_timer_start(struct vkhel_timer *);
void vkhel_timer_stop(struct vkhel_timer *);
/* waits for the stop event; milliseconds between start and stop */
double vkhel_timer_elapsed_ms(struct vkhel_timer *);
void vkhel_timer_destroy(struct vkhel_timer *);

/* number of kernels this context has launched so far (what is recorded is
 * launched first; _noflush reads the counter without doing that) */
uint64_t vkhel_ctx_launch_count(const struct vkhel_ctx *);
uint64_t vkhel_ctx_launch_count_noflush(const struct vkhel_ctx *);
/* vkhel_vector_forward_transform / _inverse_transform (one vector per call,
 * reference src/vector.c:513-657) are recorded and consecutive independent
 * calls with the same tables go out as one batched launch when the context is
 * next needed (see vector.cu).  These report how many such batches were
 * launched and how many transforms they carried; vkhel_ctx_flush launches
 * what is recorded without waiting for it.  $VKHEL_NO_DEFER=1 disables
 * recording. */
void vkhel_ctx_deferred_stats(const struct vkhel_ctx *, uint64_t *batches,
		uint64_t *transforms);
void vkhel_ctx_flush(struct vkhel_ctx *);
/* vkhel_vector_elemmul (reference src/vector.c:388-427) is recorded as well:
 * when the next call is the in-place vkhel_vector_inverse_transform of its
 * result with the same modulus -- the reference's polynomial product -- the
 * multiplication happens inside the inverse transform's first pass; any other
 * call launches the product first.  Number of products fused so far: */
uint64_t vkhel_ctx_fused_products(const struct vkhel_ctx *);
/* Integer-pipe probes on the context's device (probe.cu): the denominators of
 * the transform's binding roofline, measured in the run that reports them.
 * Writes up to `count` doubles and returns how many:
 *   [0] SM count                      [1] SM clock during the probes, MHz
 *   [2] IMAD        thread-instructions per clock per SM (mad.lo.u32)
 *   [3] IMAD.WIDE   (mad.wide.u32)    [4] IMAD.HI (mad.hi.u32)
 *   [5] LOP3        (the ALU pipe)
 *   [6] forward / [7] inverse lazy butterflies per clock per SM, the library's
 *       own butterfly code on register operands only
 *   [8] SHF (funnel shift)   [9] IADD3 with three addends
 *   [10] 64-bit addition of three values (IADD3 + IADD3.X, two carries)
 *   [11] 64-bit add + conditional subtraction (the butterflies' csub) */
int vkhel_ctx_probe_int_peaks(struct vkhel_ctx *, double *out, int count);
/* write a buffer larger than L2 so the next kernel starts cold */
void vkhel_ctx_flush_l2(struct vkhel_ctx *);

#ifdef __cplusplus
}
#endif

#endif
