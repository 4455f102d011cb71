/*
 * Shared declarations of the CUDA side (device layer, vectors, kernels).
 */
#ifndef VKHEL_COMMON_CUH
#define VKHEL_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "priv/vkhel.h"
#include "priv/ntt_tables.h"
#include "priv/numbers.h"
#include "modarith.cuh"

/* Error convention of the reference (assert -> abort): print where
 * and what, then abort.  Unlike assert() this is never compiled out. */
#define VK_DIE(...) do { \
		fprintf(stderr, "vkhel: %s:%d: ", __FILE__, __LINE__); \
		fprintf(stderr, __VA_ARGS__); \
		fputc('\n', stderr); \
		abort(); \
	} while (0)

#define VK_REQUIRE(cond, ...) do { if (!(cond)) VK_DIE(__VA_ARGS__); } while (0)

#define CUDA_CHECK(call) do { \
		cudaError_t err__ = (call); \
		if (err__ != cudaSuccess) \
			VK_DIE("%s failed: %s", #call, cudaGetErrorString(err__)); \
	} while (0)

static inline cudaStream_t ctx_stream(const struct vkhel_ctx *ctx) {
	return (cudaStream_t) ctx->dev.stream;
}

/* ---- device view of one table (one RNS limb) -------------------------------
 * tw[k]     = (root[k],     floor(root[k]*2^64/q))      k in [0,n)
 * tw[n + k] = (inv_root[k], floor(inv_root[k]*2^64/q))
 * in the reference's bit-reversed order (stage m uses the slice [m, 2m)).
 * tw[2n + k] = the pair of inv_root[k] * n^-1 mod q, k in [0, scaled_tw_pairs(n)):
 * the top of the inverse twiddle heap with the inverse transform's n^-1 factor
 * multiplied in (column pass of the inverse, kernels_ntt.cu). */
#define SCALED_TW_MAX_LOG2 10
static inline uint64_t scaled_tw_pairs(uint64_t n) {
	return n < (1ull << SCALED_TW_MAX_LOG2) ? n : (1ull << SCALED_TW_MAX_LOG2);
}
/* bytes of the (w, w') pairs of one device mirror */
static inline size_t mirror_pair_bytes(uint64_t n) {
	return (size_t) (2 * n + scaled_tw_pairs(n)) * sizeof(ulonglong2);
}

struct limb_desc {
	const ulonglong2 *tw;
	u64 q;
	u64 inv_n, inv_n_shoup;       /* n^-1 and its Shoup companion */
	u64 inv_w1n, inv_w1n_shoup;   /* inv_root[1] * n^-1: last inverse stage */
	/* a*b mod q for canonical a, b (fused point-wise product): divisor
	 * q << mm_s and its Moller-Granlund reciprocal, see struct modulus */
	u64 mm_d, mm_v;
	unsigned mm_s, pad_[3];
};

/* one polynomial of an indirect batch (deferred single-vector transforms that
 * are launched together, vector.cu) */
struct ntt_ptrs {
	const u64 *src;
	u64 *dst;
	const u64 *src2;   /* second factor of a recorded inverse-of-product, else unused */
};

/* device.cu */
extern "C" void *device_alloc(struct vkhel_ctx *ctx, size_t bytes);
extern "C" void device_free(struct vkhel_ctx *ctx, void *ptr);
extern "C" void *device_scratch(struct vkhel_ctx *ctx, size_t bytes);
extern "C" void *pinned_acquire(struct vkhel_ctx *ctx, size_t bytes);
extern "C" void pinned_release(struct vkhel_ctx *ctx, void *ptr);
extern "C" void pinned_release_after(struct vkhel_ctx *ctx, void *ptr,
		void *stream);
/* memcpy, shared with helper threads from 128 KiB on (hostcopy.cu) */
extern "C" void host_copy(void *dst, const void *src, size_t bytes);
/* device pointer to the limb_desc of `ntt` on ctx's device (uploads the
 * mirror on first use) */
const limb_desc *ntt_tables_device_desc(struct vkhel_ctx *ctx,
		struct vkhel_ntt_tables *ntt);
/* device array of limb_desc for an RNS basis (cached per context) */
const limb_desc *rns_plan_device_descs(struct vkhel_ctx *ctx,
		struct vkhel_ntt_tables *const *ntt, uint64_t limbs);

/* kernels_elem.cu */
struct modulus make_modulus(uint64_t q);
void launch_elemmul(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *out, uint64_t len, uint64_t q);
void launch_elemmul_rns(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *out, const uint64_t *mods, uint64_t limbs, uint64_t n,
		uint64_t batch);
void launch_elemfma(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *out, uint64_t len, uint64_t mult, uint64_t q);
void launch_elemmulconst(struct vkhel_ctx *ctx, const u64 *in, u64 *out,
		uint64_t len, uint64_t b, uint64_t q);
void launch_elemgtadd(struct vkhel_ctx *ctx, const u64 *in, u64 *out,
		uint64_t len, uint64_t bound, uint64_t diff);
void launch_elemgtsub(struct vkhel_ctx *ctx, const u64 *in, u64 *out,
		uint64_t len, uint64_t bound, uint64_t diff, uint64_t q);
void launch_elemmodbytwo(struct vkhel_ctx *ctx, const u64 *in, u64 *out,
		uint64_t len, uint64_t signed_bound);

/* kernels_ntt.cu: transform `polys` polynomials of n = 2^log2n coefficients,
 * polynomial p using descs[p % limbs].  dst may equal src. */
void launch_ntt(struct vkhel_ctx *ctx, bool inverse, const u64 *src, u64 *dst,
		const limb_desc *descs, uint64_t limbs, uint64_t polys,
		unsigned log2n, uint64_t q_max, bool lazy_out = false);
/* lazy_out (forward only, where ntt_lazy_forward_supported): dst receives
 * values in [0,3q) that only the inverse transform's kernels may read */
bool ntt_lazy_forward_supported(unsigned log2n, uint64_t q_max);
/* kernel launches of one transform launched by launch_ntt on the fast path */
unsigned ntt_launches_per_transform(unsigned log2n);
/* inverse transform of the point-wise product src * src2 (any 64-bit values,
 * reduced like the reference's elemmul) or, with fma_mult != 0 (already
 * reduced mod q, q itself standing for 0; one modulus only), of
 * src * fma_mult + src2 (elemfma); returns false when the fused kernel does not
 * apply and the caller has to run the point-wise kernel separately */
bool launch_ntt_inverse_of_product(struct vkhel_ctx *ctx, const u64 *src,
		const u64 *src2, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, uint64_t q_max, uint64_t fma_mult = 0);

/* c = INTT(NTT(a) (*) NTT(b)) per polynomial with the row passes of the three
 * transforms and the product in one kernel; tmp holds polys << log2n words
 * (unused for n <= 256).  dst may alias a and/or b.  Returns false when the
 * fast path does not apply (nothing has been launched then). */
bool launch_ntt_polymul(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *tmp, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, uint64_t q_max);

/* the same transform on `polys` separate polynomials, polynomial i read from
 * tab[i].src, written to tab[i].dst and using descs[i % limbs]; only where
 * ntt_indirect_supported().  The pointer table is either `tab` in device
 * memory or, for at most NTT_INLINE_PTRS polynomials, `host_tab` in host
 * memory, which travels in the kernel parameters (exactly one is non-NULL). */
#define NTT_INLINE_PTRS 8
bool ntt_indirect_supported(unsigned log2n, uint64_t q);
void launch_ntt_indirect(struct vkhel_ctx *ctx, bool inverse,
		const ntt_ptrs *tab, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, uint64_t q_max,
		const ntt_ptrs *host_tab = NULL, bool product = false);
/* `product` (inverse only): polynomial i is the inverse transform of the
 * point-wise product tab[i].src * tab[i].src2 (the reference's elemmul followed
 * by the in-place inverse transform), where ntt_indirect_product_supported() */
bool ntt_indirect_product_supported(unsigned log2n, uint64_t q);

/* kernels_ntt_cluster.cu: single-pass transform of 2^14 <= n <= 2^16 on a
 * thread-block cluster (polynomial distributed over the CTAs' shared memory);
 * same arguments as the fast path of launch_ntt */
bool ntt_cluster_enabled(unsigned log2n);
void launch_ntt_cluster(struct vkhel_ctx *ctx, bool inverse, bool apx,
		const u64 *src, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, unsigned limbs_total, unsigned limb0);

/* kernels_ntt_tma.cu: the forward column pass of n = 2^16 with its tile loaded
 * by one TMA tensor-map copy instead of per-thread loads (opt-in variant) */
bool ntt_cols_tma_enabled(unsigned log2n, unsigned kcol, unsigned s0);
void launch_ntt_cols_tma(struct vkhel_ctx *ctx, bool apx, const u64 *src,
		u64 *dst, const limb_desc *descs, uint64_t limbs, uint64_t polys,
		unsigned limbs_total, unsigned limb0);

/* kernels_ntt.cu: make the context's stream wait for the slices a sliced
 * transform has left on the auxiliary stream (no-op when there are none) */
void ntt_split_join(struct vkhel_ctx *ctx);

/* kernels_ntt_small.cu: whole products c = INTT(NTT(a) (*) NTT(b)) of small
 * polynomials (8 <= n <= 2^11), one CTA each, `count` of them in one launch;
 * the forward transforms are stored too (a_dst, b_dst).  Pointer table in
 * device memory (`tab`) or, for at most 4 products, in host memory
 * (`host_tab`, travels in the kernel parameters). */
#define SMALL_PRODUCT_MAX_LOG2N 11
struct small_product {
	const u64 *a_src, *b_src;
	u64 *a_dst, *b_dst, *c;
};
bool ntt_small_product_supported(unsigned log2n, uint64_t q);
void launch_ntt_small_products(struct vkhel_ctx *ctx, const small_product *tab,
		const small_product *host_tab, unsigned count, const limb_desc *desc,
		unsigned log2n, uint64_t q);

/* vector.cu: launch the deferred single-vector transforms of the context (all
 * of them, or only if they use `ntt`) */
void defer_flush(struct vkhel_ctx *ctx);
/* only a held forward transform (the RNS plan cache, before it evicts) */
void defer_flush_held(struct vkhel_ctx *ctx);
void defer_flush_tables(struct vkhel_ctx *ctx,
		const struct vkhel_ntt_tables *ntt);
void defer_destroy(struct vkhel_ctx *ctx);
void readahead_destroy(struct vkhel_ctx *ctx);

#endif
