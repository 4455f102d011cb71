/*
 * Twiddle tables generated on the GPU.
 *
 * The reference builds its tables on the host, serially: a running product
 * for the powers of psi and one extended-Euclid inversion per element
 * (src/ntt_tables.c:17-44; n = 2^16 takes ~21 ms, 32 limbs x 2^17 ~1.4 s).
 * vkhel_ntt_tables_create keeps that contract on the host (one inversion in
 * total, ~4 ms).  vkhel_ntt_tables_create_on computes the same four arrays on
 * the device of a context, one thread per power:
 *     roots[brv(i)]     = psi^i           inv_roots[brv(i)] = psi^-i
 *     *_barrett_factors = floor(value * 2^64 / q)
 * directly in the device mirror the transforms use ((w, w') pairs, common.cuh),
 * and copies them back into the host arrays the reference's struct exposes
 * (include/priv/ntt_tables.h:6-15, read by test/ntt.c:19-23).  Values are
 * identical to the host path (tests/test_gpu_tables.py).
 */
#include <string.h>
#include <time.h>

#include "common.cuh"
#include "vkhel_ext.h"

#define TABGEN_THREADS 256

static double tb_now(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

struct tabgen_params {
	u64 pw[32];    /* psi^(2^k) */
	u64 ipw[32];   /* psi^-(2^k) */
	modulus m;
	unsigned log2n;
};

/* floor(r * 2^64 / q) for r < q: the quotient of the two-word division that
 * reduce128 (modarith.cuh) only takes the remainder of */
__device__ __forceinline__ u64 shoup_companion(u64 r, const modulus &m) {
	const u64 u1 = r << m.s;      /* r < q, so u1 < d */
	u64 q0 = m.v * u1;
	u64 q1 = __umul64hi(m.v, u1);
	q1 += u1 + 1;                 /* (q1,q0) = v*u1 + (u1,0), then q1 + 1 */
	u64 rem = 0 - q1 * m.d;
	if (rem > q0) {
		q1--;
		rem += m.d;
	}
	if (rem >= m.d) {
		q1++;
	}
	return q1;
}

/* pairs: the device mirror, (w, w') interleaved; planar: the same values as
 * the four arrays of the host struct, [roots][inv_roots][roots'][inv_roots'] */
__global__ void __launch_bounds__(TABGEN_THREADS)
tables_generate_kernel(ulonglong2 *pairs, u64 *planar, const tabgen_params p) {
	const u64 n = (u64) 1 << p.log2n;
	const u64 i = (u64) blockIdx.x * TABGEN_THREADS + threadIdx.x;
	if (i >= n) {
		return;
	}
	u64 r = 1 % p.m.q, ri = r;
	for (unsigned k = 0; k < p.log2n; k++) {
		if ((i >> k) & 1) {
			r = mulmod(r, p.pw[k], p.m);
			ri = mulmod(ri, p.ipw[k], p.m);
		}
	}
	const u64 idx = p.log2n ? __brevll(i) >> (64 - p.log2n) : 0;
	const u64 rs = shoup_companion(r, p.m), ris = shoup_companion(ri, p.m);
	pairs[idx] = make_ulonglong2(r, rs);
	pairs[n + idx] = make_ulonglong2(ri, ris);
	planar[idx] = r;
	planar[n + idx] = ri;
	planar[2 * n + idx] = rs;
	planar[3 * n + idx] = ris;
}

/* device.cu */
void *ntt_tables_mirror_alloc(struct vkhel_ctx *ctx, size_t bytes);
void ntt_tables_adopt_mirror(struct vkhel_ctx *ctx,
		struct vkhel_ntt_tables *ntt, char *dev_buf);

extern "C" struct vkhel_ntt_tables *vkhel_ntt_tables_create_on(
		struct vkhel_ctx *ctx, uint64_t n, uint64_t q, uint64_t w) {
	VK_REQUIRE(ctx, "vkhel_ntt_tables_create_on: NULL context");
	VK_REQUIRE(n >= 1 && (n & (n - 1)) == 0, "n must be a power of two");
	if (n < 2 || q >= (1ull << 63) || n > (1ull << 30)) {
		/* nothing to parallelise / outside the device arithmetic */
		return vkhel_ntt_tables_create(n, q, w);
	}
	CUDA_CHECK(cudaSetDevice(ctx->dev.device));
	static const bool trace = getenv("VKHEL_TRACE_TABLES") != NULL;
	double t0 = tb_now(), t1;
#define TB_MARK(what) do { if (trace) { t1 = tb_now(); \
		fprintf(stderr, "tables_create_on: %-12s %.3f ms\n", what, t1 - t0); \
		t0 = t1; } } while (0)
	struct vkhel_ntt_tables *ntt = ntt_tables_alloc(n, q, w);
	TB_MARK("alloc");

	tabgen_params p;
	memset(&p, 0, sizeof(p));
	p.m = make_modulus(q);
	p.log2n = (unsigned) ntt->log2n;
	u64 b = w % q, bi = nt_inverse_mod(w % q, q);
	for (unsigned k = 0; k < p.log2n; k++) {
		p.pw[k] = b;
		p.ipw[k] = bi;
		b = nt_multiply_mod(b, b, q, 0);
		bi = nt_multiply_mod(bi, bi, q, 0);
	}

	TB_MARK("powers");
	const size_t pair_bytes = 2 * n * sizeof(ulonglong2);
	char *dev_buf = (char *) ntt_tables_mirror_alloc(ctx,
			sizeof(limb_desc) + mirror_pair_bytes(n));
	TB_MARK("mirror alloc");
	ulonglong2 *pairs = (ulonglong2 *) (dev_buf + sizeof(limb_desc));
	u64 *planar = (u64 *) device_alloc(ctx, pair_bytes);
	const unsigned blocks = (unsigned) ((n + TABGEN_THREADS - 1) / TABGEN_THREADS);
	tables_generate_kernel<<<blocks, TABGEN_THREADS, 0, ctx_stream(ctx)>>>(
			pairs, planar, p);
	CUDA_CHECK(cudaGetLastError());
	ctx->dev.launches++;
	TB_MARK("launch");

	/* host copies of the four arrays (the reference's struct exposes them) */
	u64 *host = (u64 *) pinned_acquire(ctx, pair_bytes);
	TB_MARK("pinned");
	CUDA_CHECK(cudaMemcpyAsync(host, planar, pair_bytes, cudaMemcpyDeviceToHost,
				ctx_stream(ctx)));
	device_free(ctx, planar);
	CUDA_CHECK(cudaStreamSynchronize(ctx_stream(ctx)));
	TB_MARK("kernel+d2h");
	memcpy(ntt->roots_of_unity, host, n * sizeof(u64));
	memcpy(ntt->inv_roots_of_unity, host + n, n * sizeof(u64));
	memcpy(ntt->roots_barrett_factors, host + 2 * n, n * sizeof(u64));
	memcpy(ntt->inv_roots_barrett_factors, host + 3 * n, n * sizeof(u64));
	pinned_release(ctx, host);
	TB_MARK("host copies");
	/* descriptor in front of the pairs; the buffer becomes this device's
	 * mirror of the tables */
	ntt_tables_adopt_mirror(ctx, ntt, dev_buf);
	TB_MARK("adopt");
	return ntt;
}
