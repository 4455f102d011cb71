/*
 * Whole negacyclic products of small polynomials, one CTA each, many per
 * launch: c = INTT(NTT(a) (*) NTT(b)) for 8 <= n <= 2^11.
 *
 * The reference writes a product as four calls -- forward(a), forward(b),
 * elemmul(a, b, c), inverse(c) (examples/example.c:18-60;
 * src/vector.c:388-427, 513-657) -- and at these sizes each of them is a kernel
 * launch that takes longer to start than to run.  vector.cu recognises the
 * four-call sequence, records it, and hands the recorded products of a loop to
 * this kernel as ONE launch: a CTA of n/8 threads loads a and b, runs both
 * forward transforms (whose results the caller can still observe, so they are
 * stored), multiplies, runs the inverse transform with n^-1 folded into its
 * last stage and stores c.  All stages of all three transforms stay in
 * registers and shared memory.
 *
 * Arithmetic and value ranges are those of the single-pass kernel
 * (kernels_ntt.cu, ntt_single_kernel): lazy Harvey butterflies, canonical
 * stores, hence results bit-identical to the reference's four calls.
 */
#include "ntt_device.cuh"

#define SMALL_INLINE 4

struct small_pass {
	const small_product *tab;          /* device table, or NULL: inl[] */
	small_product inl[SMALL_INLINE];
	const limb_desc *desc;             /* one modulus for the whole batch */
	u64 zero;
};

/* all K stages of one transform on the CTA's polynomial: radix-8 rounds with
 * exchanges through the CTA's padded buffer */
template <bool INV, int K, bool APX>
__device__ __forceinline__ void small_rounds(u64 (&x)[1][8], u64 *sm_x, int t,
		const ulonglong2 *tw, u64 q, u64 bq, ulonglong2 fold_a,
		ulonglong2 fold_b, u64 zr) {
	using G = tile_geom<K>;
	static_for<0, G::rounds>([&](auto rrc) {
		constexpr int rr = decltype(rrc)::value;
		constexpr int r = INV ? G::rounds - 1 - rr : rr;
		if constexpr (rr > 0) {
			constexpr int prev = INV ? r + 1 : r - 1;
			u64 *xw = sm_x + xpad(G::tbase(prev, t));
#pragma unroll
			for (int e = 0; e < 8; e++) {
				xw[xpad(G::eoff(prev, e))] = x[0][e];
			}
			__syncthreads();
			const u64 *xr = sm_x + xpad(G::tbase(r, t));
#pragma unroll
			for (int e = 0; e < 8; e++) {
				x[0][e] = xr[xpad(G::eoff(r, e))];
			}
		}
		tile_round<K, INV, INV ? FOLD_LAST : FOLD_NONE, 1, APX>(x, r, t, tw, q, bq,
				fold_a, fold_b, nullptr, zr);
	});
}

template <int K, bool APX>
__global__ void __launch_bounds__(K > 3 ? (1 << (K - 3)) : 1)
ntt_small_product_kernel(const __grid_constant__ small_pass p) {
	using G = tile_geom<K>;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	ulonglong2 *sm_twf = (ulonglong2 *) smem_raw;          /* [2^K] */
	ulonglong2 *sm_twi = sm_twf + (1 << K);                /* [2^K] */
	u64 *sm_x = (u64 *) (sm_twi + (1 << K));               /* xpad(2^K) words */
	__shared__ __align__(8) u64 tw_bar;

	const limb_desc &d = *p.desc;
	const u64 q = d.q, bq = APX ? 3 * q : 2 * q;
	const int t = threadIdx.x;
	if (t == 0) {
		mbar_init(&tw_bar, 2);
	}
	__syncthreads();
	if (t == 0) {
		stage_twiddles_tma<K>(sm_twf, d.tw, 0, 0, 1, &tw_bar);
		stage_twiddles_tma<K>(sm_twi, d.tw + ((u64) 1 << K), 0, 0, 1, &tw_bar);
	}
	pdl_wait();
	const small_product ent = p.tab ? p.tab[blockIdx.x] : p.inl[blockIdx.x];
	const ulonglong2 fold_a = make_ulonglong2(d.inv_n, d.inv_n_shoup);
	const ulonglong2 fold_b = make_ulonglong2(d.inv_w1n, d.inv_w1n_shoup);
	const ulonglong2 none = make_ulonglong2(0, 0);
	modulus m;
	m.q = q;
	m.d = d.mm_d;
	m.v = d.mm_v;
	m.s = d.mm_s;
	m.mu = 0;
	constexpr int lastf = G::rounds - 1;       /* layout the forward ends in */
	const int tb0 = G::tbase(0, t), tbl = G::tbase(lastf, t);

	u64 xa[1][8], x[1][8];
#pragma unroll
	for (int e = 0; e < 8; e++) {
		xa[0][e] = ent.a_src[tb0 + G::eoff(0, e)];
	}
#pragma unroll
	for (int e = 0; e < 8; e++) {
		x[0][e] = ent.b_src[tb0 + G::eoff(0, e)];
	}
	mbar_wait(&tw_bar, 0);
	small_rounds<false, K, APX>(xa, sm_x, t, sm_twf, q, bq, none, none, p.zero);
#pragma unroll
	for (int e = 0; e < 8; e++) {
		xa[0][e] = tile_canon<false, APX>(xa[0][e], q, bq);
		ent.a_dst[tbl + G::eoff(lastf, e)] = xa[0][e];
	}
	__syncthreads();   /* the exchange buffer changes hands */
	small_rounds<false, K, APX>(x, sm_x, t, sm_twf, q, bq, none, none, p.zero);
#pragma unroll
	for (int e = 0; e < 8; e++) {
		const u64 cb = tile_canon<false, APX>(x[0][e], q, bq);
		ent.b_dst[tbl + G::eoff(lastf, e)] = cb;
		/* both factors canonical: the reference's elemmul (elemmul.comp:62-73) */
		x[0][e] = mulmod(xa[0][e], cb, m);
	}
	__syncthreads();
	small_rounds<true, K, APX>(x, sm_x, t, sm_twi, q, bq, fold_a, fold_b, p.zero);
	pdl_launch_dependents();
#pragma unroll
	for (int e = 0; e < 8; e++) {
		ent.c[tb0 + G::eoff(0, e)] = tile_canon<true, APX>(x[0][e], q, bq);
	}
}

template <int K, bool APX>
static void run_small(struct vkhel_ctx *ctx, const small_pass &p, unsigned count) {
	const size_t smem = 2 * (sizeof(ulonglong2) << K)
		+ sizeof(u64) * (size_t) (xpad(1 << K) + 4);
	auto kernel = ntt_small_product_kernel<K, APX>;
	if (smem + 256 > 48 * 1024) {
		CUDA_CHECK(cudaFuncSetAttribute(kernel,
					cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	}
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(count);
	cfg.blockDim = dim3(K > 3 ? (1u << (K - 3)) : 1u);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = ctx_stream(ctx);
	cudaLaunchAttribute attr;
	attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr.val.programmaticStreamSerializationAllowed = FAST_PDL;
	cfg.attrs = &attr;
	cfg.numAttrs = 1;
	CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, p));
	ctx->dev.launches++;
}

bool ntt_small_product_supported(unsigned log2n, uint64_t q) {
	static const bool off = getenv("VKHEL_FORCE_GENERIC") != NULL
		|| getenv("VKHEL_NO_SMALL_PRODUCT") != NULL;
	return !off && log2n >= 3 && log2n <= SMALL_PRODUCT_MAX_LOG2N
		&& q < (1ull << 62);
}

void launch_ntt_small_products(struct vkhel_ctx *ctx, const small_product *tab,
		const small_product *host_tab, unsigned count, const limb_desc *desc,
		unsigned log2n, uint64_t q) {
	VK_REQUIRE(ntt_small_product_supported(log2n, q),
			"internal: small product of 2^%u points", log2n);
	VK_REQUIRE((tab != NULL) != (host_tab != NULL) && count >= 1,
			"internal: small products need exactly one pointer table");
	small_pass p;
	p.tab = tab;
	p.desc = desc;
	p.zero = 0;
	if (host_tab) {
		VK_REQUIRE(count <= SMALL_INLINE, "internal: inline batch too long");
		for (unsigned i = 0; i < count; i++) {
			p.inl[i] = host_tab[i];
		}
	}
	/* the approximate-quotient butterflies need 6q < 2^64 */
	const bool apx = q < 0xffffffffffffffffull / 6
		&& getenv("VKHEL_EXACT_QUOTIENT") == NULL;
#define SMALL_CASE(K_) \
	case K_: \
		if (apx) run_small<K_, true>(ctx, p, count); \
		else run_small<K_, false>(ctx, p, count); \
		break;
	switch (log2n) {
	SMALL_CASE(3) SMALL_CASE(4) SMALL_CASE(5) SMALL_CASE(6) SMALL_CASE(7)
	SMALL_CASE(8) SMALL_CASE(9) SMALL_CASE(10) SMALL_CASE(11)
	default: VK_DIE("internal: small product of 2^%u points", log2n);
	}
#undef SMALL_CASE
}
