/*
 * Negacyclic NTT kernels for sm_100a.
 *
 * Replaces the reference's per-group butterfly dispatches
 * (src/vector.c:536-566 forward, :599-639 inverse + n^-1 scaling;
 * shaders nttfwdbutterfly.comp:31-58, nttrevbutterfly.comp:31-58,
 * elemmulconst.comp:35-49).
 *
 * Index facts used throughout (derived from the reference's loops): with
 * L = log2 n, stage s (m = 2^s groups, t = n / 2^(s+1)) pairs elements j and
 * j + t, i.e. elements that differ in bit L-1-s, and uses twiddle number
 *     2^s + (j >> (L - s))
 * of the bit-reversed table.  Forward runs s = 0..L-1, inverse s = L-1..0 with
 * the inverse table and the same numbering.  A run of k stages [s0, s0+k)
 * therefore decomposes into independent 2^k-point "tiles": the elements that
 * share the s0 high bits H and the L-s0-k low bits; inside a tile, local stage
 * u and local group g use twiddle ((2^s0 + H) << u) + g -- the tile is the
 * subtree of the twiddle heap rooted at node 2^s0 + H.
 *
 * Value ranges (q < 2^62, Harvey lazy butterflies): forward values live in
 * [0,4q) and inverse values in [0,2q) between stages and between passes; the
 * pass that contains the last stage stores canonical residues, which is what
 * the reference stores after every stage -- hence bit-identical output.  For
 * 2^62 <= q < 2^63 the strict butterflies keep every value canonical.  The
 * n^-1 scaling of the inverse costs no pass of its own: the column pass takes
 * it from a scaled copy of the top of the inverse twiddle heap (FOLD_TWID,
 * ntt_engine.cuh), the single-pass, row-only and generic kernels fold it into
 * their last stage (twiddles n^-1 and inv_root[1] * n^-1, FOLD_LAST).
 */
#include <string.h>
#include "ntt_device.cuh"

/* ======================================================================================
 * Generic path: any n >= 2, any batch.  One CTA stages GEN_ELEMS coefficients
 * in shared memory (G tiles x K points x C columns), runs the k stages of the
 * pass with one __syncthreads per stage, and writes back.  Passes that do not
 * contain the last k stages read C adjacent columns per row (C*8 contiguous
 * bytes), the final pass reads G whole tiles back to back.
 * ====================================================================================== */
#define GEN_THREADS 256
#define GEN_LOG2_ELEMS 12
#define GEN_ELEMS (1u << GEN_LOG2_ELEMS)

struct gen_pass {
	const u64 *src;
	u64 *dst;
	const limb_desc *descs;
	unsigned limbs;
	unsigned log2n;
	unsigned s0;     /* first stage of the pass */
	unsigned k;      /* number of stages; tile size K = 2^k */
	unsigned log2c;  /* adjacent columns per CTA */
	unsigned log2g;  /* tiles per CTA (consecutive in (poly, H)) */
	u64 tiles;       /* polys << s0 */
	bool last;       /* pass stores canonical residues */
};

template <bool INVERSE, bool STRICT>
__global__ void __launch_bounds__(GEN_THREADS)
ntt_generic_kernel(const gen_pass p) {
	extern __shared__ u64 sm[];

	const unsigned L = p.log2n, s0 = p.s0, k = p.k;
	const unsigned s1 = s0 + k;
	const unsigned low_bits = L - s1;             /* bits below the tile */
	const unsigned log2c = p.log2c, log2g = p.log2g;
	const unsigned col_groups_log2 = low_bits - log2c;
	const unsigned elems_log2 = log2g + k + log2c;
	const unsigned elems = 1u << elems_log2;

	/* blockIdx.x -> (first tile, column group) */
	const u64 block = blockIdx.x;
	const u64 tile0 = (block >> col_groups_log2) << log2g;
	const u64 col0 = (block & ((1ull << col_groups_log2) - 1)) << log2c;

	/* element e of the CTA: (t, mid, c) = (tile, point, column) */
	auto global_index = [&](unsigned e, u64 &tile) -> u64 {
		const unsigned c = e & ((1u << log2c) - 1);
		const unsigned mid = (e >> log2c) & ((1u << k) - 1);
		const unsigned t = e >> (log2c + k);
		tile = tile0 + t;
		const u64 poly = tile >> s0;
		const u64 H = tile & ((1ull << s0) - 1);
		return (poly << L) | (H << (L - s0)) | ((u64) mid << low_bits)
			| (col0 + c);
	};

	for (unsigned e = threadIdx.x; e < elems; e += GEN_THREADS) {
		u64 tile;
		const u64 j = global_index(e, tile);
		sm[e] = tile < p.tiles ? p.src[j] : 0;
	}
	__syncthreads();

	const unsigned bf = elems >> 1;
	for (unsigned step = 0; step < k; step++) {
		const unsigned u = INVERSE ? k - 1 - step : step;
		const unsigned half_log2 = k - 1 - u;     /* log2 of the pair stride */
		for (unsigned x = threadIdx.x; x < bf; x += GEN_THREADS) {
			const unsigned c = x & ((1u << log2c) - 1);
			const unsigned b = (x >> log2c) & ((1u << (k - 1)) - 1);
			const unsigned t = x >> (log2c + k - 1);
			const unsigned g = b >> half_log2;
			const unsigned pos = b & ((1u << half_log2) - 1);
			const unsigned mid0 = (g << (half_log2 + 1)) | pos;
			const unsigned i0 = (((t << k) | mid0) << log2c) | c;
			const unsigned i1 = i0 + (1u << (half_log2 + log2c));

			const u64 tile = tile0 + t;
			if (tile >= p.tiles) {
				continue;
			}
			const u64 poly = tile >> s0;
			const u64 H = tile & ((1ull << s0) - 1);
			const limb_desc &d = p.descs[poly % p.limbs];
			const u64 q = d.q;
			const u64 node = ((((u64) 1 << s0) | H) << u) + g;

			u64 X = sm[i0], Y = sm[i1];
			if (INVERSE && s0 == 0 && u == 0) {
				/* last inverse stage with n^-1 folded in */
				if (STRICT) {
					const u64 s = csub(X + Y, q);
					const u64 df = X >= Y ? X - Y : X - Y + q;
					X = shoup_canon(s, d.inv_n, d.inv_n_shoup, q);
					Y = shoup_canon(df, d.inv_w1n, d.inv_w1n_shoup, q);
				} else {
					const u64 s = X + Y;
					const u64 df = X - Y + 2 * q;
					X = shoup_lazy(s, d.inv_n, d.inv_n_shoup, q);
					Y = shoup_lazy(df, d.inv_w1n, d.inv_w1n_shoup, q);
				}
			} else {
				const ulonglong2 w = d.tw[(INVERSE ? ((u64) 1 << L) : 0) + node];
				if (INVERSE) {
					if (STRICT) gs_strict(X, Y, w.x, w.y, q);
					else gs_lazy(X, Y, w.x, w.y, q, 2 * q);
				} else {
					if (STRICT) ct_strict(X, Y, w.x, w.y, q);
					else ct_lazy(X, Y, w.x, w.y, q, 2 * q);
				}
			}
			sm[i0] = X;
			sm[i1] = Y;
		}
		__syncthreads();
	}

	for (unsigned e = threadIdx.x; e < elems; e += GEN_THREADS) {
		u64 tile;
		const u64 j = global_index(e, tile);
		if (tile >= p.tiles) {
			continue;
		}
		u64 v = sm[e];
		if (p.last && !STRICT) {
			const u64 q = p.descs[(tile >> s0) % p.limbs].q;
			if (!INVERSE) {
				v = csub(v, 2 * q);
			}
			v = csub(v, q);
		}
		p.dst[j] = v;
	}
}

template <bool INVERSE, bool STRICT>
static void run_generic_pass(struct vkhel_ctx *ctx, gen_pass p) {
	const unsigned low_bits = p.log2n - p.s0 - p.k;
	/* fill the CTA's GEN_ELEMS budget: columns first, then extra tiles when
	 * the tile already spans whole rows */
	unsigned log2c = GEN_LOG2_ELEMS > p.k ? GEN_LOG2_ELEMS - p.k : 0;
	if (log2c > low_bits) {
		log2c = low_bits;
	}
	unsigned log2g = 0;
	if (low_bits == log2c) {
		log2g = GEN_LOG2_ELEMS > p.k + log2c ? GEN_LOG2_ELEMS - p.k - log2c : 0;
		while (log2g > 0 && (1ull << (log2g - 1)) >= p.tiles) {
			log2g--;
		}
	}
	p.log2c = log2c;
	p.log2g = log2g;
	const u64 tile_groups = (p.tiles + (1ull << log2g) - 1) >> log2g;
	const u64 blocks = tile_groups << (low_bits - log2c);
	VK_REQUIRE(blocks <= 0x7fffffffull, "transform too large for one launch");
	const size_t smem = sizeof(u64) << (log2g + p.k + log2c);
	if (smem > 48 * 1024) {
		/* the leading pass of n = 2^29, 2^30 on the fast path holds 2^13 / 2^14
		 * points per tile: above the default dynamic shared memory limit */
		VK_REQUIRE(smem <= ctx->dev.smem_optin,
				"transform pass of 2^%u points does not fit in shared memory",
				p.k);
		CUDA_CHECK(cudaFuncSetAttribute(ntt_generic_kernel<INVERSE, STRICT>,
					cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	}
	ntt_generic_kernel<INVERSE, STRICT>
		<<<(unsigned) blocks, GEN_THREADS, smem, ctx_stream(ctx)>>>(p);
	CUDA_CHECK(cudaGetLastError());
	ctx->dev.launches++;
}

template <bool INVERSE, bool STRICT>
static void run_generic(struct vkhel_ctx *ctx, const u64 *src, u64 *dst,
		const limb_desc *descs, uint64_t limbs, uint64_t polys,
		unsigned log2n) {
	/* split the log2n stages into passes: the pass over the deepest stages
	 * takes up to GEN_LOG2_ELEMS of them, the rest goes into strided passes
	 * of at most 8 stages each */
	unsigned ks[16];
	unsigned npass = 0;
	const unsigned kf = log2n < GEN_LOG2_ELEMS ? log2n : GEN_LOG2_ELEMS;
	unsigned rest = log2n - kf;
	const unsigned nstrided = (rest + 7) / 8;
	for (unsigned i = 0; i < nstrided; i++) {
		const unsigned take = (rest + (nstrided - i) - 1) / (nstrided - i);
		ks[npass++] = take;
		rest -= take;
	}
	ks[npass++] = kf;

	gen_pass p;
	p.descs = descs;
	p.limbs = (unsigned) limbs;
	p.log2n = log2n;
	for (unsigned i = 0; i < npass; i++) {
		/* forward walks the passes top-down, inverse bottom-up */
		const unsigned idx = INVERSE ? npass - 1 - i : i;
		unsigned s0 = 0;
		for (unsigned j = 0; j < idx; j++) {
			s0 += ks[j];
		}
		p.src = i == 0 ? src : dst;
		p.dst = dst;
		p.s0 = s0;
		p.k = ks[idx];
		p.tiles = polys << s0;
		p.last = i == npass - 1;
		run_generic_pass<INVERSE, STRICT>(ctx, p);
	}
}

/* ======================================================================================
 * Fast path (q < 2^62, n >= 8): register radix-8 engine (ntt_engine.cuh).
 *
 *   row pass    : the deepest K_row stages; a tile is 2^K_row CONTIGUOUS
 *                 coefficients.  One warp-group of 2^(K_row-3) lanes per tile:
 *                 global -> registers -> (shared-memory exchanges inside the
 *                 warp, __syncwarp only) -> global.  A CTA stages the twiddle
 *                 subtrees of its tiles in shared memory once and reuses them
 *                 over the batch.
 *   column pass : the K_col stages above them; a tile is 2^K_col coefficients
 *                 at stride 2^(L-s0-K_col).  A CTA takes C adjacent columns;
 *                 thread = (column, row group), lanes along the columns, so
 *                 every global and shared access is contiguous.
 *
 * Forward: column pass (stages 0..) then row pass; inverse: row pass then
 * column pass, which also applies the n^-1 scaling (through its twiddles).
 * Values between the passes stay lazy ([0,4q) forward, [0,2q) inverse; [0,6q)
 * and [0,3q) with the approximate quotient) in the result vector.
 * ====================================================================================== */
#define FAST_THREADS 256
/* Tuning knobs: tiles per thread (NP) and occupancy targets.
 * Measured on B200 (tools/build_variant.sh + tools/bench_variants.sh, n = 2^16
 * x 512): the row pass is fastest with one tile per thread and 6 CTAs per SM
 * (40 registers, a few spill slots); the column pass with two adjacent columns
 * per thread (128-bit accesses, each twiddle fetch feeding two butterflies) and
 * 4 CTAs per SM (64 registers). */
#ifndef ROWS_NP
#define ROWS_NP 1
#endif
#ifndef COLS_NP
#define COLS_NP 2
#endif
#ifndef ROWS_MIN_CTAS_NP1
#define ROWS_MIN_CTAS_NP1 6
#endif
#ifndef ROWS_MIN_CTAS_NP2
#define ROWS_MIN_CTAS_NP2 4
#endif
#ifndef COLS_MIN_THREADS_NP1
#define COLS_MIN_THREADS_NP1 1280
#endif
#ifndef COLS_MIN_THREADS_NP2
#define COLS_MIN_THREADS_NP2 1024
#endif
#ifndef ROWS_ROUNDS_PER_CTA
#define ROWS_ROUNDS_PER_CTA 2
#endif
#ifndef ROWS_PREFETCH
#define ROWS_PREFETCH 1
#endif
#ifndef VKHEL_DEFAULT_SLICE_MIB
#define VKHEL_DEFAULT_SLICE_MIB 32
#endif
#ifndef FAST_PDL_EARLY
#define FAST_PDL_EARLY 0
#endif
/* three-input additions with a zero only the host knows keep the butterflies'
 * additions off the fmaheavy pipe (modarith.cuh, ct_lazy); 0 = plain sums */
#ifndef FAST_OPAQUE_ZERO
#define FAST_OPAQUE_ZERO 0
#endif
/* Stores of two adjacent 64-bit values as two 64-bit stores instead of one
 * 128-bit store: a 128-bit store wants its four source registers in one
 * aligned quad, and ptxas assembles that quad with up to three register moves
 * per store (58-169 IMAD.MOV per column kernel, on the fmaheavy pipe); a pair
 * of 64-bit stores takes the values where the butterflies left them.
 * XST: shared-memory exchange stores, GST: global stores. */
#ifndef COLS_XST64
#define COLS_XST64 0
#endif
#ifndef COLS_GST64
#define COLS_GST64 0
#endif
#ifndef ROWS_XST64
#define ROWS_XST64 0
#endif
#ifndef ROWS_GST64
#define ROWS_GST64 0
#endif
#if FAST_OPAQUE_ZERO
#define OPAQUE_ZERO(p) ((p).zero)
#else
#define OPAQUE_ZERO(p) ((u64) 0)
#endif

struct fast_pass {
	const u64 *src;
	const u64 *src2;      /* row pass, MUL: second operand of the fused point-wise op */
	u64 fma_mult;         /* MUL: 0 = product src*src2, else src*fma_mult + src2 */
	bool mul;             /* the row pass of this transform multiplies while it loads */
	bool lazy;            /* forward: the row pass stores [0,bq), see LAZY below */
	u64 *dst;
	const limb_desc *descs;
	unsigned limbs;
	unsigned log2n;
	unsigned s0;          /* first stage of the pass */
	u64 polys;            /* batch * limbs */
	unsigned hgroup_log2; /* row pass: consecutive H per CTA */
	unsigned bchunk;      /* row pass: batch entries per CTA */
	/* indirect batch (kernels instantiated with IND): polynomial i lives at
	 * tab[i].src / tab[i].dst instead of src + i*n / dst + i*n; the second
	 * pass of a transform reads what the first one wrote (tab[i].dst) */
	const ntt_ptrs *tab;
	unsigned tab_second;
	/* a short indirect batch carries its pointers in the kernel parameters
	 * (tab_inline of them, tab == NULL): no pointer-table copy to the device */
	unsigned tab_inline;
	ntt_ptrs inl[NTT_INLINE_PTRS];
	/* limb slice: this launch covers limbs [limb0, limb0 + limbs) of vectors
	 * laid out [batch][limbs_total][n]; descs already points at limb0.  The
	 * whole vector: limb0 = 0, limbs_total = limbs. */
	unsigned limbs_total, limb0;
	/* always 0, but only the host knows: three-input additions with it stay
	 * on the ALU pipe (modarith.cuh, ct_lazy) */
	u64 zero;

	__host__ __device__ bool indirect() const {
		return tab != NULL || tab_inline != 0;
	}
	/* pointers of polynomial i of an indirect batch; p is a __grid_constant__
	 * kernel parameter, so the inline table is read from constant memory */
	__device__ __forceinline__ ntt_ptrs entry(u64 i) const {
		if (tab_inline) {
			return inl[i];
		}
		return tab[i];
	}
};

/* The point-wise operation fused into the load of an inverse transform's first
 * pass: the reference's elemmul (elemmul.comp:62-73) or elemfma (contract
 * (a*mult + b) mod q) between the forward and the inverse
 * transform (src/vector.c:298-340,388-427).  Exact for ANY 64-bit operands,
 * like the stand-alone kernels (kernels_elem.cu): the result is the canonical
 * residue those would have stored.  `fma_mult` = 0 selects the product,
 * otherwise x*fma_mult + y with fma_mult already reduced mod q (a multiplier
 * that is 0 mod q is passed as q: a*q + b = b mod q). */
__device__ __forceinline__ u64 fused_pointwise(u64 x, u64 y, u64 fma_mult,
		const modulus &m) {
	if (fma_mult) {
		/* the high word of x*mult is at most mult - 1 <= q - 1 and the carry
		 * of the addition raises it to at most q: one conditional reduction */
		const u64 lo = x * fma_mult;
		const u64 sum = lo + y;
		u64 hi = __umul64hi(x, fma_mult) + (sum < lo);
		if (hi >= m.q) {
			hi -= m.q;
		}
		return reduce128(hi, sum, m);
	}
	/* (x mod q)(y mod q) mod q: the high word is reduced first when it is not
	 * already below q -- a branch canonical factors never take */
	u64 hi = __umul64hi(x, y);
	if (hi >= m.q) {
		hi = reduce128(0, hi, m);
	}
	return reduce128(hi, x * y, m);
}

/* the stream the fast kernels go to: the context's, or the auxiliary one
 * while launch_ntt alternates slices between the two */
static inline cudaStream_t launch_stream(struct vkhel_ctx *ctx) {
	return ctx->dev.launch_stream ? (cudaStream_t) ctx->dev.launch_stream
		: ctx_stream(ctx);
}

/* Dynamic shared memory above the default 48 KiB limit needs the per-kernel
 * opt-in; the kernels' few static bytes (their mbarriers) count towards the
 * same limit, hence the margin. */
static inline bool smem_needs_optin(size_t dynamic_bytes) {
	return dynamic_bytes + 256 > 48 * 1024;
}

template <class... KArgs, class... Args>
static void launch_fast(struct vkhel_ctx *ctx, void (*kernel)(KArgs...),
		unsigned grid, unsigned block, size_t smem, Args... args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(grid);
	cfg.blockDim = dim3(block);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = launch_stream(ctx);
	cudaLaunchAttribute attr;
	attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr.val.programmaticStreamSerializationAllowed = FAST_PDL;
	cfg.attrs = &attr;
	cfg.numAttrs = 1;
	CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, args...));
	ctx->dev.launches++;
}

/* launch_fast with the shared-memory opt-in where the kernel needs it (per
 * device, and cheap: set on every such launch) */
static void launch_fast_optin(struct vkhel_ctx *ctx, void (*kernel)(fast_pass),
		unsigned grid, unsigned block, size_t smem, const fast_pass &p) {
	if (smem_needs_optin(smem)) {
		CUDA_CHECK(cudaFuncSetAttribute(kernel,
					cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	}
	launch_fast(ctx, kernel, grid, block, smem, p);
}

template <int K>
struct row_cfg {
	static constexpr int tile = 1 << K;
	static constexpr int xbuf = tile + ((tile >> 5) << 2) + 4; /* words per tile */
	static constexpr int group = 1 << (K - 3);                 /* lanes per tile */
	static constexpr int groups_per_warp = 32 / group;
	static constexpr int groups_per_cta = (FAST_THREADS / 32) * groups_per_warp;
};

/* Row pass.  CTA = (batch chunk, limb, H group): `bchunk` batch entries of
 * 2^hgroup_log2 consecutive tiles sharing one staged twiddle set.  A warp-group
 * of 2^(K-3) lanes carries NP batch entries of one tile position at a time. */
/* TOP (inverse only): the pass holds stage 0 (s0 == 0, i.e. the row pass is the
 * whole transform) and applies n^-1 in it.  A template parameter rather than
 * a run-time branch per round: with both variants in one kernel ptxas merged
 * their register assignments after every round with some 20 moves each. */
/* LAZY (forward only): the outputs are stored after the first of the three
 * (exact quotient: two) conditional subtractions, i.e. in [0,bq) -- the range
 * the inverse transform's butterflies accept -- instead of canonical.  Only for
 * a forward transform whose very next use is the in-place inverse transform
 * with the same tables (vector.cu, "held forward transform"): the lazy values
 * are overwritten before anything else can read them. */
template <bool INV, int K, int NP, bool MUL, bool APX, bool IND, bool TOP,
	bool LAZY = false>
__global__ void __launch_bounds__(FAST_THREADS,
		NP == 2 ? ROWS_MIN_CTAS_NP2 : ROWS_MIN_CTAS_NP1)
ntt_rows_kernel(const __grid_constant__ fast_pass p) {
	static_assert(INV || !TOP, "TOP distinguishes inverse passes only");
	static_assert(!LAZY || (!INV && !IND), "LAZY: direct forward passes only");
	using G = tile_geom<K>;
	using C = row_cfg<K>;
	extern __shared__ __align__(16) unsigned char smem_raw[];

	const unsigned L = p.log2n, s0 = p.s0;       /* s0 + K == L */
	const unsigned hgroup = 1u << p.hgroup_log2;
	ulonglong2 *sm_tw = (ulonglong2 *) smem_raw;                 /* [hgroup][2^K] */
	u64 *sm_x = (u64 *) (sm_tw + ((size_t) hgroup << K));        /* per group, NP tiles */

	/* blockIdx.x -> (batch chunk, limb, H group) */
	const unsigned hgroups = (1u << s0) >> p.hgroup_log2;
	unsigned blk = blockIdx.x;
	const unsigned hg = blk % hgroups;
	blk /= hgroups;
	const unsigned limb = blk % p.limbs;
	const unsigned bc = blk / p.limbs;
	const u64 batch = p.polys / p.limbs;
	const u64 b0 = (u64) bc * p.bchunk;
	const unsigned nb = (unsigned) (batch - b0 < p.bchunk ? batch - b0 : p.bchunk);
	const unsigned H0 = hg << p.hgroup_log2;

	const limb_desc &d = p.descs[limb];
	const u64 q = d.q, bq = APX ? 3 * q : 2 * q;   /* butterfly bound */
	__shared__ __align__(8) u64 tw_bar;
	if (threadIdx.x == 0) {
		mbar_init(&tw_bar, 1);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		stage_twiddles_tma<K>(sm_tw, d.tw + (INV ? ((u64) 1 << L) : 0), s0, H0,
				hgroup, &tw_bar);
	}
	bool tw_ready = false;
	pdl_wait();   /* the coefficients come from the previous kernel */
#if FAST_PDL_EARLY
	pdl_launch_dependents();
#endif

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int t = lane & (C::group - 1);                 /* thread within group */
	const int slot = warp * C::groups_per_warp + (lane >> (K - 3));
	u64 *xb = sm_x + (size_t) slot * (NP * C::xbuf);
	constexpr int first = INV ? G::rounds - 1 : 0;
	constexpr int last = INV ? 0 : G::rounds - 1;
	constexpr bool fold = INV && TOP;
	constexpr bool canon = INV ? TOP : true;     /* pass with the final stage */
	ulonglong2 fold_a = make_ulonglong2(0, 0), fold_b = fold_a;
	if (fold) {
		fold_a = make_ulonglong2(d.inv_n, d.inv_n_shoup);
		fold_b = make_ulonglong2(d.inv_w1n, d.inv_w1n_shoup);
	}
	/* per-thread bases; every access below is base + compile-time constant */
	const int tb_first = G::tbase(first, t), tb_last = G::tbase(last, t);

	/* work items of the CTA: (group of NP batch entries, h), h fastest; slot
	 * `slot` takes every groups_per_cta-th one.  Uniform loop count. */
	const unsigned nitems = ((nb + NP - 1) / NP) << p.hgroup_log2;
	for (unsigned base = 0; base < nitems; base += C::groups_per_cta) {
		const unsigned idx = base + slot;
		const unsigned h = idx & (hgroup - 1), bg = idx >> p.hgroup_log2;
		const ulonglong2 *twt = sm_tw + ((size_t) h << K);
		bool active[NP];
		u64 off[NP];
		u64 *dbase[NP];   /* IND only */
		u64 x[NP][8];
#if ROWS_PREFETCH
		/* pull the tile this slot takes in the next iteration towards the
		 * SM while this one is being computed: one 128-byte line per lane
		 * of the group, no registers held */
		if (!IND) {
			const unsigned nidx = idx + C::groups_per_cta;
			const unsigned nh = nidx & (hgroup - 1);
			const unsigned nbl = (nidx >> p.hgroup_log2) * NP;
			if (nidx < nitems && nbl < nb && (t << 4) < (1 << K)) {
				const u64 npoly = (b0 + nbl) * p.limbs_total + p.limb0 + limb;
				const u64 *np_ = p.src + (npoly << L) + ((u64) (H0 + nh) << K)
					+ (t << 4);
				asm volatile("prefetch.global.L2 [%0];" :: "l"(np_));
			}
		}
#endif
#pragma unroll
		for (int pp = 0; pp < NP; pp++) {
			const unsigned bl = bg * NP + pp;
			active[pp] = idx < nitems && bl < nb;
			const u64 poly = (b0 + bl) * p.limbs_total + p.limb0 + limb;
			const u64 *sp;
			const u64 *src2base = p.src2;   /* MUL only */
			if (IND) {
				off[pp] = (u64) (H0 + h) << K;
				ntt_ptrs ent = { NULL, NULL, NULL };
				if (active[pp]) {
					ent = p.entry(poly);
				}
				dbase[pp] = ent.dst;
				src2base = ent.src2;
				sp = (p.tab_second ? ent.dst : ent.src) + off[pp] + tb_first;
			} else {
				off[pp] = (poly << L) + ((u64) (H0 + h) << K);
				sp = p.src + off[pp] + tb_first;
			}
			if (G::eoff(first, 1) == 1) {
				/* registers (e, e+1) are adjacent coefficients (the layout of
				 * the deepest round): 128-bit loads */
#pragma unroll
				for (int e = 0; e < 8; e += 2) {
					ulonglong2 v = make_ulonglong2(0, 0);
					if (active[pp]) {
						v = *(const ulonglong2 *) (sp + G::eoff(first, e));
					}
					x[pp][e] = v.x;
					x[pp][e + 1] = v.y;
				}
			} else {
#pragma unroll
				for (int e = 0; e < 8; e++) {
					x[pp][e] = active[pp] ? sp[G::eoff(first, e)] : 0;
				}
			}
			if (MUL) {
				/* fused point-wise product (the reference's elemmul between
				 * the forward and the inverse transform, src/vector.c:388-427) */
				modulus m;
				m.q = q;
				m.d = d.mm_d;
				m.v = d.mm_v;
				m.s = d.mm_s;
				m.mu = 0;
				const u64 *sp2 = src2base + off[pp] + tb_first;
#pragma unroll
				for (int e = 0; e < 8; e++) {
					const u64 y = active[pp] ? sp2[G::eoff(first, e)] : 0;
					x[pp][e] = fused_pointwise(x[pp][e], y, p.fma_mult, m);
				}
			}
		}

		if (!tw_ready) {
			mbar_wait(&tw_bar, 0);   /* twiddles have landed */
			tw_ready = true;
		}
		static_for<0, G::rounds>([&](auto rrc) {
			constexpr int rr = decltype(rrc)::value;
			constexpr int r = INV ? G::rounds - 1 - rr : rr;
			if constexpr (rr > 0) {
				/* redistribute: previous round's layout -> this round's */
				constexpr int prev = INV ? r + 1 : r - 1;
				/* where registers (e, e+1) are adjacent coefficients (the
				 * deepest round) the exchange moves 128 bits at a time: half
				 * the shared-memory instructions and half the bank conflicts
				 * of that layout's 32-byte lane stride */
				u64 *xw = xb + xpad(G::tbase(prev, t));
#pragma unroll
				for (int pp = 0; pp < NP; pp++) {
					if (ROWS_XST64) {
						const unsigned xa =
							(unsigned) __cvta_generic_to_shared(xw + pp * C::xbuf)
							+ (unsigned) OPAQUE_ZERO(p);
						static_for<0, 8>([&](auto ec) {
							constexpr int e = decltype(ec)::value;
							sts64_at<8 * xpad(G::eoff(prev, e))>(xa, x[pp][e]);
						});
					} else if (G::eoff(prev, 1) == 1) {
#pragma unroll
						for (int e = 0; e < 8; e += 2) {
							*(ulonglong2 *) (xw + pp * C::xbuf
									+ xpad(G::eoff(prev, e))) =
								make_ulonglong2(x[pp][e], x[pp][e + 1]);
						}
					} else {
#pragma unroll
						for (int e = 0; e < 8; e++) {
							xw[pp * C::xbuf + xpad(G::eoff(prev, e))] = x[pp][e];
						}
					}
				}
				__syncwarp();
				const u64 *xr = xb + xpad(G::tbase(r, t));
#pragma unroll
				for (int pp = 0; pp < NP; pp++) {
					if (G::eoff(r, 1) == 1) {
#pragma unroll
						for (int e = 0; e < 8; e += 2) {
							const ulonglong2 v = *(const ulonglong2 *) (xr
									+ pp * C::xbuf + xpad(G::eoff(r, e)));
							x[pp][e] = v.x;
							x[pp][e + 1] = v.y;
						}
					} else {
#pragma unroll
						for (int e = 0; e < 8; e++) {
							x[pp][e] = xr[pp * C::xbuf + xpad(G::eoff(r, e))];
						}
					}
				}
			}
			tile_round<K, INV, fold ? FOLD_LAST : FOLD_NONE, NP, APX>(x, r, t,
					twt, q, bq, fold_a, fold_b, nullptr, OPAQUE_ZERO(p));
		});
		__syncwarp();   /* the exchange buffer is reused by the next item */
		if (base + C::groups_per_cta >= nitems) {
			pdl_launch_dependents();   /* only this CTA's last stores remain */
		}

#pragma unroll
		for (int pp = 0; pp < NP; pp++) {
			if (active[pp]) {
				u64 *dp = (IND ? dbase[pp] : p.dst) + off[pp] + tb_last;
#pragma unroll
				for (int e = 0; e < 8; e++) {
					if (LAZY) {
						x[pp][e] = csub(x[pp][e], bq);   /* [0,2bq) -> [0,bq) */
					} else if (canon) {
						x[pp][e] = tile_canon<INV, APX>(x[pp][e], q, bq);
					}
				}
				if (G::eoff(last, 1) == 1 && !ROWS_GST64) {
					/* adjacent coefficients in (e, e+1): 128-bit stores */
#pragma unroll
					for (int e = 0; e < 8; e += 2) {
						*(ulonglong2 *) (dp + G::eoff(last, e)) =
							make_ulonglong2(x[pp][e], x[pp][e + 1]);
					}
				} else {
#pragma unroll
					for (int e = 0; e < 8; e++) {
						dp[G::eoff(last, e)] = x[pp][e];
					}
				}
			}
		}
	}
}

/* ---- fused negacyclic product: the middle of c = INTT(NTT(a) (*) NTT(b)) -------------
 * The last K stages of the forward transform and the first K stages of the
 * inverse transform work on the same 2^K contiguous coefficients, so for the
 * reference's call sequence forward, forward, elemmul, inverse
 * (src/vector.c:388-427,513-657) the row passes of both forward transforms, the
 * point-wise product and the row pass of the inverse transform are one kernel:
 * a tile of a and the matching tile of b are transformed in registers,
 * multiplied, and taken through the inverse stages before anything is stored.
 * Against the separate passes this saves the write + read of NTT(a), NTT(b)
 * and of the product (4 of the 13 vector sweeps of the unfused sequence). */
template <bool INV, int K, bool FOLD, bool APX>
__device__ __forceinline__ void rows_rounds(u64 (&x)[1][8], u64 *xb, int t,
		const ulonglong2 *twt, u64 q, u64 bq, ulonglong2 fold_a,
		ulonglong2 fold_b, u64 zr) {
	using G = tile_geom<K>;
#pragma unroll
	for (int rr = 0; rr < G::rounds; rr++) {
		const int r = INV ? G::rounds - 1 - rr : rr;
		if (rr > 0) {
			const int prev = INV ? r + 1 : r - 1;
			u64 *xw = xb + xpad(G::tbase(prev, t));
			if (G::eoff(prev, 1) == 1) {
#pragma unroll
				for (int e = 0; e < 8; e += 2) {
					*(ulonglong2 *) (xw + xpad(G::eoff(prev, e))) =
						make_ulonglong2(x[0][e], x[0][e + 1]);
				}
			} else {
#pragma unroll
				for (int e = 0; e < 8; e++) {
					xw[xpad(G::eoff(prev, e))] = x[0][e];
				}
			}
			__syncwarp();
			const u64 *xr = xb + xpad(G::tbase(r, t));
			if (G::eoff(r, 1) == 1) {
#pragma unroll
				for (int e = 0; e < 8; e += 2) {
					const ulonglong2 v =
						*(const ulonglong2 *) (xr + xpad(G::eoff(r, e)));
					x[0][e] = v.x;
					x[0][e + 1] = v.y;
				}
			} else {
#pragma unroll
				for (int e = 0; e < 8; e++) {
					x[0][e] = xr[xpad(G::eoff(r, e))];
				}
			}
		}
		tile_round<K, INV, FOLD ? FOLD_LAST : FOLD_NONE, 1, APX>(x, r, t, twt, q,
				bq, fold_a, fold_b, nullptr, zr);
	}
}

#ifndef PMUL_MIN_CTAS
#define PMUL_MIN_CTAS 4
#endif

template <int K, bool APX>
__global__ void __launch_bounds__(FAST_THREADS, PMUL_MIN_CTAS)
ntt_rows_polymul_kernel(const __grid_constant__ fast_pass p) {
	using G = tile_geom<K>;
	using C = row_cfg<K>;
	extern __shared__ __align__(16) unsigned char smem_raw[];

	const unsigned L = p.log2n, s0 = p.s0;       /* s0 + K == L */
	const unsigned hgroup = 1u << p.hgroup_log2;
	ulonglong2 *sm_twf = (ulonglong2 *) smem_raw;                /* [hgroup][2^K] */
	ulonglong2 *sm_twi = sm_twf + ((size_t) hgroup << K);        /* [hgroup][2^K] */
	u64 *sm_x = (u64 *) (sm_twi + ((size_t) hgroup << K));       /* per group */

	const unsigned hgroups = (1u << s0) >> p.hgroup_log2;
	unsigned blk = blockIdx.x;
	const unsigned hg = blk % hgroups;
	blk /= hgroups;
	const unsigned limb = blk % p.limbs;
	const unsigned bc = blk / p.limbs;
	const u64 batch = p.polys / p.limbs;
	const u64 b0 = (u64) bc * p.bchunk;
	const unsigned nb = (unsigned) (batch - b0 < p.bchunk ? batch - b0 : p.bchunk);
	const unsigned H0 = hg << p.hgroup_log2;

	const limb_desc &d = p.descs[limb];
	const u64 q = d.q, bq = APX ? 3 * q : 2 * q;
	__shared__ __align__(8) u64 tw_bar[2];
	if (threadIdx.x == 0) {
		mbar_init(&tw_bar[0], 1);
		mbar_init(&tw_bar[1], 1);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		stage_twiddles_tma<K>(sm_twf, d.tw, s0, H0, hgroup, &tw_bar[0]);
		stage_twiddles_tma<K>(sm_twi, d.tw + ((u64) 1 << L), s0, H0, hgroup,
				&tw_bar[1]);
	}
	bool tw_ready = false;
	pdl_wait();   /* the coefficients come from the previous kernels */

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int t = lane & (C::group - 1);
	const int slot = warp * C::groups_per_warp + (lane >> (K - 3));
	u64 *xb = sm_x + (size_t) slot * C::xbuf;
	const bool fold = s0 == 0;          /* single pass: the whole product here */
	ulonglong2 fold_a = make_ulonglong2(0, 0), fold_b = fold_a;
	if (fold) {
		fold_a = make_ulonglong2(d.inv_n, d.inv_n_shoup);
		fold_b = make_ulonglong2(d.inv_w1n, d.inv_w1n_shoup);
	}
	modulus m;
	m.q = q;
	m.d = d.mm_d;
	m.v = d.mm_v;
	m.s = d.mm_s;
	m.mu = 0;
	/* both directions enter and leave in the layout of round 0 */
	const int tb0 = G::tbase(0, t);

	const unsigned nitems = nb << p.hgroup_log2;
	for (unsigned base = 0; base < nitems; base += C::groups_per_cta) {
		const unsigned idx = base + slot;
		const unsigned h = idx & (hgroup - 1), bl = idx >> p.hgroup_log2;
		const bool active = idx < nitems;
		const u64 poly = (b0 + bl) * p.limbs + limb;
		const u64 off = (poly << L) + ((u64) (H0 + h) << K) + tb0;
		const ulonglong2 *twf = sm_twf + ((size_t) h << K);
		const ulonglong2 *twi = sm_twi + ((size_t) h << K);
		u64 xa[1][8], x[1][8];
#pragma unroll
		for (int e = 0; e < 8; e++) {
			xa[0][e] = active ? p.src[off + G::eoff(0, e)] : 0;
		}
#pragma unroll
		for (int e = 0; e < 8; e++) {
			x[0][e] = active ? p.src2[off + G::eoff(0, e)] : 0;
		}
		if (!tw_ready) {
			mbar_wait(&tw_bar[0], 0);
			mbar_wait(&tw_bar[1], 0);
			tw_ready = true;
		}
		rows_rounds<false, K, false, APX>(xa, xb, t, twf, q, bq, fold_a, fold_b,
				OPAQUE_ZERO(p));
		__syncwarp();   /* the exchange buffer changes hands */
		rows_rounds<false, K, false, APX>(x, xb, t, twf, q, bq, fold_a, fold_b,
				OPAQUE_ZERO(p));
		/* point-wise product (reference elemmul.comp:62-73): one factor is
		 * made canonical, the other stays lazy (below 2*bq), so the high word
		 * of the product is below q as reduce128 requires; the result is the
		 * canonical residue the reference's elemmul would have stored */
#pragma unroll
		for (int e = 0; e < 8; e++) {
			const u64 ca = tile_canon<false, APX>(xa[0][e], q, bq);
			x[0][e] = mulmod(ca, x[0][e], m);
		}
		if (fold) {
			rows_rounds<true, K, true, APX>(x, xb, t, twi, q, bq, fold_a, fold_b,
					OPAQUE_ZERO(p));
		} else {
			rows_rounds<true, K, false, APX>(x, xb, t, twi, q, bq, fold_a, fold_b,
					OPAQUE_ZERO(p));
		}
		__syncwarp();   /* the exchange buffer is reused by the next item */
		if (base + C::groups_per_cta >= nitems) {
			pdl_launch_dependents();
		}
		if (active) {
#pragma unroll
			for (int e = 0; e < 8; e++) {
				u64 v = x[0][e];
				if (fold) {
					v = tile_canon<true, APX>(v, q, bq);
				}
				p.dst[off + G::eoff(0, e)] = v;
			}
		}
	}
}

/* ---- single pass for 2^9 <= n <= 2^11 ---------------------------------------------------
 * Two passes move 32n bytes per transform; at n = 2^9 .. 2^12 the batched
 * transform then runs at 4.3-4.6 TB/s of real traffic, i.e. it is bound by HBM,
 * not by the butterflies.  A polynomial of up to 4096 coefficients fits one
 * CTA: 2^(K-3) threads hold 8 coefficients each and run all K = log2 n stages,
 * exchanging through (padded) shared memory between rounds.  The CTA stages
 * the whole twiddle table of its limb once (TMA, one copy per level) and
 * reuses it over `bchunk` batch entries.  16n bytes per transform, one launch.
 * Measured on B200 at 2^27 coefficients, forward / inverse, against the
 * two-pass split: n = 2^9 0.74 / 0.76 ms (0.99 / 0.97), 2^10 0.81 / 0.78
 * (0.93 / 0.94), 2^11 0.97 / 0.89 (1.00 / 1.00), 2^12 1.18 / 1.20 (1.00 /
 * 1.11: 512 threads and 96 KB per CTA, two CTAs per SM -- slower, so the
 * default stops at 2^11; $VKHEL_SINGLE_MAX_LOG2N moves the limit, 8 = off). */
#ifndef SINGLE_MAX_LOG2N
#define SINGLE_MAX_LOG2N 11
#endif

template <bool INV, int K, bool APX, bool IND, bool MUL = false>
__global__ void __launch_bounds__(1 << (K - 3), (1024 >> (K - 3)) > 16 ? 16 : (1024 >> (K - 3)))
ntt_single_kernel(const __grid_constant__ fast_pass p) {
	static_assert(!MUL || (INV && !IND), "fused product: direct inverse only");
	using G = tile_geom<K>;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	ulonglong2 *sm_tw = (ulonglong2 *) smem_raw;           /* [2^K] */
	u64 *sm_x = (u64 *) (sm_tw + (1 << K));                /* xpad(2^K) words */

	/* blockIdx.x -> (batch chunk, limb) */
	const unsigned limb = blockIdx.x % p.limbs;
	const unsigned bc = blockIdx.x / p.limbs;
	const u64 batch = p.polys / p.limbs;
	const u64 b0 = (u64) bc * p.bchunk;
	const unsigned nb = (unsigned) (batch - b0 < p.bchunk ? batch - b0 : p.bchunk);

	const limb_desc &d = p.descs[limb];
	const u64 q = d.q, bq = APX ? 3 * q : 2 * q;
	__shared__ __align__(8) u64 tw_bar;
	if (threadIdx.x == 0) {
		mbar_init(&tw_bar, 1);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		stage_twiddles_tma<K>(sm_tw, d.tw + (INV ? ((u64) 1 << K) : 0), 0, 0, 1,
				&tw_bar);
	}
	pdl_wait();

	const int t = threadIdx.x;
	constexpr int first = INV ? G::rounds - 1 : 0;
	constexpr int last = INV ? 0 : G::rounds - 1;
	ulonglong2 fold_a = make_ulonglong2(0, 0), fold_b = fold_a;
	if (INV) {
		fold_a = make_ulonglong2(d.inv_n, d.inv_n_shoup);
		fold_b = make_ulonglong2(d.inv_w1n, d.inv_w1n_shoup);
	}
	const int tb_first = G::tbase(first, t), tb_last = G::tbase(last, t);
	bool tw_ready = false;

	for (unsigned bl = 0; bl < nb; bl++) {
		const u64 poly = (b0 + bl) * p.limbs_total + p.limb0 + limb;
		const u64 *sp;
		u64 *dp;
		if (IND) {
			const ntt_ptrs ent = p.entry(poly);
			sp = ent.src;
			dp = ent.dst;
		} else {
			sp = p.src + (poly << K);
			dp = p.dst + (poly << K);
		}
		u64 x[1][8];
		if (G::eoff(first, 1) == 1) {
#pragma unroll
			for (int e = 0; e < 8; e += 2) {
				const ulonglong2 v =
					*(const ulonglong2 *) (sp + tb_first + G::eoff(first, e));
				x[0][e] = v.x;
				x[0][e + 1] = v.y;
			}
		} else {
#pragma unroll
			for (int e = 0; e < 8; e++) {
				x[0][e] = sp[tb_first + G::eoff(first, e)];
			}
		}
		if (MUL) {
			/* fused point-wise product, as in the row pass: (x mod q)(y mod q)
			 * mod q for any 64-bit factors */
			modulus m;
			m.q = q;
			m.d = d.mm_d;
			m.v = d.mm_v;
			m.s = d.mm_s;
			m.mu = 0;
			const u64 *sp2 = p.src2 + (poly << K) + tb_first;
#pragma unroll
			for (int e = 0; e < 8; e++) {
				const u64 y = sp2[G::eoff(first, e)];
				x[0][e] = fused_pointwise(x[0][e], y, p.fma_mult, m);
			}
		}
		if (!tw_ready) {
			mbar_wait(&tw_bar, 0);
			tw_ready = true;
		}
#pragma unroll
		for (int rr = 0; rr < G::rounds; rr++) {
			const int r = INV ? G::rounds - 1 - rr : rr;
			if (rr > 0) {
				const int prev = INV ? r + 1 : r - 1;
				u64 *xw = sm_x + xpad(G::tbase(prev, t));
				if (G::eoff(prev, 1) == 1) {
#pragma unroll
					for (int e = 0; e < 8; e += 2) {
						*(ulonglong2 *) (xw + xpad(G::eoff(prev, e))) =
							make_ulonglong2(x[0][e], x[0][e + 1]);
					}
				} else {
#pragma unroll
					for (int e = 0; e < 8; e++) {
						xw[xpad(G::eoff(prev, e))] = x[0][e];
					}
				}
				__syncthreads();
				const u64 *xr = sm_x + xpad(G::tbase(r, t));
				if (G::eoff(r, 1) == 1) {
#pragma unroll
					for (int e = 0; e < 8; e += 2) {
						const ulonglong2 v =
							*(const ulonglong2 *) (xr + xpad(G::eoff(r, e)));
						x[0][e] = v.x;
						x[0][e + 1] = v.y;
					}
				} else {
#pragma unroll
					for (int e = 0; e < 8; e++) {
						x[0][e] = xr[xpad(G::eoff(r, e))];
					}
				}
				/* the next exchange writes exactly the words this thread has
				 * just read: no second barrier */
			}
			tile_round<K, INV, INV ? FOLD_LAST : FOLD_NONE, 1, APX>(x, r, t,
					sm_tw, q, bq, fold_a, fold_b, nullptr, OPAQUE_ZERO(p));
		}
		if (bl + 1 == nb) {
			pdl_launch_dependents();   /* only this CTA's last stores remain */
		} else {
			__syncthreads();           /* the next entry reuses the buffer */
		}
#pragma unroll
		for (int e = 0; e < 8; e++) {
			x[0][e] = tile_canon<INV, APX>(x[0][e], q, bq);
		}
		if (G::eoff(last, 1) == 1) {
#pragma unroll
			for (int e = 0; e < 8; e += 2) {
				*(ulonglong2 *) (dp + tb_last + G::eoff(last, e)) =
					make_ulonglong2(x[0][e], x[0][e + 1]);
			}
		} else {
#pragma unroll
			for (int e = 0; e < 8; e++) {
				dp[tb_last + G::eoff(last, e)] = x[0][e];
			}
		}
	}
}

/* Column pass.  CTA = (poly, H, column group): one tile group of 2^K rows at
 * stride 2^(L-s0-K) by 2^CL adjacent columns.  thread = (row group, NP adjacent
 * columns), lanes along the columns: every global access is a run of 2^CL * 8
 * contiguous bytes (128-bit per thread for NP = 2) and every shared-memory
 * access is conflict-free without padding. */
/* How the inverse column pass applies n^-1 when it holds the last stage
 * (ntt_engine.cuh): through the scaled top of the inverse twiddle heap
 * (FOLD_TWID, the default: n/2 fewer modular products per transform) or by
 * multiplying both outputs of the last stage (FOLD_LAST). */
#ifndef COLS_FOLD
#define COLS_FOLD FOLD_TWID
#endif

template <int K, int CL, int NP>
struct col_cfg {
	static constexpr int cthreads_log2 = CL - (NP == 2 ? 1 : 0);
	static constexpr int threads = 1 << (K - 3 + cthreads_log2);
	static constexpr int cols = 1 << CL;
};

template <int NP> struct col_vec;
template <> struct col_vec<1> { typedef u64 type; };
template <> struct col_vec<2> { typedef ulonglong2 type; };

/* (Tried: the row stride as a template parameter, so that every global access
 * is "per-thread base + immediate" -- 80 instructions fewer per thread, and the
 * forward pass 7 % SLOWER, 156.4 against 146.0 us, inverse unchanged:
 * profiles/r02_kernel_ab.txt, variant n0 against the LB8 build.) */
template <bool INV, int K, int CL, int NP, bool APX, bool IND, bool TOP>
__global__ void __launch_bounds__(1 << (K - 3 + CL - (NP == 2 ? 1 : 0)),
		((NP == 2 ? COLS_MIN_THREADS_NP2 : COLS_MIN_THREADS_NP1)
			>> (K - 3 + CL - (NP == 2 ? 1 : 0))) > 0
		? ((NP == 2 ? COLS_MIN_THREADS_NP2 : COLS_MIN_THREADS_NP1)
			>> (K - 3 + CL - (NP == 2 ? 1 : 0))) : 1)
ntt_cols_kernel(const __grid_constant__ fast_pass p) {
	using G = tile_geom<K>;
	using C = col_cfg<K, CL, NP>;
	typedef typename col_vec<NP>::type vec_t;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	ulonglong2 *sm_tw = (ulonglong2 *) smem_raw;           /* [2^K] */
	u64 *sm_x = (u64 *) (sm_tw + (1 << K));                /* [2^K][2^CL] */

	const unsigned L = p.log2n, s0 = p.s0;
	const unsigned low_bits = L - s0 - K;                  /* log2 of the row stride */
	const unsigned cgroups_log2 = low_bits - CL;

	/* blockIdx.x -> (poly, H, column group), column group fastest */
	u64 blk = blockIdx.x;
	const u64 cg = blk & (((u64) 1 << cgroups_log2) - 1);
	blk >>= cgroups_log2;
	const u64 H = blk & (((u64) 1 << s0) - 1);
	const u64 lpoly = blk >> s0;

	/* blk counts the polynomials of this launch: (batch entry, limb of the
	 * slice) -> position in the [batch][limbs_total] layout */
	const unsigned pb = (unsigned) lpoly / p.limbs;   /* grid below 2^31 */
	const unsigned pl = (unsigned) lpoly - pb * p.limbs;
	const u64 poly = (u64) pb * p.limbs_total + p.limb0 + pl;
	const limb_desc &d = p.descs[pl];
	const u64 q = d.q, bq = APX ? 3 * q : 2 * q;   /* butterfly bound */

	const int c = (threadIdx.x & ((1 << C::cthreads_log2) - 1)) * NP;
	const int t = threadIdx.x >> C::cthreads_log2;         /* row group */
	constexpr int first = INV ? G::rounds - 1 : 0;
	constexpr int last = INV ? 0 : G::rounds - 1;
	const u64 base = (IND ? 0 : (poly << L)) + (H << (L - s0)) + (cg << CL) + c;
	const u64 *src_base = p.src;
	u64 *dst_base = p.dst;
	if (IND) {
		const ntt_ptrs ent = p.entry(poly);
		src_base = p.tab_second ? ent.dst : ent.src;
		dst_base = ent.dst;
	}

	/* the forward transform always ends in a row pass; the inverse ends here
	 * when this pass holds stage 0.  (A one-CTA-per-polynomial single-launch
	 * variant of this kernel, K = log2 n, was measured no faster than the
	 * two-pass split even for a single polynomial.) */
	static_assert(INV || !TOP, "TOP distinguishes inverse passes only");
	constexpr bool fold = INV && TOP;            /* TOP: s0 == 0 */
	constexpr int FM = fold ? COLS_FOLD : FOLD_NONE;
	static_assert(K <= SCALED_TW_MAX_LOG2, "scaled twiddles cover 2^10 nodes");
	/* FOLD_TWID: the tile is the top of the heap (root 1); its scaled copy
	 * lies behind the coefficients' exchange buffer */
	const ulonglong2 *sm_tws = (const ulonglong2 *) (sm_x + ((size_t) 1 << (K + CL)));
	constexpr bool scaled_tw = FM == FOLD_TWID && fold;

	__shared__ __align__(8) u64 tw_bar;
	if (threadIdx.x == 0) {
		mbar_init(&tw_bar, scaled_tw ? 2 : 1);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		stage_twiddles_tma<K>(sm_tw, d.tw + (INV ? ((u64) 1 << L) : 0), s0, H, 1,
				&tw_bar);
		if (scaled_tw) {
			stage_twiddles_tma<K>((ulonglong2 *) sm_tws, d.tw + ((u64) 2 << L),
					0, 0, 1, &tw_bar);
		}
	}
	pdl_wait();   /* the coefficients come from the previous kernel */
#if FAST_PDL_EARLY
	pdl_launch_dependents();
#endif
	u64 x[NP][8];
	{
		const u64 *sp = src_base + base + ((u64) G::tbase(first, t) << low_bits);
#pragma unroll
		for (int e = 0; e < 8; e++) {
			const vec_t v = *(const vec_t *) (sp + ((u64) G::eoff(first, e) << low_bits));
			if (NP == 2) {
				x[0][e] = ((const u64 *) &v)[0];
				x[NP - 1][e] = ((const u64 *) &v)[NP - 1];
			} else {
				x[0][e] = ((const u64 *) &v)[0];
			}
		}
	}
	constexpr bool canon = fold;
	ulonglong2 fold_a = make_ulonglong2(0, 0), fold_b = fold_a;
	if (fold) {
		fold_a = make_ulonglong2(d.inv_n, d.inv_n_shoup);
		fold_b = make_ulonglong2(d.inv_w1n, d.inv_w1n_shoup);
	}
	mbar_wait(&tw_bar, 0);   /* twiddles have landed */

	static_for<0, G::rounds>([&](auto rrc) {
		constexpr int rr = decltype(rrc)::value;
		constexpr int r = INV ? G::rounds - 1 - rr : rr;
		if constexpr (rr > 0) {
			constexpr int prev = INV ? r + 1 : r - 1;
			u64 *xw = sm_x + (G::tbase(prev, t) << CL) + c;
#pragma unroll
			for (int e = 0; e < 8; e++) {
				if (COLS_XST64) {
					break;
				}
				vec_t v;
				((u64 *) &v)[0] = x[0][e];
				if (NP == 2) {
					((u64 *) &v)[NP - 1] = x[NP - 1][e];
				}
				*(vec_t *) (xw + (G::eoff(prev, e) << CL)) = v;
			}
			if (COLS_XST64) {
				/* (the opaque zero hides the base's 16-byte alignment from
				 * ptxas, which would otherwise fuse the pairs again) */
				const unsigned xa = (unsigned) __cvta_generic_to_shared(xw)
					+ (unsigned) OPAQUE_ZERO(p);
				static_for<0, 8>([&](auto ec) {
					constexpr int e = decltype(ec)::value;
					static_for<0, NP>([&](auto pc) {
						constexpr int pp = decltype(pc)::value;
						sts64_at<8 * ((G::eoff(prev, e) << CL) + pp)>(xa, x[pp][e]);
					});
				});
			}
			__syncthreads();
			const u64 *xr = sm_x + (G::tbase(r, t) << CL) + c;
#pragma unroll
			for (int e = 0; e < 8; e++) {
				const vec_t v = *(const vec_t *) (xr + (G::eoff(r, e) << CL));
				x[0][e] = ((const u64 *) &v)[0];
				if (NP == 2) {
					x[NP - 1][e] = ((const u64 *) &v)[NP - 1];
				}
			}
		}
		tile_round<K, INV, FM, NP, APX>(x, r, t, sm_tw, q, bq, fold_a, fold_b,
				sm_tws, OPAQUE_ZERO(p));
	});

	pdl_launch_dependents();   /* only this CTA's stores remain */
	u64 *dp = dst_base + base + ((u64) G::tbase(last, t) << low_bits);
#pragma unroll
	for (int e = 0; e < 8; e++) {
		vec_t v;
#pragma unroll
		for (int pp = 0; pp < NP; pp++) {
			u64 w = x[pp][e];
			if (canon) {
				w = tile_canon<true, APX, FM>(w, q, bq);
			}
			if (COLS_GST64) {
				dp[((u64) G::eoff(last, e) << low_bits) + pp] = w;
			}
			((u64 *) &v)[pp] = w;
		}
		if (!COLS_GST64) {
			*(vec_t *) (dp + ((u64) G::eoff(last, e) << low_bits)) = v;
		}
	}
	if (FM == FOLD_TWID && scaled_tw && t == 0) {
		/* coefficient 0 of the tile (register 0 of row group 0) is the one
		 * value no difference has scaled (ntt_engine.cuh): the explicit
		 * product, stored over what this thread has just written there */
#pragma unroll
		for (int pp = 0; pp < NP; pp++) {
			const u64 w = shoup_lazy(x[pp][0], fold_a.x, fold_a.y, q);
			dp[pp] = csub(w, q);
		}
	}
}

template <bool INV, int K, int NP, bool MUL, bool APX>
static void run_rows_np(struct vkhel_ctx *ctx, fast_pass p) {
	using C = row_cfg<K>;
	const u64 batch = p.polys / p.limbs;
	const u64 slots = (u64) C::groups_per_cta * NP;   /* tiles in flight per CTA */
	/* enough tiles per CTA to occupy every warp-group: H values first (they
	 * are contiguous in memory), then batch entries */
	unsigned hgroup_log2 = 0;
	while ((1u << hgroup_log2) < (unsigned) C::groups_per_cta
			&& hgroup_log2 < p.s0
			&& ((u64) (batch + NP - 1) / NP << hgroup_log2) < (u64) C::groups_per_cta) {
		hgroup_log2++;
	}
	/* two rounds of items per CTA amortise the twiddle staging */
	/* One round per CTA -- twice the CTAs, half the ragged last wave -- for a
	 * launch of fewer than four waves that has the device to itself; slices
	 * on two streams fill each other's tails and keep two rounds.  Measured: n
	 * = 2^14 x 256 forward 40.8 against 42.3 us; with one round everywhere the
	 * full bench loses 0.5 % (631.3 against 627.8 us).  $VKHEL_ROWS_ROUNDS=1|2
	 * forces either. */
	static const int rounds_env = getenv("VKHEL_ROWS_ROUNDS")
		? atoi(getenv("VKHEL_ROWS_ROUNDS")) : 0;
	u64 rounds = ROWS_ROUNDS_PER_CTA;
	if (!ctx->dev.in_slices) {
		u64 bc2 = (rounds * slots) >> hgroup_log2;
		bc2 = bc2 < (u64) NP ? NP : bc2 > batch ? batch : bc2;
		const u64 ctas2 = ((batch + bc2 - 1) / bc2) * p.limbs
			* ((1ull << p.s0) >> hgroup_log2);
		if (ctas2 < 4ull * 6 * (u64) ctx->dev.sm_count) {
			rounds = 1;
		}
	}
	if (rounds_env == 1 || rounds_env == 2) {
		rounds = (u64) rounds_env;
	}
	u64 bchunk = (rounds * slots) >> hgroup_log2;
	if (bchunk < (u64) NP) {
		bchunk = NP;
	}
	if (bchunk > batch) {
		bchunk = batch;
	}
	p.hgroup_log2 = hgroup_log2;
	p.bchunk = (unsigned) bchunk;
	const u64 bchunks = (batch + bchunk - 1) / bchunk;
	const u64 blocks = bchunks * p.limbs * ((1ull << p.s0) >> hgroup_log2);
	VK_REQUIRE(blocks <= 0x7fffffffull, "transform too large for one launch");
	const size_t smem = ((size_t) sizeof(ulonglong2) << (hgroup_log2 + K))
		+ (size_t) C::groups_per_cta * NP * C::xbuf * sizeof(u64);
	/* inverse: the variant that holds stage 0 (and n^-1), or the plain one */
	const bool top = INV && p.s0 == 0;
	void (*kernel)(fast_pass);
	if (p.indirect()) {
		if constexpr (INV) {
			kernel = top ? ntt_rows_kernel<INV, K, NP, MUL, APX, true, true>
				: ntt_rows_kernel<INV, K, NP, MUL, APX, true, false>;
		} else {
			kernel = ntt_rows_kernel<INV, K, NP, false, APX, true, false>;
		}
		launch_fast_optin(ctx, kernel, (unsigned) blocks, FAST_THREADS,
				smem, p);
		return;
	}
	if constexpr (INV) {
		kernel = top ? ntt_rows_kernel<INV, K, NP, MUL, APX, false, true>
			: ntt_rows_kernel<INV, K, NP, MUL, APX, false, false>;
	} else {
		kernel = ntt_rows_kernel<INV, K, NP, MUL, APX, false, false>;
		if constexpr (!MUL && NP == 1) {
			if (p.lazy) {
				kernel = ntt_rows_kernel<INV, K, NP, false, APX, false, false, true>;
			}
		}
	}
	launch_fast_optin(ctx, kernel, (unsigned) blocks, FAST_THREADS, smem, p);
}

template <bool INV, int K, bool APX>
static void run_rows(struct vkhel_ctx *ctx, const fast_pass &p) {
	if constexpr (INV) {
		if (p.mul) {
			run_rows_np<INV, K, 1, true, APX>(ctx, p);
			return;
		}
	}
	/* two batch entries per thread when there are two to pair up */
	if constexpr (ROWS_NP == 2) {
		if (p.polys / p.limbs >= 2) {
			run_rows_np<INV, K, 2, false, APX>(ctx, p);
			return;
		}
	}
	run_rows_np<INV, K, 1, false, APX>(ctx, p);
}

template <bool INV, int K, int CL, int NP, bool APX>
static void run_cols_cl(struct vkhel_ctx *ctx, const fast_pass &p) {
	using C = col_cfg<K, CL, NP>;
	const unsigned low_bits = p.log2n - p.s0 - K;
	VK_REQUIRE((int) low_bits >= CL,
			"internal: column pass narrower than its CTA");
	const u64 blocks = (p.polys << p.s0) << (low_bits - CL);
	VK_REQUIRE(blocks <= 0x7fffffffull, "transform too large for one launch");
	/* twiddle subtree, exchange buffer, and for the inverse's last pass the
	 * scaled subtree (COLS_FOLD) */
	const size_t smem = (sizeof(ulonglong2) << K) + (sizeof(u64) << (K + CL))
		+ (INV && COLS_FOLD == FOLD_TWID && p.s0 == 0
				? sizeof(ulonglong2) << K : 0);
	const bool top = INV && p.s0 == 0;
	void (*kernel)(fast_pass);
	if (p.indirect()) {
		if constexpr (INV) {
			kernel = top ? ntt_cols_kernel<INV, K, CL, NP, APX, true, true>
				: ntt_cols_kernel<INV, K, CL, NP, APX, true, false>;
		} else {
			kernel = ntt_cols_kernel<INV, K, CL, NP, APX, true, false>;
		}
	} else {
		if constexpr (INV) {
			kernel = top ? ntt_cols_kernel<INV, K, CL, NP, APX, false, true>
				: ntt_cols_kernel<INV, K, CL, NP, APX, false, false>;
		} else {
			kernel = ntt_cols_kernel<INV, K, CL, NP, APX, false, false>;
		}
	}
	launch_fast_optin(ctx, kernel, (unsigned) blocks, C::threads, smem, p);
}

/* 256 threads per CTA: 2^(12-K) columns with two columns per thread, 2^(11-K)
 * with one; the 8-point column pass of n = 2^9 and 2^10 has only 64 / 128
 * columns to offer and runs with fewer threads */
#ifndef COLS_THREADS_LOG2
#define COLS_THREADS_LOG2 8
#endif
template <bool INV, int K, bool APX>
static void run_cols(struct vkhel_ctx *ctx, const fast_pass &p) {
	const unsigned low_bits = p.log2n - p.s0 - K;
	if constexpr (COLS_NP == 2 && K <= 8) {
		/* CTA of 2^COLS_THREADS_LOG2 threads: 2^(K-3) row groups by
		 * 2^(COLS_THREADS_LOG2 + 4 - K) columns, two per thread */
		constexpr int CLW = COLS_THREADS_LOG2 + 4 - K;
		if (CLW >= 1 && low_bits >= (unsigned) CLW) {
			run_cols_cl<INV, K, CLW < 1 ? 1 : CLW, 2, APX>(ctx, p);
			return;
		}
		if (low_bits >= 12 - K) {
			run_cols_cl<INV, K, 12 - K, 2, APX>(ctx, p);
			return;
		}
	}
	if constexpr (K >= 9) {
		/* 512- and 1024-point column tiles (n = 2^17, 2^18), 512 threads: two
		 * adjacent columns per thread here as well (16 / 8 columns per CTA);
		 * measured against one column per thread at 2^27 coefficients:
		 * n = 2^17 1.348 / 1.470 ms forward / inverse instead of 1.418 /
		 * 1.557, n = 2^18 1.502 / 1.646 instead of 1.773 / 1.882 */
		if constexpr (COLS_NP == 2) {
			run_cols_cl<INV, K, K == 9 ? 4 : 3, 2, APX>(ctx, p);
		} else {
			run_cols_cl<INV, K, 3, 1, APX>(ctx, p);
		}
	} else if (low_bits >= 11 - K) {
		run_cols_cl<INV, K, 11 - K, 1, APX>(ctx, p);
	} else if (K == 3 && low_bits == 7) {
		run_cols_cl<INV, 3, 7, 1, APX>(ctx, p);
	} else if (K == 3 && low_bits == 6) {
		run_cols_cl<INV, 3, 6, 1, APX>(ctx, p);
	} else {
		VK_DIE("internal: no column kernel for K=%d, row stride 2^%u", K,
				low_bits);
	}
}

template <bool INV, bool APX>
static void run_rows_k(struct vkhel_ctx *ctx, const fast_pass &p, unsigned k) {
	switch (k) {
	case 3: run_rows<INV, 3, APX>(ctx, p); break;
	case 4: run_rows<INV, 4, APX>(ctx, p); break;
	case 5: run_rows<INV, 5, APX>(ctx, p); break;
	case 6: run_rows<INV, 6, APX>(ctx, p); break;
	case 7: run_rows<INV, 7, APX>(ctx, p); break;
	case 8: run_rows<INV, 8, APX>(ctx, p); break;
	default: VK_DIE("internal: row pass of %u stages", k);
	}
}

template <bool INV, bool APX>
static void run_cols_k(struct vkhel_ctx *ctx, const fast_pass &p, unsigned k) {
	switch (k) {
	case 3: run_cols<INV, 3, APX>(ctx, p); break;
	case 4: run_cols<INV, 4, APX>(ctx, p); break;
	case 5: run_cols<INV, 5, APX>(ctx, p); break;
	case 6: run_cols<INV, 6, APX>(ctx, p); break;
	case 7: run_cols<INV, 7, APX>(ctx, p); break;
	case 8: run_cols<INV, 8, APX>(ctx, p); break;
	case 9: run_cols<INV, 9, APX>(ctx, p); break;
	case 10: run_cols<INV, 10, APX>(ctx, p); break;
	default: VK_DIE("internal: column pass of %u stages", k);
	}
}

template <bool INV, int K, bool APX>
static void run_single(struct vkhel_ctx *ctx, fast_pass p) {
	const u64 batch = p.polys / p.limbs;
	constexpr unsigned threads = 1u << (K - 3);
	/* batch entries per CTA: reuse the staged twiddles, but keep at least four
	 * waves of CTAs */
	const u64 resident = (u64) ctx->dev.sm_count * (2048 / threads > 16 ? 16 : 2048 / threads);
	u64 bchunk = batch * p.limbs / (4 * resident);
	if (bchunk > 16) {
		bchunk = 16;
	}
	if (bchunk < 1) {
		bchunk = 1;
	}
	p.bchunk = (unsigned) bchunk;
	const u64 blocks = ((batch + bchunk - 1) / bchunk) * p.limbs;
	VK_REQUIRE(blocks <= 0x7fffffffull, "transform too large for one launch");
	const size_t smem = (sizeof(ulonglong2) << K)
		+ sizeof(u64) * (size_t) (xpad(1 << K) + 4);
	if constexpr (INV) {
		if (p.mul) {
			VK_REQUIRE(!p.indirect(), "internal: indirect product in one pass");
			if (smem_needs_optin(smem)) {
				CUDA_CHECK(cudaFuncSetAttribute(
							ntt_single_kernel<true, K, APX, false, true>,
							cudaFuncAttributeMaxDynamicSharedMemorySize,
							(int) smem));
			}
			launch_fast(ctx, ntt_single_kernel<true, K, APX, false, true>,
					(unsigned) blocks, threads, smem, p);
			return;
		}
	}
	if (p.indirect()) {
		if (smem_needs_optin(smem)) {
			CUDA_CHECK(cudaFuncSetAttribute(ntt_single_kernel<INV, K, APX, true>,
						cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		}
		launch_fast(ctx, ntt_single_kernel<INV, K, APX, true>, (unsigned) blocks,
				threads, smem, p);
		return;
	}
	if (smem_needs_optin(smem)) {
		CUDA_CHECK(cudaFuncSetAttribute(ntt_single_kernel<INV, K, APX, false>,
					cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	}
	launch_fast(ctx, ntt_single_kernel<INV, K, APX, false>, (unsigned) blocks,
			threads, smem, p);
}

template <bool INV, bool APX>
static void run_single_k(struct vkhel_ctx *ctx, const fast_pass &p, unsigned k) {
	switch (k) {
	case 9: run_single<INV, 9, APX>(ctx, p); break;
	case 10: run_single<INV, 10, APX>(ctx, p); break;
	case 11: run_single<INV, 11, APX>(ctx, p); break;
	case 12: run_single<INV, 12, APX>(ctx, p); break;
	case 13: run_single<INV, 13, APX>(ctx, p); break;
	default: VK_DIE("internal: single pass of %u stages", k);
	}
}

/* n in [2^9, 2^single_max]: one pass ($VKHEL_SINGLE_MAX_LOG2N, default 11, at
 * most 13; 8 disables it) */
static unsigned single_max_log2n() {
	static int v = -1;
	if (v < 0) {
		const char *env = getenv("VKHEL_SINGLE_MAX_LOG2N");
		v = env && *env ? atoi(env) : SINGLE_MAX_LOG2N;
		if (v > 13) {
			v = 13;   /* 2^13: 128 KB of twiddles + 72 KB of coefficients */
		}
	}
	return (unsigned) v;
}

/* kernel launches of one directly launched fast-path transform */
unsigned ntt_launches_per_transform(unsigned log2n) {
	if (log2n <= 8 || log2n <= single_max_log2n()) {
		return 1;
	}
	return log2n <= 18 ? 2 : 3;
}

/* stage split of the fast path: [lead (generic, strided)] [col] [row] */
struct fast_plan {
	unsigned lead, kcol, krow;
};

/* $VKHEL_KROW="12:6,15:6": stages of the row pass for given log2 n (tuning) */
static unsigned krow_override(unsigned log2n) {
	static int table[32];
	static bool parsed = false;
	if (!parsed) {
		parsed = true;
		const char *env = getenv("VKHEL_KROW");
		while (env && *env) {
			char *end;
			const long l = strtol(env, &end, 10);
			if (*end != ':') {
				break;
			}
			const long k = strtol(end + 1, &end, 10);
			if (l >= 9 && l <= 18 && k >= 3 && k <= 8 && l - k >= 3
					&& l - k <= 10) {
				table[l] = (int) k;
			}
			env = *end == ',' ? end + 1 : end;
			if (*end != ',') {
				break;
			}
		}
	}
	return log2n < 32 ? (unsigned) table[log2n] : 0;
}

/* `forward_only`: the plan of a plain forward transform, which may differ from
 * the split the inverse and the fused product use */
static fast_plan plan_fast(unsigned log2n, bool forward_only = false) {
	fast_plan pl = { 0, 0, 0 };
	if (log2n <= 8) {
		pl.krow = log2n;
	} else if (log2n <= 18) {
		pl.krow = log2n - 8 >= 3 ? 8 : log2n - 3;
		/* measured at 2^27 coefficients (tools/split_bench.py): forward
		 * n = 2^13 as 6 + 7 1.067 ms against 1.094 as 5 + 8, n = 2^15 as 8 + 7
		 * 1.156 against 1.207 as 7 + 8; the inverse gains nothing from
		 * either, and every other size is best with the 8-stage row pass */
		if (forward_only && (log2n == 13 || log2n == 15)) {
			pl.krow = 7;
		}
		if (krow_override(log2n)) {
			pl.krow = krow_override(log2n);
		}
		pl.kcol = log2n - pl.krow;
	} else {
		pl.krow = 8;
		pl.kcol = 8;
		pl.lead = log2n - 16;
	}
	return pl;
}

/* $VKHEL_ONLY_PASS=cols|rows launches only that pass of a two-pass transform:
 * results are INVALID, the switch exists to time the passes apart
 * (tools/kernel_ab.py) */
static int only_pass() {
	static int v = -1;
	if (v < 0) {
		const char *env = getenv("VKHEL_ONLY_PASS");
		v = !env ? 0 : !strcmp(env, "cols") ? 1 : !strcmp(env, "rows") ? 2 : 0;
	}
	return v;
}

template <bool INV, bool APX>
static void run_fast(struct vkhel_ctx *ctx, const u64 *src, u64 *dst,
		const limb_desc *descs, uint64_t limbs, uint64_t polys,
		unsigned log2n, const u64 *src2 = NULL, const ntt_ptrs *tab = NULL,
		unsigned limbs_total = 0, unsigned limb0 = 0,
		const ntt_ptrs *inline_tab = NULL, u64 fma_mult = 0,
		bool ind_product = false) {
	const fast_plan pl = plan_fast(log2n, !INV);
	fast_pass p;
	p.zero = 0;
	p.fma_mult = fma_mult;
	p.mul = false;
	p.lazy = !INV && ctx->dev.lazy_out && !tab && !inline_tab;
	p.tab = tab;
	p.tab_second = 0;
	p.tab_inline = 0;
	if (inline_tab) {
		VK_REQUIRE(polys <= NTT_INLINE_PTRS, "internal: inline batch too long");
		p.tab_inline = (unsigned) polys;
		for (uint64_t i = 0; i < polys; i++) {
			p.inl[i] = inline_tab[i];
		}
	}
	p.limbs_total = limbs_total ? limbs_total : (unsigned) limbs;
	p.limb0 = limb0;
	VK_REQUIRE(!p.indirect() || !pl.lead,
			"internal: indirect batch of n > 2^18");
	p.src2 = NULL;
	p.descs = descs;
	p.limbs = (unsigned) limbs;
	p.log2n = log2n;
	p.polys = polys;
	p.hgroup_log2 = 0;
	p.bchunk = 1;

	if ((!src2 || INV) && log2n >= 9 && log2n <= single_max_log2n()) {
		p.src = src;
		p.src2 = INV ? src2 : NULL;
		p.mul = INV && src2 != NULL;
		p.dst = dst;
		p.s0 = 0;
		run_single_k<INV, APX>(ctx, p, log2n);
		return;
	}

	gen_pass lead;
	lead.descs = descs;
	lead.limbs = (unsigned) limbs;
	lead.log2n = log2n;
	lead.s0 = 0;
	lead.k = pl.lead;
	lead.tiles = polys;
	lead.last = false;

	if (!INV) {
		const u64 *cur = src;
		if (pl.lead) {
			lead.src = cur;
			lead.dst = dst;
			run_generic_pass<false, false>(ctx, lead);
			cur = dst;
		}
		if (pl.kcol && only_pass() != 2) {
			p.src = cur;
			p.dst = dst;
			p.s0 = pl.lead;
			if (!p.indirect() && ntt_cols_tma_enabled(log2n, pl.kcol, pl.lead)) {
				/* opt-in variant: the tile arrives by one TMA tensor-map load */
				launch_ntt_cols_tma(ctx, APX, cur, dst, descs, limbs, polys,
						p.limbs_total, p.limb0);
			} else {
				run_cols_k<false, APX>(ctx, p, pl.kcol);
			}
			cur = dst;
			p.tab_second = 1;
		}
		p.src = cur;
		p.dst = dst;
		p.s0 = pl.lead + pl.kcol;
		if (!pl.kcol || only_pass() != 1) {
			run_rows_k<false, APX>(ctx, p, pl.krow);
		}
	} else {
		p.src = src;
		p.src2 = src2;
		p.mul = src2 != NULL || ind_product;
		p.dst = dst;
		p.s0 = pl.lead + pl.kcol;
		if (!pl.kcol || only_pass() != 1) {
			run_rows_k<true, APX>(ctx, p, pl.krow);
		}
		p.src2 = NULL;
		p.mul = false;
		p.tab_second = 1;
		if (pl.kcol && only_pass() != 2) {
			p.src = dst;
			p.s0 = pl.lead;
			run_cols_k<true, APX>(ctx, p, pl.kcol);
		}
		if (pl.lead) {
			lead.src = dst;
			lead.dst = dst;
			lead.last = true;
			run_generic_pass<true, false>(ctx, lead);
		}
	}
}

template <int K, bool APX>
static void run_rows_polymul(struct vkhel_ctx *ctx, fast_pass p) {
	using C = row_cfg<K>;
	const u64 batch = p.polys / p.limbs;
	unsigned hgroup_log2 = 0;
	while ((1u << hgroup_log2) < (unsigned) C::groups_per_cta
			&& hgroup_log2 < p.s0
			&& (batch << hgroup_log2) < (u64) C::groups_per_cta) {
		hgroup_log2++;
	}
	u64 bchunk = (ROWS_ROUNDS_PER_CTA * (u64) C::groups_per_cta) >> hgroup_log2;
	if (bchunk < 1) {
		bchunk = 1;
	}
	if (bchunk > batch) {
		bchunk = batch;
	}
	p.hgroup_log2 = hgroup_log2;
	p.bchunk = (unsigned) bchunk;
	const u64 bchunks = (batch + bchunk - 1) / bchunk;
	const u64 blocks = bchunks * p.limbs * ((1ull << p.s0) >> hgroup_log2);
	VK_REQUIRE(blocks <= 0x7fffffffull, "product too large for one launch");
	const size_t smem = 2 * ((size_t) sizeof(ulonglong2) << (hgroup_log2 + K))
		+ (size_t) C::groups_per_cta * C::xbuf * sizeof(u64);
	if (smem_needs_optin(smem)) {
		CUDA_CHECK(cudaFuncSetAttribute(ntt_rows_polymul_kernel<K, APX>,
					cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	}
	launch_fast(ctx, ntt_rows_polymul_kernel<K, APX>, (unsigned) blocks,
			FAST_THREADS, smem, p);
}

template <bool APX>
static void run_rows_polymul_k(struct vkhel_ctx *ctx, const fast_pass &p,
		unsigned k) {
	switch (k) {
	case 3: run_rows_polymul<3, APX>(ctx, p); break;
	case 4: run_rows_polymul<4, APX>(ctx, p); break;
	case 5: run_rows_polymul<5, APX>(ctx, p); break;
	case 6: run_rows_polymul<6, APX>(ctx, p); break;
	case 7: run_rows_polymul<7, APX>(ctx, p); break;
	case 8: run_rows_polymul<8, APX>(ctx, p); break;
	default: VK_DIE("internal: row pass of %u stages", k);
	}
}

/* c = INTT(NTT(a) (*) NTT(b)): the strided passes of both forward transforms
 * (b into tmp first, so that dst may alias a, b or both), the fused middle
 * kernel, the strided passes of the inverse.  n <= 256: the middle kernel is
 * the whole product and tmp is not used. */
template <bool APX>
static void run_fast_polymul(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *tmp, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n) {
	const fast_plan pl = plan_fast(log2n);
	fast_pass p;
	p.zero = 0;
	p.fma_mult = 0;
	p.mul = false;
	p.lazy = false;
	p.tab = NULL;
	p.tab_second = 0;
	p.tab_inline = 0;
	p.limbs_total = (unsigned) limbs;
	p.limb0 = 0;
	p.src2 = NULL;
	p.descs = descs;
	p.limbs = (unsigned) limbs;
	p.log2n = log2n;
	p.polys = polys;
	p.hgroup_log2 = 0;
	p.bchunk = 1;

	gen_pass lead;
	lead.descs = descs;
	lead.limbs = (unsigned) limbs;
	lead.log2n = log2n;
	lead.s0 = 0;
	lead.k = pl.lead;
	lead.tiles = polys;
	lead.last = false;

	const u64 *cur_a = a, *cur_b = b;
	if (pl.lead) {
		lead.src = b;
		lead.dst = tmp;
		run_generic_pass<false, false>(ctx, lead);
		lead.src = a;
		lead.dst = dst;
		run_generic_pass<false, false>(ctx, lead);
		cur_a = dst;
		cur_b = tmp;
	}
	if (pl.kcol) {
		p.s0 = pl.lead;
		p.src = cur_b;
		p.dst = tmp;
		run_cols_k<false, APX>(ctx, p, pl.kcol);
		p.src = cur_a;
		p.dst = dst;
		run_cols_k<false, APX>(ctx, p, pl.kcol);
		cur_a = dst;
		cur_b = tmp;
	}
	p.src = cur_a;
	p.src2 = cur_b;
	p.dst = dst;
	p.s0 = pl.lead + pl.kcol;
	run_rows_polymul_k<APX>(ctx, p, pl.krow);
	p.src2 = NULL;
	if (pl.kcol) {
		p.src = dst;
		p.s0 = pl.lead;
		run_cols_k<true, APX>(ctx, p, pl.kcol);
	}
	if (pl.lead) {
		lead.src = dst;
		lead.dst = dst;
		lead.last = true;
		run_generic_pass<true, false>(ctx, lead);
	}
}

/* ======================================================================================
 * Dispatch
 * ====================================================================================== */
/* The approximate-quotient butterflies keep forward values below 6q, so they
 * need 6q < 2^64; they are not combined with the leading generic pass of
 * n > 2^18, whose inverse expects values below 2q. */
static bool use_approx(uint64_t q_max, unsigned log2n) {
	/* $VKHEL_EXACT_QUOTIENT forces the exact-quotient kernels (tests run
	 * both families on the same moduli) */
	static const bool force_exact = getenv("VKHEL_EXACT_QUOTIENT") != NULL;
	return !force_exact && q_max < 0xffffffffffffffffull / 6 && log2n <= 18;
}

static void run_fast_any(struct vkhel_ctx *ctx, bool inverse, bool apx,
		const u64 *src, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, unsigned limbs_total, unsigned limb0) {
	if (ntt_cluster_enabled(log2n)) {
		/* opt-in: one launch on a thread-block cluster instead of two passes */
		launch_ntt_cluster(ctx, inverse, apx, src, dst, descs, limbs, polys,
				log2n, limbs_total, limb0);
		return;
	}
	if (apx) {
		if (inverse) run_fast<true, true>(ctx, src, dst, descs, limbs, polys, log2n, NULL, NULL, limbs_total, limb0);
		else run_fast<false, true>(ctx, src, dst, descs, limbs, polys, log2n, NULL, NULL, limbs_total, limb0);
	} else {
		if (inverse) run_fast<true, false>(ctx, src, dst, descs, limbs, polys, log2n, NULL, NULL, limbs_total, limb0);
		else run_fast<false, false>(ctx, src, dst, descs, limbs, polys, log2n, NULL, NULL, limbs_total, limb0);
	}
}

/* ---- L2-resident intermediate -------------------------------------------------------
 * A two-pass transform writes the whole batch between its passes and reads it
 * back: 32n bytes of HBM traffic per transform for 16n algorithmic.  When the
 * batch is larger than L2 it is therefore cut into slices small enough that a
 * slice written by the first pass is still in L2 when the second pass reads
 * it: by limb ranges for an RNS batch (a row-pass CTA shares its twiddles over
 * the batch entries of one limb, so whole limbs stay together), by batch
 * ranges for a single modulus.  The slices alternate between the context's
 * stream and an auxiliary one, so that the ragged end of one slice's kernels
 * overlaps the next slice instead of leaving SMs idle.  $VKHEL_SLICE_MIB sets
 * the slice size for both cases (0 turns slicing off; a set value also slices
 * batches that would fit in L2, which is what the tests use). */
static size_t slice_bytes_setting(bool by_limb, bool *forced) {
	static double mib = -2;
	if (mib == -2) {
		const char *env = getenv("VKHEL_SLICE_MIB");
		mib = env && *env ? atof(env) : -1;
	}
	*forced = mib >= 0;
	if (mib >= 0) {
		return (size_t) (mib * 1048576.0);
	}
	/* measured on B200: slicing an RNS batch by limbs costs
	 * no time (n = 2^16, 32 limbs x 16: 1.54 M NTT/s either way) and takes
	 * the DRAM traffic from 2.0x to 1.1x the algorithmic bytes; slicing a
	 * single-modulus batch by batch ranges costs 1-3 %, so it is opt-in */
	return by_limb ? (size_t) VKHEL_DEFAULT_SLICE_MIB << 20 : 0;
}

/* streams the slices of a transform alternate between: the context's, the
 * auxiliary one, and up to two more ($VKHEL_SLICE_STREAMS, default 2) */
static int slice_streams() {
	static int v = -1;
	if (v < 0) {
		const char *env = getenv("VKHEL_SLICE_STREAMS");
		v = env && *env ? atoi(env) : 2;
		v = v < 2 ? 2 : v > 4 ? 4 : v;
	}
	return v;
}

/* stream number k (0 = the context's stream) and the event that joins it */
static cudaStream_t slice_stream(struct vkhel_ctx *ctx, int k, cudaEvent_t *ev) {
	struct device_ctx *dev = &ctx->dev;
	if (k == 0) {
		return ctx_stream(ctx);
	}
	if (k == 1) {
		*ev = (cudaEvent_t) dev->ev_aux;
		return (cudaStream_t) dev->stream_aux;
	}
	if (!dev->stream_more[k - 2]) {
		cudaStream_t st;
		cudaEvent_t e;
		CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
		CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		dev->stream_more[k - 2] = st;
		dev->ev_more[k - 2] = e;
	}
	*ev = (cudaEvent_t) dev->ev_more[k - 2];
	return (cudaStream_t) dev->stream_more[k - 2];
}

void ntt_split_join(struct vkhel_ctx *ctx) {
	if (!ctx->dev.split_active) {
		return;
	}
	ctx->dev.split_active = 0;
	CUDA_CHECK(cudaSetDevice(ctx->dev.device));
	for (int k = 1; k < slice_streams(); k++) {
		cudaEvent_t ev;
		cudaStream_t st = slice_stream(ctx, k, &ev);
		CUDA_CHECK(cudaEventRecord(ev, st));
		CUDA_CHECK(cudaStreamWaitEvent(ctx_stream(ctx), ev, 0));
	}
}

/* $VKHEL_LAZY_JOIN=0: join the two streams at the end of every sliced
 * transform (the round-1 behaviour) */
static bool lazy_join() {
	static int v = -1;
	if (v < 0) {
		const char *env = getenv("VKHEL_LAZY_JOIN");
		v = !(env && !strcmp(env, "0"));
	}
	return v;
}

/* $VKHEL_SPLIT_SMALL_MIB (default 8): an RNS batch that fits in L2 is still cut
 * in two limb slices, one per stream, when it holds at least this much: the
 * launch and drain of one slice's kernels overlap the other slice's
 * butterflies (0 = never) */
static size_t split_small_bytes() {
	static double mib = -1;
	if (mib < 0) {
		const char *env = getenv("VKHEL_SPLIT_SMALL_MIB");
		mib = env && *env ? atof(env) : 8;
	}
	return (size_t) (mib * 1048576.0);
}

static bool run_fast_sliced(struct vkhel_ctx *ctx, bool inverse, bool apx,
		const u64 *src, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n) {
	const bool by_limb = limbs > 1;
	bool forced;
	const size_t slice = slice_bytes_setting(by_limb, &forced);
	const fast_plan pl = plan_fast(log2n);
	const size_t poly_bytes = sizeof(u64) << log2n;
	const size_t total = polys * poly_bytes;
	const uint64_t batch = polys / limbs;
	/* units: limbs (all batch entries of each) or, for one modulus, batch
	 * entries */
	const uint64_t units = by_limb ? limbs : batch;
	const size_t unit_bytes = (by_limb ? batch : 1) * poly_bytes;
	uint64_t per = 0;
	if (!pl.kcol || pl.lead || (log2n >= 9 && log2n <= single_max_log2n())) {
		return false;   /* one pass, or a leading generic pass */
	}
	if (slice && total >= 3 * slice
			&& (forced || total > ctx->dev.l2_bytes)) {
		/* larger than L2: slices small enough to stay in it between passes */
		per = slice / unit_bytes;
		if (per < 1) {
			per = 1;
		}
		if (!by_limb && per < 16 && per < units) {
			per = 16 < units ? 16 : units;   /* keep the twiddle sharing of the row pass */
		}
	} else if (!forced && by_limb && units >= 2 && lazy_join()
			&& split_small_bytes() && total >= split_small_bytes()) {
		/* fits in L2: two limb halves, one per stream.  (A single-modulus
		 * batch cut in two batch ranges gains nothing: n = 2^14 x 256, 43.4 us
		 * against 42.1 us forward.) */
		per = (units + 1) / 2;
	}
	const uint64_t nslices = per ? (units + per - 1) / per : 0;
	if (nslices < 2) {
		return false;
	}
	cudaStream_t main_stream = ctx_stream(ctx);
	cudaEvent_t ev = (cudaEvent_t) ctx->dev.ev_scratch;   /* fork */
	struct device_ctx *dev = &ctx->dev;
	const bool continues = dev->split_active && dev->split_dst == (const void *) dst
		&& dev->split_bytes == total && dev->split_per == per
		&& dev->split_units == units && dev->split_by_limb == (int) by_limb
		&& dev->split_log2n == log2n;
	if (!continues) {
		/* (join what another partition has left,) fork: the auxiliary
		 * stream continues from here */
		ntt_split_join(ctx);
		CUDA_CHECK(cudaEventRecord(ev, main_stream));
		for (int k = 1; k < slice_streams(); k++) {
			cudaEvent_t unused;
			CUDA_CHECK(cudaStreamWaitEvent(slice_stream(ctx, k, &unused), ev, 0));
		}
	}
	/* else: slice i follows slice i of the transform before it on the same
	 * stream -- it reads and writes only what that one wrote (or reads a
	 * vector no pending slice writes) */
	dev->split_active = 0;
	dev->in_slices = 1;
	for (uint64_t i = 0; i < nslices; i++) {
		const uint64_t u0 = i * per;
		const uint64_t cnt = units - u0 < per ? units - u0 : per;
		{
			cudaEvent_t unused;
			const int k = (int) (i % (uint64_t) slice_streams());
			ctx->dev.launch_stream = k ? (void *) slice_stream(ctx, k, &unused)
				: NULL;
		}
		if (by_limb) {
			run_fast_any(ctx, inverse, apx, src, dst, descs + u0, cnt,
					cnt * batch, log2n, (unsigned) limbs, (unsigned) u0);
		} else {
			const size_t off = (size_t) u0 << log2n;
			run_fast_any(ctx, inverse, apx, src + off, dst + off, descs, 1,
					cnt, log2n, 0, 0);
		}
	}
	ctx->dev.launch_stream = NULL;
	dev->in_slices = 0;
	if (lazy_join() && !ctx->dev.stream_exposed) {
		/* (a caller that holds the stream orders its own work by stream
		 * position: for it every call joins before it returns.)
		 * The join is left to whoever needs the context's stream next
		 * (defer_flush -> ntt_split_join), or to nobody if the next call is
		 * the same partition of the same vector again */
		dev->split_active = 1;
		dev->split_dst = dst;
		dev->split_bytes = total;
		dev->split_per = per;
		dev->split_units = units;
		dev->split_by_limb = by_limb;
		dev->split_log2n = log2n;
	} else {
		dev->split_active = 1;
		ntt_split_join(ctx);
	}
	return true;
}

/* A forward transform may store lazy values ([0,3q) approximate family,
 * [0,2q) exact) when the in-place inverse transform with the same tables is
 * the only thing that will ever read them: both directions must take the
 * two-pass lazy-butterfly kernels.  $VKHEL_LAZY_FORWARD=0 turns it off. */
bool ntt_lazy_forward_supported(unsigned log2n, uint64_t q_max) {
	static const bool off = getenv("VKHEL_FORCE_GENERIC") != NULL
		|| (getenv("VKHEL_LAZY_FORWARD") && !strcmp(getenv("VKHEL_LAZY_FORWARD"), "0"));
	if (off || q_max >= (1ull << 62) || log2n < 9 || log2n > 30
			|| log2n <= single_max_log2n() || ntt_cluster_enabled(log2n)
			|| only_pass()) {
		return false;
	}
	const fast_plan pl = plan_fast(log2n);
	return pl.kcol && pl.krow;
}

void launch_ntt(struct vkhel_ctx *ctx, bool inverse, const u64 *src, u64 *dst,
		const limb_desc *descs, uint64_t limbs, uint64_t polys,
		unsigned log2n, uint64_t q_max, bool lazy_out) {
	VK_REQUIRE(log2n >= 1 && log2n <= 30, "unsupported transform size 2^%u",
			log2n);
	VK_REQUIRE(q_max < (1ull << 63), "NTT modulus must be below 2^63");
	VK_REQUIRE(!lazy_out || (!inverse && ntt_lazy_forward_supported(log2n, q_max)),
			"internal: lazy forward store of 2^%u points", log2n);
	/* (read by run_fast; cleared again by whoever set it, below) */
	struct lazy_scope {
		struct vkhel_ctx *ctx;
		~lazy_scope() { ctx->dev.lazy_out = 0; }
	} scope = { ctx };
	ctx->dev.lazy_out = lazy_out;
	const bool strict = q_max >= (1ull << 62);
	static const bool force_generic = getenv("VKHEL_FORCE_GENERIC") != NULL;
	if (!strict && log2n >= 3 && !force_generic) {
		const bool apx = use_approx(q_max, log2n);
		if (run_fast_sliced(ctx, inverse, apx, src, dst, descs, limbs, polys,
					log2n)) {
			return;
		}
	}
	/* not a sliced transform: slices a previous one left on the auxiliary
	 * stream (the caller may have held the join back, sliced_operands in
	 * vector.cu) are joined before anything is launched */
	ntt_split_join(ctx);
	if (!strict && log2n >= 3 && !force_generic) {
		const bool apx = use_approx(q_max, log2n);
		run_fast_any(ctx, inverse, apx, src, dst, descs, limbs, polys, log2n,
				0, 0);
		return;
	}
	if (inverse) {
		if (strict) run_generic<true, true>(ctx, src, dst, descs, limbs, polys, log2n);
		else run_generic<true, false>(ctx, src, dst, descs, limbs, polys, log2n);
	} else {
		if (strict) run_generic<false, true>(ctx, src, dst, descs, limbs, polys, log2n);
		else run_generic<false, false>(ctx, src, dst, descs, limbs, polys, log2n);
	}
}

bool launch_ntt_inverse_of_product(struct vkhel_ctx *ctx, const u64 *src,
		const u64 *src2, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, uint64_t q_max, uint64_t fma_mult) {
	static const bool force_generic = getenv("VKHEL_FORCE_GENERIC") != NULL;
	if (q_max >= (1ull << 62) || log2n < 3 || force_generic) {
		return false;   /* generic path: no fused product */
	}
	VK_REQUIRE(!fma_mult || limbs == 1, "internal: fused fma is single-modulus");
	if (use_approx(q_max, log2n)) {
		run_fast<true, true>(ctx, src, dst, descs, limbs, polys, log2n, src2,
				NULL, 0, 0, NULL, fma_mult);
	} else {
		run_fast<true, false>(ctx, src, dst, descs, limbs, polys, log2n, src2,
				NULL, 0, 0, NULL, fma_mult);
	}
	return true;
}

bool launch_ntt_polymul(struct vkhel_ctx *ctx, const u64 *a, const u64 *b,
		u64 *tmp, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, uint64_t q_max) {
	static const bool force_generic = getenv("VKHEL_FORCE_GENERIC") != NULL;
	static const bool unfused = getenv("VKHEL_POLYMUL_UNFUSED") != NULL;
	if (q_max >= (1ull << 62) || log2n < 3 || force_generic || unfused) {
		return false;
	}
	if (use_approx(q_max, log2n)) {
		run_fast_polymul<true>(ctx, a, b, tmp, dst, descs, limbs, polys, log2n);
	} else {
		run_fast_polymul<false>(ctx, a, b, tmp, dst, descs, limbs, polys, log2n);
	}
	return true;
}

bool ntt_indirect_supported(unsigned log2n, uint64_t q) {
	static const bool force_generic = getenv("VKHEL_FORCE_GENERIC") != NULL;
	return !force_generic && q < (1ull << 62) && log2n >= 3 && log2n <= 18;
}

bool ntt_indirect_product_supported(unsigned log2n, uint64_t q) {
	/* two-pass sizes: below them the whole product is one CTA
	 * (kernels_ntt_small.cu), and the single-pass kernel has no indirect
	 * product */
	return ntt_indirect_supported(log2n, q) && log2n > single_max_log2n()
		&& log2n > 8;
}

void launch_ntt_indirect(struct vkhel_ctx *ctx, bool inverse,
		const ntt_ptrs *tab, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, uint64_t q_max,
		const ntt_ptrs *host_tab, bool product) {
	VK_REQUIRE(ntt_indirect_supported(log2n, q_max),
			"internal: indirect batch outside the fast path");
	VK_REQUIRE((tab != NULL) != (host_tab != NULL),
			"internal: indirect batch needs exactly one pointer table");
	VK_REQUIRE(!product || (inverse && ntt_indirect_product_supported(log2n, q_max)),
			"internal: indirect product outside its range");
	if (use_approx(q_max, log2n)) {
		if (inverse) run_fast<true, true>(ctx, NULL, NULL, descs, limbs, polys, log2n, NULL, tab, 0, 0, host_tab, 0, product);
		else run_fast<false, true>(ctx, NULL, NULL, descs, limbs, polys, log2n, NULL, tab, 0, 0, host_tab);
	} else {
		if (inverse) run_fast<true, false>(ctx, NULL, NULL, descs, limbs, polys, log2n, NULL, tab, 0, 0, host_tab, 0, product);
		else run_fast<false, false>(ctx, NULL, NULL, descs, limbs, polys, log2n, NULL, tab, 0, 0, host_tab);
	}
}
