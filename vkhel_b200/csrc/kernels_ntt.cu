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
 * n^-1 scaling of the inverse is folded into its last stage (twiddles n^-1 and
 * inv_root[1] * n^-1).
 */
#include "common.cuh"

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
 * Dispatch
 * ====================================================================================== */
void launch_ntt(struct vkhel_ctx *ctx, bool inverse, const u64 *src, u64 *dst,
		const limb_desc *descs, uint64_t limbs, uint64_t polys,
		unsigned log2n, uint64_t q_max) {
	VK_REQUIRE(log2n >= 1 && log2n <= 30, "unsupported transform size 2^%u",
			log2n);
	VK_REQUIRE(q_max < (1ull << 63), "NTT modulus must be below 2^63");
	const bool strict = q_max >= (1ull << 62);
	if (inverse) {
		if (strict) run_generic<true, true>(ctx, src, dst, descs, limbs, polys, log2n);
		else run_generic<true, false>(ctx, src, dst, descs, limbs, polys, log2n);
	} else {
		if (strict) run_generic<false, true>(ctx, src, dst, descs, limbs, polys, log2n);
		else run_generic<false, false>(ctx, src, dst, descs, limbs, polys, log2n);
	}
}
