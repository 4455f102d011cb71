/*
 * Column pass with a TMA tensor-map tile load (BASELINE north_star: "staged
 * tile-by-tile into shared memory (TMA bulk copies where tile shape allows)";
 * "the tile [2^k1 rows x C cols] is a strided 2-D box -> a TMA
 * tensor-map tile").
 *
 * The shipped column kernel (kernels_ntt.cu, ntt_cols_kernel) brings its tile
 * -- 256 rows at stride 256 coefficients by 16 adjacent columns for n = 2^16 --
 * into registers with eight 128-bit loads per thread and as many 64-bit
 * address computations.  Here one thread issues ONE cp.async.bulk.tensor.2d
 * (SASS UTMALDG) for the whole 32 KB box: the batch is described to the TMA
 * unit as a 2-D tensor of 64-bit elements, [polys * 256 rows][256 columns],
 * and the box [256][16] lands in shared memory in exactly the layout of the
 * kernel's exchange buffer; the threads then take their coefficients from
 * shared memory.  Everything after the load is the shipped kernel's code
 * (same engine, same exchanges, same stores).
 *
 * Scope of the variant: the forward transform's column pass of n = 2^16
 * (8 + 8 stages), direct batches.  $VKHEL_COLS_TMA=1 selects it; default off
 * (measured against the register loads, profiles/r02_cols_tma.txt).
 * Reference semantics as in kernels_ntt.cu.
 */
#include <cuda.h>
#include <string.h>

#include "ntt_device.cuh"

#define TMA_K 8          /* stages of the pass */
#define TMA_CL 4         /* log2 columns per CTA */
#define TMA_LB 8         /* log2 row stride: the row pass below holds 8 stages */
#define TMA_THREADS (1 << (TMA_K - 3 + TMA_CL - 1))

struct tma_pass {
	u64 *dst;
	const limb_desc *descs;
	unsigned limbs, limbs_total, limb0;
};

__device__ __forceinline__ void tma_load_2d(void *smem, const CUtensorMap *map,
		int x, int y, u64 *bar) {
	const unsigned dst = (unsigned) __cvta_generic_to_shared(smem);
	const unsigned mb = (unsigned) __cvta_generic_to_shared(bar);
	asm volatile(
		"cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
		"[%0], [%1, {%2, %3}], [%4];"
		:: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(mb) : "memory");
}

template <bool APX>
__global__ void __launch_bounds__(TMA_THREADS, 1024 / TMA_THREADS)
ntt_cols_tma_kernel(const __grid_constant__ CUtensorMap map,
		const __grid_constant__ tma_pass p) {
	constexpr int K = TMA_K, CL = TMA_CL, NP = 2;
	using G = tile_geom<K>;
	extern __shared__ __align__(128) unsigned char smem_raw[];
	u64 *sm_x = (u64 *) smem_raw;                            /* [2^K][2^CL], TMA target */
	ulonglong2 *sm_tw = (ulonglong2 *) (sm_x + (1 << (K + CL)));   /* [2^K] */
	__shared__ __align__(8) u64 bars[2];

	/* blockIdx.x -> (polynomial of the launch, column group) */
	const unsigned cg = blockIdx.x & ((1u << (TMA_LB - CL)) - 1);
	const unsigned lpoly = blockIdx.x >> (TMA_LB - CL);
	const unsigned pb = lpoly / p.limbs, pl = lpoly - pb * p.limbs;
	const u64 poly = (u64) pb * p.limbs_total + p.limb0 + pl;
	const limb_desc &d = p.descs[pl];
	const u64 q = d.q, bq = APX ? 3 * q : 2 * q;

	if (threadIdx.x == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		stage_twiddles_tma<K>(sm_tw, d.tw, 0, 0, 1, &bars[0]);
	}
	pdl_wait();   /* the coefficients may come from the previous kernel */
	if (threadIdx.x == 0) {
		/* the whole tile: rows poly*256 .. +255, columns cg*16 .. +15 */
		mbar_expect_tx(&bars[1], sizeof(u64) << (K + CL));
		tma_load_2d(sm_x, &map, (int) (cg << CL), (int) (poly << K), &bars[1]);
	}
	const int c = (threadIdx.x & ((1 << (CL - 1)) - 1)) * NP;
	const int t = threadIdx.x >> (CL - 1);
	mbar_wait(&bars[1], 0);
	u64 x[NP][8];
	{
		const u64 *xr = sm_x + (G::tbase(0, t) << CL) + c;
#pragma unroll
		for (int e = 0; e < 8; e++) {
			const ulonglong2 v = *(const ulonglong2 *) (xr + (G::eoff(0, e) << CL));
			x[0][e] = v.x;
			x[1][e] = v.y;
		}
	}
	mbar_wait(&bars[0], 0);
	const ulonglong2 none = make_ulonglong2(0, 0);
	static_for<0, G::rounds>([&](auto rc) {
		constexpr int r = decltype(rc)::value;
		if constexpr (r > 0) {
			/* (the first exchange writes the words this thread has read from
			 * the TMA-filled buffer: no barrier in between) */
			u64 *xw = sm_x + (G::tbase(r - 1, t) << CL) + c;
#pragma unroll
			for (int e = 0; e < 8; e++) {
				*(ulonglong2 *) (xw + (G::eoff(r - 1, e) << CL)) =
					make_ulonglong2(x[0][e], x[1][e]);
			}
			__syncthreads();
			const u64 *xr = sm_x + (G::tbase(r, t) << CL) + c;
#pragma unroll
			for (int e = 0; e < 8; e++) {
				const ulonglong2 v = *(const ulonglong2 *) (xr + (G::eoff(r, e) << CL));
				x[0][e] = v.x;
				x[1][e] = v.y;
			}
		}
		tile_round<K, false, FOLD_NONE, NP, APX>(x, r, t, sm_tw, q, bq, none, none);
	});
	pdl_launch_dependents();
	u64 *dp = p.dst + (poly << (K + TMA_LB)) + (cg << CL) + c
		+ ((u64) G::tbase(G::rounds - 1, t) << TMA_LB);
#pragma unroll
	for (int e = 0; e < 8; e++) {
		*(ulonglong2 *) (dp + ((u64) G::eoff(G::rounds - 1, e) << TMA_LB)) =
			make_ulonglong2(x[0][e], x[1][e]);
	}
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType,
		cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
		const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
		CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn encode_tiled() {
	static encode_tiled_fn fn = NULL;
	if (!fn) {
		void *sym = NULL;
		cudaDriverEntryPointQueryResult res;
		CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym,
					cudaEnableDefault, &res));
		VK_REQUIRE(sym && res == cudaDriverEntryPointSuccess,
				"cuTensorMapEncodeTiled is not available in this driver");
		fn = (encode_tiled_fn) sym;
	}
	return fn;
}

/* $VKHEL_COLS_TMA=1: the forward column pass of n = 2^16 loads its tiles by
 * TMA tensor map (default off) */
bool ntt_cols_tma_enabled(unsigned log2n, unsigned kcol, unsigned s0) {
	static int on = -1;
	if (on < 0) {
		const char *env = getenv("VKHEL_COLS_TMA");
		on = env && *env && strcmp(env, "0") != 0;
	}
	return on && log2n == TMA_K + TMA_LB && kcol == TMA_K && s0 == 0;
}

void launch_ntt_cols_tma(struct vkhel_ctx *ctx, bool apx, const u64 *src,
		u64 *dst, const limb_desc *descs, uint64_t limbs, uint64_t polys,
		unsigned limbs_total, unsigned limb0) {
	const u64 batch = polys / limbs;
	const u64 lt = limbs_total ? limbs_total : limbs;
	/* the whole [batch][limbs_total] vector as rows of 256 coefficients */
	const cuuint64_t dims[2] = { 1u << TMA_LB, (cuuint64_t) (batch * lt) << TMA_K };
	const cuuint64_t strides[1] = { sizeof(u64) << TMA_LB };
	const cuuint32_t box[2] = { 1u << TMA_CL, 1u << TMA_K };
	const cuuint32_t estr[2] = { 1, 1 };
	CUtensorMap map;
	const CUresult res = encode_tiled()(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2,
			(void *) src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
			CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
			CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	VK_REQUIRE(res == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int) res);
	tma_pass p;
	p.dst = dst;
	p.descs = descs;
	p.limbs = (unsigned) limbs;
	p.limbs_total = (unsigned) lt;
	p.limb0 = limb0;
	const size_t smem = (sizeof(u64) << (TMA_K + TMA_CL)) + (sizeof(ulonglong2) << TMA_K);
	const u64 blocks = polys << (TMA_LB - TMA_CL);
	VK_REQUIRE(blocks <= 0x7fffffffull, "transform too large for one launch");
	void (*kernel)(CUtensorMap, tma_pass) = apx ? ntt_cols_tma_kernel<true>
		: ntt_cols_tma_kernel<false>;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned) blocks);
	cfg.blockDim = dim3(TMA_THREADS);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = ctx->dev.launch_stream ? (cudaStream_t) ctx->dev.launch_stream
		: ctx_stream(ctx);
	cudaLaunchAttribute attr;
	attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr.val.programmaticStreamSerializationAllowed = FAST_PDL;
	cfg.attrs = &attr;
	cfg.numAttrs = 1;
	CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, map, p));
	ctx->dev.launches++;
}
