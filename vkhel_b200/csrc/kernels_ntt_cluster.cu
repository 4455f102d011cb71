/*
 * Single-pass transform for 2^14 <= n <= 2^16 on a thread-block cluster with
 * distributed shared memory (BASELINE north_star: "single-pass when a whole
 * polynomial fits in shared memory").
 *
 * A polynomial of n = 2^(13+CB) coefficients does not fit the 227 KB of one
 * SM, but it fits the shared memory of a cluster of C = 2^CB CTAs (C = 2, 4,
 * 8): block c of 8192 consecutive coefficients lives in CTA c, next to the
 * twiddle subtree of that block (128 KB: the sub-heap rooted at node C + c).
 * The stage structure (index facts at the top of kernels_ntt.cu) splits as
 *
 *   top CB stages   pair coefficients of DIFFERENT blocks: (k, j) with
 *                   (k ^ half, j), same position j in blocks k.  Per position
 *                   j that is a C-point transform over k with the C - 1
 *                   twiddles at the top of the heap.
 *   low 13 stages   stay inside a block: the radix-8 engine of the one-CTA
 *                   kernel (ntt_single_kernel), 1024 threads x 8 coefficients,
 *                   exchanges through the CTA's own shared memory.
 *
 * forward: CTA c reads the positions j of ITS share (j in [c*8192/C, ...)) of
 * all C blocks straight from global memory (coalesced: consecutive threads,
 * consecutive j), runs the top stages in registers and writes value (k, j)
 * into the exchange buffer of CTA k -- an all-to-all through DSMEM
 * (st.shared::cluster) --, cluster barrier, then every CTA runs the low 13
 * stages on its block and stores it.  inverse: the mirror image -- low stages
 * first, block left in shared memory, cluster barrier, CTA c gathers (k, j)
 * for its positions from the C blocks (ld.shared::cluster), runs the top
 * stages with n^-1 folded into the last one and stores.
 *
 * One launch and 16n bytes of global traffic per transform instead of two
 * launches and 32n (of which the second 16n are L2 hits in the two-pass
 * path).  Reference semantics: src/vector.c:536-566 (forward), :599-639
 * (inverse + n^-1), butterflies nttfwdbutterfly.comp:41-57 /
 * nttrevbutterfly.comp:41-57; outputs are canonical, hence bit-identical.
 *
 * Whether this path is used: $VKHEL_CLUSTER=1 (default off; measured against
 * the two-pass path in profiles/r02_cluster.txt and DESIGN.md 5.8).
 */
#include <string.h>

#include "ntt_device.cuh"

#define CL_K 13                     /* stages inside a block */
#define CL_BLOCK (1 << CL_K)        /* coefficients per CTA */
#define CL_THREADS (CL_BLOCK / 8)

struct cluster_pass {
	const u64 *src;
	u64 *dst;
	const limb_desc *descs;
	unsigned limbs;        /* limbs of this launch */
	unsigned limbs_total;  /* layout [batch][limbs_total][n] */
	unsigned limb0;
	unsigned batch;
	unsigned bchunk;       /* batch entries per cluster */
	u64 zero;
};

__device__ __forceinline__ unsigned cluster_ctarank() {
	unsigned r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}

__device__ __forceinline__ void cluster_arrive() {
	asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}

__device__ __forceinline__ void cluster_wait() {
	asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

/* address of the same shared-memory location in CTA `rank` of the cluster */
__device__ __forceinline__ unsigned dsmem_addr(const void *local, unsigned rank) {
	const unsigned a = (unsigned) __cvta_generic_to_shared(local);
	unsigned r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
	return r;
}

__device__ __forceinline__ void dsmem_st(unsigned addr, u64 v) {
	asm volatile("st.shared::cluster.b64 [%0], %1;" :: "r"(addr), "l"(v) : "memory");
}

__device__ __forceinline__ u64 dsmem_ld(unsigned addr) {
	u64 v;
	asm volatile("ld.shared::cluster.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
	return v;
}

/* the C-point transform over the block index k of the top CB stages, on the
 * registers v[k]; tw[node] = the top of the twiddle heap (nodes 1 .. C-1).
 * forward: stages u = 0 .. CB-1; inverse: u = CB-1 .. 1 here, stage 0 (with
 * n^-1) by the caller */
template <int CB, bool INV, bool APX>
__device__ __forceinline__ void top_stages(u64 (&v)[1 << CB],
		const ulonglong2 *tw, u64 q, u64 bq, u64 zr) {
	constexpr int C = 1 << CB;
#pragma unroll
	for (int step = 0; step < CB; step++) {
		const int u = INV ? CB - 1 - step : step;
		if (INV && u == 0) {
			break;
		}
		const int half = 1 << (CB - 1 - u);
#pragma unroll
		for (int k = 0; k < C; k++) {
			if (k & half) {
				continue;
			}
			const ulonglong2 w = tw[(1 << u) + (k >> (CB - u))];
			if (INV) {
				if (APX) gs_lazy3(v[k], v[k | half], w.x, w.y, q, bq, zr);
				else gs_lazy(v[k], v[k | half], w.x, w.y, q, bq, zr);
			} else {
				if (APX) ct_lazy3(v[k], v[k | half], w.x, w.y, q, bq, zr);
				else ct_lazy(v[k], v[k | half], w.x, w.y, q, bq, zr);
			}
		}
	}
}

template <bool INV, int CB, bool APX>
__global__ void __launch_bounds__(CL_THREADS, 1)
ntt_cluster_kernel(const __grid_constant__ cluster_pass p) {
	using G = tile_geom<CL_K>;
	constexpr int C = 1 << CB;
	constexpr int L = CL_K + CB;
	constexpr int COLS = 8 / C;           /* positions j per thread in the top stages */
	extern __shared__ __align__(16) unsigned char smem_raw[];
	ulonglong2 *sm_tw = (ulonglong2 *) smem_raw;                   /* [2^13] */
	u64 *sm_x = (u64 *) (sm_tw + CL_BLOCK);                        /* xpad(2^13) words */
	__shared__ ulonglong2 sm_top[C];                               /* heap nodes 1 .. C-1 */
	__shared__ __align__(8) u64 tw_bar;

	const unsigned c = cluster_ctarank();
	const unsigned cluster_id = blockIdx.x >> CB;
	const unsigned limb = cluster_id % p.limbs;
	const unsigned bc = cluster_id / p.limbs;
	const unsigned b0 = bc * p.bchunk;
	const unsigned nb = p.batch - b0 < p.bchunk ? p.batch - b0 : p.bchunk;

	const limb_desc &d = p.descs[limb];
	const u64 q = d.q, bq = APX ? 3 * q : 2 * q;
	const ulonglong2 *tw_g = d.tw + (INV ? ((u64) 1 << L) : 0);
	const int t = threadIdx.x;
	if (t == 0) {
		mbar_init(&tw_bar, 1);
	}
	if (t < C) {
		sm_top[t] = tw_g[t ? t : 1];
	}
	__syncthreads();
	if (t == 0) {
		/* the block's sub-heap: root C + c, 13 levels */
		stage_twiddles_tma<CL_K>(sm_tw, tw_g, CB, c, 1, &tw_bar);
	}
	pdl_wait();
	/* every CTA of the cluster is running before anyone touches a peer */
	cluster_arrive();
	cluster_wait();

	constexpr int first = INV ? G::rounds - 1 : 0;
	constexpr int last = INV ? 0 : G::rounds - 1;
	const int tb_first = G::tbase(first, t), tb_last = G::tbase(last, t);
	ulonglong2 fold_a = make_ulonglong2(d.inv_n, d.inv_n_shoup);
	ulonglong2 fold_b = make_ulonglong2(d.inv_w1n, d.inv_w1n_shoup);
	const u64 zr = p.zero;
	bool tw_ready = false;
	/* positions of this CTA's share in the top stages: j = j0 + t + m*1024 */
	const unsigned j0 = c * (CL_BLOCK / C);

	for (unsigned bl = 0; bl < nb; bl++) {
		const u64 poly = (u64) (b0 + bl) * p.limbs_total + p.limb0 + limb;
		const u64 *sp = p.src + (poly << L);
		u64 *dp = p.dst + (poly << L);
		u64 x[1][8];

		if (!INV) {
			/* ---- top stages: C blocks x COLS positions per thread ---------- */
#pragma unroll
			for (int m = 0; m < COLS; m++) {
				const unsigned j = j0 + t + m * CL_THREADS;
				u64 v[C];
#pragma unroll
				for (int k = 0; k < C; k++) {
					v[k] = sp[((u64) k << CL_K) + j];
				}
				top_stages<CB, false, APX>(v, sm_top, q, bq, zr);
				/* all-to-all: (k, j) goes to block k */
#pragma unroll
				for (int k = 0; k < C; k++) {
					dsmem_st(dsmem_addr(sm_x + xpad(j), k), v[k]);
				}
			}
			cluster_arrive();
			cluster_wait();
			/* ---- low 13 stages on this CTA's block -------------------------- */
#pragma unroll
			for (int e = 0; e < 8; e++) {
				x[0][e] = sm_x[xpad(tb_first + G::eoff(first, e))];
			}
			/* (the first exchange below writes exactly the words this thread
			 * has just read: no barrier needed) */
		} else {
#pragma unroll
			for (int e = 0; e < 8; e++) {
				x[0][e] = sp[((u64) c << CL_K) + tb_first + G::eoff(first, e)];
			}
		}
		if (!tw_ready) {
			mbar_wait(&tw_bar, 0);
			tw_ready = true;
		}
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
			tile_round<CL_K, INV, FOLD_NONE, 1, APX>(x, r, t, sm_tw, q, bq, fold_a,
					fold_b, nullptr, zr);
		});
		if (!INV) {
			/* peers may start the next polynomial's all-to-all into this
			 * CTA's buffer once every thread here has done its last read of it
			 * (each thread arrives after its own) */
			cluster_arrive();
#pragma unroll
			for (int e = 0; e < 8; e++) {
				dp[((u64) c << CL_K) + tb_last + G::eoff(last, e)] =
					tile_canon<false, APX>(x[0][e], q, bq);
			}
			cluster_wait();
		} else {
			/* ---- block left in shared memory: every thread overwrites the
			 * words it read in the last exchange (layout of round 0) ---------- */
#pragma unroll
			for (int e = 0; e < 8; e++) {
				sm_x[xpad(tb_last + G::eoff(last, e))] = x[0][e];
			}
			cluster_arrive();
			cluster_wait();
			/* ---- top stages: gather (k, j) from the C blocks ---------------- */
#pragma unroll
			for (int m = 0; m < COLS; m++) {
				const unsigned j = j0 + t + m * CL_THREADS;
				u64 v[C];
#pragma unroll
				for (int k = 0; k < C; k++) {
					v[k] = dsmem_ld(dsmem_addr(sm_x + xpad(j), k));
				}
				top_stages<CB, true, APX>(v, sm_top, q, bq, zr);
				/* stage 0 with n^-1: both outputs multiplied (exact quotient:
				 * results below 2q, one subtraction to the canonical residue) */
				constexpr int half = C / 2;
#pragma unroll
				for (int k = 0; k < half; k++) {
					const u64 s = v[k] + v[k | half];
					const u64 df = v[k] - v[k | half] + bq;
					v[k] = csub(shoup_lazy(s, fold_a.x, fold_a.y, q), q);
					v[k | half] = csub(shoup_lazy(df, fold_b.x, fold_b.y, q), q);
				}
#pragma unroll
				for (int k = 0; k < C; k++) {
					dp[((u64) k << CL_K) + j] = v[k];
				}
			}
			/* the next polynomial overwrites the buffers the peers have been
			 * reading */
			cluster_arrive();
			cluster_wait();
		}
	}
	pdl_launch_dependents();
}

static size_t cluster_smem_bytes() {
	return (sizeof(ulonglong2) << CL_K) + sizeof(u64) * (size_t) (xpad(CL_BLOCK) + 4);
}

template <bool INV, int CB, bool APX>
static void run_cluster(struct vkhel_ctx *ctx, cluster_pass p) {
	const size_t smem = cluster_smem_bytes();
	/* batch entries per cluster: reuse the staged twiddles, but keep at least
	 * four clusters per resident slot of the device */
	const unsigned resident = (unsigned) ctx->dev.sm_count >> CB;
	unsigned bchunk = (p.batch * p.limbs) / (4 * (resident ? resident : 1));
	bchunk = bchunk < 1 ? 1 : bchunk > 16 ? 16 : bchunk;
	bchunk = bchunk > p.batch ? p.batch : bchunk;
	p.bchunk = bchunk;
	const unsigned clusters = ((p.batch + bchunk - 1) / bchunk) * p.limbs;
	auto kernel = ntt_cluster_kernel<INV, CB, APX>;
	CUDA_CHECK(cudaFuncSetAttribute(kernel,
				cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(clusters << CB);
	cfg.blockDim = dim3(CL_THREADS);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = ctx->dev.launch_stream ? (cudaStream_t) ctx->dev.launch_stream
		: ctx_stream(ctx);
	cudaLaunchAttribute attr[2];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = 1 << CB;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[1].val.programmaticStreamSerializationAllowed = FAST_PDL;
	cfg.attrs = attr;
	cfg.numAttrs = 2;
	CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, p));
	ctx->dev.launches++;
}

/* $VKHEL_CLUSTER=1 routes 2^14 <= n <= 2^16 (lazy-path moduli) through the
 * cluster kernel; default off */
bool ntt_cluster_enabled(unsigned log2n) {
	static int on = -1;
	if (on < 0) {
		const char *env = getenv("VKHEL_CLUSTER");
		on = env && *env && strcmp(env, "0") != 0;
	}
	return on && log2n >= CL_K + 1 && log2n <= CL_K + 3;
}

void launch_ntt_cluster(struct vkhel_ctx *ctx, bool inverse, bool apx,
		const u64 *src, u64 *dst, const limb_desc *descs, uint64_t limbs,
		uint64_t polys, unsigned log2n, unsigned limbs_total, unsigned limb0) {
	VK_REQUIRE(log2n >= CL_K + 1 && log2n <= CL_K + 3,
			"internal: cluster transform of 2^%u points", log2n);
	VK_REQUIRE(cluster_smem_bytes() + 1024 <= ctx->dev.smem_optin,
			"cluster transform needs %zu bytes of shared memory per CTA",
			cluster_smem_bytes());
	cluster_pass p;
	p.src = src;
	p.dst = dst;
	p.descs = descs;
	p.limbs = (unsigned) limbs;
	p.limbs_total = limbs_total ? limbs_total : (unsigned) limbs;
	p.limb0 = limb0;
	p.batch = (unsigned) (polys / limbs);
	p.bchunk = 1;
	p.zero = 0;
	const int cb = (int) log2n - CL_K;
#define CLUSTER_CASE(CB_) \
	case CB_: \
		if (inverse) { \
			if (apx) run_cluster<true, CB_, true>(ctx, p); \
			else run_cluster<true, CB_, false>(ctx, p); \
		} else { \
			if (apx) run_cluster<false, CB_, true>(ctx, p); \
			else run_cluster<false, CB_, false>(ctx, p); \
		} \
		break;
	switch (cb) {
	CLUSTER_CASE(1)
	CLUSTER_CASE(2)
	CLUSTER_CASE(3)
	default: VK_DIE("internal: cluster of 2^%d CTAs", cb);
	}
#undef CLUSTER_CASE
}
