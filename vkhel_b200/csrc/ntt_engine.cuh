/*
 * Register-resident radix-8 NTT engine (device side).
 *
 * A "tile" is a 2^K-point sub-transform: the K consecutive stages of one pass
 * (see the index facts at the top of kernels_ntt.cu).  It is computed by a
 * group of 2^(K-3) threads that hold 8 coefficients each.  The K stages are cut
 * into rounds of up to 3; inside a round a thread owns all 8 coefficients that
 * differ in the round's 3 "in-thread" index bits, so the round's butterflies
 * are pure register work (12 butterflies per thread for a full round).
 * Between rounds the group re-distributes coefficients through shared memory.
 *
 * Bit bookkeeping for tile index i (K bits), forward direction:
 *   stage u pairs on bit K-1-u and uses twiddle node (1<<u) + (i >> (K-u))
 *   round r covers stages 3r .. 3r+cnt-1   (cnt = min(3, K-3r))
 *   full round   : in-thread bits are p+2,p+1,p with p = K-3(r+1)
 *                  i = ((t >> p) << (p+3)) | (e << p) | (t & ((1<<p)-1))
 *   partial round: (last, cnt = 1 or 2) pair bits cnt-1..0; the other 3-cnt
 *                  in-thread bits are the top bits of i (passengers)
 *                  i = ((e >> cnt) << (K-3+cnt)) | (t << cnt) | (e & ((1<<cnt)-1))
 * with t the thread's number inside its group and e = 0..7 the register.
 * The inverse direction walks the same rounds and stages backwards.
 *
 * A thread may carry NP = 2 such tiles at once when they share their twiddles
 * (two batch entries of one limb, or two adjacent columns): the shared-memory
 * twiddle fetch, the dominant non-arithmetic cost (tools/bfly_bench.cu: 2.46
 * butterflies/clk/SM with an LDS.128 per butterfly against 2.85 without), is
 * then paid once per two butterflies.
 *
 * Arithmetic: Harvey lazy butterflies (modarith.cuh).  Exact quotient (APX =
 * false, q < 2^62): forward values stay in [0,4q), inverse values in [0,2q).
 * Approximate quotient (APX = true, 6q < 2^64): [0,6q) and [0,3q).  `bq` is
 * the butterfly's bound, 2q or 3q.
 */
#ifndef VKHEL_NTT_ENGINE_CUH
#define VKHEL_NTT_ENGINE_CUH

#include "modarith.cuh"

template <int K>
struct tile_geom {
	static constexpr int rounds = (K + 2) / 3;
	static constexpr int group_log2 = K - 3;          /* threads per tile */
	static constexpr int last_cnt = K - 3 * (rounds - 1);

	__host__ __device__ static constexpr int cnt(int r) {
		return r == rounds - 1 ? last_cnt : 3;
	}
	__host__ __device__ static constexpr bool full(int r) {
		return cnt(r) == 3;
	}
	/* The tile index of register e of group-thread t in round r is
	 * tbase(r,t) + eoff(r,e): the two parts occupy disjoint bits, so every
	 * address derived from it is "per-thread base + compile-time constant"
	 * and the loads/stores carry immediate offsets. */
	__device__ __forceinline__ static int tbase(int r, int t) {
		if (full(r)) {
			const int p = K - 3 * (r + 1);
			return ((t >> p) << (p + 3)) | (t & ((1 << p) - 1));
		}
		return t << last_cnt;
	}
	__host__ __device__ static constexpr int eoff(int r, int e) {
		if (full(r)) {
			return e << (K - 3 * (r + 1));
		}
		return ((e >> last_cnt) << (K - 3 + last_cnt))
			| (e & ((1 << last_cnt) - 1));
	}
	__device__ __forceinline__ static int index(int r, int t, int e) {
		return tbase(r, t) + eoff(r, e);
	}

	/* Twiddle group g = i >> (K-u) of stage j of round r, same split:
	 * gbase(r,j,t) + goff(r,j,e) (e: the butterfly's upper register) */
	__device__ __forceinline__ static int gbase(int r, int j, int t) {
		if (full(r)) {
			return (t >> (K - 3 * (r + 1))) << j;
		}
		return t << j;
	}
	__host__ __device__ static constexpr int goff(int r, int j, int e) {
		if (full(r)) {
			return e >> (3 - j);
		}
		const int c = last_cnt;
		return ((e >> c) << (K - 3 + j)) | ((e & ((1 << c) - 1)) >> (c - j));
	}
};

/* How the inverse transform's n^-1 factor is applied (inverse only; the pass
 * that holds global stage 0, i.e. local stage 0 of a tile rooted at node 1):
 *   FOLD_NONE   not in this pass
 *   FOLD_LAST   in the last stage: both outputs of every butterfly of stage 0
 *               are multiplied (n^-1 and inv_root[1] * n^-1) -- n/2 extra
 *               products per transform
 *   FOLD_TWID   in the twiddles.  A Gentleman-Sande butterfly multiplies only
 *               its difference, so a coefficient is "unscaled" as long as it
 *               has been on the sum side of every stage of this pass so far,
 *               i.e. while the index bits paired so far are all zero.  A
 *               butterfly whose inputs are unscaled takes its twiddle from the
 *               table of inv_root * n^-1 (`tws`): its difference leaves scaled,
 *               its sum stays unscaled.  Both inputs of a butterfly always have
 *               the same history, and scaled inputs use the plain twiddle.
 *               After the last stage exactly one coefficient of the tile
 *               (index 0) is still unscaled and gets the one explicit product
 *               (by the kernel, when it stores).
 *               Which butterflies are concerned is static per register (the
 *               pair bits of this round below the current one are zero) and per
 *               thread (the bits paired in earlier rounds are zero), so the
 *               only run-time cost is one pointer select per round. */
enum { FOLD_NONE = 0, FOLD_LAST = 1, FOLD_TWID = 2 };

/* One round of butterflies on NP interleaved tiles x[p][0..7] that share their
 * twiddles (NP batch entries of the same limb and tile position, or NP adjacent
 * columns): every (w, w') pair fetched from shared memory feeds NP butterflies.
 *   twt: the tile's twiddle subtree in shared memory, twt[node] = (w, w')
 *   tws: FOLD_TWID only -- the same subtree of the scaled inverse table
 *   FOLD: see above; fold_a = n^-1, fold_b = inv_root[1] * n^-1
 *   APX:  butterflies around the approximate Shoup product (bq = 3q)
 *   zr:   an opaque zero (modarith.cuh, ct_lazy) */
template <int K, bool INV, int FOLD, int NP, bool APX>
__device__ __forceinline__ void tile_round(u64 (&x)[NP][8], int r, int t,
		const ulonglong2 *twt, u64 q, u64 bq, ulonglong2 fold_a,
		ulonglong2 fold_b, const ulonglong2 *tws = nullptr, u64 zr = 0) {
	using G = tile_geom<K>;
	static_assert(INV || FOLD == FOLD_NONE, "only the inverse carries n^-1");
	const int cnt = G::cnt(r);
	/* FOLD_TWID: the index bits paired in earlier rounds are the low
	 * K - 3(r+1) bits of the tile index, all taken from t (none for the round
	 * the inverse starts with) */
	const int hist_bits = G::full(r) ? K - 3 * (r + 1) : 0;
	const bool unscaled_thread = FOLD == FOLD_TWID
		&& (t & ((1 << hist_bits) - 1)) == 0;
	const ulonglong2 *twu = FOLD == FOLD_TWID && unscaled_thread ? tws : twt;
#pragma unroll
	for (int step = 0; step < 3; step++) {
		if (step >= cnt) {
			break;
		}
		const int j = INV ? cnt - 1 - step : step;  /* stage within the round */
		const int u = 3 * r + j;                    /* local stage */
		const int beta = cnt - 1 - j;               /* pair bit inside e */
		const ulonglong2 *twp = twt + (1 << u) + G::gbase(r, j, t);
		const ulonglong2 *twpu = twu + (1 << u) + G::gbase(r, j, t);
#pragma unroll
		for (int e = 0; e < 8; e++) {
			if (e & (1 << beta)) {
				continue;
			}
			/* inputs that have only seen sums inside this round so far */
			const bool upos = FOLD == FOLD_TWID && (e & ((1 << beta) - 1)) == 0;
			if (INV && FOLD == FOLD_LAST && u == 0) {
#pragma unroll
				for (int p = 0; p < NP; p++) {
					u64 &X = x[p][e];
					u64 &Y = x[p][e | (1 << beta)];
					const u64 s = X + Y;
					const u64 d = X - Y + bq;
					/* exact quotient here even in the approximate family: the
					 * outputs of this stage are the transform's outputs, and
					 * [0,2q) needs one conditional subtraction instead of two */
					X = shoup_lazy(s, fold_a.x, fold_a.y, q);
					Y = shoup_lazy(d, fold_b.x, fold_b.y, q);
				}
			} else {
				const ulonglong2 w = upos ? twpu[G::goff(r, j, e)]
					: twp[G::goff(r, j, e)];
#pragma unroll
				for (int p = 0; p < NP; p++) {
					u64 &X = x[p][e];
					u64 &Y = x[p][e | (1 << beta)];
					if (INV) {
						if (APX) gs_lazy3(X, Y, w.x, w.y, q, bq, zr);
						else gs_lazy(X, Y, w.x, w.y, q, bq, zr);
					} else {
						if (APX) ct_lazy3(X, Y, w.x, w.y, q, bq, zr);
						else ct_lazy(X, Y, w.x, w.y, q, bq, zr);
					}
				}
				/* FOLD_TWID: after stage 0 the tile's coefficient 0 (register 0
				 * of group-thread 0) is the one value no difference has scaled;
				 * the caller multiplies it by n^-1 when it stores (a branch
				 * here would make ptxas shuffle the whole register file to
				 * merge the two paths: 34 moves per thread) */
			}
		}
	}
}

/* canonical residue of a value at the end of a transform: forward values are
 * below 2*bq, inverse values below bq (bq = 2q exact, 3q approximate); the
 * FOLD_LAST stage leaves [0,2q) in both families */
/* FWD_LAZY_STORE=1 (timing experiment only, tools/build_variant.sh): forward
 * transforms store [0,bq) instead of canonical values -- what a "lazy vector"
 * design would save.  The inverse kernels accept that range, so round trips
 * stay exact, but forward outputs are no longer the reference's. */
#ifndef FWD_LAZY_STORE
#define FWD_LAZY_STORE 0
#endif

template <bool INV, bool APX, int FOLD = FOLD_LAST>
__device__ __forceinline__ u64 tile_canon(u64 v, u64 q, u64 bq) {
	if (!INV) {
		v = csub(v, bq);        /* [0,2bq) -> [0,bq) */
		if (FWD_LAZY_STORE) {
			return v;
		}
	}
	if (APX && (!INV || FOLD == FOLD_TWID)) {
		v = csub(v, q);         /* [0,3q) -> [0,2q) */
	}
	return csub(v, q);
}

#endif
