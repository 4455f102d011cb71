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
 * Arithmetic: Harvey lazy butterflies (modarith.cuh).  Forward values stay in
 * [0,4q), inverse values in [0,2q); q < 2^62.
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
	/* tile index of register e of group-thread t in round r */
	__device__ __forceinline__ static int index(int r, int t, int e) {
		if (full(r)) {
			const int p = K - 3 * (r + 1);
			return ((t >> p) << (p + 3)) | (e << p) | (t & ((1 << p) - 1));
		}
		const int c = last_cnt;
		return ((e >> c) << (K - 3 + c)) | (t << c) | (e & ((1 << c) - 1));
	}
};

/* One round of butterflies on x[0..7].
 *   TW(node) returns the (w, w') pair of local twiddle node `node`.
 *   FOLD: inverse only -- local stage 0 is global stage 0: multiply by n^-1
 *         (fold_a = n^-1, fold_b = inv_root[1] * n^-1) instead of node 1. */
template <int K, bool INV, bool FOLD, class TW>
__device__ __forceinline__ void tile_round(u64 (&x)[8], int r, int t,
		const TW &tw, u64 q, u64 twoq, ulonglong2 fold_a, ulonglong2 fold_b) {
	using G = tile_geom<K>;
	const int cnt = G::cnt(r);
#pragma unroll
	for (int step = 0; step < 3; step++) {
		if (step >= cnt) {
			break;
		}
		const int j = INV ? cnt - 1 - step : step;  /* stage within the round */
		const int u = 3 * r + j;                    /* local stage */
		const int beta = cnt - 1 - j;               /* pair bit inside e */
		/* group number g = i >> (K-u) splits into a thread part and the
		 * bits of e above the pair bit */
		int g_thread;
		if (G::full(r)) {
			const int p = K - 3 * (r + 1);
			g_thread = (t >> p) << j;
		} else {
			g_thread = t << j;
		}
#pragma unroll
		for (int e = 0; e < 8; e++) {
			if (e & (1 << beta)) {
				continue;
			}
			int g;
			if (G::full(r)) {
				g = g_thread | (e >> (beta + 1));
			} else {
				/* passengers (top bits) then thread bits then e's upper
				 * pair bits */
				const int c = cnt;
				const int e_hi = e >> c;
				const int e_lo = e & ((1 << c) - 1);
				g = (e_hi << (K - 3 + c - (c - j))) | g_thread
					| (e_lo >> (beta + 1));
			}
			u64 &X = x[e];
			u64 &Y = x[e | (1 << beta)];
			if (INV && FOLD && u == 0) {
				const u64 s = X + Y;
				const u64 d = X - Y + twoq;
				X = shoup_lazy(s, fold_a.x, fold_a.y, q);
				Y = shoup_lazy(d, fold_b.x, fold_b.y, q);
			} else {
				const ulonglong2 w = tw((1 << u) + g);
				if (INV) {
					gs_lazy(X, Y, w.x, w.y, q, twoq);
				} else {
					ct_lazy(X, Y, w.x, w.y, q, twoq);
				}
			}
		}
	}
}

#endif
