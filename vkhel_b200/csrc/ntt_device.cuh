/*
 * Device-side helpers shared by the transform kernels (kernels_ntt.cu,
 * kernels_ntt_cluster.cu): exchange-buffer padding, programmatic dependent
 * launch, mbarrier + TMA bulk staging of twiddle subtrees, compile-time loops.
 */
#ifndef VKHEL_NTT_DEVICE_CUH
#define VKHEL_NTT_DEVICE_CUH

#include <type_traits>

#include "common.cuh"
#include "ntt_engine.cuh"

#ifndef FAST_PDL
#define FAST_PDL 1
#endif

/* padded position of tile element i in a warp-group's exchange buffer: 4 words
 * of padding per 32 keep the strided reads of the second round conflict-free.
 * Additive over disjoint bit fields, like the tile index itself. */
#ifndef ROWS_XPAD_HALF
#define ROWS_XPAD_HALF 0
#endif
__host__ __device__ constexpr int xpad(int i) {
#if ROWS_XPAD_HALF
	/* 2 words per 16 instead of 4 per 32 (same buffer size): the 128-bit
	 * accesses of the deepest round's layout -- 16 bytes at a 32-byte lane
	 * stride -- then take 4 wavefronts instead of 8, everything else stays
	 * minimal (tools/bank_model.py).  Measured: forward 0.3191 ms against
	 * 0.3174, inverse 0.3454 against 0.3463 per 512 transforms of n = 2^16 --
	 * nothing either way (the shared-memory pipe is at 18-20 %), so the
	 * shipped padding stays */
	return i + ((i >> 4) << 1);
#else
	return i + ((i >> 5) << 2);
#endif
}

/* 64-bit shared-memory store at a compile-time offset from a per-thread base,
 * as inline PTX so that the compiler's store vectoriser cannot merge two of
 * them back into one 128-bit store (see COLS_XST64) */
template <int OFF_BYTES>
__device__ __forceinline__ void sts64_at(unsigned base, u64 v) {
	asm volatile("st.shared.b64 [%0+%1], %2;"
			:: "r"(base), "n"(OFF_BYTES), "l"(v) : "memory");
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
	if constexpr (I < N) {
		f(std::integral_constant<int, I>());
		static_for<I + 1, N>(f);
	}
}

/* ---- programmatic dependent launch (PDL) ------------------------------------------
 * The fast kernels are launched with programmaticStreamSerialization: a CTA of
 * the next kernel may start once every CTA of the current one has executed
 * pdl_launch_dependents() (or exited), runs its prologue -- barrier init and
 * the TMA staging of its twiddles, which do not depend on the previous kernel
 * -- and blocks in pdl_wait() until the previous kernel has completed and its
 * writes are visible.  This overlaps one kernel's prologue with the tail of
 * the one before it. */
__device__ __forceinline__ void pdl_wait() {
#if FAST_PDL
	asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__device__ __forceinline__ void pdl_launch_dependents() {
#if FAST_PDL
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

/* ---- twiddle staging by TMA bulk copies --------------------------------------------
 * The twiddle subtree of a tile root is one contiguous run per level in the
 * bit-reversed table (2^u pairs of 16 bytes at level u), so staging it is K
 * one-dimensional bulk copies (cp.async.bulk, SASS UBLKCP) issued by a single
 * thread and tracked by an mbarrier; the other threads go straight to their
 * coefficient loads and only wait on the barrier before the first butterfly. */
__device__ __forceinline__ void mbar_init(u64 *bar, unsigned count) {
	const unsigned addr = (unsigned) __cvta_generic_to_shared(bar);
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(addr), "r"(count));
	/* make the initialised barrier visible to the async (TMA) proxy */
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(u64 *bar, unsigned bytes) {
	const unsigned addr = (unsigned) __cvta_generic_to_shared(bar);
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
			:: "r"(addr), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(u64 *bar, unsigned parity) {
	const unsigned addr = (unsigned) __cvta_generic_to_shared(bar);
	asm volatile(
		"{\n\t"
		".reg .pred done;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n\t"
		"@!done bra WAIT_%=;\n\t"
		"}" :: "r"(addr), "r"(parity) : "memory");
}

__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem,
		unsigned bytes, u64 *bar) {
	const unsigned dst = (unsigned) __cvta_generic_to_shared(smem);
	const unsigned mb = (unsigned) __cvta_generic_to_shared(bar);
	asm volatile(
		"cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
		"[%0], [%1], %2, [%3];"
		:: "r"(dst), "l"(gmem), "r"(bytes), "r"(mb) : "memory");
}

/* issue (one thread) the copies of the subtrees rooted at nodes
 * 2^s0 + H0 .. + hgroup - 1 into sm_tw[h << K | node] */
template <int K>
__device__ __forceinline__ void stage_twiddles_tma(ulonglong2 *sm_tw,
		const ulonglong2 *tw_g, unsigned s0, u64 H0, unsigned hgroup,
		u64 *bar) {
	/* every level u moves 2^u pairs: (2^K - 1) pairs per root */
	mbar_expect_tx(bar, hgroup * ((1u << K) - 1) * (unsigned) sizeof(ulonglong2));
	for (unsigned h = 0; h < hgroup; h++) {
		const u64 root = ((u64) 1 << s0) + H0 + h;
#pragma unroll
		for (unsigned u = 0; u < (unsigned) K; u++) {
			bulk_g2s(sm_tw + ((size_t) h << K) + (1u << u), tw_g + (root << u),
					(unsigned) sizeof(ulonglong2) << u, bar);
		}
	}
}

#endif
