/*
 * Host-side modular arithmetic: table generation and checking only -- nothing
 * here runs on the GPU.
 *
 * The five nt_* functions and the two constants keep the names, signatures and
 * results of the reference's include/priv/numbers.h:7-22, because its
 * white-box test test/numbers.c calls them directly (lines 12-94).  The
 * implementations (vkhel_b200/csrc/numbers.c) are written from that contract
 * with 128-bit integers rather than from the reference's Barrett code.
 */
#ifndef PRIV_NUMBERS_H
#define PRIV_NUMBERS_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Parameters of the reference shaders' Barrett reduction (alpha - beta = 64).
 * Nothing in this library depends on them; callers of the reference API
 * compute `1 << (bits + nt_alpha - 64)` with them (test/numbers.c:24-26). */
static const int64_t nt_alpha = 62;
static const int64_t nt_beta = -2;

/* Bit length of v, i.e. floor(log2 v) + 1 -- NOT the ceiling of log2, despite
 * the name it inherits (test/numbers.c:12-18 pins 1 -> 1, 2 -> 2, 4 -> 3). */
static inline uint64_t nt_ceil_log2(uint64_t v) {
	return 64 - (uint64_t) __builtin_clzll(v);
}

/* a^-1 mod `mod` for a < mod, gcd(a, mod) = 1 */
uint64_t nt_inverse_mod(const uint64_t a, const uint64_t mod);

/* base^exp mod `mod` */
uint64_t nt_power_mod(uint64_t base, uint64_t exp, const uint64_t mod);

/* a * b mod `mod`.  The fourth argument is accepted and ignored, as in the
 * reference (src/numbers.c:36-40 recomputes its own constant). */
uint64_t nt_multiply_mod(const uint64_t a, const uint64_t b,
		const uint64_t mod, const uint64_t barrett_factor);

/* floor(factor * 2^64 / mod): the Shoup companion of `factor`.  The third
 * argument is unused (the reference only asserts on it, src/numbers.c:30-34). */
uint64_t nt_compute_barrett_factor(uint64_t factor, uint64_t mod, uint64_t n);

/* true iff root has multiplicative order exactly `degree` (a power of two),
 * tested as root^(degree/2) == mod - 1 (src/numbers.c:61-69) */
bool nt_is_primitive_root(const uint64_t root, const uint64_t degree,
		const uint64_t mod);

#ifdef __cplusplus
}
#endif

#endif
