/*
 * Host-side modular arithmetic used for table generation and checking.
 * Same names, signatures and results as the reference's
 * include/priv/numbers.h:7-22 (pinned by its test/numbers.c:12-94).
 */
#ifndef PRIV_NUMBERS_H
#define PRIV_NUMBERS_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Barrett parameters of the reference shaders (alpha - beta = 64);
 * kept because callers compute `1 << (bits + nt_alpha - 64)` with them
 * (reference test/numbers.c:24-26). */
static const int64_t nt_alpha	= 62;
static const int64_t nt_beta	= -2;

/* NB: despite its name this is the bit length of v (floor(log2 v) + 1),
 * reference include/priv/numbers.h:11-13, test/numbers.c:12-18. */
static inline uint64_t nt_ceil_log2(uint64_t v) {
	return 64 - (uint64_t) __builtin_clzll(v);
}

/* floor(factor * 2^64 / mod): the Shoup companion of `factor`
 * (reference src/numbers.c:30-34; third argument unused there as well). */
uint64_t nt_compute_barrett_factor(uint64_t factor, uint64_t mod, uint64_t n);
/* a * b mod `mod`; barrett_factor is ignored, as in src/numbers.c:36-40 */
uint64_t nt_multiply_mod(const uint64_t a, const uint64_t b,
		const uint64_t mod, const uint64_t barrett_factor);
uint64_t nt_power_mod(uint64_t base, uint64_t exp, const uint64_t mod);
/* root^(degree/2) == mod - 1 (src/numbers.c:61-69) */
bool nt_is_primitive_root(const uint64_t root, const uint64_t degree,
		const uint64_t mod);
uint64_t nt_inverse_mod(const uint64_t a, const uint64_t mod);

#ifdef __cplusplus
}
#endif

#endif
