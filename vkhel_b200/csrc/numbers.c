/*
 * Host-side modular arithmetic: table generation and checking.
 *
 * Same contract as the reference's src/numbers.c:5-99 (names, argument
 * meaning, results -- pinned by the reference's test/numbers.c), written from
 * the contract rather than from its Barrett code: the compiler's 128-bit
 * arithmetic gives the canonical residue directly.
 */
#include <assert.h>
#include "priv/numbers.h"

typedef unsigned __int128 u128;

uint64_t nt_compute_barrett_factor(uint64_t factor, uint64_t mod, uint64_t n) {
	(void) n; /* only asserted on in the reference (src/numbers.c:31) */
	assert(mod != 0);
	return (uint64_t) ((((u128) factor) << 64) / mod);
}

uint64_t nt_multiply_mod(const uint64_t a, const uint64_t b,
		const uint64_t mod, const uint64_t barrett_factor) {
	(void) barrett_factor; /* ignored by the reference too (numbers.c:36-40) */
	return (uint64_t) (((u128) a * b) % mod);
}

uint64_t nt_power_mod(uint64_t base, uint64_t exp, const uint64_t mod) {
	/* square-and-multiply, least significant exponent bit first */
	uint64_t acc = 1 % mod;
	uint64_t sq = base % mod;
	for (; exp != 0; exp >>= 1) {
		if (exp & 1) {
			acc = (uint64_t) (((u128) acc * sq) % mod);
		}
		sq = (uint64_t) (((u128) sq * sq) % mod);
	}
	return acc;
}

bool nt_is_primitive_root(const uint64_t root, const uint64_t degree,
		const uint64_t mod) {
	/* for a power-of-two degree: root has order exactly `degree`
	 * iff root^(degree/2) = -1 (reference src/numbers.c:61-69) */
	if (root == 0) {
		return false;
	}
	assert(degree != 0 && (degree & (degree - 1)) == 0);
	return nt_power_mod(root, degree / 2, mod) == mod - 1;
}

uint64_t nt_inverse_mod(const uint64_t a, const uint64_t mod) {
	/* extended Euclid on (mod, a), tracking only the coefficient of a.
	 * The reference (src/numbers.c:71-99) works in int64_t, hence its
	 * mod < 2^63 limit; the coefficients here are kept in 128 bits. */
	assert(a < mod);
	if (mod == 1) {
		return 0;
	}
	if (a <= 1) {
		return 1; /* the reference also returns 1 for a == 0 */
	}
	uint64_t r0 = mod, r1 = a;
	__int128 t0 = 0, t1 = 1;
	while (r1 > 1) {
		const uint64_t quot = r0 / r1;
		const uint64_t r2 = r0 - quot * r1;
		const __int128 t2 = t0 - (__int128) quot * t1;
		r0 = r1; r1 = r2;
		t0 = t1; t1 = t2;
		assert(r1 != 0 && "nt_inverse_mod: operand not invertible");
	}
	__int128 res = t1 % (__int128) mod;
	if (res < 0) {
		res += mod;
	}
	return (uint64_t) res;
}
