"""Workload parameters: NTT-friendly primes and 2n-th roots.

Host-side integer helpers only (Python ints); used by tests and bench.py to
build inputs.  The recipe is the survey's: P[k] is the k-th largest prime below
2^60 with P[k] = 1 (mod 2^18); psi_n = x^((q-1)/2n) for the smallest x >= 2 for
which psi_n is a primitive 2n-th root (psi^n = -1).
"""

# 61-bit prime of the reference's test/numbers.c:48 (2^61 - 2^21 + 1)
Q61 = 2305843009211596801
# reference KAT moduli (test/vector.c)
Q_KAT_52 = 2251799813685313
W_KAT_52_N16 = 110968848420801


def is_prime(n):
    """Deterministic Miller-Rabin for n < 2^64."""
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def ntt_primes(count, bits=60, two_adicity=18):
    """`count` largest primes below 2^bits of the form k * 2^two_adicity + 1"""
    out = []
    step = 1 << two_adicity
    cand = (1 << bits) - step + 1
    while len(out) < count and cand > step:
        if is_prime(cand):
            out.append(cand)
        cand -= step
    return out


def find_psi(n, q):
    """primitive 2n-th root of unity mod q (n a power of two)"""
    assert (q - 1) % (2 * n) == 0, "q - 1 must be divisible by 2n"
    x = 2
    while True:
        psi = pow(x, (q - 1) // (2 * n), q)
        if pow(psi, n, q) == q - 1:
            return psi
        x += 1


# largest NTT prime (q = 1 mod 2^18) below 2^62: top of the lazy-butterfly range
Q62_LAZY_MAX = 4611686018425815041
# largest below 2^63: exercises the strict (always canonical) butterflies
Q63_STRICT = 9223372036836950017

# first of the RNS primes, used by most single-prime configurations
P0 = 1152921504606584833


def xorshift64_stream(seed, count, q):
    """SURVEY 8(d) input generator: xorshift64, each coefficient s mod q.
    Vectorised with numpy by running `lanes` independent generators."""
    import numpy as np
    lanes = min(count, 4096) or 1
    state = (np.arange(lanes, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(seed)) | np.uint64(1)
    rows = (count + lanes - 1) // lanes
    out = np.empty((rows, lanes), np.uint64)
    for r in range(rows):
        state ^= state << np.uint64(13)
        state ^= state >> np.uint64(7)
        state ^= state << np.uint64(17)
        out[r] = state
    return (out.reshape(-1)[:count] % np.uint64(q)).astype(np.uint64)
