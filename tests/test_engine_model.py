"""CPU suite: a Python model of the radix-8 engine's index geometry
(vkhel_b200/csrc/ntt_engine.cuh: tile_geom, tile_round) running the forward
transform against the reference's stage loops (src/vector.c:536-566), and the
inverse transform with FOLD_TWID -- n^-1 taken from a table of scaled inverse twiddles
by the butterflies whose inputs have only been on the sum side so far -- against
the reference's stage loops and final scaling (src/vector.c:599-639).  It pins
the static "which register / which thread is still unscaled" rules the CUDA
code relies on, for every tile size the column pass is instantiated with."""
import random

import pytest

Q61 = 2305843009211596801      # 2^61 - 2^21 + 1 (reference test/numbers.c:48)


def brv(i, bits):
    r = 0
    for k in range(bits):
        if i >> k & 1:
            r |= 1 << (bits - 1 - k)
    return r


def inverse_tables(n, q, psi):
    bits = n.bit_length() - 1
    roots = [1] * n
    p = 1
    for i in range(1, n):
        p = p * psi % q
        roots[brv(i, bits)] = p
    return [pow(r, -1, q) for r in roots]


def reference_inverse(x, inv, q):
    """src/vector.c:599-639: GS stages m = n/2 .. 1, then every element * n^-1"""
    n = len(x)
    x = list(x)
    t, m = 1, n // 2
    while m >= 1:
        for i in range(m):
            w = inv[m + i]
            for p in range(t):
                j = 2 * t * i + p
                a, b = x[j], x[j + t]
                x[j], x[j + t] = (a + b) % q, (a - b) * w % q
        t, m = t * 2, m // 2
    ninv = pow(n, -1, q)
    return [v * ninv % q for v in x]


class TileGeom:
    """tile_geom<K> of ntt_engine.cuh"""

    def __init__(self, k):
        self.k = k
        self.rounds = (k + 2) // 3
        self.last_cnt = k - 3 * (self.rounds - 1)

    def cnt(self, r):
        return self.last_cnt if r == self.rounds - 1 else 3

    def full(self, r):
        return self.cnt(r) == 3

    def tbase(self, r, t):
        if self.full(r):
            p = self.k - 3 * (r + 1)
            return ((t >> p) << (p + 3)) | (t & ((1 << p) - 1))
        return t << self.last_cnt

    def eoff(self, r, e):
        if self.full(r):
            return e << (self.k - 3 * (r + 1))
        c = self.last_cnt
        return ((e >> c) << (self.k - 3 + c)) | (e & ((1 << c) - 1))

    def gbase(self, r, j, t):
        if self.full(r):
            return (t >> (self.k - 3 * (r + 1))) << j
        return t << j

    def goff(self, r, j, e):
        if self.full(r):
            return e >> (3 - j)
        c = self.last_cnt
        return ((e >> c) << (self.k - 3 + j)) | ((e & ((1 << c) - 1)) >> (c - j))


def engine_inverse_fold_twid(x, inv, q):
    """tile_round<K, INV = true, FOLD_TWID> over a tile rooted at node 1"""
    n = len(x)
    k = n.bit_length() - 1
    g = TileGeom(k)
    ninv = pow(n, -1, q)
    scaled = [w * ninv % q for w in inv]
    mem = list(x)
    explicit = 0
    for rr in range(g.rounds):
        r = g.rounds - 1 - rr
        cnt = g.cnt(r)
        out = list(mem)
        hist_bits = k - 3 * (r + 1) if g.full(r) else 0
        for t in range(1 << (k - 3)):
            reg = [mem[g.tbase(r, t) + g.eoff(r, e)] for e in range(8)]
            unscaled_thread = (t & ((1 << hist_bits) - 1)) == 0
            for step in range(cnt):
                j = cnt - 1 - step
                u = 3 * r + j
                beta = cnt - 1 - j
                for e in range(8):
                    if e & (1 << beta):
                        continue
                    upos = (e & ((1 << beta) - 1)) == 0
                    node = (1 << u) + g.gbase(r, j, t) + g.goff(r, j, e)
                    w = scaled[node] if upos and unscaled_thread else inv[node]
                    a, b = reg[e], reg[e | (1 << beta)]
                    reg[e], reg[e | (1 << beta)] = (a + b) % q, (a - b) * w % q
                    if u == 0 and upos and unscaled_thread:
                        assert (t, e) == (0, 0)
                        reg[e] = reg[e] * ninv % q
                        explicit += 1
            for e in range(8):
                out[g.tbase(r, t) + g.eoff(r, e)] = reg[e]
        mem = out
    return mem, explicit


@pytest.mark.parametrize("k", range(3, 11))
def test_scaled_twiddles_reproduce_the_reference_inverse(k):
    n, q = 1 << k, Q61
    psi = pow(13, (q - 1) // (2 * n), q)
    assert pow(psi, n, q) == q - 1
    inv = inverse_tables(n, q, psi)
    rnd = random.Random(k)
    x = [rnd.randrange(q) for _ in range(n)]
    got, explicit = engine_inverse_fold_twid(x, inv, q)
    assert got == reference_inverse(x, inv, q)
    assert explicit == 1      # one coefficient per tile is scaled explicitly


def forward_tables(n, q, psi):
    bits = n.bit_length() - 1
    roots = [1] * n
    p = 1
    for i in range(1, n):
        p = p * psi % q
        roots[brv(i, bits)] = p
    return roots


def reference_forward(x, roots, q):
    """src/vector.c:536-566: CT stages m = 1 .. n/2, t = n/2 .. 1"""
    n = len(x)
    x = list(x)
    t, m = n // 2, 1
    while m < n:
        for i in range(m):
            w = roots[m + i]
            for p in range(t):
                j = 2 * t * i + p
                a, b = x[j], x[j + t] * w % q
                x[j], x[j + t] = (a + b) % q, (a - b) % q
        t, m = t // 2, m * 2
    return x


def engine_forward(x, roots, q):
    """tile_round<K, INV = false> over a tile rooted at node 1: rounds in
    order, local stage u = 3r + j pairs register bit cnt-1-j and takes twiddle
    node 2^u + gbase + goff"""
    n = len(x)
    k = n.bit_length() - 1
    g = TileGeom(k)
    mem = list(x)
    for r in range(g.rounds):
        cnt = g.cnt(r)
        out = list(mem)
        for t in range(1 << (k - 3)):
            reg = [mem[g.tbase(r, t) + g.eoff(r, e)] for e in range(8)]
            for j in range(cnt):
                u = 3 * r + j
                beta = cnt - 1 - j
                for e in range(8):
                    if e & (1 << beta):
                        continue
                    node = (1 << u) + g.gbase(r, j, t) + g.goff(r, j, e)
                    a, b = reg[e], reg[e | (1 << beta)] * roots[node] % q
                    reg[e], reg[e | (1 << beta)] = (a + b) % q, (a - b) % q
            for e in range(8):
                out[g.tbase(r, t) + g.eoff(r, e)] = reg[e]
        mem = out
    return mem


@pytest.mark.parametrize("k", range(3, 11))
def test_engine_geometry_reproduces_the_reference_forward(k):
    n, q = 1 << k, Q61
    psi = pow(13, (q - 1) // (2 * n), q)
    roots = forward_tables(n, q, psi)
    rnd = random.Random(100 + k)
    x = [rnd.randrange(q) for _ in range(n)]
    assert engine_forward(x, roots, q) == reference_forward(x, roots, q)
    # every tile index is owned by exactly one (thread, register) per round
    g = TileGeom(k)
    for r in range(g.rounds):
        owned = sorted(g.tbase(r, t) + g.eoff(r, e)
                       for t in range(1 << (k - 3)) for e in range(8))
        assert owned == list(range(n))
