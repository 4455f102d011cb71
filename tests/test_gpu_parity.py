"""GPU suite: parity of the CUDA path with the oracle, through the C-ABI.

Every call below goes through libvkhel.so's exported entry points (ctypes),
i.e. the boundary a C program links against.  Bit-exact comparison (integer
work): known-answer vectors of the reference's tests, seeded random inputs
against the oracle at sizes it finishes in seconds, and size-independent
properties at the BASELINE configuration sizes.
"""
import numpy as np
import pytest

import oracle
import vkhel_b200 as vk
from vkhel_b200 import params
from conftest import u64, rand_mod, rand_u64

pytestmark = pytest.mark.gpu


# ---- helpers ---------------------------------------------------------------------------
class TablePair:
    """library tables + oracle tables for the same (n, q, psi)"""

    def __init__(self, n, q, w=None):
        self.n, self.q = n, q
        self.w = params.find_psi(n, q) if w is None else w
        self.lib = vk.NttTables(n, q, self.w)
        self.ora = oracle.Tables(n, q, self.w)

    def destroy(self):
        self.lib.destroy()


def run_forward(ctx, x, tp, in_place=False):
    a = ctx.from_host(x)
    b = a if in_place else ctx.vector(len(x))
    ctx.forward_transform(a, b, tp.lib)
    out = b.to_host()
    if not in_place:
        assert np.array_equal(a.to_host(), x), "operand modified"
        b.destroy()
    a.destroy()
    return out


def run_inverse(ctx, x, tp, in_place=False):
    a = ctx.from_host(x)
    b = a if in_place else ctx.vector(len(x))
    ctx.inverse_transform(a, b, tp.lib)
    out = b.to_host()
    if not in_place:
        b.destroy()
    a.destroy()
    return out


# ---- vector lifecycle (reference test/vector.c:23-37,331-346) -------------------------
def test_copy_from_host_and_map_roundtrip(ctx):
    for length in (1, 2, 10, 1000, 65537):
        data = np.arange(length, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
        v = ctx.vector(length)
        assert not v.to_host().any(), "create must zero-fill"
        v.copy_from_host(data)
        assert np.array_equal(v.to_host(), data)
        v.destroy()


def test_dup(ctx):
    data = np.full(64, 5, np.uint64)
    a = ctx.from_host(data)
    b = a.dup()
    a.destroy()
    assert np.array_equal(b.to_host(), data)
    b.destroy()


def test_map_write_back(ctx):
    """unmap writes the host view back (reference vector.c:285-296)"""
    import ctypes
    lib = vk.lib()
    v = ctx.from_host(np.arange(16, dtype=np.uint64))
    mem = ctypes.c_void_p()
    lib.vkhel_vector_map(v.handle, ctypes.byref(mem), 16 * 8)
    view = (ctypes.c_uint64 * 16).from_address(mem.value)
    for i in range(16):
        view[i] = 100 + i
    lib.vkhel_vector_unmap(v.handle)
    assert v.to_host().tolist() == list(range(100, 116))
    v.destroy()


def test_pinned_upload_download(ctx):
    n = 1 << 16
    host = vk.host_alloc(n)
    back = vk.host_alloc(n)
    host.array[:] = np.arange(n, dtype=np.uint64)
    v = ctx.vector(n, zero=False)
    v.upload(host)
    v.download(back)
    ctx.sync()
    assert np.array_equal(back.array, host.array)
    v.destroy()
    host.free()
    back.free()


# ---- known-answer vectors of the reference ---------------------------------------------
@pytest.mark.parametrize("idx", [0, 1])
def test_forward_kat(ctx, kats, idx):
    k = kats["forward_transform"][idx]
    tp = TablePair(k["n"], k["q"], k["w"])
    for in_place in (False, True):
        got = run_forward(ctx, u64(k["operand"]), tp, in_place)
        assert got.tolist() == k["expected"]
    tp.destroy()


@pytest.mark.parametrize("idx", [0, 1])
def test_inverse_kat(ctx, kats, idx):
    k = kats["inverse_transform"][idx]
    tp = TablePair(k["n"], k["q"], k["w"])
    for in_place in (False, True):
        got = run_inverse(ctx, u64(k["operand"]), tp, in_place)
        assert got.tolist() == k["expected"]
    tp.destroy()


def test_elementwise_kats(ctx, kats):
    k = kats["elemfma"]
    a, b = ctx.from_host(k["a"]), ctx.from_host(k["b"])
    for case in k["cases"]:
        c = ctx.vector(len(k["a"]))
        ctx.elemfma(a, b, c, case["multiplier"], k["mod"])
        assert c.to_host().tolist() == case["expected"]
        c.destroy()
    a.destroy(), b.destroy()

    for k in kats["elemmul"]:
        a, b = ctx.from_host(k["a"]), ctx.from_host(k["b"])
        c = ctx.vector(len(k["a"]))
        ctx.elemmul(a, b, c, k["mod"])
        assert c.to_host().tolist() == k["expected"]
        a.destroy(), b.destroy(), c.destroy()

    k = kats["elemgtadd"]
    a = ctx.from_host(k["a"])
    for case in k["cases"]:
        c = ctx.vector(len(k["a"]))
        ctx.elemgtadd(a, c, case["bound"], case["diff"])
        assert c.to_host().tolist() == case["expected"]
        c.destroy()
    a.destroy()

    k = kats["elemgtsub"]
    a = ctx.from_host(k["a"])
    for case in k["cases"]:
        c = ctx.vector(len(k["a"]))
        ctx.elemgtsub(a, c, case["bound"], case["diff"], case["mod"])
        assert c.to_host().tolist() == case["expected"]
        c.destroy()
    a.destroy()

    for case in kats["elemmod"]:
        a = ctx.from_host(case["a"])
        c = ctx.vector(len(case["a"]))
        ctx.elemmod(a, c, case["mod"], case["q"])
        assert c.to_host().tolist() == case["expected"]
        a.destroy(), c.destroy()

    ex = kats["example"]
    a, b = ctx.from_host(ex["a"]), ctx.from_host(ex["b"])
    c = ctx.vector(4)
    ctx.elemmul(a, b, c, ex["mod"])
    assert c.to_host().tolist() == [4, 0, 0, 4]
    a.destroy(), b.destroy(), c.destroy()


# ---- randomised element-wise parity ------------------------------------------------------
ELEM_MODULI = [2, 3, 5, 10, 17, 769, 1125891450734593, params.Q_KAT_52,
               params.P0, params.Q61, params.Q62_LAZY_MAX, params.Q63_STRICT,
               (1 << 63) + 29, (1 << 64) - 59]


@pytest.mark.parametrize("q", ELEM_MODULI)
def test_elemmul_random(ctx, q):
    rng = np.random.default_rng(q % 9973)
    for length in (1, 7, 4096 + 3, 100001):
        # arbitrary 64-bit operands: the shader reduces both first
        a, b = rand_u64(rng, length), rand_u64(rng, length)
        if length == 7:
            a[:4] = u64([0, q - 1, q, (1 << 64) - 1])
            b[:4] = u64([q - 1, q - 1, q + 1 if q < (1 << 64) - 1 else 0,
                         (1 << 64) - 1])
        va, vb, vc = ctx.from_host(a), ctx.from_host(b), ctx.vector(length)
        ctx.elemmul(va, vb, vc, q)
        assert np.array_equal(vc.to_host(), oracle.elemmul(a, b, q))
        # in place on an operand
        ctx.elemmul(va, vb, va, q)
        assert np.array_equal(va.to_host(), oracle.elemmul(a, b, q))
        va.destroy(), vb.destroy(), vc.destroy()


@pytest.mark.parametrize("q", ELEM_MODULI)
def test_elemfma_random(ctx, q):
    rng = np.random.default_rng(q % 9967)
    length = 50001
    a, b = rand_mod(rng, length, q), rand_mod(rng, length, q)
    va, vb, vc = ctx.from_host(a), ctx.from_host(b), ctx.vector(length)
    for mult in (0, 1, 2, q // 2, q - 1):
        ctx.elemfma(va, vb, vc, mult, q)
        # wrap-around cases included: the contract, not the shader defect
        assert np.array_equal(vc.to_host(), oracle.elemfma(a, b, mult, q))
    va.destroy(), vb.destroy(), vc.destroy()


@pytest.mark.parametrize("q", ELEM_MODULI)
def test_elemgtsub_elemmod_random(ctx, q):
    rng = np.random.default_rng(q % 9949)
    length = 30011
    a = rand_u64(rng, length)
    va, vc = ctx.from_host(a), ctx.vector(length)
    for bound, diff in ((0, 1), (1 << 63, q - 1), ((1 << 64) - 1, 12345),
                        (q // 2, q)):
        ctx.elemgtsub(va, vc, bound, diff, q)
        assert np.array_equal(vc.to_host(),
                              oracle.elemgtsub(a, bound, diff, q))
    # elemmod: values in [0, Q) reduced to a small plaintext modulus t
    Q = params.P0
    x = rand_mod(rng, length, Q)
    vx = ctx.from_host(x)
    for t in (2, 3, 5, 65537, q):
        ctx.elemmod(vx, vc, t, Q)
        assert np.array_equal(vc.to_host(), oracle.elemmod(x, t, Q))
    va.destroy(), vc.destroy(), vx.destroy()


def test_elemgtadd_random(ctx):
    rng = np.random.default_rng(1)
    a = rand_u64(rng, 77777)
    va, vc = ctx.from_host(a), ctx.vector(len(a))
    for bound, diff in ((0, 0), (1 << 63, 1 << 63), ((1 << 64) - 2, 5)):
        ctx.elemgtadd(va, vc, bound, diff)
        assert np.array_equal(vc.to_host(), oracle.elemgtadd(a, bound, diff))
    va.destroy(), vc.destroy()


# ---- randomised NTT parity vs the oracle ---------------------------------------------------
NTT_MODULI = [params.P0, params.Q61, params.Q62_LAZY_MAX, params.Q63_STRICT]


@pytest.mark.parametrize("log2n", list(range(1, 16)))
def test_ntt_random_all_sizes(ctx, log2n):
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    moduli = list(NTT_MODULI)
    if log2n <= 10:
        moduli.append(1125891450734593)       # 50-bit, 2-adicity 11
    if log2n <= 3:
        moduli.append(113)                    # 7-bit KAT modulus
    for q in moduli:
        tp = TablePair(n, q)
        x = rand_mod(rng, n, q)
        x[:2] = u64([0, q - 1])
        want = oracle.forward(x, tp.ora)
        for in_place in (False, True):
            assert np.array_equal(run_forward(ctx, x, tp, in_place), want), \
                ("forward", n, q, in_place)
        y = rand_mod(rng, n, q)
        want_inv = oracle.inverse(y, tp.ora)
        for in_place in (False, True):
            assert np.array_equal(run_inverse(ctx, y, tp, in_place),
                                  want_inv), ("inverse", n, q, in_place)
        tp.destroy()


@pytest.mark.parametrize("log2n", [16, 17])
def test_ntt_random_large(ctx, log2n):
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    for q in (params.P0, params.Q63_STRICT):
        tp = TablePair(n, q)
        x = rand_mod(rng, n, q)
        fwd = run_forward(ctx, x, tp)
        assert np.array_equal(fwd, oracle.forward(x, tp.ora))
        assert np.array_equal(run_inverse(ctx, fwd, tp, in_place=True), x)
        assert np.array_equal(run_inverse(ctx, x, tp),
                              oracle.inverse(x, tp.ora))
        tp.destroy()


@pytest.mark.parametrize("log2n", [18, 19, 20])
def test_ntt_random_huge(ctx, log2n):
    """n = 2^18: 1024-point column tiles; n > 2^18: leading strided pass.
    Q61 = 2^61 - 2^21 + 1 is the only KAT modulus with enough 2-adicity."""
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    tp = TablePair(n, params.Q61)
    x = rand_mod(rng, n, tp.q)
    fwd = run_forward(ctx, x, tp)
    assert np.array_equal(fwd, oracle.forward(x, tp.ora))
    assert np.array_equal(run_inverse(ctx, fwd, tp, in_place=True), x)
    assert np.array_equal(run_inverse(ctx, x, tp), oracle.inverse(x, tp.ora))
    tp.destroy()


def test_inverse_scales_tail_like_reference(ctx):
    """result longer than n: the reference multiplies the tail by n^-1 too
    (SURVEY App. B, Q4); the forward transform leaves the tail alone"""
    n, q = 64, params.P0
    tp = TablePair(n, q)
    rng = np.random.default_rng(2)
    x = rand_mod(rng, n, q)
    tail = rand_u64(rng, 37)
    full = np.concatenate([x, tail])
    a, b = ctx.from_host(full), ctx.from_host(full)
    ctx.inverse_transform(a, b, tp.lib)
    want = oracle.inverse(x, tp.ora, out_len=n + 37, out_init=full)
    assert np.array_equal(b.to_host(), want)
    ctx.forward_transform(a, b, tp.lib)
    got = b.to_host()
    assert np.array_equal(got[:n], oracle.forward(x, tp.ora))
    assert np.array_equal(got[n:], want[n:])
    a.destroy(), b.destroy(), tp.destroy()


# ---- batched / RNS extensions ------------------------------------------------------------------
@pytest.mark.parametrize("log2n,batch", [(2, 1000), (6, 33), (10, 17),
                                          (12, 9), (14, 5), (16, 3)])
def test_batch_matches_oracle(ctx, log2n, batch):
    n = 1 << log2n
    tp = TablePair(n, params.P0)
    rng = np.random.default_rng(batch)
    x = rand_mod(rng, n * batch, tp.q)
    a, b = ctx.from_host(x), ctx.vector(n * batch)
    ctx.forward_transform_batch(a, b, tp.lib, batch)
    want = oracle.forward_batch(x, [tp.ora], threads=8)
    assert np.array_equal(b.to_host(), want)
    ctx.inverse_transform_batch(b, b, tp.lib, batch)
    assert np.array_equal(b.to_host(), x)
    a.destroy(), b.destroy(), tp.destroy()


@pytest.mark.parametrize("log2n,limbs,batch", [(4, 3, 5), (11, 4, 3),
                                                (13, 5, 2), (16, 3, 2)])
def test_rns_matches_oracle(ctx, log2n, limbs, batch):
    n = 1 << log2n
    primes = params.ntt_primes(limbs)
    tps = [TablePair(n, q) for q in primes]
    rng = np.random.default_rng(limbs)
    x = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                        for p in range(limbs * batch)])
    a, b = ctx.from_host(x), ctx.vector(x.size)
    ctx.forward_transform_rns(a, b, [t.lib for t in tps], batch)
    want = oracle.forward_batch(x, [t.ora for t in tps], threads=8)
    assert np.array_equal(b.to_host(), want)
    ctx.inverse_transform_rns(b, a, [t.lib for t in tps], batch)
    assert np.array_equal(a.to_host(), x)
    # element-wise product with one modulus per limb
    y = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                        for p in range(limbs * batch)])
    c = ctx.from_host(y)
    ctx.elemmul_rns(a, c, b, primes, n, batch)
    want_mul = np.concatenate([
        oracle.elemmul(x[p * n:(p + 1) * n], y[p * n:(p + 1) * n],
                       primes[p % limbs]) for p in range(limbs * batch)])
    assert np.array_equal(b.to_host(), want_mul)
    a.destroy(), b.destroy(), c.destroy()
    for t in tps:
        t.destroy()


@pytest.mark.parametrize("log2n,limbs,batch", [(3, 2, 9), (5, 2, 3), (8, 1, 4),
                                                (9, 2, 5), (12, 3, 2),
                                                (16, 2, 3), (19, 1, 2)])
def test_polymul_matches_schoolbook_and_oracle(ctx, log2n, limbs, batch):
    n = 1 << log2n
    # n > 2^17 needs more 2-adicity than the RNS primes have
    primes = params.ntt_primes(limbs) if log2n <= 17 else [params.Q61]
    tps = [TablePair(n, q) for q in primes]
    rng = np.random.default_rng(n)
    polys = limbs * batch
    a = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                        for p in range(polys)])
    b = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                        for p in range(polys)])
    va, vb, vc = ctx.from_host(a), ctx.from_host(b), ctx.vector(a.size)
    ctx.polymul_rns(va, vb, vc, [t.lib for t in tps], batch)
    got = vc.to_host()
    for p in range(polys):
        t = tps[p % limbs]
        sl = slice(p * n, (p + 1) * n)
        prod = oracle.elemmul(oracle.forward(a[sl], t.ora),
                              oracle.forward(b[sl], t.ora), t.q)
        assert np.array_equal(got[sl], oracle.inverse(prod, t.ora))
        if n <= 256:
            assert np.array_equal(
                got[sl], oracle.negacyclic_schoolbook(a[sl], b[sl], t.q))
    assert np.array_equal(va.to_host(), a) and np.array_equal(vb.to_host(), b)
    va.destroy(), vb.destroy(), vc.destroy()
    for t in tps:
        t.destroy()


@pytest.mark.parametrize("log2n", [6, 13])
@pytest.mark.parametrize("alias", ["a", "b", "square"])
def test_polymul_result_may_alias_operands(ctx, log2n, alias):
    """result == a, result == b, and a == b == result (in-place squaring)"""
    n, limbs, batch = 1 << log2n, 2, 3
    primes = params.ntt_primes(limbs)
    tps = [TablePair(n, q) for q in primes]
    rng = np.random.default_rng(log2n)
    polys = limbs * batch
    a = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                        for p in range(polys)])
    b = a if alias == "square" else np.concatenate(
        [rand_mod(rng, n, primes[p % limbs]) for p in range(polys)])
    va = ctx.from_host(a)
    vb = va if alias == "square" else ctx.from_host(b)
    vc = vb if alias == "b" else va
    ctx.polymul_rns(va, vb, vc, [t.lib for t in tps], batch)
    got = vc.to_host()
    for p in range(polys):
        t = tps[p % limbs]
        sl = slice(p * n, (p + 1) * n)
        prod = oracle.elemmul(oracle.forward(a[sl], t.ora),
                              oracle.forward(b[sl], t.ora), t.q)
        assert np.array_equal(got[sl], oracle.inverse(prod, t.ora)), p
    va.destroy()
    if vb is not va:
        vb.destroy()
    for t in tps:
        t.destroy()


# ---- BASELINE configurations at full size: size-independent properties --------------------------
def _roundtrip_and_linearity(ctx, tables, n, batch, q_of_poly):
    """inverse(forward(x)) == x and forward(x + y) == forward(x) + forward(y)
    on the full-size batch, plus a sampled oracle comparison."""
    polys = len(tables) * batch
    rng = np.random.default_rng(polys)
    total = n * polys
    qs = np.repeat(u64([q_of_poly(p) for p in range(polys)]), n)
    x = rand_u64(rng, total) % qs
    y = rand_u64(rng, total) % qs
    s = (x + y) % qs                      # q < 2^63: no 64-bit overflow
    libs = [t.lib for t in tables]
    vx, vy, vs = ctx.from_host(x), ctx.from_host(y), ctx.from_host(s)
    fx, fy = ctx.vector(total, zero=False), ctx.vector(total, zero=False)
    ctx.forward_transform_rns(vx, fx, libs, batch)
    ctx.forward_transform_rns(vy, fy, libs, batch)
    ctx.forward_transform_rns(vs, vs, libs, batch)
    hx, hy, hs = fx.to_host(), fy.to_host(), vs.to_host()
    assert np.array_equal((hx + hy) % qs, hs), "linearity"
    assert (hx < qs).all(), "canonical output"
    # sampled polynomials against the oracle
    for p in sorted(set([0, 1, polys // 2, polys - 1])):
        sl = slice(p * n, (p + 1) * n)
        assert np.array_equal(
            hx[sl], oracle.forward(x[sl], tables[p % len(tables)].ora))
    ctx.inverse_transform_rns(fx, fx, libs, batch)
    assert np.array_equal(fx.to_host(), x), "round trip"
    for v in (vx, vy, vs, fx, fy):
        v.destroy()


def test_config1_example_shape(ctx):
    """BASELINE configs[0]: n=4096, 61-bit prime, round trip + pointwise mul"""
    n, q = 4096, params.Q61
    tp = TablePair(n, q)
    assert tp.w == 700439432845261874
    rng = np.random.default_rng(4096)
    a, b = rand_mod(rng, n, q), rand_mod(rng, n, q)
    fa, fb = run_forward(ctx, a, tp), run_forward(ctx, b, tp)
    assert np.array_equal(fa, oracle.forward(a, tp.ora))
    assert np.array_equal(run_inverse(ctx, fa, tp), a)
    va, vb, vc = ctx.from_host(fa), ctx.from_host(fb), ctx.vector(n)
    ctx.elemmul(va, vb, vc, q)
    prod = vc.to_host()
    assert np.array_equal(prod, oracle.elemmul(fa, fb, q))
    ctx.inverse_transform(vc, vc, tp.lib)
    assert np.array_equal(vc.to_host(), oracle.inverse(prod, tp.ora))
    va.destroy(), vb.destroy(), vc.destroy(), tp.destroy()


def test_config2_n14_batch256(ctx):
    """BASELINE configs[1]: n=2^14, one prime, batch 256"""
    n = 1 << 14
    tp = TablePair(n, params.P0)
    _roundtrip_and_linearity(ctx, [tp], n, 256, lambda p: params.P0)
    tp.destroy()


def test_config3_n16_rns32_batch16(ctx):
    """BASELINE configs[2]: n=2^16, 32 RNS limbs x batch 16 (256 MiB)"""
    n = 1 << 16
    primes = params.ntt_primes(32)
    tps = [TablePair(n, q) for q in primes]
    _roundtrip_and_linearity(ctx, tps, n, 16, lambda p: primes[p % 32])
    for t in tps:
        t.destroy()


def test_config4_polymul_n16(ctx):
    """BASELINE configs[3] (per-GPU share, batch 128): c = INTT(NTT(a)*NTT(b));
    checked through a size-independent identity: multiplying by the
    monomial x^k rotates the coefficients negacyclically"""
    n, batch = 1 << 16, 128
    tp = TablePair(n, params.P0)
    q = tp.q
    rng = np.random.default_rng(16)
    a = rand_mod(rng, n * batch, q)
    b = np.zeros(n * batch, np.uint64)
    shifts = rng.integers(0, n, size=batch)
    for p, k in enumerate(shifts):
        b[p * n + k] = 1
    va, vb, vc = ctx.from_host(a), ctx.from_host(b), ctx.vector(n * batch)
    ctx.polymul_rns(va, vb, vc, [tp.lib], batch)
    got = vc.to_host().reshape(batch, n)
    a2 = a.reshape(batch, n)
    for p, k in enumerate(shifts):
        k = int(k)
        rolled = np.roll(a2[p], k)
        neg = (np.uint64(q) - rolled[:k]) % np.uint64(q)
        want = np.concatenate([neg, rolled[k:]])
        assert np.array_equal(got[p], want), p
    va.destroy(), vb.destroy(), vc.destroy(), tp.destroy()


@pytest.mark.parametrize("log2n", [10, 12, 13, 15, 17])
def test_config5_sweep_roundtrip(ctx, log2n):
    """BASELINE configs[4] shape (reduced to 2^24 coefficients per size)"""
    n = 1 << log2n
    batch = (1 << 24) >> log2n
    tp = TablePair(n, params.P0)
    _roundtrip_and_linearity(ctx, [tp], n, batch, lambda p: params.P0)
    tp.destroy()


# ---- BASELINE configs[3] and [4] at FULL size ------------------------------------------------------
def _psi_powers(n, q, psi):
    """psi^0 .. psi^(n-1) mod q"""
    out = np.empty(n, np.uint64)
    v = 1
    for i in range(n):
        out[i] = v
        v = v * psi % q
    return out


def _evaluate_at_psi(polys, n, q, powers):
    """a(psi) mod q for every polynomial of a [batch][n] array: the dot product
    with the powers of psi (oracle mulmod, then an exact sum in two halves)"""
    batch = polys.size // n
    out = []
    chunk = 64
    for b0 in range(0, batch, chunk):
        nb = min(chunk, batch - b0)
        prod = oracle.elemmul(polys[b0 * n:(b0 + nb) * n], np.tile(powers, nb), q)
        prod = prod.reshape(nb, n)
        hi = (prod >> np.uint64(32)).sum(axis=1, dtype=np.uint64)
        lo = (prod & np.uint64(0xffffffff)).sum(axis=1, dtype=np.uint64)
        out += [((int(h) << 32) + int(l)) % q for h, l in zip(hi, lo)]
    return out


def test_config4_polymul_full_batch1024(ctx):
    """BASELINE configs[3] at full size: n = 2^16, batch 1024, random a AND b
    (512 MiB per operand).  Eight sampled products bit for bit against the
    oracle's forward / elemmul / inverse, and EVERY product through a
    size-independent identity: x^n + 1 vanishes at psi, so
    c(psi) = a(psi) * b(psi) mod q."""
    n, batch = 1 << 16, 1024
    tp = TablePair(n, params.P0)
    q = tp.q
    rng = np.random.default_rng(1024)
    a = rand_mod(rng, n * batch, q)
    b = rand_mod(rng, n * batch, q)
    va, vb, vc = ctx.from_host(a), ctx.from_host(b), ctx.vector(n * batch, zero=False)
    ctx.polymul_rns(va, vb, vc, [tp.lib], batch)
    got = vc.to_host()
    assert np.array_equal(va.to_host(), a) and np.array_equal(vb.to_host(), b), \
        "operands modified"
    for v in (va, vb, vc):
        v.destroy()
    assert (got < np.uint64(q)).all(), "canonical output"
    for p in (0, 1, 127, 128, 500, 777, 1022, 1023):
        sl = slice(p * n, (p + 1) * n)
        prod = oracle.elemmul(oracle.forward(a[sl], tp.ora),
                              oracle.forward(b[sl], tp.ora), q)
        assert np.array_equal(got[sl], oracle.inverse(prod, tp.ora)), p
    powers = _psi_powers(n, q, tp.w)
    ea = _evaluate_at_psi(a, n, q, powers)
    eb = _evaluate_at_psi(b, n, q, powers)
    ec = _evaluate_at_psi(got, n, q, powers)
    bad = [p for p in range(batch) if ec[p] != ea[p] * eb[p] % q]
    assert not bad, bad[:8]
    tp.destroy()


@pytest.mark.parametrize("log2n", [10, 17])
def test_config5_sweep_full_size(ctx, log2n):
    """BASELINE configs[4] at full size: 2^27 coefficients (1 GiB) at the two
    ends of the sweep.  Every polynomial: canonical output, the evaluation
    identity out[0] = a(psi) (output index 0 is the evaluation at psi^1,
    SURVEY 8 a1) and the round trip; eight sampled polynomials bit for bit
    against the oracle."""
    n = 1 << log2n
    batch = (1 << 27) >> log2n
    tp = TablePair(n, params.P0)
    q = tp.q
    rng = np.random.default_rng(log2n)
    x = rand_mod(rng, n * batch, q)
    vx = ctx.from_host(x)
    vf = ctx.vector(n * batch, zero=False)
    ctx.forward_transform_batch(vx, vf, tp.lib, batch)
    f = vf.to_host()
    assert (f < np.uint64(q)).all(), "canonical output"
    for p in sorted({0, 1, 2, batch // 3, batch // 2, batch - 3, batch - 2,
                     batch - 1}):
        sl = slice(p * n, (p + 1) * n)
        assert np.array_equal(f[sl], oracle.forward(x[sl], tp.ora)), p
    ex = _evaluate_at_psi(x, n, q, _psi_powers(n, q, tp.w))
    first = f.reshape(batch, n)[:, 0]
    bad = [p for p in range(batch) if int(first[p]) != ex[p]]
    assert not bad, bad[:8]
    del f
    ctx.inverse_transform_batch(vf, vf, tp.lib, batch)
    assert np.array_equal(vf.to_host(), x), "round trip"
    vx.destroy(), vf.destroy(), tp.destroy()
