"""CPU suite, part 1: pin the oracle.

The oracle (oracle/oracle.c) is checked against every known-answer vector the
reference's own tests hold for the path (tests/golden/reference_kats.json,
extracted from test/vector.c, test/ntt.c, examples/example.c), against golden
tables produced by the reference's compiled host code
(tests/golden/reference_tables.json) and, when oracle/_ref is present, against
that code live.  Then its internal consistency: literal shader arithmetic ==
the canonical contract, transform identities, NTT product == schoolbook.
"""
import hashlib

import numpy as np
import pytest

import oracle
from vkhel_b200 import params
from conftest import u64, rand_mod, rand_u64


def brv(v, bits):
    return int(format(v, "0%db" % bits)[::-1], 2) if bits else 0


# ---- golden vectors of the reference tests ------------------------------------------
def test_tables_kat(kats):
    k = kats["tables"]
    t = oracle.Tables(k["n"], k["q"], k["w"])
    assert t.roots.tolist() == k["roots_of_unity"]
    # the reference only checks truthiness (test/ntt.c:14); assert the product
    for r, i in zip(t.roots.tolist(), t.inv_roots.tolist()):
        assert r * i % k["q"] == 1


@pytest.mark.parametrize("idx", [0, 1])
def test_forward_kat(kats, idx):
    k = kats["forward_transform"][idx]
    t = oracle.Tables(k["n"], k["q"], k["w"])
    assert oracle.forward(k["operand"], t).tolist() == k["expected"]


@pytest.mark.parametrize("idx", [0, 1])
def test_inverse_kat(kats, idx):
    k = kats["inverse_transform"][idx]
    t = oracle.Tables(k["n"], k["q"], k["w"])
    assert oracle.inverse(k["operand"], t).tolist() == k["expected"]


def test_elemmul_kat(kats):
    for k in kats["elemmul"]:
        for literal in (False, True):
            got = oracle.elemmul(k["a"], k["b"], k["mod"], literal=literal)
            assert got.tolist() == k["expected"]
    ex = kats["example"]
    assert oracle.elemmul(ex["a"], ex["b"], ex["mod"]).tolist() == [4, 0, 0, 4]


def test_elemfma_kat(kats):
    k = kats["elemfma"]
    for case in k["cases"]:
        got = oracle.elemfma(k["a"], k["b"], case["multiplier"], k["mod"])
        assert got.tolist() == case["expected"]


def test_elemfma_literal_defect(kats):
    """The shader is wrong on wrap-around: the literal
    restatement reproduces the KATs without wrap but not 8*100+16 mod 769."""
    k = kats["elemfma"]
    lit = oracle.elemfma(k["a"], k["b"], 2, k["mod"], literal=True)
    assert lit.tolist() == k["cases"][1]["expected"]
    q = params.P0
    rng = np.random.default_rng(7)
    a, b = rand_mod(rng, 4096, q), rand_mod(rng, 4096, q)
    good = oracle.elemfma(a, b, q // 3, q)
    bad = oracle.elemfma(a, b, q // 3, q, literal=True)
    assert 0.3 < float(np.mean(good != bad)) < 0.8


def test_elemgtadd_kat(kats):
    k = kats["elemgtadd"]
    for case in k["cases"]:
        got = oracle.elemgtadd(k["a"], case["bound"], case["diff"])
        assert got.tolist() == case["expected"]


def test_elemgtsub_kat(kats):
    k = kats["elemgtsub"]
    for case in k["cases"]:
        for literal in (False, True):
            got = oracle.elemgtsub(k["a"], case["bound"], case["diff"],
                                   case["mod"], literal=literal)
            assert got.tolist() == case["expected"]


def test_elemmod_kat(kats):
    for case in kats["elemmod"]:
        got = oracle.elemmod(case["a"], case["mod"], case["q"])
        assert got.tolist() == case["expected"]


# ---- against the reference's compiled host code --------------------------------------
def test_tables_match_reference_fixtures(table_fixtures):
    for e in table_fixtures["tables"]:
        if e["n"] > (1 << 14):
            continue  # the literal per-element Euclid is slow; covered below
        t = oracle.Tables(e["n"], e["q"], e["w"])
        for name in ("roots", "inv_roots", "roots_shoup", "inv_roots_shoup"):
            arr = getattr(t, name)
            assert hashlib.sha256(arr.tobytes()).hexdigest() == \
                e[name + "_sha256"], (e["n"], e["q"], name)
        assert oracle.inverse_mod(e["n"] % e["q"], e["q"]) == e["inv_n"]


def test_scalars_match_reference_fixtures(table_fixtures):
    for s in table_fixtures["scalars"]:
        q, a, b = s["q"], s["a"], s["b"]
        assert oracle.multiply_mod(a, b, q) == s["multiply_mod"]
        assert oracle.power_mod(a, b, q) == s["power_mod"]
        assert oracle.inverse_mod(a, q) == s["inverse_mod"]
        assert oracle.shoup_factor(a, q) == s["shoup_factor"]
        # and the mathematics
        assert s["multiply_mod"] == a * b % q
        assert s["power_mod"] == pow(a, b, q)
        assert s["inverse_mod"] * a % q == 1
        assert s["shoup_factor"] == (a << 64) // q


def test_tables_match_reference_live():
    ref = oracle.reference_host()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(11)
    primes = params.ntt_primes(4, bits=59, two_adicity=12)
    for n in (2, 8, 64, 1024):
        for q in primes + [params.Q61, 769 if n <= 128 else params.P0]:
            w = params.find_psi(n, q)
            got = oracle.Tables(n, q, w)
            want = oracle.reference_tables(ref, n, q, w)
            for name, arr in zip(("roots", "inv_roots", "roots_shoup",
                                  "inv_roots_shoup"), want):
                assert np.array_equal(getattr(got, name), arr), (n, q, name)
    for q in (769, params.Q61, params.P0):
        for _ in range(200):
            a, b = (int(x) for x in rand_mod(rng, 2, q))
            assert oracle.multiply_mod(a, b, q) == \
                ref.nt_multiply_mod(a, b, q, 0)


# ---- internal consistency --------------------------------------------------------------
@pytest.mark.parametrize("q", [769, 1125891450734593, params.Q_KAT_52,
                               params.P0, params.Q61, (1 << 62) - 57])
def test_elemmul_literal_equals_contract(q):
    assert oracle.barrett_defined(q)
    rng = np.random.default_rng(q % 1000)
    a, b = rand_u64(rng, 20000), rand_u64(rng, 20000)
    assert np.array_equal(oracle.elemmul(a, b, q, literal=True),
                          oracle.elemmul(a, b, q))
    want = [int(x) % q * (int(y) % q) % q for x, y in zip(a[:200], b[:200])]
    assert oracle.elemmul(a[:200], b[:200], q).tolist() == want


@pytest.mark.parametrize("q", [5, 10, 769, params.P0])
def test_elemgtsub_literal_equals_contract(q):
    rng = np.random.default_rng(3)
    a = rand_u64(rng, 5000)
    bound = 1 << 63
    for diff in (3, q + 3, (1 << 64) - 1):
        assert np.array_equal(
            oracle.elemgtsub(a, bound, diff, q, literal=True),
            oracle.elemgtsub(a, bound, diff, q))


@pytest.mark.parametrize("log2n", [1, 2, 3, 5, 8, 11])
@pytest.mark.parametrize("q", [params.P0, params.Q61, 1125891450734593])
def test_transform_identities(log2n, q):
    n = 1 << log2n
    if (q - 1) % (2 * n):
        pytest.skip("2n does not divide q - 1")
    psi = params.find_psi(n, q)
    t = oracle.Tables(n, q, psi)
    rng = np.random.default_rng(log2n)
    a = rand_mod(rng, n, q)
    fwd = oracle.forward(a, t)
    # round trip
    assert np.array_equal(oracle.inverse(fwd, t), a)
    # evaluation identity: out[j] = a(psi^(2*brv(j)+1))
    if n <= 256:
        coeffs = [int(x) for x in a]
        for j in range(n):
            point = pow(psi, 2 * brv(j, log2n) + 1, q)
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * point + c) % q
            assert acc == int(fwd[j])


def test_inverse_scales_whole_result_vector():
    """reference quirk Q4: n^-1 is applied to result->length elements"""
    n, q = 8, 769
    t = oracle.Tables(n, q, params.find_psi(n, q))
    x = np.arange(1, n + 1, dtype=np.uint64)
    tail = u64([5, 700, 768, 1000])
    out = oracle.inverse(x, t, out_len=n + 4,
                         out_init=np.concatenate([np.zeros(n, np.uint64),
                                                  tail]))
    inv_n = pow(n, -1, q)
    assert out[n:].tolist() == [int(v) * inv_n % q for v in tail]
    assert np.array_equal(out[:n], oracle.inverse(x, t))


@pytest.mark.parametrize("n,q", [(8, 769), (64, 1125891450734593),
                                 (128, params.P0)])
def test_ntt_product_equals_schoolbook(n, q):
    t = oracle.Tables(n, q, params.find_psi(n, q))
    rng = np.random.default_rng(n)
    a, b = rand_mod(rng, n, q), rand_mod(rng, n, q)
    prod = oracle.elemmul(oracle.forward(a, t), oracle.forward(b, t), q)
    assert np.array_equal(oracle.inverse(prod, t),
                          oracle.negacyclic_schoolbook(a, b, q))


def test_batch_driver_matches_single():
    n = 64
    primes = params.ntt_primes(3, bits=50, two_adicity=10)
    tables = [oracle.Tables(n, q, params.find_psi(n, q)) for q in primes]
    rng = np.random.default_rng(5)
    x = np.concatenate([rand_mod(rng, n, tables[p % 3].q) for p in range(7)])
    fwd = oracle.forward_batch(x, tables, threads=2)
    for p in range(7):
        assert np.array_equal(fwd[p * n:(p + 1) * n],
                              oracle.forward(x[p * n:(p + 1) * n],
                                             tables[p % 3]))
    assert np.array_equal(oracle.inverse_batch(fwd, tables, threads=2), x)


def test_reference_host_barrett_limit():
    """Domain of the reference: its host Barrett (numbers.c:5-28) is exact up
    to 62-bit moduli and overflows for 63-bit ones, so its tables are only
    defined for q < 2^62.  Above that the oracle states the contract instead
    (oracle.Tables._fill_exact) -- checked here to coincide with the literal
    restatement where both are defined."""
    rng = np.random.default_rng(63)
    q62 = params.Q62_LAZY_MAX
    for _ in range(2000):
        a, b = (int(v) for v in rand_mod(rng, 2, q62))
        assert oracle.multiply_mod(a, b, q62) == a * b % q62
    q63 = params.Q63_STRICT
    wrong = 0
    for _ in range(2000):
        a, b = (int(v) for v in rand_mod(rng, 2, q63))
        wrong += oracle.multiply_mod(a, b, q63) != a * b % q63
    assert wrong > 0
    n, q = 128, params.P0
    w = params.find_psi(n, q)
    lit = oracle.Tables(n, q, w)
    exact = oracle.Tables(n, q, w)
    exact._fill_exact()
    for f in ("roots", "inv_roots", "roots_shoup", "inv_roots_shoup"):
        assert np.array_equal(getattr(lit, f), getattr(exact, f))
    # 63-bit modulus: evaluation identity holds with the contract tables
    n = 8
    w = params.find_psi(n, q63)
    t = oracle.Tables(n, q63, w)
    x = rand_mod(rng, n, q63)
    fwd = oracle.forward(x, t)
    for j in range(n):
        point = pow(w, 2 * brv(j, 3) + 1, q63)
        assert int(fwd[j]) == sum(int(c) * pow(point, k, q63)
                                  for k, c in enumerate(x)) % q63
    assert np.array_equal(oracle.inverse(fwd, t), x)
