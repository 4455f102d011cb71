"""GPU suite: the reference's one-vector-per-call transforms
(vkhel_vector_forward_transform / _inverse_transform, src/vector.c:513-657)
are recorded and launched as indirect batches (vkhel_b200/csrc/vector.cu).
Whatever the call pattern, the results must be those of immediate launches:
every case is checked against the CPU oracle."""
import numpy as np
import pytest

import oracle
import vkhel_b200 as vk
from vkhel_b200 import params
from conftest import rand_mod

pytestmark = pytest.mark.gpu


class Tables:
    def __init__(self, n, q):
        self.n, self.q = n, q
        self.w = params.find_psi(n, q)
        self.lib = vk.NttTables(n, q, self.w)
        self.ora = oracle.Tables(n, q, self.w)

    def destroy(self):
        self.lib.destroy()


@pytest.mark.parametrize("log2n,count", [(3, 40), (8, 33), (9, 17), (12, 64),
                                         (16, 5)])
def test_loop_over_vectors_is_one_batch(ctx, log2n, count):
    n = 1 << log2n
    t = Tables(n, params.P0)
    rng = np.random.default_rng(log2n)
    xs = [rand_mod(rng, n, t.q) for _ in range(count)]
    vs = [ctx.from_host(x) for x in xs]
    outs = [ctx.vector(n, zero=False) for _ in range(count)]
    ctx.sync()
    b0, t0 = ctx.deferred_stats
    for v, o in zip(vs, outs):
        ctx.forward_transform(v, o, t.lib)
    for o in outs:                      # in place, other direction: new batch
        ctx.inverse_transform(o, o, t.lib)
    for v in vs:                        # in place forward
        ctx.forward_transform(v, v, t.lib)
    ctx.sync()
    b1, t1 = ctx.deferred_stats
    assert (b1 - b0, t1 - t0) == (3, 3 * count)
    for x, v, o in zip(xs, vs, outs):
        assert np.array_equal(o.to_host(), x)
        assert np.array_equal(v.to_host(), oracle.forward(x, t.ora))
    for v in vs + outs:
        v.destroy()
    t.destroy()


def test_dependent_transforms_are_not_reordered(ctx):
    n = 1 << 10
    t = Tables(n, params.P0)
    rng = np.random.default_rng(1)
    x, y = rand_mod(rng, n, t.q), rand_mod(rng, n, t.q)
    a, b, c = ctx.from_host(x), ctx.vector(n), ctx.from_host(y)
    ctx.forward_transform(a, b, t.lib)      # b = F(x)
    ctx.forward_transform(b, b, t.lib)      # RAW + WAW on b: b = F(F(x))
    ctx.forward_transform(c, a, t.lib)      # WAR on a: a = F(y)
    ctx.forward_transform(a, c, t.lib)      # RAW on a: c = F(F(y))
    fx, fy = oracle.forward(x, t.ora), oracle.forward(y, t.ora)
    assert np.array_equal(b.to_host(), oracle.forward(fx, t.ora))
    assert np.array_equal(a.to_host(), fy)
    assert np.array_equal(c.to_host(), oracle.forward(fy, t.ora))
    # the same vector three times in place
    v = ctx.from_host(x)
    for _ in range(3):
        ctx.forward_transform(v, v, t.lib)
    want = x
    for _ in range(3):
        want = oracle.forward(want, t.ora)
    assert np.array_equal(v.to_host(), want)
    for vec in (a, b, c, v):
        vec.destroy()
    t.destroy()


def test_shared_operand_and_mixed_tables(ctx):
    n1, n2 = 1 << 7, 1 << 11
    t1, t2 = Tables(n1, params.P0), Tables(n2, params.ntt_primes(2)[1])
    rng = np.random.default_rng(2)
    x1, x2 = rand_mod(rng, n1, t1.q), rand_mod(rng, n2, t2.q)
    a1, a2 = ctx.from_host(x1), ctx.from_host(x2)
    outs1 = [ctx.vector(n1) for _ in range(4)]
    outs2 = [ctx.vector(n2) for _ in range(4)]
    for o1, o2 in zip(outs1, outs2):        # tables alternate: batches of 1
        ctx.forward_transform(a1, o1, t1.lib)
        ctx.forward_transform(a2, o2, t2.lib)
    f1, f2 = oracle.forward(x1, t1.ora), oracle.forward(x2, t2.ora)
    for o1, o2 in zip(outs1, outs2):
        assert np.array_equal(o1.to_host(), f1)
        assert np.array_equal(o2.to_host(), f2)
    # one operand read by many recorded transforms (no hazard)
    b0, _ = ctx.deferred_stats
    for o in outs1:
        ctx.inverse_transform(a1, o, t1.lib)
    ctx.flush()
    assert ctx.deferred_stats[0] == b0 + 1
    i1 = oracle.inverse(x1, t1.ora)
    for o in outs1:
        assert np.array_equal(o.to_host(), i1)
    for vec in [a1, a2] + outs1 + outs2:
        vec.destroy()
    t1.destroy(), t2.destroy()


@pytest.mark.parametrize("log2n,limbs,batch,extra", [(10, 5, 7, 0),
                                                     (13, 3, 4, 2),
                                                     (16, 4, 2, 0)])
def test_rns_polynomials_as_one_vector_per_limb(ctx, log2n, limbs, batch,
                                                 extra):
    """the HE pattern of the reference API: every limb of every polynomial is
    its own vector with its own tables.  Equal counts per table: one launch
    laid out [batch][limbs]; unequal: one launch per table."""
    n = 1 << log2n
    primes = params.ntt_primes(limbs)
    ts = [Tables(n, q) for q in primes]
    rng = np.random.default_rng(limbs * batch)
    xs = [[rand_mod(rng, n, primes[l]) for l in range(limbs)]
          for _ in range(batch + extra)]
    vs = [[ctx.from_host(x) for x in row] for row in xs[:batch]]
    vs += [[ctx.from_host(xs[batch + e][0])] for e in range(extra)]
    ctx.sync()
    b0, t0 = ctx.deferred_stats
    for row in vs:
        for l, v in enumerate(row):
            ctx.forward_transform(v, v, ts[l].lib)
    ctx.flush()
    b1, t1 = ctx.deferred_stats
    assert t1 - t0 == batch * limbs + extra
    assert b1 - b0 == (1 if extra == 0 else limbs)
    for row, xrow in zip(vs, xs):
        for l, (v, x) in enumerate(zip(row, xrow)):
            assert np.array_equal(v.to_host(), oracle.forward(x, ts[l].ora))
    for row in vs:
        for l, v in enumerate(row):
            ctx.inverse_transform(v, v, ts[l].lib)
    for row, xrow in zip(vs, xs):
        for v, x in zip(row, xrow):
            assert np.array_equal(v.to_host(), x)
            v.destroy()
    for t in ts:
        t.destroy()


def test_recorded_transforms_meet_other_operations(ctx):
    """element-wise ops, transfers, dup, destroy of vectors and of the tables
    while transforms are still only recorded"""
    n = 1 << 9
    t = Tables(n, params.P0)
    q = t.q
    rng = np.random.default_rng(3)
    xs = [rand_mod(rng, n, q) for _ in range(6)]
    vs = [ctx.from_host(x) for x in xs]
    fs = [oracle.forward(x, t.ora) for x in xs]
    for v in vs:
        ctx.forward_transform(v, v, t.lib)
    prod = ctx.vector(n)
    ctx.elemmul(vs[0], vs[1], prod, q)           # element-wise right after
    assert np.array_equal(prod.to_host(), oracle.elemmul(fs[0], fs[1], q))
    for v in vs:
        ctx.inverse_transform(v, v, t.lib)
    d = vs[2].dup()                              # dup sees the transform
    assert np.array_equal(d.to_host(), xs[2])
    # download / upload of a vector with a recorded transform
    host = vk.host_alloc(n)
    ctx.forward_transform(vs[3], vs[3], t.lib)
    vs[3].download(host)
    ctx.sync()
    assert np.array_equal(host.array, fs[3])
    ctx.forward_transform(vs[4], vs[4], t.lib)
    host.array[:] = xs[5]
    vs[4].upload(host)                           # overwrites the result
    ctx.forward_transform(vs[4], vs[4], t.lib)
    assert np.array_equal(vs[4].to_host(), fs[5])
    # destroy a vector, then the tables, with transforms recorded
    out = ctx.vector(n)
    ctx.forward_transform(vs[0], out, t.lib)
    tmp = ctx.from_host(xs[1])
    ctx.forward_transform(tmp, tmp, t.lib)
    tmp.destroy()
    t.destroy()
    assert np.array_equal(out.to_host(), fs[0])
    for vec in vs + [prod, d, out]:
        vec.destroy()
    host.free()


def test_more_vectors_than_one_batch_holds(ctx):
    n, count = 8, 4096 + 700
    t = Tables(n, params.P0)
    rng = np.random.default_rng(4)
    x = rand_mod(rng, n * count, t.q)
    vs = [ctx.from_host(x[i * n:(i + 1) * n]) for i in range(count)]
    for v in vs:
        ctx.forward_transform(v, v, t.lib)
    want = oracle.forward_batch(x, [t.ora], threads=8)
    got = np.concatenate([v.to_host() for v in vs])
    assert np.array_equal(got, want)
    for v in vs:
        v.destroy()
    t.destroy()


def test_inverse_with_tail_and_unsupported_moduli_bypass_the_queue(ctx):
    n = 64
    t = Tables(n, params.P0)
    ts = Tables(n, params.Q63_STRICT)
    rng = np.random.default_rng(5)
    x = rand_mod(rng, n + 9, t.q)
    v = ctx.from_host(x)
    ctx.inverse_transform(v, v, t.lib)           # result longer than n
    want = oracle.inverse(x[:n], t.ora, out_len=n + 9, out_init=x)
    assert np.array_equal(v.to_host(), want)
    y = rand_mod(rng, n, ts.q)
    w = ctx.from_host(y)
    ctx.forward_transform(w, w, ts.lib)          # strict path, 63-bit modulus
    assert np.array_equal(w.to_host(), oracle.forward(y, ts.ora))
    v.destroy(), w.destroy(), t.destroy(), ts.destroy()


# ---- recorded product: elemmul fused into the in-place inverse that follows --------------
def product_oracle(a, b, t):
    return oracle.inverse(oracle.elemmul(a, b, t.q), t.ora)


@pytest.mark.parametrize("log2n", [3, 5, 8, 9, 11, 12, 14, 16])
@pytest.mark.parametrize("q", [params.P0, params.Q61, params.Q62_LAZY_MAX])
def test_reference_polymul_sequence_is_fused(ctx, log2n, q):
    """forward, forward, elemmul, inverse in place -- the reference's product
    (examples/example.c:18-60): same result as the separate launches, and the
    product went out inside the inverse transform"""
    n = 1 << log2n
    if (q - 1) % (2 * n):
        pytest.skip("modulus without a 2n-th root of unity")
    t = Tables(n, q)
    rng = np.random.default_rng(log2n)
    x, y = rand_mod(rng, n, q), rand_mod(rng, n, q)
    a, b = ctx.from_host(x), ctx.from_host(y)
    c = ctx.vector(n)
    f0 = ctx.fused_products
    ctx.forward_transform(a, a, t.lib)
    ctx.forward_transform(b, b, t.lib)
    ctx.elemmul(a, b, c, q)
    ctx.inverse_transform(c, c, t.lib)
    fx, fy = oracle.forward(x, t.ora), oracle.forward(y, t.ora)
    assert np.array_equal(c.to_host(), product_oracle(fx, fy, t))
    assert np.array_equal(a.to_host(), fx)      # the factors are untouched
    assert np.array_equal(b.to_host(), fy)
    assert ctx.fused_products == f0 + 1
    for v in (a, b, c):
        v.destroy()
    t.destroy()


@pytest.mark.parametrize("alias", ["a", "b", "both"])
def test_fused_product_in_place(ctx, alias):
    n = 1 << 12
    t = Tables(n, params.P0)
    rng = np.random.default_rng(7)
    x, y = rand_mod(rng, n, t.q), rand_mod(rng, n, t.q)
    a, b = ctx.from_host(x), ctx.from_host(y)
    f0 = ctx.fused_products
    if alias == "a":
        ctx.elemmul(a, b, a, t.q)
        ctx.inverse_transform(a, a, t.lib)
        got, want = a.to_host(), product_oracle(x, y, t)
        assert np.array_equal(b.to_host(), y)
    elif alias == "b":
        ctx.elemmul(a, b, b, t.q)
        ctx.inverse_transform(b, b, t.lib)
        got, want = b.to_host(), product_oracle(x, y, t)
        assert np.array_equal(a.to_host(), x)
    else:
        ctx.elemmul(a, a, a, t.q)
        ctx.inverse_transform(a, a, t.lib)
        got, want = a.to_host(), product_oracle(x, x, t)
    assert np.array_equal(got, want)
    assert ctx.fused_products == f0 + 1
    a.destroy()
    b.destroy()
    t.destroy()


def test_fused_product_of_arbitrary_64_bit_factors(ctx):
    """elemmul reduces both factors (reference elemmul.comp:62-73); the fused
    kernel must do the same for operands that are not canonical residues"""
    n = 1 << 10
    t = Tables(n, params.P0)
    rng = np.random.default_rng(11)
    x = rng.integers(0, 1 << 64, n, dtype=np.uint64)
    y = rng.integers(0, 1 << 64, n, dtype=np.uint64)
    x[:3] = np.array([2**64 - 1, t.q, t.q - 1], dtype=np.uint64)
    y[:3] = np.array([2**64 - 1, 2**64 - 1, t.q + 1], dtype=np.uint64)
    a, b = ctx.from_host(x), ctx.from_host(y)
    c = ctx.vector(n)
    f0 = ctx.fused_products
    ctx.elemmul(a, b, c, t.q)
    ctx.inverse_transform(c, c, t.lib)
    assert np.array_equal(c.to_host(), product_oracle(x, y, t))
    assert ctx.fused_products == f0 + 1
    for v in (a, b, c):
        v.destroy()
    t.destroy()


def test_recorded_product_is_launched_when_anything_else_follows(ctx):
    n = 1 << 11
    t = Tables(n, params.P0)
    t2 = Tables(n, params.Q61)
    rng = np.random.default_rng(13)
    x, y = rand_mod(rng, n, t.q), rand_mod(rng, n, t.q)
    prod = oracle.elemmul(x, y, t.q)
    a, b = ctx.from_host(x), ctx.from_host(y)
    f0 = ctx.fused_products

    # read back right away
    c = ctx.vector(n)
    ctx.elemmul(a, b, c, t.q)
    assert np.array_equal(c.to_host(), prod)
    ctx.inverse_transform(c, c, t.lib)
    assert np.array_equal(c.to_host(), oracle.inverse(prod, t.ora))

    # inverse out of place: the product stays visible
    d = ctx.vector(n)
    ctx.elemmul(a, b, c, t.q)
    ctx.inverse_transform(c, d, t.lib)
    assert np.array_equal(d.to_host(), oracle.inverse(prod, t.ora))
    assert np.array_equal(c.to_host(), prod)

    # another modulus' tables
    ctx.elemmul(a, b, c, t.q)
    ctx.inverse_transform(c, c, t2.lib)
    assert np.array_equal(c.to_host(), oracle.inverse(prod, t2.ora))

    # a second product on top of the first, a forward transform, elemfma, dup
    ctx.elemmul(a, b, c, t.q)
    ctx.elemmul(c, b, d, t.q)
    ctx.forward_transform(d, d, t.lib)
    assert np.array_equal(c.to_host(), prod)
    assert np.array_equal(
        d.to_host(), oracle.forward(oracle.elemmul(prod, y, t.q), t.ora))
    ctx.elemmul(a, b, c, t.q)
    ctx.elemfma(c, b, d, 3, t.q)
    assert np.array_equal(d.to_host(), oracle.elemfma(prod, y, 3, t.q))
    ctx.elemmul(a, b, c, t.q)
    e = c.dup()
    assert np.array_equal(e.to_host(), prod)

    # an operand is overwritten / destroyed before the inverse
    ctx.elemmul(a, b, c, t.q)
    a.copy_from_host(y)
    ctx.inverse_transform(c, c, t.lib)
    assert np.array_equal(c.to_host(), oracle.inverse(prod, t.ora))
    ctx.elemmul(a, b, c, t.q)           # a == y now
    b.destroy()
    ctx.inverse_transform(c, c, t.lib)
    assert np.array_equal(c.to_host(),
                          oracle.inverse(oracle.elemmul(y, y, t.q), t.ora))
    assert ctx.fused_products == f0

    # result longer than the transform: the tail is scaled, nothing is fused
    big = ctx.from_host(np.concatenate([x, x]))
    big2 = ctx.from_host(np.concatenate([y, y]))
    ctx.elemmul(big, big2, big, t.q)
    ctx.inverse_transform(big, big, t.lib)
    want = oracle.inverse(prod, t.ora, out_len=2 * n,
                          out_init=np.concatenate([prod, prod]))
    assert np.array_equal(big.to_host(), want)
    assert ctx.fused_products == f0
    for v in (a, c, d, e, big, big2):
        v.destroy()
    t.destroy()
    t2.destroy()


@pytest.mark.parametrize("log2n,mult", [(3, 1), (8, 1), (10, 5), (12, 1),
                                        (14, 0), (16, 1), (16, 123456789)])
def test_elemfma_followed_by_inverse_is_fused(ctx, log2n, mult):
    """The point-wise ADD of the transform domain (reference
    src/vector.c:298-340: elemfma, "modular add" when the multiplier is 1)
    followed by the in-place inverse transform of its result: recorded like the
    product and applied while the inverse transform loads, same result as the
    separate launches.  Arbitrary 64-bit a, multipliers 0, 1, > 1 and >= q."""
    n, q = 1 << log2n, params.P0
    t = Tables(n, q)
    rng = np.random.default_rng(100 + log2n + mult)
    x = rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    y = rand_mod(rng, n, q)
    a, b, c = ctx.from_host(x), ctx.from_host(y), ctx.vector(n)
    f0 = ctx.fused_products
    for m in (mult, mult + q):                   # the multiplier is reduced mod q
        ctx.elemfma(a, b, c, m % (1 << 64), q)
        ctx.inverse_transform(c, c, t.lib)
        want = oracle.inverse(oracle.elemfma(x, y, mult % q, q), t.ora)
        assert np.array_equal(c.to_host(), want)
    assert ctx.fused_products == f0 + 2
    assert np.array_equal(a.to_host(), x) and np.array_equal(b.to_host(), y)
    # accumulate in place, then transform: result aliases the addend
    ctx.elemfma(a, b, b, 1, q)
    ctx.inverse_transform(b, b, t.lib)
    assert np.array_equal(
        b.to_host(), oracle.inverse(oracle.elemfma(x, y, 1, q), t.ora))
    assert ctx.fused_products == f0 + 3
    # anything else in between: the sum is launched on its own
    ctx.elemfma(a, b, c, 1, q)
    got = c.to_host()
    assert np.array_equal(got, oracle.elemfma(x, b.to_host(), 1, q))
    assert ctx.fused_products == f0 + 3
    for v in (a, b, c):
        v.destroy()
    t.destroy()


# ---- recorded whole products (n <= 2^11) ----------------------------------------------------------
def product_quad(ctx, a, b, c, t, swap=False):
    """the reference's four calls (examples/example.c:18-60)"""
    ctx.forward_transform(a, a, t.lib)
    ctx.forward_transform(b, b, t.lib)
    if swap:
        ctx.elemmul(b, a, c, t.q)
    else:
        ctx.elemmul(a, b, c, t.q)
    ctx.inverse_transform(c, c, t.lib)


@pytest.mark.parametrize("log2n,count", [(12, 1), (12, 9), (13, 20), (14, 5),
                                         (16, 3), (17, 2)])
def test_loop_of_products_at_two_pass_sizes_is_four_launches(ctx, log2n, count):
    """The same loop above n = 2^11: the forward transforms of all products
    stay in the record and go out as one indirect batch, the inverse
    transforms of the products as a second one that multiplies while it loads
    -- four launches for the whole loop instead of four per product."""
    n, q = 1 << log2n, params.P0
    t = Tables(n, q)
    rng = np.random.default_rng(7200 + log2n)
    xs = [rand_mod(rng, n, q) for _ in range(count)]
    ys = [rand_mod(rng, n, q) for _ in range(count)]
    va = [ctx.from_host(x) for x in xs]
    vb = [ctx.from_host(y) for y in ys]
    vc = [ctx.vector(n) for _ in range(count)]
    ctx.sync()
    l0, f0 = ctx.launch_count, ctx.fused_products
    for i in range(count):
        product_quad(ctx, va[i], vb[i], vc[i], t, swap=bool(i & 1))
    assert ctx.launch_count_noflush == l0
    ctx.flush()
    assert ctx.launch_count == l0 + 4
    assert ctx.fused_products == f0 + count
    for i in range(count):
        fx, fy = oracle.forward(xs[i], t.ora), oracle.forward(ys[i], t.ora)
        assert np.array_equal(vc[i].to_host(), product_oracle(fx, fy, t)), i
        assert np.array_equal(va[i].to_host(), fx), i
        assert np.array_equal(vb[i].to_host(), fy), i
    for v in va + vb + vc:
        v.destroy()
    t.destroy()


@pytest.mark.parametrize("log2n,count", [(3, 5), (4, 3), (6, 9), (8, 40),
                                         (9, 7), (10, 33), (11, 6)])
def test_loop_of_small_products_is_one_launch(ctx, log2n, count):
    """A loop of four-call products over separate vector triples at n <= 2^11:
    every product is recorded as a unit and the whole loop goes out as ONE
    launch (kernels_ntt_small.cu); products and the (still observable) forward
    transforms equal the oracle's."""
    n, q = 1 << log2n, params.P0
    t = Tables(n, q)
    rng = np.random.default_rng(7000 + log2n)
    xs = [rand_mod(rng, n, q) for _ in range(count)]
    ys = [rand_mod(rng, n, q) for _ in range(count)]
    va = [ctx.from_host(x) for x in xs]
    vb = [ctx.from_host(y) for y in ys]
    vc = [ctx.vector(n) for _ in range(count)]
    ctx.sync()
    l0, f0 = ctx.launch_count, ctx.fused_products
    for i in range(count):
        product_quad(ctx, va[i], vb[i], vc[i], t, swap=bool(i & 1))
    assert ctx.launch_count_noflush == l0          # nothing launched yet
    ctx.flush()
    assert ctx.launch_count == l0 + 1
    assert ctx.fused_products == f0 + count
    for i in range(count):
        fx, fy = oracle.forward(xs[i], t.ora), oracle.forward(ys[i], t.ora)
        assert np.array_equal(vc[i].to_host(), product_oracle(fx, fy, t)), i
        assert np.array_equal(va[i].to_host(), fx), i
        assert np.array_equal(vb[i].to_host(), fy), i
    for v in va + vb + vc:
        v.destroy()
    t.destroy()


@pytest.mark.parametrize("log2n", [10, 13])
def test_recorded_products_out_of_place_reused_vectors_and_interruptions(ctx, log2n):
    """whole products (n = 2^10) and batched inverse-of-products (n = 2^13)
    under every way of not completing the four-call sequence"""
    n, q = 1 << log2n, params.P0
    t = Tables(n, q)
    t2 = Tables(n, params.Q61)
    rng = np.random.default_rng(7100)
    x, y, z = (rand_mod(rng, n, q) for _ in range(3))
    fx, fy, fz = (oracle.forward(v, t.ora) for v in (x, y, z))
    a, b, w = ctx.from_host(x), ctx.from_host(y), ctx.from_host(z)
    fa, fb, c, c2 = (ctx.vector(n) for _ in range(4))
    # forward transforms out of place: the sources stay what they were
    ctx.forward_transform(a, fa, t.lib)
    ctx.forward_transform(b, fb, t.lib)
    ctx.elemmul(fa, fb, c, q)
    ctx.inverse_transform(c, c, t.lib)
    # a recorded transform of an unrelated vector shares the record
    ctx.forward_transform(w, w, t.lib)
    # the same operands again, into another result: depends on nothing recorded
    # that it overwrites, but re-transforms fa/fb -> launched in order
    ctx.forward_transform(a, fa, t.lib)
    ctx.forward_transform(b, fb, t.lib)
    ctx.elemmul(fa, fb, c2, q)
    ctx.inverse_transform(c2, c2, t.lib)
    want = product_oracle(fx, fy, t)
    assert np.array_equal(c.to_host(), want)
    assert np.array_equal(c2.to_host(), want)
    assert np.array_equal(a.to_host(), x) and np.array_equal(fa.to_host(), fx)
    assert np.array_equal(w.to_host(), fz)
    # interrupted: the product is read before any inverse transform
    f0 = ctx.fused_products
    ctx.forward_transform(a, fa, t.lib)
    ctx.forward_transform(b, fb, t.lib)
    ctx.elemmul(fa, fb, c, q)
    assert np.array_equal(c.to_host(), oracle.elemmul(fx, fy, q))
    # ... the inverse goes elsewhere, or uses other tables
    ctx.forward_transform(a, fa, t.lib)
    ctx.forward_transform(b, fb, t.lib)
    ctx.elemmul(fa, fb, c, q)
    ctx.inverse_transform(c, c2, t.lib)
    assert np.array_equal(c2.to_host(), want)
    assert np.array_equal(c.to_host(), oracle.elemmul(fx, fy, q))
    ctx.forward_transform(a, fa, t.lib)
    ctx.forward_transform(b, fb, t.lib)
    ctx.elemmul(fa, fb, c, q)
    ctx.inverse_transform(c, c, t2.lib)
    assert np.array_equal(c.to_host(),
                          oracle.inverse(oracle.elemmul(fx, fy, q), t2.ora))
    # ... another transform is recorded, or an operand is destroyed, in between
    ctx.forward_transform(a, fa, t.lib)
    ctx.forward_transform(b, fb, t.lib)
    ctx.elemmul(fa, fb, c, q)
    ctx.forward_transform(w, w, t.lib)          # w = forward(forward(z))
    ctx.inverse_transform(c, c, t.lib)
    assert np.array_equal(c.to_host(), want)
    assert np.array_equal(w.to_host(), oracle.forward(fz, t.ora))
    ctx.forward_transform(a, fa, t.lib)
    ctx.forward_transform(b, fb, t.lib)
    ctx.elemmul(fa, fb, c, q)
    fb.destroy()
    ctx.inverse_transform(c, c, t.lib)
    assert np.array_equal(c.to_host(), want)
    assert ctx.fused_products == f0
    # squaring and in-place products are not whole-product candidates
    ctx.forward_transform(a, fa, t.lib)
    ctx.elemmul(fa, fa, c, q)
    ctx.inverse_transform(c, c, t.lib)
    assert np.array_equal(c.to_host(), product_oracle(fx, fx, t))
    # tables destroyed while a recorded product still needs them
    ctx.forward_transform(a, fa, t.lib)
    ctx.forward_transform(w, w, t.lib)
    ctx.elemmul(fa, w, c, q)
    ctx.inverse_transform(c, c, t.lib)
    fw = oracle.forward(oracle.forward(fz, t.ora), t.ora)
    t.destroy()
    t3 = Tables(n, q)
    assert np.array_equal(c.to_host(), product_oracle(fx, fw, t3))
    for v in (a, b, w, fa, c, c2):
        v.destroy()
    t2.destroy(), t3.destroy()
