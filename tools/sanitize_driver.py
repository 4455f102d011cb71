#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck):
every kernel family once, checked against the oracle.

    compute-sanitizer --tool racecheck python tools/sanitize_driver.py
"""
import os
import sys

import numpy as np

# small RNS batches already take the two-stream limb slices with the lazy join
os.environ.setdefault("VKHEL_SPLIT_SMALL_MIB", "0.25")

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import vkhel_b200 as vk  # noqa: E402
from vkhel_b200 import params  # noqa: E402


def variants():
    """the opt-in kernels: thread-block-cluster single pass ($VKHEL_CLUSTER=1)
    and the TMA tensor-map column load ($VKHEL_COLS_TMA=1); run with both set"""
    rng = np.random.default_rng(1)
    ctx = vk.Context(0)
    for log2n, limbs, batch in [(14, 2, 2), (15, 1, 3), (16, 2, 1)]:
        n = 1 << log2n
        primes = params.ntt_primes(limbs)
        tabs = [vk.NttTables(n, q, params.find_psi(n, q)) for q in primes]
        oras = [oracle.Tables(n, q, t.w) for q, t in zip(primes, tabs)]
        x = np.concatenate([rng.integers(0, primes[p % limbs], n, dtype=np.uint64)
                            for p in range(limbs * batch)])
        a, b = ctx.from_host(x), ctx.vector(x.size)
        ctx.forward_transform_rns(a, b, tabs, batch)
        assert np.array_equal(b.to_host(), oracle.forward_batch(x, oras, threads=4))
        ctx.inverse_transform_rns(b, b, tabs, batch)
        assert np.array_equal(b.to_host(), x)
        a.destroy(), b.destroy()
        for t in tabs:
            t.destroy()
    ctx.destroy()
    print("sanitize driver ok")


def main():
    if "--variants" in sys.argv:
        return variants()
    rng = np.random.default_rng(0)
    ctx = vk.Context(0)
    # (log2n, limbs, batch): row-only, column+row (several tile shapes),
    # 512-point column tiles, leading generic pass, generic tiny sizes
    shapes = [(2, 1, 3), (5, 1, 9), (8, 2, 3), (10, 1, 5), (12, 3, 2),
              (14, 1, 2), (16, 2, 1), (17, 1, 1)]
    for log2n, limbs, batch in shapes:
        n = 1 << log2n
        primes = params.ntt_primes(limbs)
        tabs = [vk.NttTables(n, q, params.find_psi(n, q)) for q in primes]
        oras = [oracle.Tables(n, q, t.w) for q, t in zip(primes, tabs)]
        polys = limbs * batch
        x = np.concatenate([rng.integers(0, primes[p % limbs], n, dtype=np.uint64)
                            for p in range(polys)])
        a, b = ctx.from_host(x), ctx.vector(x.size)
        ctx.forward_transform_rns(a, b, tabs, batch)
        assert np.array_equal(b.to_host(), oracle.forward_batch(x, oras, threads=4))
        ctx.inverse_transform_rns(b, b, tabs, batch)
        assert np.array_equal(b.to_host(), x)
        ctx.elemmul_rns(a, b, b, primes, n, batch) if n >= 2 else None
        a.destroy(), b.destroy()
        for t in tabs:
            t.destroy()
    # strict path (63-bit modulus) and the legacy entry points
    n, q = 256, params.Q63_STRICT
    t = vk.NttTables(n, q, params.find_psi(n, q))
    x = rng.integers(0, q, n + 5, dtype=np.uint64)
    v = ctx.from_host(x)
    ctx.forward_transform(v, v, t)
    ctx.inverse_transform(v, v, t)
    v.to_host()
    # recorded one-vector transforms going out as indirect batches, and the
    # fused polynomial product
    for log2n in (5, 9, 13):
        n, q = 1 << log2n, params.P0
        t2 = vk.NttTables(n, q, params.find_psi(n, q))
        o2 = oracle.Tables(n, q, t2.w)
        xs = [rng.integers(0, q, n, dtype=np.uint64) for _ in range(5)]
        vs = [ctx.from_host(x) for x in xs]
        for vec in vs:
            ctx.forward_transform(vec, vec, t2)
        for x, vec in zip(xs, vs):
            assert np.array_equal(vec.to_host(), oracle.forward(x, o2))
        for vec in vs:
            ctx.inverse_transform(vec, vec, t2)
        assert np.array_equal(vs[4].to_host(), xs[4])
        ctx.polymul_rns(vs[0], vs[1], vs[2], [t2], 1)
        prod = oracle.elemmul(oracle.forward(xs[0], o2),
                              oracle.forward(xs[1], o2), q)
        assert np.array_equal(vs[2].to_host(), oracle.inverse(prod, o2))
        # the reference's product sequence: two recorded forward transforms,
        # the product fused into the in-place inverse (row pass / single pass)
        fused0 = ctx.fused_products
        ctx.forward_transform(vs[3], vs[3], t2)
        ctx.forward_transform(vs[4], vs[4], t2)
        ctx.elemmul(vs[3], vs[4], vs[0], q)
        ctx.inverse_transform(vs[0], vs[0], t2)
        prod = oracle.elemmul(oracle.forward(xs[3], o2),
                              oracle.forward(xs[4], o2), q)
        assert np.array_equal(vs[0].to_host(), oracle.inverse(prod, o2))
        assert ctx.fused_products == fused0 + 1
        # a record longer than the inline pointer table (device-side table)
        more = [ctx.from_host(xs[i % 5]) for i in range(11)]
        for vec in more:
            ctx.forward_transform(vec, vec, t2)
        assert np.array_equal(more[10].to_host(), oracle.forward(xs[0], o2))
        for vec in vs + more:
            vec.destroy()
        t2.destroy()
    # a loop of the reference's four-call products at small n: recorded whole
    # products, one launch (inline pointer table for 3, device table for 7)
    for log2n, count in ((5, 3), (9, 7)):
        n, q = 1 << log2n, params.P0
        t3 = vk.NttTables(n, q, params.find_psi(n, q))
        o3 = oracle.Tables(n, q, t3.w)
        xs = [rng.integers(0, q, n, dtype=np.uint64) for _ in range(2 * count)]
        va = [ctx.from_host(x) for x in xs[:count]]
        vb = [ctx.from_host(x) for x in xs[count:]]
        vc = [ctx.vector(n) for _ in range(count)]
        for i in range(count):
            ctx.forward_transform(va[i], va[i], t3)
            ctx.forward_transform(vb[i], vb[i], t3)
            ctx.elemmul(va[i], vb[i], vc[i], q)
            ctx.inverse_transform(vc[i], vc[i], t3)
        for i in range(count):                     # maps with read-ahead
            prod = oracle.elemmul(oracle.forward(xs[i], o3),
                                  oracle.forward(xs[count + i], o3), q)
            assert np.array_equal(vc[i].to_host(), oracle.inverse(prod, o3))
        # the point-wise add folded into the inverse transform
        ctx.elemfma(va[0], vb[0], vc[0], 1, q)
        ctx.inverse_transform(vc[0], vc[0], t3)
        want = oracle.inverse(oracle.elemfma(va[0].to_host(), vb[0].to_host(),
                                             1, q), o3)
        assert np.array_equal(vc[0].to_host(), want)
        for vec in va + vb + vc:
            vec.destroy()
        t3.destroy()
    # limb slices on two streams with the lazy join: a chain of transforms of
    # the same vector, then an element-wise operation that has to join
    n, limbs, batch = 1 << 12, 4, 4
    primes = params.ntt_primes(limbs)
    tabs = [vk.NttTables(n, q, params.find_psi(n, q)) for q in primes]
    x = np.concatenate([rng.integers(0, primes[p % limbs], n, dtype=np.uint64)
                        for p in range(limbs * batch)])
    a, b = ctx.from_host(x), ctx.vector(x.size)
    for _ in range(3):
        ctx.forward_transform_rns(a, b, tabs, batch)
        ctx.inverse_transform_rns(b, b, tabs, batch)
    # (the second and third forward transform were held and stored lazily)
    assert ctx.lazy_forwards == 2 or os.environ.get("VKHEL_LAZY_FORWARD") == "0"
    ctx.elemgtadd(b, b, 0, 0)
    assert np.array_equal(b.to_host(), x)
    ctx.forward_transform_rns(a, b, tabs, batch)      # held, then read
    assert np.array_equal(b.to_host()[:n],
                          oracle.forward(x[:n], oracle.Tables(n, primes[0], tabs[0].w)))
    a.destroy(), b.destroy()
    for t4 in tabs:
        t4.destroy()
    q = 769
    a = ctx.from_host(rng.integers(0, 1 << 62, 1001, dtype=np.uint64))
    c = ctx.vector(1001)
    ctx.elemmul(a, a, c, q)
    ctx.elemfma(a, a, c, 5, q)
    ctx.elemgtadd(a, c, 7, 9)
    ctx.elemgtsub(a, c, 7, 9, q)
    ctx.elemmod(a, c, 2, q)
    ctx.elemmod(a, c, 3, q)
    c.to_host()
    for vec in (v, a, c):
        vec.destroy()
    t.destroy()
    ctx.destroy()
    print("sanitize driver ok")


if __name__ == "__main__":
    main()
