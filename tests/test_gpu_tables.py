"""GPU suite: twiddle tables generated on the device
(vkhel_ntt_tables_create_on, vkhel_b200/csrc/tables_device.cu) against the
host generator, whose contract is the reference's src/ntt_tables.c:17-44 and
which tests/test_host.py pins to the reference's own compiled code."""
import numpy as np
import pytest

import oracle
import vkhel_b200 as vk
from vkhel_b200 import params
from conftest import rand_mod

pytestmark = pytest.mark.gpu

CASES = [
    (4, 113, 18),                                  # reference test/ntt.c:19
    (16, params.Q_KAT_52, params.W_KAT_52_N16),    # reference test/vector.c:285
    (2, 769, None), (8, 769, None), (128, 769, None),
    (4096, params.Q61, None),
    (1 << 14, params.P0, None),
    (1 << 16, params.ntt_primes(3)[2], None),
    (1 << 17, params.P0, None),
    (1 << 12, params.Q62_LAZY_MAX, None),
    (1 << 10, params.Q63_STRICT, None),
]


@pytest.mark.parametrize("n,q,w", CASES)
def test_device_tables_equal_host_tables(ctx, n, q, w):
    w = w if w is not None else params.find_psi(n, q)
    host = vk.NttTables(n, q, w)
    dev = vk.NttTables(n, q, w, ctx=ctx)
    for name in ("roots_of_unity", "inv_roots_of_unity",
                 "roots_barrett_factors", "inv_roots_barrett_factors"):
        assert np.array_equal(getattr(dev, name), getattr(host, name)), name
    # and the mirror it left on the device drives the transforms
    ora = oracle.Tables(n, q, w)
    rng = np.random.default_rng(n)
    x = rand_mod(rng, n, q)
    v = ctx.from_host(x)
    ctx.forward_transform(v, v, dev)
    assert np.array_equal(v.to_host(), oracle.forward(x, ora))
    ctx.inverse_transform(v, v, dev)
    assert np.array_equal(v.to_host(), x)
    v.destroy(), host.destroy(), dev.destroy()


def test_reference_kat_roots(ctx):
    """test/ntt.c:19-21: roots_of_unity of (4, 113, 18) are {1, 98, 18, 69}"""
    t = vk.NttTables(4, 113, 18, ctx=ctx)
    assert t.roots_of_unity.tolist() == [1, 98, 18, 69]
    t.destroy()


def test_degenerate_and_wide_moduli_fall_back_to_the_host(ctx):
    t = vk.NttTables(1, 113, 18, ctx=ctx)
    assert t.roots_of_unity.tolist() == [1]
    t.destroy()


def test_rns_basis_generated_on_device(ctx):
    n, limbs, batch = 1 << 13, 4, 2
    primes = params.ntt_primes(limbs)
    psis = [params.find_psi(n, q) for q in primes]
    tabs = [vk.NttTables(n, q, w, ctx=ctx) for q, w in zip(primes, psis)]
    oras = [oracle.Tables(n, q, w) for q, w in zip(primes, psis)]
    rng = np.random.default_rng(13)
    x = np.concatenate([rand_mod(rng, n, primes[p % limbs])
                        for p in range(limbs * batch)])
    a = ctx.from_host(x)
    ctx.forward_transform_rns(a, a, tabs, batch)
    assert np.array_equal(a.to_host(),
                          oracle.forward_batch(x, oras, threads=4))
    # the inverse's column pass reads the scaled top of the inverse table,
    # which for device-generated tables is appended when the mirror is adopted
    ctx.inverse_transform_rns(a, a, tabs, batch)
    assert np.array_equal(a.to_host(), x)
    y = ctx.from_host(x)
    ctx.inverse_transform_rns(y, y, tabs, batch)
    assert np.array_equal(y.to_host(),
                          oracle.inverse_batch(x, oras, threads=4))
    y.destroy()
    a.destroy()
    for t in tabs:
        t.destroy()
