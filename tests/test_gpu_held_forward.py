"""GPU suite: the held forward transform (vkhel_b200/csrc/vector.cu).

A batched forward transform that is followed at once by the in-place inverse
transform of its result with the same tables may store lazy residues, because
the inverse overwrites them before anything can read them; once a
forward/inverse pair has been seen, the next forward transform is held back
until the call after it shows which it is.  Whatever follows, every value the
API hands out must be the reference's canonical one
(nttfwdbutterfly.comp:41-57, nttrevbutterfly.comp:41-57): all cases are
compared with the CPU oracle."""
import numpy as np
import pytest

import oracle
import vkhel_b200 as vk
from vkhel_b200 import params
from conftest import rand_mod

pytestmark = pytest.mark.gpu


class Basis:
    def __init__(self, ctx, log2n, limbs, first=0):
        self.n = 1 << log2n
        self.qs = params.ntt_primes(first + limbs)[first:]
        ws = [params.find_psi(self.n, q) for q in self.qs]
        self.lib = [vk.NttTables(self.n, q, w, ctx=ctx)
                    for q, w in zip(self.qs, ws)]
        self.ora = [oracle.Tables(self.n, q, w) for q, w in zip(self.qs, ws)]

    def random(self, rng, batch):
        return np.concatenate([rand_mod(rng, self.n, q)
                               for _ in range(batch) for q in self.qs])

    def destroy(self):
        for t in self.lib:
            t.destroy()


@pytest.fixture
def own_ctx():
    context = vk.Context(0)
    yield context
    context.destroy()


def trained(ctx, basis, a, b, batch):
    """one forward/inverse pair: the next forward transform will be held"""
    ctx.forward_transform_rns(a, b, basis.lib, batch)
    ctx.inverse_transform_rns(b, b, basis.lib, batch)


@pytest.mark.parametrize("log2n,limbs,batch", [(12, 3, 5), (14, 4, 2),
                                               (16, 2, 3), (16, 8, 4), (17, 2, 2)])
def test_loop_of_pairs_stores_lazily_and_round_trips(own_ctx, log2n, limbs, batch):
    ctx = own_ctx
    basis = Basis(ctx, log2n, limbs)
    rng = np.random.default_rng(log2n * 100 + limbs)
    x = basis.random(rng, batch)
    a, b = ctx.from_host(x), ctx.vector(x.size, zero=False)
    steps = 5
    for _ in range(steps):
        ctx.forward_transform_rns(a, b, basis.lib, batch)
        ctx.inverse_transform_rns(b, b, basis.lib, batch)
    assert np.array_equal(b.to_host(), x)
    # the first forward is launched at once, the others are held and lazy
    assert ctx.lazy_forwards == steps - 1
    # in place as well: a -> a -> a (the pattern is known: all three are held)
    for _ in range(3):
        ctx.forward_transform_rns(a, a, basis.lib, batch)
        ctx.inverse_transform_rns(a, a, basis.lib, batch)
    assert np.array_equal(a.to_host(), x)
    assert ctx.lazy_forwards == steps - 1 + 3
    a.destroy(), b.destroy(), basis.destroy()


def test_held_forward_read_by_anything_else_is_canonical(own_ctx):
    ctx = own_ctx
    batch = 3
    basis = Basis(ctx, 13, 4)
    other = Basis(ctx, 13, 4, first=4)
    rng = np.random.default_rng(7)
    x = basis.random(rng, batch)
    want = oracle.forward_batch(x, basis.ora)
    a, b, c = ctx.from_host(x), ctx.vector(x.size), ctx.vector(x.size)

    def held_forward():
        trained(ctx, basis, a, b, batch)
        before = ctx.launch_count_noflush
        ctx.forward_transform_rns(a, b, basis.lib, batch)
        assert ctx.launch_count_noflush == before      # held back
        return ctx.lazy_forwards

    # read through map
    lazy = held_forward()
    assert np.array_equal(b.to_host(), want)
    assert ctx.lazy_forwards == lazy
    # the pattern is broken: the next forward transform is launched at once
    before = ctx.launch_count_noflush
    ctx.forward_transform_rns(a, b, basis.lib, batch)
    assert ctx.launch_count_noflush > before
    assert np.array_equal(b.to_host(), want)

    # inverse out of place: b stays observable
    lazy = held_forward()
    ctx.inverse_transform_rns(b, c, basis.lib, batch)
    assert ctx.lazy_forwards == lazy
    assert np.array_equal(b.to_host(), want)
    assert np.array_equal(c.to_host(), x)

    # inverse in place with other tables
    lazy = held_forward()
    ctx.inverse_transform_rns(b, b, other.lib, batch)
    assert ctx.lazy_forwards == lazy
    assert np.array_equal(b.to_host(), oracle.inverse_batch(want, other.ora))

    # inverse of fewer polynomials
    lazy = held_forward()
    ctx.inverse_transform_rns(b, b, basis.lib, batch - 1)
    assert ctx.lazy_forwards == lazy
    got = b.to_host()
    cut = (batch - 1) * len(basis.qs) * basis.n
    assert np.array_equal(got[:cut], x[:cut])
    assert np.array_equal(got[cut:], want[cut:])

    # an element-wise operation on the result
    lazy = held_forward()
    ctx.elemmul_rns(b, b, c, basis.qs, basis.n, batch)
    assert ctx.lazy_forwards == lazy
    assert np.array_equal(b.to_host(), want)
    squares = np.concatenate([
        oracle.elemmul(p, p, basis.qs[i % len(basis.qs)])
        for i, p in enumerate(want.reshape(-1, basis.n))])
    assert np.array_equal(c.to_host(), squares)

    # a single-vector transform recorded after the held one reads its result
    single = vk.NttTables(basis.n, basis.qs[0], params.find_psi(basis.n, basis.qs[0]),
                          ctx=ctx)
    lazy = held_forward()
    d = ctx.vector(basis.n)
    ctx.forward_transform(b, d, single)
    assert ctx.lazy_forwards == lazy
    first = want[:basis.n]
    assert np.array_equal(d.to_host(), oracle.forward(first, basis.ora[0]))
    assert np.array_equal(b.to_host(), want)

    # ... and the matching inverse arrives while that transform is still only
    # recorded: it must read the canonical result, whatever its modulus
    foreign = vk.NttTables(basis.n, other.qs[0],
                           params.find_psi(basis.n, other.qs[0]), ctx=ctx)
    lazy = held_forward()
    ctx.forward_transform(b, d, foreign)
    ctx.inverse_transform_rns(b, b, basis.lib, batch)
    assert ctx.lazy_forwards == lazy
    assert np.array_equal(d.to_host(), oracle.forward(first, other.ora[0]))
    assert np.array_equal(b.to_host(), x)
    # the same with a recorded point-wise product of the result
    lazy = held_forward()
    e = ctx.vector(basis.n)
    ctx.elemmul(b, b, e, other.qs[0])
    ctx.inverse_transform_rns(b, b, basis.lib, batch)
    assert ctx.lazy_forwards == lazy
    assert np.array_equal(e.to_host(), oracle.elemmul(first, first, other.qs[0]))
    assert np.array_equal(b.to_host(), x)
    e.destroy(), foreign.destroy()
    for v in (a, b, c, d):
        v.destroy()
    single.destroy(), basis.destroy(), other.destroy()


def test_operand_overwritten_or_destroyed_while_held(own_ctx):
    ctx = own_ctx
    batch = 2
    basis = Basis(ctx, 14, 3)
    rng = np.random.default_rng(8)
    x, y = basis.random(rng, batch), basis.random(rng, batch)
    a, b = ctx.from_host(x), ctx.vector(x.size)
    trained(ctx, basis, a, b, batch)
    ctx.forward_transform_rns(a, b, basis.lib, batch)   # held, reads x
    a.copy_from_host(y)                                 # must come after it
    ctx.inverse_transform_rns(b, b, basis.lib, batch)
    assert np.array_equal(b.to_host(), x)
    assert np.array_equal(a.to_host(), y)
    # asynchronous upload into the operand
    trained(ctx, basis, a, b, batch)
    ctx.forward_transform_rns(a, b, basis.lib, batch)   # held, reads y
    pinned = vk.host_alloc(x.size)
    pinned.array[:] = x
    a.upload(pinned)
    ctx.inverse_transform_rns(b, b, basis.lib, batch)
    assert np.array_equal(b.to_host(), y)
    assert np.array_equal(a.to_host(), x)
    pinned.free()
    # destroyed operand, destroyed tables
    trained(ctx, basis, a, b, batch)
    ctx.forward_transform_rns(a, b, basis.lib, batch)   # held, reads x
    a.destroy()
    want = oracle.forward_batch(x, basis.ora)
    basis.destroy()
    assert np.array_equal(b.to_host(), want)
    b.destroy()


def test_single_modulus_batch_is_held_too(own_ctx):
    ctx = own_ctx
    n, batch = 1 << 14, 6
    q = params.P0
    w = params.find_psi(n, q)
    lib, ora = vk.NttTables(n, q, w, ctx=ctx), oracle.Tables(n, q, w)
    rng = np.random.default_rng(9)
    x = rand_mod(rng, n * batch, q)
    a, b = ctx.from_host(x), ctx.vector(x.size)
    for _ in range(4):
        ctx.forward_transform_batch(a, b, lib, batch)
        ctx.inverse_transform_batch(b, b, lib, batch)
    assert np.array_equal(b.to_host(), x)
    assert ctx.lazy_forwards == 3
    ctx.forward_transform_batch(a, b, lib, batch)       # held, then read
    assert np.array_equal(b.to_host(), oracle.forward_batch(x, [ora]))
    a.destroy(), b.destroy(), lib.destroy()


def test_exposed_stream_is_never_deferred(own_ctx):
    ctx = own_ctx
    batch = 2
    basis = Basis(ctx, 13, 2)
    rng = np.random.default_rng(10)
    x = basis.random(rng, batch)
    a, b = ctx.from_host(x), ctx.vector(x.size)
    trained(ctx, basis, a, b, batch)
    assert ctx.stream is not None
    for _ in range(3):
        before = ctx.launch_count_noflush
        ctx.forward_transform_rns(a, b, basis.lib, batch)
        assert ctx.launch_count_noflush > before
        ctx.inverse_transform_rns(b, b, basis.lib, batch)
    assert ctx.lazy_forwards == 0
    assert np.array_equal(b.to_host(), x)
    a.destroy(), b.destroy(), basis.destroy()


@pytest.mark.parametrize("q", [params.Q61, params.Q62_LAZY_MAX])
def test_exact_quotient_family_pairs(own_ctx, q):
    """q above 2^64/6: the exact-quotient butterflies store [0,2q) lazily"""
    ctx = own_ctx
    n, batch = 1 << 13, 4
    w = params.find_psi(n, q)
    lib = vk.NttTables(n, q, w, ctx=ctx)
    rng = np.random.default_rng(11)
    x = rand_mod(rng, n * batch, q)
    a, b = ctx.from_host(x), ctx.vector(x.size)
    for _ in range(3):
        ctx.forward_transform_batch(a, b, lib, batch)
        ctx.inverse_transform_batch(b, b, lib, batch)
    assert np.array_equal(b.to_host(), x)
    assert ctx.lazy_forwards == 2
    a.destroy(), b.destroy(), lib.destroy()


def test_switch_turns_it_off():
    """$VKHEL_LAZY_FORWARD=0 (read once per process): nothing is held"""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
import vkhel_b200 as vk
from vkhel_b200 import params
n, limbs, batch = 1 << 13, 3, 2
ctx = vk.Context(0)
qs = params.ntt_primes(limbs)
tabs = [vk.NttTables(n, q, params.find_psi(n, q)) for q in qs]
rng = np.random.default_rng(3)
x = np.concatenate([rng.integers(0, q, n, dtype=np.uint64)
                    for _ in range(batch) for q in qs])
a, b = ctx.from_host(x), ctx.vector(x.size)
for _ in range(4):
    before = ctx.launch_count_noflush
    ctx.forward_transform_rns(a, b, tabs, batch)
    assert ctx.launch_count_noflush > before
    ctx.inverse_transform_rns(b, b, tabs, batch)
assert ctx.lazy_forwards == 0
assert np.array_equal(b.to_host(), x)
print("off-ok")
""" % ROOT
    env = dict(os.environ, VKHEL_LAZY_FORWARD="0")
    res = subprocess.run([sys.executable, "-c", code], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "off-ok" in res.stdout, res.stdout + res.stderr
