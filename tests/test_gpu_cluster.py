"""GPU suite: the single-pass transform on a thread-block cluster
(vkhel_b200/csrc/kernels_ntt_cluster.cu, $VKHEL_CLUSTER=1) against the CPU
oracle: n = 2^14, 2^15, 2^16 (clusters of 2, 4, 8 CTAs), forward and inverse,
in place and out of place, ragged batches, RNS bases and limb slices.  The
switch is read once per process, hence the child process."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CHILD = r"""
import sys
sys.path.insert(0, %r)
import numpy as np
import oracle
import vkhel_b200 as vk
from vkhel_b200 import params

import os
SLICED = bool(os.environ.get("VKHEL_SLICE_MIB"))
ctx = vk.Context(0)
rng = np.random.default_rng(2026)
l0 = ctx.launch_count
for log2n, limbs, batch in [(14, 1, 1), (14, 3, 5), (15, 1, 7), (15, 2, 3),
                            (16, 1, 2), (16, 5, 3), (16, 2, 19)]:
    n = 1 << log2n
    primes = params.ntt_primes(limbs) if limbs > 1 else [params.Q61 if log2n == 15 else params.P0]
    tabs = [vk.NttTables(n, q, params.find_psi(n, q)) for q in primes]
    oras = [oracle.Tables(n, q, t.w) for q, t in zip(primes, tabs)]
    polys = limbs * batch
    x = np.concatenate([rng.integers(0, primes[p %% limbs], n, dtype=np.uint64)
                        for p in range(polys)])
    want = oracle.forward_batch(x, oras, threads=8)
    a, b = ctx.from_host(x), ctx.vector(x.size + 3)
    before = ctx.launch_count
    ctx.forward_transform_rns(a, b, tabs, batch)          # out of place
    # one launch per transform -- two when the RNS batch is cut in two limb
    # slices for the two streams (8 MiB and more, kernels_ntt.cu)
    launched = ctx.launch_count - before
    assert launched == 1 or SLICED or (limbs > 1 and launched == 2), launched
    assert np.array_equal(b.to_host()[:x.size], want), ("forward", log2n, limbs, batch)
    assert np.array_equal(a.to_host(), x), "operand modified"
    ctx.inverse_transform_rns(b, b, tabs, batch)          # in place
    assert np.array_equal(b.to_host()[:x.size], x), ("inverse", log2n, limbs, batch)
    ctx.forward_transform_rns(a, a, tabs, batch)          # in place
    assert np.array_equal(a.to_host(), want)
    ctx.inverse_transform_rns(a, b, tabs, batch)          # out of place
    assert np.array_equal(b.to_host()[:x.size], x)
    # single-vector entry points (recorded, then launched)
    if limbs == 1:
        v = ctx.from_host(x[:n])
        ctx.forward_transform(v, v, tabs[0])
        assert np.array_equal(v.to_host(), want[:n])
        ctx.inverse_transform(v, v, tabs[0])
        assert np.array_equal(v.to_host(), x[:n])
        v.destroy()
    a.destroy(), b.destroy()
    for t in tabs:
        t.destroy()
# a modulus above 2^61.4: the exact-quotient family
n = 1 << 14
q = params.Q62_LAZY_MAX
t = vk.NttTables(n, q, params.find_psi(n, q))
o = oracle.Tables(n, q, t.w)
x = rng.integers(0, q, 2 * n, dtype=np.uint64)
a = ctx.from_host(x)
ctx.forward_transform_batch(a, a, t, 2)
assert np.array_equal(a.to_host(), oracle.forward_batch(x, [o], threads=2))
ctx.inverse_transform_batch(a, a, t, 2)
assert np.array_equal(a.to_host(), x)
print("cluster ok")
""" % ROOT


@pytest.mark.parametrize("slice_mib", [None, "1"])
def test_cluster_transform_matches_the_oracle(slice_mib):
    env = dict(os.environ, VKHEL_CLUSTER="1")
    if slice_mib:
        env["VKHEL_SLICE_MIB"] = slice_mib       # limb slices on two streams
    res = subprocess.run([sys.executable, "-c", CHILD], env=env,
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "cluster ok" in res.stdout
