"""CPU suite, part 3: the N > 1 path with two gloo ranks.

The data path has no collective; what exists at N > 1 is the partition
(vkhel_b200/shard.py), the optional result gather and the max-over-ranks
timing reduction.  Two processes, gloo backend, 127.0.0.1; each rank runs the
transform of its limb shard with the CPU oracle standing in for the device and
the gathered result must equal the unsharded transform.
"""
import os
import socket
import subprocess
import sys

import numpy as np

from vkhel_b200 import shard
from conftest import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["VKHEL_ROOT"])
import numpy as np
import torch.distributed as dist
import oracle
from vkhel_b200 import params, shard

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
n, limbs, batch = 256, 5, 3
primes = params.ntt_primes(limbs, bits=50, two_adicity=12)
tables = [oracle.Tables(n, q, params.find_psi(n, q)) for q in primes]
rng = np.random.default_rng(1234)                      # same data on every rank
full = np.stack([[rng.integers(0, primes[l], n, dtype=np.uint64) for l in range(limbs)]
                 for b in range(batch)])               # [batch][limbs][n]
lo, hi = shard.limb_shard(limbs, world, rank)
mine = np.ascontiguousarray(full[:, lo:hi, :]).reshape(-1)
local = oracle.forward_batch(mine, tables[lo:hi], threads=1)
gathered = shard.gather_limb_sharded(local, limbs, n, batch, dist)
want = oracle.forward_batch(full.reshape(-1), tables, threads=1)
assert np.array_equal(gathered, want), "rank %d: gathered transform differs" % rank
slowest = shard.max_over_ranks(10.0 + rank, dist)
assert slowest == 10.0 + world - 1
# batch sharding covers every polynomial exactly once
owned = [shard.batch_shard(7, world, r) for r in range(world)]
assert owned[0][0] == 0 and owned[-1][1] == 7
assert all(owned[i][1] == owned[i + 1][0] for i in range(world - 1))
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
"""


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_split_range_properties():
    for total in (0, 1, 4, 7, 32, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.split_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert [shard.limb_shard(32, 8, r) for r in (0, 7)] == [(0, 4), (28, 32)]


def test_two_rank_gloo_limb_shard_and_gather():
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", PORT=str(port),
                   VKHEL_ROOT=ROOT, MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env,
                                      stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert "rank %d ok" % rank in out


def test_e2e_plan_covers_every_polynomial_once_and_follows_the_rates():
    """bench.py's end-to-end sharding (shard.e2e_plan): whatever the measured
    rates, every (limb, batch entry) is moved by exactly one rank; a pair's
    batch split follows its two rates; equal rates give equal work."""
    from vkhel_b200 import shard
    limbs, batch = 32, 16
    cases = {
        1: [[48.0]],
        2: [[42.5, 31.1], [40.0, 40.0], [0.0, 0.0]],
        4: [[17.5, 17.5, 23.5, 23.5], [20.0] * 4],
        8: [[8.0] * 4 + [11.3] * 4, [11.3, 8.0] * 4, [10.0] * 8,
            [1.0, 100.0] + [10.0] * 6],
    }
    for world, rate_sets in cases.items():
        for rates in rate_sets:
            plan = shard.e2e_plan(rates, limbs, batch)
            seen = {}
            for r, (ls, b0, b1) in enumerate(plan):
                assert 0 <= b0 < b1 <= batch and ls == sorted(set(ls))
                for l in ls:
                    for b in range(b0, b1):
                        assert (l, b) not in seen, (world, rates, l, b)
                        seen[(l, b)] = r
            assert len(seen) == limbs * batch
            work = [len(ls) * (b1 - b0) for ls, b0, b1 in plan]
            if len(set(rates)) == 1:
                assert set(work) == {limbs * batch // world}
            elif world > 1:
                # the faster rank of the extreme pair moves more
                slow = min(range(world), key=lambda r: (rates[r], r))
                fast = max(range(world), key=lambda r: (rates[r], -r))
                assert work[fast] >= work[slow]
    # 8.0 against 11.3 GB/s: 7 and 9 batch entries of the pair's 8 limbs
    plan = shard.e2e_plan([8.0] * 4 + [11.3] * 4, limbs, batch)
    assert [b1 - b0 for _, b0, b1 in plan] == [7] * 4 + [9] * 4
    assert all(len(ls) == 8 for ls, _, _ in plan)
    assert shard.chunk_sizes(7, 4) == [2, 2, 2, 1]
    assert shard.chunk_sizes(9, 4) == [3, 2, 2, 2]
    assert shard.chunk_sizes(16, 4) == [4, 4, 4, 4]
    assert shard.chunk_sizes(2, 4) == [1, 1]
