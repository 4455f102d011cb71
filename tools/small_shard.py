#!/usr/bin/env python3
"""The per-GPU share of BASELINE configs[2] under limb sharding, on ONE GPU:
n = 2^16, batch 16, `limbs` = 32 / 16 / 8 / 4 limbs (the shard of a 1-, 2-, 4-,
8-GPU run): step time of forward + inverse and the efficiency against the
32-limb run, i.e. what strong scaling can reach at best on N GPUs.

    python tools/small_shard.py            # one JSON line per shard size
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhel_b200 as vk  # noqa: E402
from vkhel_b200 import params  # noqa: E402

N, BATCH = 1 << 16, 16


def main():
    ctx = vk.Context(0)
    timer = ctx.timer()
    primes = params.ntt_primes(32)
    tables = [vk.NttTables(N, q, params.find_psi(N, q), ctx=ctx) for q in primes]
    rng = np.random.default_rng(5)
    base = None
    for limbs in (32, 16, 8, 4):
        tabs = tables[:limbs]
        host = np.concatenate([rng.integers(0, primes[p % limbs], N, dtype=np.uint64)
                               for p in range(limbs * BATCH)])
        a = ctx.from_host(host)
        b = ctx.vector(host.size, zero=False)

        def step():
            ctx.forward_transform_rns(a, b, tabs, BATCH)
            ctx.inverse_transform_rns(b, b, tabs, BATCH)

        for _ in range(20):
            step()
        ctx.sync()
        iters = 300
        timer.start()
        for _ in range(iters):
            step()
        timer.stop()
        us = timer.elapsed_ms() / iters * 1e3
        ok = bool(np.array_equal(b.to_host(), host))
        per_ntt = us / (2 * limbs * BATCH)
        base = base or per_ntt
        print(json.dumps({"limbs": limbs, "gpus_this_models": 32 // limbs,
                          "step_us": us, "ntt_per_s": 1e6 / per_ntt,
                          "efficiency_vs_32_limbs": base / per_ntt,
                          "round_trip_exact": ok,
                          "slice_mib": os.environ.get("VKHEL_SLICE_MIB")}),
              flush=True)
        a.destroy(), b.destroy()
    timer.destroy()
    for t in tables:
        t.destroy()
    ctx.destroy()


if __name__ == "__main__":
    main()
