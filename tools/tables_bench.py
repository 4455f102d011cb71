#!/usr/bin/env python3
"""Twiddle-table generation: vkhel_ntt_tables_create (host; contract of the
reference's src/ntt_tables.c:17-44) against vkhel_ntt_tables_create_on (GPU),
32 RNS limbs per size.  Prints one JSON object per size."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhel_b200 as vk  # noqa: E402
from vkhel_b200 import params  # noqa: E402


def main():
    ctx = vk.Context(0)
    primes = params.ntt_primes(32, two_adicity=18)
    for log2n in (12, 14, 16, 17):
        n = 1 << log2n
        psis = [params.find_psi(n, q) for q in primes]
        vk.NttTables(n, primes[0], psis[0], ctx=ctx).destroy()   # warm up
        t0 = time.perf_counter()
        host = [vk.NttTables(n, q, w) for q, w in zip(primes, psis)]
        t_host = time.perf_counter() - t0
        t0 = time.perf_counter()
        dev = [vk.NttTables(n, q, w, ctx=ctx) for q, w in zip(primes, psis)]
        t_dev = time.perf_counter() - t0
        same = all((h.roots_barrett_factors == d.roots_barrett_factors).all()
                   and (h.inv_roots_of_unity == d.inv_roots_of_unity).all()
                   for h, d in zip(host, dev))
        print(json.dumps({"config": "32 tables", "log2n": log2n,
                          "host_ms_per_table": t_host / 32 * 1e3,
                          "device_ms_per_table": t_dev / 32 * 1e3,
                          "identical": bool(same)}))
        for t in host + dev:
            t.destroy()
    ctx.destroy()


if __name__ == "__main__":
    main()
