#!/usr/bin/env python3
"""Forward / inverse time of one degree at 2^27 coefficients for the stage
split given in $VKHEL_KROW (tuning of plan_fast, kernels_ntt.cu)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhel_b200 as vk  # noqa: E402
from vkhel_b200 import params  # noqa: E402
from sweep import time_ms  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [12, 13, 14, 15]
    ctx = vk.Context(0)
    timer = ctx.timer()
    q = params.P0
    total = 1 << 27
    rng = np.random.default_rng(0)
    host = rng.integers(0, 1 << 62, total, dtype=np.uint64) % np.uint64(q)
    a = ctx.from_host(host)
    b = ctx.vector(total, zero=False)
    for log2n in sizes:
        n, batch = 1 << log2n, total >> log2n
        t = vk.NttTables(n, q, params.find_psi(n, q))
        f = time_ms(ctx, timer, lambda: ctx.forward_transform_batch(a, b, t, batch), 10)
        i = time_ms(ctx, timer, lambda: ctx.inverse_transform_batch(b, b, t, batch), 10)
        ctx.forward_transform_batch(a, b, t, batch)
        ctx.inverse_transform_batch(b, b, t, batch)
        ok = bool(np.array_equal(b.to_host(), host))
        print(json.dumps({"krow": os.environ.get("VKHEL_KROW", "default"),
                          "log2n": log2n, "fwd_ms": round(f, 4),
                          "inv_ms": round(i, 4), "round_trip_exact": ok}))
        t.destroy()
    ctx.destroy()


if __name__ == "__main__":
    main()
