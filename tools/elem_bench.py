#!/usr/bin/env python3
"""Element-wise kernels at 2^27 elements (HBM-bound): GB/s per operation
(algorithmic bytes: 24 per element for the two-input kernels, 16 otherwise)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhel_b200 as vk  # noqa: E402
from vkhel_b200 import params  # noqa: E402
from sweep import time_ms  # noqa: E402


def main():
    ctx = vk.Context(0)
    timer = ctx.timer()
    q = params.P0
    total = 1 << 27
    rng = np.random.default_rng(0)
    host = rng.integers(0, 1 << 62, total, dtype=np.uint64) % np.uint64(q)
    a = ctx.from_host(host)
    b = ctx.from_host(host[::-1].copy())
    c = ctx.vector(total, zero=False)
    res = {"lib": os.path.basename(vk.LIB_PATH), "elements": total}
    for name, fn, nbytes in (
            ("elemmul", lambda: ctx.elemmul(a, b, c, q), 24),
            ("elemfma", lambda: ctx.elemfma(a, b, c, 12345, q), 24),
            ("elemgtsub", lambda: ctx.elemgtsub(a, c, q // 2, 7, q), 16),
            ("elemgtadd", lambda: ctx.elemgtadd(a, c, q // 2, 7), 16),
            ("elemmod2", lambda: ctx.elemmod(a, c, 2, q), 16)):
        ms = time_ms(ctx, timer, fn, 20, warmup=3)
        res[name] = round(nbytes * total / ms / 1e6)
    print(json.dumps(res))
    ctx.destroy()


if __name__ == "__main__":
    main()
