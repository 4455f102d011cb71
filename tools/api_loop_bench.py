#!/usr/bin/env python3
"""The reference API called in a loop over separate vectors (an RNS polynomial
is one vector per limb there), forward_transform on each, then one sync.
The calls are recorded and go out as one indirect batched launch
(vkhel_b200/csrc/vector.cu); VKHEL_NO_DEFER=1 shows one launch pair per call.
Host wall clock around calls + sync, since the cost is on the host side."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhel_b200 as vk  # noqa: E402
from vkhel_b200 import params  # noqa: E402


def main():
    ctx = vk.Context(0)
    q = params.P0
    for log2n, count in ((10, 1024), (12, 1024), (14, 256)):
        n = 1 << log2n
        t1 = vk.NttTables(n, q, params.find_psi(n, q))
        vs = [ctx.vector(n, zero=False) for _ in range(count)]
        for _ in range(2):
            for v in vs:
                ctx.forward_transform(v, v, t1)
            ctx.sync()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            for v in vs:
                ctx.forward_transform(v, v, t1)
            ctx.sync()
        us = (time.perf_counter() - t0) / (reps * count) * 1e6
        print(json.dumps({"config": "reference API loop over %d vectors, "
                          "forward_transform each, then sync" % count,
                          "log2n": log2n, "us_per_transform": us,
                          "deferred": os.environ.get("VKHEL_NO_DEFER") is None,
                          "deferred_stats": ctx.deferred_stats}))
        for v in vs:
            v.destroy()
        t1.destroy()
    ctx.destroy()


if __name__ == "__main__":
    main()
