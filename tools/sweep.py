#!/usr/bin/env python3
"""BASELINE configs[4]: degree sweep n = 2^10 .. 2^17 at a fixed number of
coefficients (default 2^27 = 1 GiB), single prime, forward + inverse.
Also configs[1] (n = 2^14, batch 256) and configs[3] (polymul n = 2^16).
Prints one JSON object per measurement (CUDA-event timing, data resident)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhel_b200 as vk  # noqa: E402
from vkhel_b200 import params  # noqa: E402


def time_ms(ctx, timer, fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    ctx.sync()
    timer.start()
    for _ in range(iters):
        fn()
    timer.stop()
    return timer.elapsed_ms() / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-total", type=int, default=27)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    ctx = vk.Context(0)
    timer = ctx.timer()
    q = params.P0
    rng = np.random.default_rng(0)
    total = 1 << args.log2_total
    host = (rng.integers(0, 1 << 62, total, dtype=np.uint64) % np.uint64(q))
    a = ctx.from_host(host)
    b = ctx.vector(total, zero=False)
    for log2n in range(10, 18):
        n = 1 << log2n
        batch = total >> log2n
        t = vk.NttTables(n, q, params.find_psi(n, q))
        fwd = time_ms(ctx, timer, lambda: ctx.forward_transform_batch(a, b, t, batch), args.iters)
        inv = time_ms(ctx, timer, lambda: ctx.inverse_transform_batch(b, b, t, batch), args.iters)
        ctx.forward_transform_batch(a, b, t, batch)
        ctx.inverse_transform_batch(b, b, t, batch)
        ok = bool(np.array_equal(b.to_host(), host))
        bfly = batch * (n // 2) * log2n
        print(json.dumps({
            "config": "sweep", "log2n": log2n, "batch": batch,
            "fwd_ms": fwd, "inv_ms": inv,
            "fwd_ntt_per_s": batch / fwd * 1e3, "inv_ntt_per_s": batch / inv * 1e3,
            "fwd_GBps": 16 * n * batch / fwd / 1e6, "inv_GBps": 16 * n * batch / inv / 1e6,
            "fwd_Gbfly_per_s": bfly / fwd / 1e6, "inv_Gbfly_per_s": bfly / inv / 1e6,
            "round_trip_exact": ok}))
        t.destroy()
    # configs[1]: n = 2^14, batch 256
    n, batch = 1 << 14, 256
    t = vk.NttTables(n, q, params.find_psi(n, q))
    fwd = time_ms(ctx, timer, lambda: ctx.forward_transform_batch(a, b, t, batch), 50)
    inv = time_ms(ctx, timer, lambda: ctx.inverse_transform_batch(b, b, t, batch), 50)
    print(json.dumps({"config": "n=2^14 batch 256 (32 MiB, L2-resident)",
                      "fwd_ms": fwd, "inv_ms": inv,
                      "fwd_ntt_per_s": batch / fwd * 1e3, "inv_ntt_per_s": batch / inv * 1e3}))
    t.destroy()
    # configs[3] per-GPU share: polymul n = 2^16, batch 128, single prime
    n, batch = 1 << 16, 128
    t = vk.NttTables(n, q, params.find_psi(n, q))
    c = ctx.vector(n * batch, zero=False)
    ms = time_ms(ctx, timer, lambda: ctx.polymul_rns(a, b, c, [t], batch), 10)
    print(json.dumps({"config": "polymul n=2^16 batch 128 (per-GPU share of configs[3])",
                      "ms": ms, "polymul_per_s": batch / ms * 1e3,
                      "api_bytes_GBps": 72 * n * batch / ms / 1e6}))
    # element-wise kernels (HBM-bound): 2^27 elements, 24 or 16 bytes each
    cfull = ctx.vector(total, zero=False)
    for name, fn, nbytes in (
            ("elemmul", lambda: ctx.elemmul(a, b, cfull, q), 24),
            ("elemfma", lambda: ctx.elemfma(a, b, cfull, 12345, q), 24),
            ("elemgtsub", lambda: ctx.elemgtsub(a, cfull, q // 2, 7, q), 16),
            ("elemgtadd", lambda: ctx.elemgtadd(a, cfull, q // 2, 7), 16)):
        ms = time_ms(ctx, timer, fn, 10)
        print(json.dumps({"config": name + " 2^%d elements" % args.log2_total,
                          "ms": ms, "GBps": nbytes * total / ms / 1e6}))
    cfull.destroy()
    # legacy single-vector API, one transform per call (launch-bound regime)
    for log2n in (12, 16):
        n = 1 << log2n
        t1 = vk.NttTables(n, q, params.find_psi(n, q))
        v = ctx.vector(n, zero=False)
        ms = time_ms(ctx, timer, lambda: ctx.forward_transform(v, v, t1), 200, warmup=10)
        print(json.dumps({"config": "legacy API, one forward_transform per call",
                          "log2n": log2n, "us_per_call": ms * 1e3}))
        v.destroy()
        t1.destroy()
    ctx.destroy()


if __name__ == "__main__":
    main()
