#!/usr/bin/env python3
"""Per-pass timing of library variants on the bench workload (n = 2^16, 32
limbs x batch 16, resident): each pass of the forward and the inverse transform
launched alone ($VKHEL_ONLY_PASS, results invalid, unsliced) and the complete
transforms as shipped.  One line per library.

    python tools/kernel_ab.py vkhel_b200/lib/libvkhel.so build/variants/*.so
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, json
sys.path.insert(0, %r)
import numpy as np
import vkhel_b200 as vk
from vkhel_b200 import params
N, LIMBS, BATCH = 1 << 16, int(os.environ.get("AB_LIMBS", "32")), 16
primes = params.ntt_primes(LIMBS)
ctx = vk.Context(0)
tabs = [vk.NttTables(N, q, params.find_psi(N, q), ctx=ctx) for q in primes]
rng = np.random.default_rng(1)
host = np.concatenate([rng.integers(0, primes[p %% LIMBS], N, dtype=np.uint64)
                       for p in range(LIMBS)] * BATCH)
a = ctx.from_host(host)
b = ctx.vector(host.size, zero=False)
timer = ctx.timer()
def t(fn, iters=int(os.environ.get("AB_ITERS", "100"))):
    for _ in range(10):
        fn()
    ctx.sync()
    timer.start()
    for _ in range(iters):
        fn()
    timer.stop()
    return timer.elapsed_ms() / iters * 1e3
fwd = t(lambda: ctx.forward_transform_rns(a, b, tabs, BATCH))
inv = t(lambda: ctx.inverse_transform_rns(b, b, tabs, BATCH))
ok = None
if not os.environ.get("VKHEL_ONLY_PASS"):
    ctx.forward_transform_rns(a, b, tabs, BATCH)
    ctx.inverse_transform_rns(b, b, tabs, BATCH)
    ok = bool(np.array_equal(b.to_host(), host))
print(json.dumps({"fwd_us": fwd, "inv_us": inv, "ok": ok}))
""" % ROOT


def run(lib, only):
    env = dict(os.environ, VKHEL_LIB_PATH=lib)
    if only:
        env["VKHEL_ONLY_PASS"] = only
        env["VKHEL_SLICE_MIB"] = "0"
    out = subprocess.run([sys.executable, "-c", CHILD], env=env,
                         capture_output=True, text=True)
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            return json.loads(line)
    return {"fwd_us": float("nan"), "inv_us": float("nan"), "ok": False,
            "err": out.stderr[-300:]}


def main():
    for lib in sys.argv[1:]:
        cols, rows, full = run(lib, "cols"), run(lib, "rows"), run(lib, None)
        step = full["fwd_us"] + full["inv_us"]
        print("%-28s fwd cols %6.1f rows %6.1f | inv rows %6.1f cols %6.1f | "
              "full fwd %6.1f inv %6.1f = %6.1f us -> %.3f M NTT/s %s"
              % (os.path.basename(lib), cols["fwd_us"], rows["fwd_us"],
                 rows["inv_us"], cols["inv_us"], full["fwd_us"],
                 full["inv_us"], step, 1024 / step,
                 "ok" if full["ok"] else "MISMATCH " + full.get("err", "")),
              flush=True)


if __name__ == "__main__":
    main()
