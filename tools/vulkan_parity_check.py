#!/usr/bin/env python3
"""Compare a dump made by tools/vulkan_dump.c (a vkhel library driven through
the reference's public API on seeded inputs) with the CPU oracle.

    python tools/vulkan_parity_check.py dump.txt          # exit 0 = identical
    python tools/vulkan_parity_check.py --self-test       # oracle vs itself

Every line `<op> n=.. q=.. w=.. seed=.. [k=v] : values` is recomputed: the
inputs from the seed (the dump program's xorshift64), the expected output with
oracle/ (canonical mode).  For elemfma the literal shader arithmetic is also
evaluated: where the real shaders follow the defect documented in SURVEY
App. B (Q2) the line is reported as `defect`, not as a mismatch of the oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

MASK = (1 << 64) - 1


def fill(count, seed, mod):
    s = seed | 1
    out = np.empty(count, np.uint64)
    for i in range(count):
        s ^= (s << 13) & MASK
        s ^= s >> 7
        s ^= (s << 17) & MASK
        out[i] = s % mod if mod else s
    return out


def expected(op, kv):
    n, q, w, seed = kv["n"], kv["q"], kv["w"], kv["seed"]
    if op in ("forward", "inverse"):
        t = oracle.Tables(n, q, w)
        a = fill(n, seed, q)
        return (oracle.forward(a, t) if op == "forward" else oracle.inverse(a, t)), None
    if op == "elemmul":
        return oracle.elemmul(fill(n, seed, 0), fill(n, seed + 7777, 0), q), None
    if op == "elemfma":
        a, b = fill(n, seed, q), fill(n, seed + 7777, q)
        return (oracle.elemfma(a, b, kv["mult"], q),
                oracle.elemfma(a, b, kv["mult"], q, literal=True))
    a = fill(n, seed, q)
    if op == "elemmod":
        return oracle.elemmod(a, kv["mod"], q), None
    if op == "elemgtadd":
        return oracle.elemgtadd(a, kv["bound"], kv["diff"]), None
    if op == "elemgtsub":
        return oracle.elemgtsub(a, kv["bound"], kv["diff"], kv["mod"]), None
    raise ValueError("unknown op " + op)


def check_lines(lines):
    bad = defects = cases = 0
    for line in lines:
        if " : " not in line and not line.rstrip().endswith(":"):
            continue                      # banners of the library
        head, _, tail = line.partition(":")
        parts = head.split()
        op = parts[0]
        kv = {k: int(v) for k, v in (p.split("=") for p in parts[1:])}
        got = np.array([int(v) for v in tail.split()], dtype=np.uint64)
        want, literal = expected(op, kv)
        cases += 1
        if np.array_equal(got, want):
            continue
        if literal is not None and np.array_equal(got, literal):
            defects += 1
            print("defect  %s (matches the literal shader arithmetic, SURVEY "
                  "App. B Q2; %d of %d elements differ from the contract)"
                  % (head.strip(), int((got != want).sum()), got.size))
            continue
        bad += 1
        print("MISMATCH %s: %d of %d elements differ"
              % (head.strip(), int((got != want).sum()), got.size))
    print("%d cases, %d mismatches, %d known-defect lines" % (cases, bad, defects))
    return bad


def self_test():
    """dump lines produced by the oracle itself must pass, and a corrupted one
    must be caught"""
    lines = []
    q = 1152921504606584833
    from vkhel_b200 import params
    n = 64
    w = params.find_psi(n, q)
    for op, extra in [("forward", {}), ("inverse", {}), ("elemmul", {"w": 0}),
                      ("elemfma", {"w": 0, "mult": 12345}),
                      ("elemmod", {"w": 0, "mod": 2}),
                      ("elemgtadd", {"w": 0, "bound": q // 3, "diff": 5}),
                      ("elemgtsub", {"w": 0, "bound": q // 3, "diff": 9,
                                     "mod": 1000003})]:
        kv = {"n": n, "q": q, "w": w, "seed": 77}
        kv.update(extra)
        want, _ = expected(op, kv)
        head = op + " " + " ".join("%s=%d" % i for i in kv.items())
        lines.append(head + " : " + " ".join(str(int(v)) for v in want))
    assert check_lines(lines) == 0
    broken = lines[0].rsplit(" ", 1)[0] + " 1"
    assert check_lines([broken]) == 1
    print("self-test ok")
    return 0


def main():
    if "--self-test" in sys.argv:
        return self_test()
    with open(sys.argv[1]) as f:
        return 1 if check_lines(f.read().splitlines()) else 0


if __name__ == "__main__":
    sys.exit(main())
