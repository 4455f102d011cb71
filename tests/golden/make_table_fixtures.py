#!/usr/bin/env python3
"""Golden twiddle tables produced by the REFERENCE's own host code.

Loads oracle/_ref/libvkhel_refhost.so (the reference's src/numbers.c +
src/ntt_tables.c compiled from /root/reference by oracle/Makefile), runs its
vkhel_ntt_tables_create and nt_* functions for a list of parameter sets, and
writes tests/golden/reference_tables.json: a SHA-256 per table array plus the
leading entries, and scalar known answers.  Run in the build container only.

    make oracle && python tests/golden/make_table_fixtures.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from vkhel_b200 import params  # noqa: E402

CASES = [
    (4, 113, 18),
    (16, params.Q_KAT_52, params.W_KAT_52_N16),
    (8, 769, None),
    (256, 1125891450734593, None),
    (4096, params.Q61, None),
    (1 << 14, params.P0, None),
    (1 << 16, params.P0, None),
    (1 << 16, params.ntt_primes(32)[31], None),
    (1 << 17, params.P0, None),
]


def main():
    ref = oracle.reference_host()
    if ref is None:
        sys.exit("oracle/_ref is not built (needs /root/reference)")
    out = {"source": "reference src/ntt_tables.c + src/numbers.c via "
                     "oracle/_ref/libvkhel_refhost.so", "tables": [],
           "scalars": []}
    for n, q, w in CASES:
        if w is None:
            w = params.find_psi(n, q)
        arrays = oracle.reference_tables(ref, n, q, w)
        names = ("roots", "inv_roots", "roots_shoup", "inv_roots_shoup")
        entry = {"n": n, "q": q, "w": w}
        for name, arr in zip(names, arrays):
            entry[name + "_sha256"] = hashlib.sha256(arr.tobytes()).hexdigest()
            entry[name + "_head"] = [int(x) for x in arr[:8]]
        entry["inv_n"] = int(ref.nt_inverse_mod(n % q, q))
        out["tables"].append(entry)
    for q in (769, 1125891450734593, params.Q61, params.P0):
        for a, b in ((q - 1, q - 1), (q // 2, q // 3), (12345 % q, 67890 % q)):
            out["scalars"].append({
                "q": q, "a": a, "b": b,
                "multiply_mod": int(ref.nt_multiply_mod(a, b, q, 0)),
                "power_mod": int(ref.nt_power_mod(a, b, q)),
                "inverse_mod": int(ref.nt_inverse_mod(a, q)),
                "shoup_factor": int(ref.nt_compute_barrett_factor(a, q, 64)),
            })
    path = os.path.join(HERE, "reference_tables.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
