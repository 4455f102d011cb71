#!/usr/bin/env python3
"""Static SASS census of the hot kernels (no GPU needed): opcode counts per
kernel from `cuobjdump -sass build/obj/kernels_ntt.o`, grouped the way the
integer roofline counts them -- IMAD.WIDE / IMAD.HI / IMAD on the fmaheavy
pipe (2.18 / 2.35 / 1 IMAD slots each as measured by probe.cu), IADD3 / SEL /
LOP3 on the ALU pipe -- plus the memory and TMA (UBLKCP) instructions.  The
counts are for the whole kernel text -- prologue, tails, and for the inverse
kernels TWO copies of the rounds (the pass with and the pass without the n^-1
handling are one kernel with a uniform branch; one copy runs) -- so they bound
the per-butterfly figures from above; the dynamic figures are in the ncu
summaries.  Forward kernels: 4 IMAD.WIDE + 1 IMAD.HI + 4 IMAD per butterfly,
as written in modarith.cuh.

    python tools/sass_census.py > profiles/r01_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = [
    ("forward columns  <0,8,4,2,apx>", "ntt_cols_kernelILb0ELi8ELi4ELi2ELb1ELb0ELb0E", 64),
    ("forward rows     <0,8,1,apx>", "ntt_rows_kernelILb0ELi8ELi1ELb0ELb1ELb0ELb0ELb0E", 32),
    ("forward rows     <0,8,1,apx,lazy> (held forward transform)", "ntt_rows_kernelILb0ELi8ELi1ELb0ELb1ELb0ELb0ELb1E", 32),
    ("inverse rows     <1,8,1,apx>", "ntt_rows_kernelILb1ELi8ELi1ELb0ELb1ELb0ELb0ELb0E", 32),
    ("inverse columns  <1,8,4,2,apx,top>", "ntt_cols_kernelILb1ELi8ELi4ELi2ELb1ELb0ELb1E", 64),
]
FMA = ("IMAD.WIDE", "IMAD.HI", "IMAD", "IMAD.X", "IMAD.MOV", "IMAD.IADD",
       "IMAD.SHL")
# IMAD slots per instruction, from vkhel_ctx_probe_int_peaks on the B200
# (profiles/r02_probe_int_peaks.json): IMAD 63.2, IMAD.WIDE 29.0, IMAD.HI 26.9
# thread-instructions per clock per SM
SLOTS = {"IMAD.WIDE": 2.18, "IMAD.HI": 2.35}


def group(op):
    base = op.split(".")[0]
    if base == "IMAD":
        for mod in ("WIDE", "HI", "MOV", "IADD", "SHL", "X"):
            if "." + mod in op:
                return "IMAD." + mod
        return "IMAD"
    if base in ("LDS", "STS", "LDG", "STG"):
        return ".".join(op.split(".")[:1] + [p for p in op.split(".")[1:]
                                             if p in ("64", "128")])
    return base


def main():
    obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        ROOT, "build", "obj", "kernels_ntt.o")
    text = subprocess.check_output(["cuobjdump", "-sass", obj], text=True)
    parts = re.split(r"\n\s*Function : ", text)
    for label, key, bfly in KERNELS:
        body = next((p for p in parts[1:] if key in p.split("\n", 1)[0]), None)
        if body is None:
            print("%s: not found" % label)
            continue
        ops = collections.Counter()
        for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)",
                             body):
            ops[group(m.group(1))] += 1
        total = sum(ops.values())
        slots = sum(ops[k] * SLOTS.get(k, 1.0) for k in FMA)
        print("== %s: %d instructions in the kernel text, %d butterflies per "
              "thread per tile" % (label, total, bfly))
        print("   fmaheavy pipe: " + ", ".join("%s %d" % (k, ops[k]) for k in FMA
                                               if ops[k])
              + "  (%.0f IMAD slots = %.1f per butterfly, kernel text)"
              % (slots, slots / bfly))
        alu = ("IADD3", "SEL", "LOP3", "LEA", "SHF", "ISETP", "PLOP3", "MOV")
        print("   ALU pipe:      " + ", ".join("%s %d" % (k, ops[k]) for k in alu
                                               if ops[k]))
        mem = sorted(k for k in ops if k.split(".")[0] in
                     ("LDS", "STS", "LDG", "STG", "UBLKCP", "SYNCS", "BAR"))
        print("   memory / sync: " + ", ".join("%s %d" % (k, ops[k]) for k in mem))
    return 0


if __name__ == "__main__":
    sys.exit(main())
