#!/usr/bin/env python3
"""Extract the known-answer vectors of the reference's own tests into JSON.

Reads /root/reference/test/vector.c, test/ntt.c and examples/example.c (data
only: array literals and the scalar arguments of the calls that use them) and
writes tests/golden/reference_kats.json.  Run in the build container, where
/root/reference exists; the JSON is committed so the GPU box needs neither.

    python tests/golden/extract_kats.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                   "reference_kats.json")


def functions(src):
    """name -> body for every `void test_xxx() { ... }` at top level"""
    out = {}
    for m in re.finditer(r"^void (test_\w+)\(\) \{\n(.*?)^\}", src,
                         re.S | re.M):
        out[m.group(1)] = m.group(2)
    return out


def arrays(body):
    """every `uint64_t name[...] = { ... };` in order of appearance"""
    found = []
    for m in re.finditer(r"uint64_t (\w+)\[\w*\] = \{(.*?)\};", body, re.S):
        nums = [int(x) for x in re.findall(r"\d+", m.group(2))]
        found.append((m.group(1), nums))
    return found


def scalars(body, name):
    return [int(x) for x in
            re.findall(r"uint64_t %s = (\d+);" % name, body)]


def main():
    vec = open(os.path.join(REF, "test/vector.c")).read()
    fns = functions(vec)
    kats = {"source": "reference test/vector.c, test/ntt.c, "
                      "examples/example.c (data only)"}

    # element-wise: arrays appear as a.., b.., then expected per case
    b = fns["test_elemfma"]
    arr = arrays(b)
    kats["elemfma"] = {
        "mod": scalars(b, "modulus")[0], "a": arr[0][1], "b": arr[1][1],
        "cases": [{"multiplier": m, "expected": e[1]}
                  for m, e in zip(scalars(b, "multiplier"), arr[2:])],
    }

    b = fns["test_elemmul"]
    arr = arrays(b)
    mods = scalars(b, "modulus")
    kats["elemmul"] = [
        {"mod": mods[i], "a": arr[3 * i][1], "b": arr[3 * i + 1][1],
         "expected": arr[3 * i + 2][1]} for i in range(len(mods))]

    b = fns["test_elemgtadd"]
    arr = arrays(b)
    calls = re.findall(r"vkhel_vector_elemgtadd\(a, \w+, (\d+), (\d+)\)", b)
    kats["elemgtadd"] = {
        "a": arr[0][1],
        "cases": [{"bound": int(bd), "diff": int(df), "expected": e[1]}
                  for (bd, df), e in zip(calls, arr[1:])]}

    b = fns["test_elemgtsub"]
    arr = arrays(b)
    calls = re.findall(
        r"vkhel_vector_elemgtsub\(a, \w+, (\d+), (\d+), (\d+)\)", b)
    # the third call re-uses d_expected
    expected = [arr[1][1], arr[2][1], arr[2][1]]
    kats["elemgtsub"] = {
        "a": arr[0][1],
        "cases": [{"bound": int(bd), "diff": int(df), "mod": int(md),
                   "expected": e}
                  for (bd, df, md), e in zip(calls, expected)]}

    b = fns["test_elemmod"]
    arr = arrays(b)
    calls = re.findall(r"vkhel_vector_elemmod\(a, \w+, (\d+), (\w+)\)", b)
    groups = [(arr[0][1], [arr[1][1], arr[2][1]]),
              (arr[3][1], [arr[4][1], arr[5][1]])]
    cases = []
    ci = 0
    for a_vals, exps in groups:
        for e in exps:
            mod, q = calls[ci]
            ci += 1
            cases.append({"a": a_vals, "mod": int(mod),
                          "q": (1 << 64) - 1 if q == "UINT64_MAX" else int(q),
                          "expected": e})
    kats["elemmod"] = cases

    # transforms
    def ntt_case(name):
        b = fns[name]
        arr = dict(arrays(b))
        m = re.search(r"vkhel_ntt_tables_create\(\s*(\w+),[^\d]*(\d+),"
                      r"[^\d]*(\d+)", b, re.S)
        operand = arr.get("operand", arr.get("elements"))
        return {"n": len(operand), "q": int(m.group(2)), "w": int(m.group(3)),
                "operand": operand, "expected": arr["expected"]}

    kats["forward_transform"] = [ntt_case("test_forward_transform"),
                                 ntt_case("test_forward_transform_big")]
    kats["inverse_transform"] = [ntt_case("test_inverse_transform"),
                                 ntt_case("test_inverse_transform_big")]

    # tables: test/ntt.c
    ntt = open(os.path.join(REF, "test/ntt.c")).read()
    m = re.search(r"vkhel_ntt_tables_create\((\d+), (\d+), (\d+)\)", ntt)
    roots = re.search(r"\(uint64_t\[\]\) \{([^}]*)\}", ntt)
    kats["tables"] = {"n": int(m.group(1)), "q": int(m.group(2)),
                      "w": int(m.group(3)),
                      "roots_of_unity": [int(x) for x in
                                         re.findall(r"\d+", roots.group(1))]}

    # example: elemmul mod 17
    ex = open(os.path.join(REF, "examples/example.c")).read()
    a = re.search(r"a_elements\[4\] = \{([^}]*)\}", ex)
    bb = re.search(r"b_elements\[\] = \{([^}]*)\}", ex)
    md = re.search(r"vkhel_vector_elemmul\(a, b, c, (\d+)\)", ex)
    kats["example"] = {"a": [int(x) for x in re.findall(r"\d+", a.group(1))],
                       "b": [int(x) for x in re.findall(r"\d+", bb.group(1))],
                       "mod": int(md.group(1))}

    with open(OUT, "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
