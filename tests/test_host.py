"""CPU suite, part 2: the library's host side and its C-ABI surface.

No compute calls (there is no GPU here): the shared library loads, exports
every symbol the headers declare and nothing else, the host table generator
agrees with the oracle and with golden tables produced by the reference's own
code, and the reference's host-only test programs (test/numbers.c,
test/ntt.c), compiled unmodified against this library, pass.
"""
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

import oracle
import vkhel_b200 as vk
from vkhel_b200 import params
from vkhel_b200.api import SIGNATURES
from conftest import ROOT


def declared_symbols():
    names = set()
    for header in ("include/vkhel/vkhel.h", "include/vkhel/vkhel_ext.h",
                   "include/priv/vector.h", "include/priv/ntt_tables.h"):
        text = open(os.path.join(ROOT, header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"\b(vkhel_\w+)\s*\(", text))
    return names


def exported_symbols():
    out = subprocess.check_output(
        ["nm", "-D", "--defined-only", vk.LIB_PATH], text=True)
    return {line.split()[-1] for line in out.splitlines()
            if line.split()[1] in "TtWw"}


def test_library_loads_and_exports_declared_api():
    lib = vk.lib()
    declared = declared_symbols()
    exported = exported_symbols()
    # the 18 reference entry points (include/vkhel/vkhel.h:8-53 there)
    reference_api = {
        "vkhel_ctx_create", "vkhel_ctx_destroy", "vkhel_ntt_tables_create",
        "vkhel_ntt_tables_destroy", "vkhel_vector_create",
        "vkhel_vector_create2", "vkhel_vector_destroy", "vkhel_vector_dup",
        "vkhel_vector_copy_from_host", "vkhel_vector_map",
        "vkhel_vector_unmap", "vkhel_vector_elemfma", "vkhel_vector_elemmod",
        "vkhel_vector_elemmul", "vkhel_vector_elemgtadd",
        "vkhel_vector_elemgtsub", "vkhel_vector_forward_transform",
        "vkhel_vector_inverse_transform"}
    assert reference_api <= declared
    assert declared <= exported, declared - exported
    # version script: nothing but vkhel_* leaves the library (vkhel.syms)
    assert all(s.startswith("vkhel_") for s in exported), \
        [s for s in exported if not s.startswith("vkhel_")]
    # the binding covers every declared symbol
    assert declared <= set(SIGNATURES), declared - set(SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_device_count_is_zero_without_gpu_or_positive():
    assert vk.device_count() >= 0


def test_tables_kat(kats):
    k = kats["tables"]
    t = vk.NttTables(k["n"], k["q"], k["w"])
    assert t.roots_of_unity.tolist() == k["roots_of_unity"]
    assert [r * i % k["q"] for r, i in zip(t.roots_of_unity.tolist(),
                                            t.inv_roots_of_unity.tolist())] \
        == [1] * k["n"]
    t.destroy()


def test_tables_match_reference_fixtures(table_fixtures):
    """library tables == tables made by the reference's compiled code"""
    for e in table_fixtures["tables"]:
        t = vk.NttTables(e["n"], e["q"], e["w"])
        for name, field in (("roots", "roots_of_unity"),
                            ("inv_roots", "inv_roots_of_unity"),
                            ("roots_shoup", "roots_barrett_factors"),
                            ("inv_roots_shoup", "inv_roots_barrett_factors")):
            arr = getattr(t, field)
            assert [int(x) for x in arr[:8]] == e[name + "_head"]
            assert hashlib.sha256(arr.tobytes()).hexdigest() == \
                e[name + "_sha256"], (e["n"], e["q"], name)
        t.destroy()


@pytest.mark.parametrize("n", [1, 2, 4, 32, 512])
def test_tables_match_oracle(n):
    for q in (params.P0, params.Q61, 1125891450734593, params.Q62_LAZY_MAX,
              params.Q63_STRICT):
        if (q - 1) % (2 * n):
            continue
        w = params.find_psi(n, q) if n > 1 else 1
        got = vk.NttTables(n, q, w)
        want = oracle.Tables(n, q, w)
        assert np.array_equal(got.roots_of_unity, want.roots)
        assert np.array_equal(got.inv_roots_of_unity, want.inv_roots)
        assert np.array_equal(got.roots_barrett_factors, want.roots_shoup)
        assert np.array_equal(got.inv_roots_barrett_factors,
                              want.inv_roots_shoup)
        got.destroy()


@pytest.mark.parametrize("name", ["ref_test_numbers", "ref_test_ntt"])
def test_reference_host_programs(name):
    """the reference's test/numbers.c and test/ntt.c, compiled unmodified
    against include/ and libvkhel_priv.a by the Makefile"""
    path = os.path.join(ROOT, "build", "bin", name)
    if not os.path.exists(path):
        pytest.skip("%s not built (needs /root/reference at build time)"
                    % name)
    res = subprocess.run([path], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr


def test_params_match_survey():
    primes = params.ntt_primes(32)
    assert primes[0] == params.P0 == 1152921504606584833
    assert primes[31] == 1152921504455589889
    assert params.find_psi(1 << 16, primes[0]) == 987813353222176621
    assert params.find_psi(4096, params.Q61) == 700439432845261874
