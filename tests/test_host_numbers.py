"""CPU suite, part 4: the library's host arithmetic (vkhel_b200/csrc/numbers.c)
against Python integers and, where oracle/_ref exists, against the reference's
compiled src/numbers.c on random operands.  The nt_* symbols are not exported
from libvkhel.so (vkhel.syms), so numbers.c is compiled on its own here."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from vkhel_b200 import params
from conftest import ROOT, rand_mod

U64 = ctypes.c_uint64


@pytest.fixture(scope="module")
def nt(tmp_path_factory):
    out = tmp_path_factory.mktemp("nt") / "libnumbers.so"
    subprocess.check_call(
        ["/usr/bin/gcc", "-O2", "-fPIC", "-shared",
         "-I" + os.path.join(ROOT, "include"),
         os.path.join(ROOT, "vkhel_b200", "csrc", "numbers.c"), "-o", str(out)])
    lib = ctypes.CDLL(str(out))
    lib.nt_multiply_mod.restype = U64
    lib.nt_multiply_mod.argtypes = [U64, U64, U64, U64]
    lib.nt_power_mod.restype = U64
    lib.nt_power_mod.argtypes = [U64, U64, U64]
    lib.nt_inverse_mod.restype = U64
    lib.nt_inverse_mod.argtypes = [U64, U64]
    lib.nt_compute_barrett_factor.restype = U64
    lib.nt_compute_barrett_factor.argtypes = [U64, U64, U64]
    lib.nt_is_primitive_root.restype = ctypes.c_bool
    lib.nt_is_primitive_root.argtypes = [U64, U64, U64]
    return lib


MODULI = [2, 3, 10, 113, 769, 1125891450734593, params.Q_KAT_52, params.P0,
          params.Q61, params.Q62_LAZY_MAX, params.Q63_STRICT, (1 << 64) - 59]


@pytest.mark.parametrize("q", MODULI)
def test_against_python_integers(nt, q):
    rng = np.random.default_rng(q % 10007)
    for _ in range(300):
        a, b = (int(v) for v in rand_mod(rng, 2, q))
        assert nt.nt_multiply_mod(a, b, q, 0) == a * b % q
        assert nt.nt_compute_barrett_factor(a, q, 64) == (a << 64) // q
        e = int(rng.integers(0, 1 << 62))
        assert nt.nt_power_mod(a, e, q) == pow(a, e, q)
    if params.is_prime(q):
        for _ in range(100):
            a = int(rand_mod(rng, 1, q)[0]) or 1
            inv = nt.nt_inverse_mod(a, q)
            assert inv * a % q == 1


def test_primitive_root_test_matches_definition(nt):
    q = params.P0
    for log2n in (1, 4, 10, 16):
        n2 = 2 << log2n
        psi = params.find_psi(n2 // 2, q)
        assert nt.nt_is_primitive_root(psi, n2, q)
        assert not nt.nt_is_primitive_root(psi * psi % q, n2, q)
        assert not nt.nt_is_primitive_root(0, n2, q)
        # order exactly n2: psi^n2 = 1 and psi^(n2/2) = -1
        assert pow(psi, n2, q) == 1 and pow(psi, n2 // 2, q) == q - 1


def test_against_reference_compiled_code(nt):
    ref = oracle.reference_host()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(99)
    for q in (769, 1125891450734593, params.P0, params.Q61,
              params.Q62_LAZY_MAX):
        for _ in range(300):
            a, b = (int(v) for v in rand_mod(rng, 2, q))
            assert nt.nt_multiply_mod(a, b, q, 0) == \
                ref.nt_multiply_mod(a, b, q, 0)
            assert nt.nt_power_mod(a, b, q) == ref.nt_power_mod(a, b, q)
            assert nt.nt_compute_barrett_factor(a, q, 64) == \
                ref.nt_compute_barrett_factor(a, q, 64)
            if a:
                assert nt.nt_inverse_mod(a, q) == ref.nt_inverse_mod(a, q)


def test_golden_kats_are_current():
    """tests/golden/reference_kats.json equals a fresh extraction from the
    reference's tests (only checkable where /root/reference exists)"""
    if not os.path.exists("/root/reference/test/vector.c"):
        pytest.skip("no /root/reference here")
    import json
    import shutil
    import tempfile
    golden = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(golden, "reference_kats.json")) as f:
        committed = json.load(f)
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, "extract_kats.py")
        shutil.copy(os.path.join(golden, "extract_kats.py"), script)
        subprocess.check_call(["python", script], stdout=subprocess.DEVNULL)
        with open(os.path.join(tmp, "reference_kats.json")) as f:
            fresh = json.load(f)
    assert fresh == committed
