"""GPU suite: the reference's own programs, compiled UNMODIFIED from
/root/reference (examples/example.c, test/vector.c, test/ntt.c,
test/numbers.c) against include/ and this library by the Makefile, must run
to completion: that is the drop-in claim of the C API."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "build", "bin")


@pytest.mark.parametrize("name", ["ref_test_vector", "ref_example",
                                  "ref_test_ntt", "ref_test_numbers"])
def test_reference_program(name):
    path = os.path.join(BIN, name)
    if not os.path.exists(path):
        pytest.fail("%s missing: run `make` where /root/reference exists; "
                    "the binary travels to the GPU box" % path)
    res = subprocess.run([path], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    if name == "ref_test_vector":
        for t in ("copy_from_host", "dup", "elemfma", "elemmod", "elemmul",
                  "elemgtadd", "elemgtsub", "forward_transform",
                  "inverse_transform", "forward_transform_big",
                  "inverse_transform_big"):
            assert "%s passed" % t in res.stdout
    if name == "ref_example":
        assert res.stdout.count("{4, 0, 0, 4, }") == 2


def test_multi_gpu_example():
    """examples/multi_gpu.c: limb-sharded transform over every visible GPU
    (host thread + context per GPU), gathered with vkhel_vector_copy_peer and
    compared with the single-GPU result; degenerates gracefully to 1 GPU"""
    path = os.path.join(BIN, "multi_gpu")
    if not os.path.exists(path):
        pytest.fail("%s missing: run `make`" % path)
    for args in (["12", "8", "4"], ["16", "5", "2"], ["4", "3", "7"]):
        res = subprocess.run([path] + args, capture_output=True, text=True,
                             timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        assert "== single-GPU result" in res.stdout


@pytest.mark.parametrize("env", [{}, {"VKHEL_NO_FUSED_PRODUCT": "1"},
                                 {"VKHEL_NO_DEFER": "1"}])
def test_api_product_example(env):
    """examples/api_product.c: the reference's product sequence (forward,
    forward, elemmul, inverse in place) from C against the schoolbook product;
    by default the element-wise product goes out inside the inverse"""
    path = os.path.join(BIN, "api_product")
    if not os.path.exists(path):
        pytest.fail("%s missing: run `make`" % path)
    for args in (["3", "5"], ["8", "9"], ["10", "7"], ["12", "6"], ["16", "2"]):
        res = subprocess.run([path] + args, capture_output=True, text=True,
                             timeout=300, env=dict(os.environ, **env))
        assert res.returncode == 0, res.stdout + res.stderr
        assert '"matches_schoolbook": true' in res.stdout
        fused = '"fused_products": 0,' not in res.stdout
        assert fused == (not env), res.stdout
