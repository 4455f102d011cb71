"""GPU suite: compute-sanitizer over one pass of every kernel family
(the reference's only checker is the Vulkan validation layer;
the CUDA equivalents are memcheck and racecheck)."""
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_compute_sanitizer_clean(tool):
    exe = shutil.which("compute-sanitizer") or \
        "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    res = subprocess.run(
        [exe, "--tool", tool, "--error-exitcode", "9", sys.executable,
         os.path.join(ROOT, "tools", "sanitize_driver.py")],
        capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-4000:]
    assert "sanitize driver ok" in out
    if tool == "racecheck":
        assert "0 hazards displayed (0 errors, 0 warnings)" in out
    else:
        assert "ERROR SUMMARY: 0 errors" in out


def test_compute_sanitizer_clean_on_the_opt_in_kernels():
    """memcheck over the thread-block-cluster single pass and the TMA
    tensor-map column load (racecheck does not follow distributed shared
    memory; their results are checked against the oracle in the driver)"""
    exe = shutil.which("compute-sanitizer") or \
        "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    for switch in ("VKHEL_CLUSTER", "VKHEL_COLS_TMA"):
        env = dict(os.environ)
        env[switch] = "1"
        res = subprocess.run(
            [exe, "--tool", "memcheck", "--error-exitcode", "9", sys.executable,
             os.path.join(ROOT, "tools", "sanitize_driver.py"), "--variants"],
            capture_output=True, text=True, timeout=900, env=env)
        out = res.stdout + res.stderr
        assert res.returncode == 0, switch + out[-4000:]
        assert "sanitize driver ok" in out and "ERROR SUMMARY: 0 errors" in out
