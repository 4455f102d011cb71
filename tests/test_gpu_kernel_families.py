"""GPU suite: the three NTT kernel families on the same moduli.

By default q < 2^64/6 runs the approximate-quotient butterflies, larger
q < 2^62 the exact-quotient ones and q >= 2^62 (or n < 8) the generic kernel.
The environment switches $VKHEL_EXACT_QUOTIENT and $VKHEL_FORCE_GENERIC force
the slower families for every modulus, $VKHEL_POLYMUL_UNFUSED the polynomial
product as separate transforms, $VKHEL_NO_DEFER immediate launches of the
single-vector transforms, $VKHEL_SINGLE_MAX_LOG2N the sizes that take the
single-pass kernel (8 = none, 12 and 13 = beyond the default),
$VKHEL_CLUSTER=1 sends 2^14 <= n <= 2^16 through the thread-block-cluster
kernel (kernels_ntt_cluster.cu; more shapes in tests/test_gpu_cluster.py),
$VKHEL_COLS_TMA=1 loads the forward column tiles of n = 2^16 by TMA tensor map
(kernels_ntt_tma.cu) and
$VKHEL_SLICE_MIB=0.0625 cuts every two-pass batch above 192 KiB into slices on two
streams (four with $VKHEL_SLICE_STREAMS=4; joined after every call with
$VKHEL_LAZY_JOIN=0); they are read once per process, so the parity tests are re-run in
child processes."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SELECT = ("ntt_random_all_sizes or ntt_random_large or kat or batch_matches "
          "or rns_matches or polymul or inverse_scales_tail")


@pytest.mark.parametrize("switch", ["VKHEL_EXACT_QUOTIENT",
                                    "VKHEL_FORCE_GENERIC",
                                    "VKHEL_POLYMUL_UNFUSED",
                                    "VKHEL_NO_DEFER",
                                    "VKHEL_SINGLE_MAX_LOG2N=8",
                                    "VKHEL_SINGLE_MAX_LOG2N=12",
                                    "VKHEL_SINGLE_MAX_LOG2N=13",
                                    "VKHEL_CLUSTER=1",
                                    "VKHEL_COLS_TMA=1",
                                    "VKHEL_SLICE_MIB=0.0625",
                                    "VKHEL_SLICE_MIB=0.0625,VKHEL_SLICE_STREAMS=4",
                                    "VKHEL_SLICE_MIB=0.0625,VKHEL_LAZY_JOIN=0"])
def test_parity_with_forced_family(switch):
    env = dict(os.environ)
    for one in switch.split(","):
        name, _, value = one.partition("=")
        env[name] = value or "1"
    res = subprocess.run(
        [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu",
         "-p", "no:cacheprovider", "-k", SELECT,
         os.path.join(ROOT, "tests", "test_gpu_parity.py")],
        env=env, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert res.returncode == 0, (res.stdout + res.stderr)[-4000:]
    assert " passed" in res.stdout
