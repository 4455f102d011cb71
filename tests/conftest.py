import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def table_fixtures():
    with open(os.path.join(GOLDEN, "reference_tables.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ctx():
    """One context for the whole GPU session, like the reference's global
    g_ctx (test/vector.c:11,348-364)."""
    import vkhel_b200 as vk
    if vk.device_count() == 0:
        pytest.fail("GPU test selected but no CUDA device is visible; "
                    "vkhel has no CPU fallback")
    context = vk.Context(0)
    yield context
    context.destroy()


def u64(x):
    return np.asarray(x, dtype=np.uint64)


def rand_mod(rng, count, q):
    """uniform residues in [0, q) for any q < 2^64"""
    raw = rng.integers(0, 1 << 63, size=count, dtype=np.uint64) * np.uint64(2) \
        + rng.integers(0, 2, size=count, dtype=np.uint64)
    return (raw % np.uint64(q)).astype(np.uint64)


def rand_u64(rng, count):
    return rng.integers(0, 1 << 63, size=count, dtype=np.uint64) * np.uint64(2) \
        + rng.integers(0, 2, size=count, dtype=np.uint64)
