"""CPU suite: the parts of bench.py's contract that need no GPU -- the
reference arm's JSON line (the CPU oracle timed on the host cores), the loud
failure of the native arm without a device, and the clock sampler's choice of
samples inside the timed region."""
import json
import os
import subprocess
import sys
import threading
import types

import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")]
                          + list(args), capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    res = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == bench.METRIC and d["unit"] == "NTT/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert (d["n_gpus"], d["steps"], d["warmup"]) == (1, 2, 1)
    assert d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "u64" and d["data"] == "synthetic"
    # the reference arm names the workload exactly like the native arm does
    assert d["config"] == bench.workload_config(1)
    assert d["config"]["limbs"] == 32 and d["config"]["batch"] == 16
    cpu = d["cpu_baseline"]
    assert "all 32 limbs" in cpu["sample"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["sample"]
    assert cpu["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "NTT/s",
                        "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_native_arm_fails_loudly_without_a_device():
    import vkhel_b200 as vk
    if vk.device_count() > 0:
        pytest.skip("a CUDA device is present")
    res = run_bench("--steps", "1", "--warmup", "1")
    assert res.returncode != 0
    assert "no CUDA device" in (res.stderr + res.stdout)
    assert not [l for l in res.stdout.splitlines() if l.startswith("{")]


def fake_sampler(lines):
    s = bench.ClockSampler(0)
    s.proc = types.SimpleNamespace(terminate=lambda: None)
    s.thread = threading.Thread(target=lambda: None)
    s.thread.start()
    s.lines = lines
    return s


def test_clock_sampler_keeps_the_samples_of_the_timed_region():
    idle = "120, 1965, 140.0, Not Active, Not Active, Not Active, Not Active"
    busy = "1965, 1965, 750.5, Not Active, Not Active, Not Active, Not Active"
    capped = "1942, 1965, 1000.1, Not Active, Not Active, Not Active, Active"
    lines = [(9.0, idle), (10.10, busy), (10.30, capped), (10.50, busy),
             (12.0, idle), (10.2, "garbage")]
    got = fake_sampler(lines).stop(10.0, 10.6)
    assert got["samples"] == 3
    assert got["sm_mhz"] == 1965.0 and got["sm_max_mhz"] == 1965.0
    assert got["power_w_max"] == 1000.1
    assert got["reasons"] == ["sw_power_cap"]
    assert "inside the timed region" in got["note"]


def test_clock_sampler_falls_back_to_the_nearest_samples():
    busy = "1965, 1965, 750.5, Not Active, Not Active, Not Active, Not Active"
    got = fake_sampler([(9.95, busy), (10.15, busy), (11.0, busy)]).stop(
        10.00, 10.01)
    assert got["samples"] == 2
    assert "shorter than one sampling period" in got["note"]
    none = bench.ClockSampler(0).stop(0.0, 1.0)
    assert none["sm_mhz"] is None and none["note"] == "nvidia-smi unavailable"


def test_workload_config_is_the_same_object_for_both_arms():
    """`config` names BASELINE configs[2] and nothing run-specific; at N > 1
    it is the limb-sharded fixed-total partition (strong scaling)"""
    one, eight = bench.workload_config(1), bench.workload_config(8)
    assert one["ntts_per_step"] == eight["ntts_per_step"] == 1024
    assert "limb-sharded 32/8" in eight["parallelism"]
    assert "4 limbs" in eight["parallelism"]
    assert set(one) == set(eight)


def test_issue_roofline_arithmetic():
    """slots per butterfly and the multiplier bound follow from the probe's
    instruction rates; the register-only rate is the probe's butterfly rate"""
    peaks = {"sm_count": 148.0, "sm_clock_mhz": 1965.0, "imad": 64.0,
             "imad_wide": 32.0, "imad_hi": 16.0, "lop3": 64.0,
             "bfly_forward": 4.0, "bfly_inverse": 3.0}
    r = bench.issue_roofline(peaks, 500.0, 1900.0)
    assert r["slots_per_butterfly"] == 4 * 2 + 4 + 4
    bound = 64.0 / 16 * 148 * 1.9
    assert abs(r["peak"] - bound) < 1e-9
    assert abs(r["frac"] - 500.0 / bound) < 1e-12
    assert abs(r["register_only_butterfly_rate"] - 3.5 * 148 * 1.9) < 1e-9


def test_inputs_do_not_depend_on_the_sharding():
    """a polynomial of the workload is the same data whichever rank owns it"""
    from vkhel_b200 import params
    primes = params.ntt_primes(bench.LIMBS)
    saved = bench.BATCH
    try:
        bench.BATCH = 2
        whole = bench.make_inputs(primes, 0, 4).reshape(2, 4, bench.N)
        part = bench.make_inputs(primes, 2, 4).reshape(2, 2, bench.N)
    finally:
        bench.BATCH = saved
    import numpy as np
    assert np.array_equal(whole[:, 2:4], part)
    assert (whole[0, 0] < primes[0]).all()
