#!/usr/bin/env python3
"""Benchmark of the NTT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): 64-bit NTTs/sec at n=2^16.  Workload: the CKKS-like
shape of BASELINE configs[2] -- n = 2^16, 32 RNS limbs (the 32 largest primes
below 2^60 with q = 1 mod 2^18) x batch 16 = 512 polynomials = 256 MiB per
GPU.  One step = one forward transform of the whole batch followed by one
inverse transform of it (1024 NTTs).  With N GPUs every rank runs that
workload on its own batch shard (weak scaling, no data-path collective; the
reference has no multi-device path at all).

One JSON line is printed by rank 0.  `value` is measured with the inputs
resident in HBM; `e2e` is the same metric through the C-ABI with host buffers
(pinned host -> device copy of the step's input and device -> host copy of its
result inside the timed region).  `roofline` places the step against the HBM
roofline with the algorithmic 16*n bytes per NTT; `cpu_baseline` is the CPU
oracle (a port of the reference's arithmetic; the reference has no CPU NTT of
its own) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2N = 16
N = 1 << LOG2N
LIMBS = 32
BATCH = 16
E2E_CHUNKS = 4
POLYS = LIMBS * BATCH
METRIC = "64-bit NTTs/sec at n=2^16"
# measured per-GPU integer peaks (profiles/r01_bfly_bench.txt)
BFLY_PEAK_G = 1038.0         # lazy Harvey butterflies/s, best compiled form (v19)
BFLY_MULT_BOUND_G = 1163.0   # 16 fmaheavy slots per butterfly, nothing else
# DRAM bytes of one step as measured by ncu (profiles/r01_ntt_traffic.json,
# made by tools/ncu_traffic.py from a --cache-control none launch list of this
# very command); the constant is the value of that file at commit time
NCU_TRAFFIC_BYTES_PER_STEP = 1180000000
NCU_TRAFFIC_SOURCE = ("ncu --cache-control none --metrics dram__bytes_*, "
                      "profiles/r01_ntt_traffic.json")


def ncu_traffic():
    path = os.path.join(ROOT, "profiles", "r01_ntt_traffic.json")
    try:
        with open(path) as f:
            return int(json.load(f)["bytes_per_step"]), NCU_TRAFFIC_SOURCE
    except (OSError, KeyError, ValueError):
        return NCU_TRAFFIC_BYTES_PER_STEP, NCU_TRAFFIC_SOURCE


WORKLOAD = ("n=2^16 negacyclic NTT, 32 RNS limbs (60-bit primes) x batch 16 "
            "= 512 polys = 256 MiB per GPU (BASELINE configs[2] shape); "
            "step = forward + inverse of the whole batch")


class c_stdout_to_stderr:
    """The library prints `using physical device N: ...` on stdout like the
    reference (src/vulkan.c:171); keep this program's stdout to the one JSON
    line by pointing fd 1 at stderr while a context is created."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ---- clocks --------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region.

    One long-lived `nvidia-smi -lms` process, started well before the timed
    region (its start-up takes a few hundred milliseconds on a fresh box);
    every line is stamped on arrival and stop(t0, t1) keeps the samples that
    fall inside the timed region [t0, t1]."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,"
             "clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    PERIOD_MS = 20

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device),
                 "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", str(self.PERIOD_MS)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=5.0):
        """block (bounded) until nvidia-smi has delivered its first line"""
        t_end = time.perf_counter() + timeout
        while self.proc and not self.lines and time.perf_counter() < t_end:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                    "note": "nvidia-smi unavailable"}
        time.sleep(2.5 * self.PERIOD_MS / 1e3)
        self.proc.terminate()
        self.thread.join(timeout=2)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]

        def collect(lo, hi):
            sm, smax, reasons, power = [], [], set(), []
            for stamp, line in self.lines:
                if stamp < lo or stamp > hi:
                    continue
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    smax.append(float(parts[1]))
                    power.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, smax, reasons, power

        # a line is stamped when it arrives, one sampling period at most after
        # the reading it carries
        slack = self.PERIOD_MS / 1e3
        sm, smax, reasons, power = collect(t0, t1 + slack)
        note = "samples inside the timed region"
        if not sm:
            # timed region shorter than the sampling period: the nearest
            # samples around it (the GPU is under the same load: warm-up steps
            # before, the forward-only / inverse-only timing after)
            sm, smax, reasons, power = collect(t0 - 0.2, t1 + 0.2)
            note = "timed region shorter than one sampling period: samples within 0.2 s of it"
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "period_ms": self.PERIOD_MS,
                "reasons": sorted(reasons), "note": note}


# ---- workload ------------------------------------------------------------------------
def make_inputs(primes, seed):
    from vkhel_b200 import params
    out = np.empty(POLYS * N, np.uint64)
    for p in range(POLYS):
        q = primes[p % LIMBS]
        out[p * N:(p + 1) * N] = params.xorshift64_stream(
            0x9E3779B97F4A7C15 + seed * 1000 + p, N, q)
    return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU arm -------------------------------------------------------------------------
def cpu_arm(primes, psis, target_seconds, sample_polys=None):
    """fwd+inv on a bounded sample with the CPU oracle on all host cores.
    Returns (ntts_per_sec, cores, sample description, seconds)."""
    import oracle
    cores = os.cpu_count() or 1
    limbs = min(LIMBS, 4)
    tables = [oracle.Tables(N, primes[l], psis[l]) for l in range(limbs)]
    from vkhel_b200 import params
    one = params.xorshift64_stream(1, N, primes[0])
    t0 = time.perf_counter()
    oracle.inverse(oracle.forward(one, tables[0]), tables[0])
    t_one = time.perf_counter() - t0            # 2 NTTs, single thread
    if sample_polys is None:
        sample_polys = int(max(cores, min(POLYS, cores * target_seconds / t_one)))
        sample_polys -= sample_polys % limbs or 0
        sample_polys = max(sample_polys, limbs)
    x = np.concatenate([params.xorshift64_stream(100 + p, N, primes[p % limbs])
                        for p in range(sample_polys)])
    reps = 0
    t0 = time.perf_counter()
    while True:
        fwd = oracle.forward_batch(x, tables, threads=cores)
        back = oracle.inverse_batch(fwd, tables, threads=cores)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= target_seconds or reps >= 64:
            break
    assert np.array_equal(back, x)
    sample = ("%d x (%d polys of n=2^16 over %d limbs, forward+inverse), "
              "OpenMP over polynomials" % (reps, sample_polys, limbs))
    return 2 * sample_polys * reps / dt, cores, sample, dt, (tables, x)


def run_reference_arm(args):
    """--impl reference: the reference's CPU arithmetic (the oracle port: the
    reference's transform exists only as GLSL and no Vulkan stack is in this
    image) on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from vkhel_b200 import params
    import oracle
    primes = params.ntt_primes(LIMBS)
    psis = [params.find_psi(N, q) for q in primes[:4]]
    cores = os.cpu_count() or 1
    rate, cores, sample, dt, (tables, x) = cpu_arm(primes, psis, 2.0)
    # one step = the per-GPU batch (512 polynomials) when K + W steps of it
    # fit in about two minutes, else the largest sample that does
    limbs = len(tables)
    budget_polys = rate / 2 * 120.0 / (args.steps + max(args.warmup, 1))
    sample_polys = int(min(x.size // N, max(budget_polys, cores, limbs)))
    sample_polys = max(sample_polys - sample_polys % limbs, limbs)
    x = x[:sample_polys * N]
    sample = ("%d polys of n=2^16 over %d limbs, forward+inverse, OpenMP over "
              "polynomials" % (sample_polys, limbs))
    for _ in range(max(args.warmup - 1, 0)):
        oracle.inverse_batch(oracle.forward_batch(x, tables, threads=cores),
                             tables, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.inverse_batch(oracle.forward_batch(x, tables, threads=cores),
                             tables, threads=cores)
    elapsed = time.perf_counter() - t0
    value = 2 * sample_polys * args.steps / elapsed
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "NTT/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "sample": sample + " per step"},
        "cpu_baseline": {"value": value, "unit": "NTT/s", "cores": cores,
                         "kind": "port", "sample": sample + " per step"},
        "e2e": {"value": value, "unit": "NTT/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_near_gpu(device):
    """Multi-rank runs: keep this rank (and therefore its pinned host buffers,
    which are first-touched by it) on the NUMA node its GPU hangs off, so that
    the ranks' host <-> device copies do not all cross one inter-socket link.
    Best effort; returns what was done for the JSON line."""
    try:
        bus = subprocess.check_output(
            ["nvidia-smi", "-i", str(device), "--query-gpu=pci.bus_id",
             "--format=csv,noheader"], text=True).strip().lower()
        dom, rest = bus.split(":", 1)
        sysdir = "/sys/bus/pci/devices/%s:%s" % (dom[-4:], rest)
        with open(sysdir + "/numa_node") as f:
            node = int(f.read())
        with open(sysdir + "/local_cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except (OSError, ValueError, subprocess.SubprocessError) as err:
        return {"error": str(err)[:80]}


# ---- GPU arm -------------------------------------------------------------------------
def run_native_arm(args):
    import vkhel_b200 as vk
    from vkhel_b200 import params

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl",
                                device_id=torch.device("cuda", local_rank))

    if vk.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; vkhel has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    numa = bind_near_gpu(local_rank) if world > 1 else None

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(value):
        if dist is None:
            return value
        import torch
        from vkhel_b200 import shard
        return shard.max_over_ranks(value, dist,
                                    torch.device("cuda", local_rank))

    primes = params.ntt_primes(LIMBS)
    psis = [params.find_psi(N, q) for q in primes]
    with c_stdout_to_stderr():
        ctx = vk.Context(local_rank)
    tables = [vk.NttTables(N, q, w) for q, w in zip(primes, psis)]

    host_in = vk.host_alloc(POLYS * N)
    host_out = vk.host_alloc(POLYS * N)
    host_in.array[:] = make_inputs(primes, rank)
    data = ctx.vector(POLYS * N, zero=False)
    work = ctx.vector(POLYS * N, zero=False)
    data.upload(host_in)
    ctx.sync()

    def step_resident():
        ctx.forward_transform_rns(data, work, tables, BATCH)
        ctx.inverse_transform_rns(work, work, tables, BATCH)

    # end-to-end path: the batch moves in E2E_CHUNKS slices of whole batch
    # entries, each slice on its own device vector, two sets of slices used
    # alternately: upload (H2D copy stream), both transforms (compute stream)
    # and download (D2H copy stream) of different slices overlap.
    chunk_batch = BATCH // E2E_CHUNKS
    chunk_elems = chunk_batch * LIMBS * N
    slices = [[ctx.vector(chunk_elems, zero=False) for _ in range(E2E_CHUNKS)]
              for _ in range(2)]
    e2e_state = {"step": 0}

    def step_e2e():
        cur = slices[e2e_state["step"] & 1]
        e2e_state["step"] += 1
        for c, v in enumerate(cur):
            v.upload(host_in, count=chunk_elems, host_offset=c * chunk_elems)
            ctx.forward_transform_rns(v, v, tables, chunk_batch)
            ctx.inverse_transform_rns(v, v, tables, chunk_batch)
            v.download(host_out, count=chunk_elems,
                       host_offset=c * chunk_elems)

    timer = ctx.timer()

    timed_window = [0.0, 0.0]   # host clock around the last timed region

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        ctx.sync()
        barrier()
        l0 = ctx.launch_count
        timed_window[0] = time.perf_counter()
        timer.start()
        for _ in range(steps):
            step_fn()
        timer.stop()
        ms = timer.elapsed_ms()
        ctx.sync()
        timed_window[1] = time.perf_counter()
        barrier()
        return max_over_ranks(ms), ctx.launch_count - l0

    # bring the GPU out of its idle clocks before anything is timed: an idle
    # B200 sits at 120 MHz and needs tens of milliseconds of load to reach its
    # boost clock; W steps of 0.8 ms are too short for that
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.25:
        for _ in range(20):
            step_resident()
        ctx.sync()
    if rank == 0:
        # keep the GPU under load while nvidia-smi starts up
        while (sampler.proc and not sampler.lines
               and time.perf_counter() - t_spin < 5.0):
            for _ in range(20):
                step_resident()
            ctx.sync()
    barrier()
    ms_total, launches = timed(step_resident, args.steps, args.warmup)
    window = tuple(timed_window)
    # the same load continues while the last samples arrive
    for _ in range(60):
        step_resident()
    ctx.sync()
    clocks = sampler.stop(*window) if rank == 0 else None

    # correctness of what was just timed: the round trip returns the input
    work.download(host_out)
    ctx.sync()
    ok = bool(np.array_equal(host_out.array, host_in.array))

    # forward-only and inverse-only durations (per direction roofline)
    def fwd_only():
        ctx.forward_transform_rns(data, work, tables, BATCH)

    def inv_only():
        ctx.inverse_transform_rns(work, work, tables, BATCH)

    ms_fwd, _ = timed(fwd_only, args.steps, 1)
    ms_inv, _ = timed(inv_only, args.steps, 1)

    host_out.array[:] = 0
    e2e_steps = max(2, min(args.steps, 20))
    ms_e2e, _ = timed(step_e2e, e2e_steps, 2)
    ok = ok and bool(np.array_equal(host_out.array, host_in.array))

    # the same trip through the reference's 18 calls only (one vector per
    # polynomial, pageable host memory): copy_from_host, forward_transform,
    # inverse_transform, map (read the result), unmap.  A bounded sample of
    # the batch: LEGACY_POLYS polynomials over the first limbs.
    legacy_polys = 64
    legacy_vecs = [ctx.vector(N, zero=False) for _ in range(legacy_polys)]
    legacy_in = np.array(host_in.array[:legacy_polys * N])   # pageable copy
    legacy_out = np.empty_like(legacy_in)

    def step_legacy():
        for p, v in enumerate(legacy_vecs):
            v.copy_from_host(legacy_in[p * N:(p + 1) * N])
        for p, v in enumerate(legacy_vecs):
            ctx.forward_transform(v, v, tables[p % LIMBS])
        for p, v in enumerate(legacy_vecs):
            ctx.inverse_transform(v, v, tables[p % LIMBS])
        for p, v in enumerate(legacy_vecs):
            legacy_out[p * N:(p + 1) * N] = v.to_host()

    step_legacy()
    ctx.sync()
    t_legacy = time.perf_counter()
    legacy_steps = 3
    for _ in range(legacy_steps):
        step_legacy()
    ctx.sync()
    t_legacy = (time.perf_counter() - t_legacy) / legacy_steps
    ok = ok and bool(np.array_equal(legacy_out, legacy_in))
    legacy_value = world * 2 * legacy_polys / t_legacy

    ntts_per_step = 2 * POLYS
    ms_per_step = ms_total / args.steps
    value = world * ntts_per_step / (ms_per_step * 1e-3)
    e2e_value = world * ntts_per_step / (ms_e2e / e2e_steps * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peaks()
        algo_bytes = 16 * N * ntts_per_step
        achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "NTT/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {
                "workload": WORKLOAD, "n": N, "limbs": LIMBS,
                "batch_per_gpu": BATCH, "parallelism": "batch-sharded x%d, "
                "no collective" % world,
                "l2": "working set 512 MiB per step > 126 MB L2, no flush",
                "round_trip_exact": ok,
                "host_numa_binding": numa,
            },
            "e2e": {"value": e2e_value, "unit": "NTT/s",
                    "h2d_bytes_per_step": POLYS * N * 8,
                    "d2h_bytes_per_step": POLYS * N * 8,
                    "steps": e2e_steps,
                    "path": "per 64 MiB slice: vkhel_vector_upload (pinned) "
                            "-> forward_transform_rns -> "
                            "inverse_transform_rns -> vkhel_vector_download;"
                            " copies on the context's H2D/D2H streams "
                            "overlap with each other and with compute"},
            "e2e_reference_api": {
                "value": legacy_value, "unit": "NTT/s",
                "sample": "%d polynomials, one vkhel_vector each" % legacy_polys,
                "path": "the reference's 18 entry points only, pageable host "
                        "memory: vkhel_vector_copy_from_host -> "
                        "vkhel_vector_forward_transform -> "
                        "vkhel_vector_inverse_transform -> vkhel_vector_map "
                        "+ copy out + vkhel_vector_unmap (which writes the "
                        "vector back, as the reference does); host wall "
                        "clock"},
            # whole job: every rank launches the same kernels on its shard
            "gpu_launches": launches * world,
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic()[0],
                "traffic_source": ncu_traffic()[1] + " (two passes per "
                                  "transform, the intermediate stays in L2: "
                                  "limb slices on two streams)",
                "peak_source": peak_src,
                "kernel": "whole step (forward + inverse NTT, all passes)",
                "algorithmic_bytes_per_step": algo_bytes,
                "forward_ms": ms_fwd / args.steps,
                "inverse_ms": ms_inv / args.steps,
                "forward_GBps": 16 * N * POLYS / (ms_fwd / args.steps * 1e-3) / 1e9,
                "inverse_GBps": 16 * N * POLYS / (ms_inv / args.steps * 1e-3) / 1e9,
            },
        }
        # the binding roofline: integer (fmaheavy-pipe) butterfly rate,
        # measured by tools/bfly_bench.cu on this pool (profiles/r01_bfly_bench.txt)
        bfly_per_step = ntts_per_step * (N // 2) * LOG2N
        bfly_rate = bfly_per_step / (ms_per_step * 1e-3) / 1e9
        line["issue_roofline"] = {
            "bound": "imad (fmaheavy pipe)", "achieved": bfly_rate,
            "peak": BFLY_PEAK_G, "unit": "G butterflies/s",
            "frac": bfly_rate / BFLY_PEAK_G,
            "peak_source": "tools/bfly_bench.cu v19 (the library's butterfly: "
                           "borrow-chain csub, approximate-quotient Shoup "
                           "product as one PTX mad chain; twiddles in uniform "
                           "registers, no memory): 3.57 per clk per SM x 148 "
                           "SMs x 1.965 GHz",
            "pure_multiplier_bound": BFLY_MULT_BOUND_G,
        }
        if world == 1 and not args.no_cpu:
            rate, cores, sample, dt, _ = cpu_arm(primes, psis[:4], 12.0)
            line["cpu_baseline"] = {"value": rate, "unit": "NTT/s",
                                    "cores": cores, "kind": "port",
                                    "sample": sample, "seconds": dt}
        print(json.dumps(line))

    timer.destroy()
    for v in [data, work] + slices[0] + slices[1] + legacy_vecs:
        v.destroy()
    for t in tables:
        t.destroy()
    host_in.free()
    host_out.free()
    ctx.destroy()
    if dist is not None:
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: round trip mismatch -- result invalid")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true",
                    help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native_arm(args)


if __name__ == "__main__":
    main()
