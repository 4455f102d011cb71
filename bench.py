#!/usr/bin/env python3
"""Benchmark of the NTT hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): 64-bit NTTs/sec at n=2^16.  Workload: BASELINE
configs[2] -- n = 2^16, 32 RNS limbs (the 32 largest primes below 2^60 with
q = 1 mod 2^18) x batch 16 = 512 polynomials = 256 MiB.  One step = one forward
transform of the whole batch followed by one inverse transform of it (1024
NTTs).

N = 1: the whole workload on one GPU.  N > 1 (torchrun, one rank per GPU):
the SAME total workload, LIMB-SHARDED as BASELINE names it -- rank r owns the
contiguous limb range 32/N and uploads only those primes' tables; `value` is
whole-job NTTs over the slowest rank's time: strong scaling, no data-path
collective (the reference has no multi-device path at all).  The
weak-scaling figure (every rank runs the full one-GPU workload) is the extra
key `weak`.

One JSON line is printed by rank 0.  `value` is measured with the inputs
resident in HBM; `e2e` is the same metric through the C-ABI with host buffers
(pinned host -> device copy of the step's input and device -> host copy of its
result inside the timed region; median of three segments) next to
`e2e_ceiling`, the box's own host <-> device bandwidth measured without the
library right before and right after those segments.  At N > 1 the end-to-end
path shards by the per-GPU rates it has just measured (vkhel_b200/shard.py,
e2e_plan): the GPUs of a box do not get equal shares of the host fabric, and
with equal shards the job would end with the slowest one.  `roofline`
places the step against the HBM roofline with the algorithmic 16*n bytes per
NTT; `issue_roofline` against the integer (fmaheavy) pipe, with the pipe rates
and the register-only butterfly rate measured in this run
(vkhel_ctx_probe_int_peaks); `sustained` repeats the step for at least a second
whatever --steps says; `cpu_baseline` is the CPU oracle (a port of the
reference's arithmetic; the reference has no CPU NTT of its own) on all host
cores.  What was timed is checked against the oracle outside the timed region
(`checks`).
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
SUSTAINED_SECONDS = 1.0
# DRAM bytes of one step as measured by ncu (profiles/*_ntt_traffic.json, made
# by tools/ncu_traffic.py from a --cache-control none launch list of this very
# command); the constant is the fallback when no such file travels with the run
NCU_TRAFFIC_BYTES_PER_STEP = 1180000000
NCU_TRAFFIC_SOURCE = ("ncu --cache-control none --metrics dram__bytes_*, "
                      "profiles/%s")


def ncu_traffic():
    for name in ("r02_ntt_traffic.json", "r01_ntt_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return (int(json.load(f)["bytes_per_step"]),
                        NCU_TRAFFIC_SOURCE % name)
        except (OSError, KeyError, ValueError):
            continue
    return NCU_TRAFFIC_BYTES_PER_STEP, NCU_TRAFFIC_SOURCE % "(constant)"


WORKLOAD = ("n=2^16 negacyclic NTT, 32 RNS limbs (60-bit primes) x batch 16 "
            "= 512 polys = 256 MiB in all (BASELINE configs[2]); "
            "step = forward + inverse of the whole batch")


def workload_config(world):
    """the `config` object of BOTH arms: it names the workload, nothing that
    differs between the arms or between runs"""
    if world == 1:
        parallelism = "single GPU, all 32 limbs"
    else:
        parallelism = ("limb-sharded %d/%d: each of %d GPUs owns %d limbs x "
                       "batch %d, fixed total, no collective"
                       % (LIMBS, world, world, LIMBS // world, BATCH))
    return {
        "workload": WORKLOAD, "n": N, "limbs": LIMBS, "batch": BATCH,
        "ntts_per_step": 2 * POLYS, "parallelism": parallelism,
        "l2": "working set 512 MiB per step over all GPUs, read and written "
              "once per pass; at N <= 2 larger than the 126 MB L2, no flush; "
              "at N >= 4 the shard (<= 64 MiB) stays L2-resident between "
              "steps, as it would in an application that keeps its limbs on "
              "the GPU",
    }


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
    """nvidia-smi clocks / throttle reasons during the timed regions.

    One long-lived `nvidia-smi -lms` process, started well before the first
    timed region (its start-up takes a few hundred milliseconds on a fresh
    box); every line is stamped on arrival, window(t0, t1) summarises the
    samples that fall inside [t0, t1] and stop(t0, t1) does the same and ends
    the process."""
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

    def window(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
                    "note": "nvidia-smi unavailable"}
        # the last readings of the window are still on their way
        time.sleep(2.5 * self.PERIOD_MS / 1e3)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]

        def collect(lo, hi):
            sm, smax, reasons, power = [], [], set(), []
            for stamp, line in list(self.lines):
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
            # samples around it (the GPU is under the same load before and
            # after: warm-up steps, the sustained run)
            sm, smax, reasons, power = collect(t0 - 0.2, t1 + 0.2)
            note = "timed region shorter than one sampling period: samples within 0.2 s of it"
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "period_ms": self.PERIOD_MS,
                "reasons": sorted(reasons), "note": note}

    def stop(self, t0, t1):
        out = self.window(t0, t1)
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)
        return out


# ---- workload ------------------------------------------------------------------------
def poly_seed(batch_entry, limb):
    """the same polynomial whatever the sharding: seeded by its position in
    the [batch][32 limbs] layout (xorshift64, coefficients mod q)"""
    return 0x9E3779B97F4A7C15 + 2 * 1000 + batch_entry * LIMBS + limb


def make_inputs(primes, lo, hi):
    """[BATCH][hi - lo][N]: the limbs [lo, hi) of the workload"""
    from vkhel_b200 import params
    own = hi - lo
    out = np.empty(BATCH * own * N, np.uint64)
    for b in range(BATCH):
        for l in range(lo, hi):
            p = b * own + (l - lo)
            out[p * N:(p + 1) * N] = params.xorshift64_stream(
                poly_seed(b, l), N, primes[l])
    return out


def make_inputs_for(primes, limb_list, b0, b1):
    """[b1 - b0][limb_list][N]: the same polynomials as make_inputs, for any
    limbs and batch range (the end-to-end path's plan)"""
    from vkhel_b200 import params
    out = np.empty((b1 - b0) * len(limb_list) * N, np.uint64)
    p = 0
    for b in range(b0, b1):
        for l in limb_list:
            out[p * N:(p + 1) * N] = params.xorshift64_stream(
                poly_seed(b, l), N, primes[l])
            p += 1
    return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---- CPU arm -------------------------------------------------------------------------
class CpuArm:
    """the CPU oracle on all host cores over all 32 limbs of the workload;
    a step transforms `sample_polys` polynomials forward and back"""

    def __init__(self):
        import oracle
        from vkhel_b200 import params
        self.oracle = oracle
        self.cores = os.cpu_count() or 1
        self.primes = params.ntt_primes(LIMBS)
        psis = [params.find_psi(N, q) for q in self.primes]
        self.tables = [oracle.Tables(N, q, w) for q, w in zip(self.primes, psis)]
        one = params.xorshift64_stream(1, N, self.primes[0])
        t0 = time.perf_counter()
        oracle.inverse(oracle.forward(one, self.tables[0]), self.tables[0])
        self.t_pair = time.perf_counter() - t0          # 2 NTTs, single thread

    def sample(self, polys):
        """`polys` polynomials in the workload's [batch][32 limbs] order
        (polynomial p uses limb p % 32), a multiple of 32"""
        from vkhel_b200 import params
        polys = max(LIMBS, polys - polys % LIMBS)
        polys = min(polys, POLYS)
        x = np.concatenate([params.xorshift64_stream(
            poly_seed(p // LIMBS, p % LIMBS), N, self.primes[p % LIMBS])
            for p in range(polys)])
        return polys, x

    def step(self, x):
        fwd = self.oracle.forward_batch(x, self.tables, threads=self.cores)
        return self.oracle.inverse_batch(fwd, self.tables, threads=self.cores)

    def polys_for(self, seconds):
        """sample size whose step takes about `seconds` on all cores"""
        return int(self.cores * seconds / self.t_pair)


def cpu_baseline(target_seconds):
    """fwd+inv on a bounded sample of the workload (all 32 limbs), repeated
    until about `target_seconds` have passed"""
    arm = CpuArm()
    polys, x = arm.sample(arm.polys_for(target_seconds / 4))
    reps = 0
    t0 = time.perf_counter()
    while True:
        back = arm.step(x)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= target_seconds or reps >= 64:
            break
    assert np.array_equal(back, x)
    sample = ("%d x (%d polys of n=2^16 over all %d limbs, forward+inverse), "
              "OpenMP over polynomials" % (reps, polys, LIMBS))
    return {"value": 2 * polys * reps / dt, "unit": "NTT/s",
            "cores": arm.cores, "kind": "port", "sample": sample,
            "seconds": dt}


def run_reference_arm(args):
    """--impl reference: the reference's CPU arithmetic (the oracle port: the
    reference's transform exists only as GLSL and no Vulkan stack is in this
    image) on all host cores, on the same workload -- all 32 limbs -- each
    step a bounded sample of it."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    arm = CpuArm()
    # one step = the whole batch (512 polynomials) when K + W steps of it fit
    # in about two minutes, else the largest whole-basis sample that does
    budget = 120.0 / (args.steps + max(args.warmup, 1))
    polys, x = arm.sample(arm.polys_for(budget))
    sample = ("%d of the %d polys (n=2^16, all %d limbs), forward+inverse, "
              "OpenMP over polynomials" % (polys, POLYS, LIMBS))
    for _ in range(max(args.warmup, 0)):
        arm.step(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        back = arm.step(x)
    elapsed = time.perf_counter() - t0
    assert np.array_equal(back, x)
    value = 2 * polys * args.steps / elapsed
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "NTT/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(world),
        "cpu_baseline": {"value": value, "unit": "NTT/s", "cores": arm.cores,
                         "kind": "port", "sample": sample + " per step"},
        "e2e": {"value": value, "unit": "NTT/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


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
def issue_roofline(peaks, bfly_rate_g, kernel_clock_mhz):
    """the binding roofline: butterflies/s against the fmaheavy pipe.  `peaks`
    are this run's vkhel_ctx_probe_int_peaks: instruction rates per clock per
    SM and the library's butterfly on registers only."""
    sms = peaks["sm_count"]
    ghz = (kernel_clock_mhz or peaks["sm_clock_mhz"]) / 1e3
    # issue slots of the pipe per instruction, in units of one IMAD
    slot_wide = peaks["imad"] / peaks["imad_wide"]
    slot_hi = peaks["imad"] / peaks["imad_hi"]
    # a lazy butterfly is one Shoup product with the approximate quotient:
    # 4 IMAD.WIDE + 1 IMAD.HI + 4 IMAD (modarith.cuh, shoup_chain<true>)
    slots = 4 * slot_wide + slot_hi + 4
    mult_bound = peaks["imad"] / slots * sms * ghz            # G butterflies/s
    reg_only = 0.5 * (peaks["bfly_forward"] + peaks["bfly_inverse"]) * sms * ghz
    return {
        "bound": "imad (fmaheavy pipe)", "achieved": bfly_rate_g,
        "peak": mult_bound, "unit": "G butterflies/s",
        "frac": bfly_rate_g / mult_bound,
        "peak_source": "measured in this run: IMAD %.1f, IMAD.WIDE %.1f, "
                       "IMAD.HI %.1f thread-instructions/clk/SM; a butterfly's "
                       "multiplies (4 wide + 1 hi + 4 narrow) = %.1f IMAD "
                       "slots; x %d SMs x %.3f GHz (SM clock of the timed "
                       "region)" % (peaks["imad"], peaks["imad_wide"],
                                    peaks["imad_hi"], slots, sms, ghz),
        "slots_per_butterfly": slots,
        "register_only_butterfly_rate": reg_only,
        "frac_of_register_only": bfly_rate_g / reg_only,
        "register_only_source": "the library's ct_lazy3 / gs_lazy3 on "
                                "register operands, no memory: %.2f / %.2f per "
                                "clk per SM, measured in this run"
                                % (peaks["bfly_forward"], peaks["bfly_inverse"]),
        "probe_clock_mhz": peaks["sm_clock_mhz"],
        "alu_pipe_lop3_per_clk_sm": peaks["lop3"],
    }


def run_native_arm(args):
    import vkhel_b200 as vk
    from vkhel_b200 import params, shard

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
        if LIMBS % world:
            raise SystemExit("bench.py: %d limbs do not shard evenly over %d "
                             "GPUs" % (LIMBS, world))

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
        return shard.max_over_ranks(value, dist,
                                    torch.device("cuda", local_rank))

    def sum_over_ranks(value):
        if dist is None:
            return value
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64,
                         device=torch.device("cuda", local_rank))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    primes = params.ntt_primes(LIMBS)
    lo, hi = shard.limb_shard(LIMBS, world, rank)
    own = hi - lo                                   # limbs of this rank
    own_polys = own * BATCH
    with c_stdout_to_stderr():
        ctx = vk.Context(local_rank)
    # only this rank's primes: tables are generated on (and for) its GPU
    tables = [vk.NttTables(N, q, params.find_psi(N, q), ctx=ctx)
              for q in primes[lo:hi]]

    host_in = vk.host_alloc(own_polys * N)
    host_out = vk.host_alloc(own_polys * N)
    host_in.array[:] = make_inputs(primes, lo, hi)
    data = ctx.vector(own_polys * N, zero=False)
    work = ctx.vector(own_polys * N, zero=False)
    data.upload(host_in)
    ctx.sync()

    def step_resident():
        ctx.forward_transform_rns(data, work, tables, BATCH)
        ctx.inverse_transform_rns(work, work, tables, BATCH)

    # equal-shard slice of the end-to-end path (the ceiling probes copy these)
    chunk_batch = BATCH // E2E_CHUNKS
    chunk_elems = chunk_batch * own * N

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
    # boost clock; W steps of well under a millisecond are too short for that
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
    lazy0 = ctx.lazy_forwards
    ms_total, launches = timed(step_resident, args.steps, args.warmup)
    lazy_stores = ctx.lazy_forwards - lazy0
    ms_per_step = ms_total / args.steps
    clocks = sampler.window(*timed_window) if rank == 0 else None
    # the sustained figure: the same step for at least SUSTAINED_SECONDS,
    # whatever --steps was, with its own clock and power samples
    sus_steps = max(args.steps, int(SUSTAINED_SECONDS / (ms_per_step * 1e-3)) + 1)
    ms_sus, _ = timed(step_resident, sus_steps, 0)
    sus_clocks = sampler.stop(*timed_window) if rank == 0 else None

    # ---- what was just timed, against the oracle (outside the timed region):
    # the forward transform of one polynomial per limb of the shard, bit for
    # bit, and the round trip of the whole shard
    import oracle
    ctx.forward_transform_rns(data, work, tables, BATCH)
    work.download(host_out)
    ctx.sync()
    checked, parity_ok = 0, True
    for l in range(own):
        b = (lo + l) % BATCH                        # a different entry per limb
        p = b * own + l
        ora = oracle.Tables(N, primes[lo + l], tables[l].w)
        want = oracle.forward(host_in.array[p * N:(p + 1) * N], ora)
        parity_ok = parity_ok and bool(
            np.array_equal(host_out.array[p * N:(p + 1) * N], want))
        checked += 1
    ctx.inverse_transform_rns(work, work, tables, BATCH)
    work.download(host_out)
    ctx.sync()
    ok = parity_ok and bool(np.array_equal(host_out.array, host_in.array))

    # forward-only and inverse-only durations (per direction roofline)
    def fwd_only():
        ctx.forward_transform_rns(data, work, tables, BATCH)

    def inv_only():
        ctx.inverse_transform_rns(work, work, tables, BATCH)

    ms_fwd, _ = timed(fwd_only, args.steps, 1)
    ms_inv, _ = timed(inv_only, args.steps, 1)

    host_out.array[:] = 0
    # (at N > 1 a step is a fraction of the one-GPU step: more of them, so that
    # filling and draining the pipeline weighs the same at every N)
    e2e_steps = max(2, min(args.steps, 20) * min(world, 4))

    # The box's own host <-> device bandwidth, all ranks copying at once,
    # without the library (tools/pcie_probe.py: plain cudaMemcpyAsync).  The
    # pool's hosts are shared and their PCIe throughput moves by tens of
    # percent within seconds, so the probe runs right before and right after
    # the end-to-end segments, and the end-to-end figure is the median of three
    # segments of e2e_steps steps.
    sys.path.insert(0, os.path.join(ROOT, "tools"))

    def probe_ceiling():
        try:
            import pcie_probe
            # at N > 1 for a fixed time, not a fixed amount: every rank copies
            # against all the others for the whole measurement (pcie_probe)
            secs = 0.3 if world > 1 else None
            pcie = pcie_probe.measure(local_rank, mib=32 if secs else 128,
                                      reps=4, barrier=barrier, seconds=secs)
            piped = pcie_probe.measure_pipelined(
                local_rank, chunk_elems * 8, chunks=E2E_CHUNKS,
                steps=e2e_steps, barrier=barrier, seconds=secs)
            vals = [pcie["both_each_GBps"], pcie["h2d_alone_GBps"],
                    pcie["d2h_alone_GBps"], piped["pipelined_each_GBps"]]
            err = None
        except Exception as exc:    # noqa: BLE001 (probe is best effort)
            vals, err = [0.0] * 4, str(exc)[:120]
        # (the slowest rank's share: with equal shards and the clock stopped
        # by the last rank, world x this is what the job can reach)
        slowest = -max_over_ranks(-vals[3])
        return ([sum_over_ranks(v) for v in vals] + [slowest * world, vals[3]],
                err)

    before, err0 = probe_ceiling()

    # End-to-end path.  Who moves what follows the rates just measured
    # (shard.e2e_plan): at N = 1 the whole workload; at N > 1 pairs of ranks,
    # slowest with fastest, share the limbs of their two shards and split the
    # batch entries in proportion to their rates, because with equal shards
    # the job ends with the slowest rank (`e2e_ceiling.equal_shards`).  A
    # rank's part moves in up to E2E_CHUNKS slices of whole batch entries, each
    # slice on its own device vector, two sets of slices used alternately:
    # upload (H2D copy stream), both transforms (compute stream) and download
    # (D2H copy stream) of different slices overlap.
    rates = [0.0] * world
    rates[rank] = before[5]
    rates = [sum_over_ranks(x) for x in rates] if world > 1 else rates
    e2e_limbs, e2e_b0, e2e_b1 = shard.e2e_plan(rates, LIMBS, BATCH)[rank]
    e2e_batches = [b1 - b0 for _, b0, b1 in shard.e2e_plan(rates, LIMBS, BATCH)]
    own_tables = dict(zip(range(lo, hi), tables))
    extra_tables = {l: vk.NttTables(N, primes[l], params.find_psi(N, primes[l]),
                                    ctx=ctx)
                    for l in e2e_limbs if l not in own_tables}
    e2e_tables = [own_tables.get(l) or extra_tables[l] for l in e2e_limbs]
    e2e_sizes = shard.chunk_sizes(e2e_b1 - e2e_b0, E2E_CHUNKS)
    e2e_in = vk.host_alloc((e2e_b1 - e2e_b0) * len(e2e_limbs) * N)
    e2e_out = vk.host_alloc((e2e_b1 - e2e_b0) * len(e2e_limbs) * N)
    e2e_in.array[:] = make_inputs_for(primes, e2e_limbs, e2e_b0, e2e_b1)
    e2e_out.array[:] = 0
    slices = [[ctx.vector(size * len(e2e_limbs) * N, zero=False)
               for size in e2e_sizes] for _ in range(2)]
    e2e_state = {"step": 0}

    def step_e2e():
        cur = slices[e2e_state["step"] & 1]
        e2e_state["step"] += 1
        offset = 0
        for v, size in zip(cur, e2e_sizes):
            v.upload(e2e_in, count=v.length, host_offset=offset)
            ctx.forward_transform_rns(v, v, e2e_tables, size)
            ctx.inverse_transform_rns(v, v, e2e_tables, size)
            v.download(e2e_out, count=v.length, host_offset=offset)
            offset += v.length

    e2e_segments = []
    for seg in range(3):
        ms_seg, _ = timed(step_e2e, e2e_steps, 2 if seg == 0 else 0)
        e2e_segments.append(ms_seg)
    ms_e2e = sorted(e2e_segments)[1]
    ok = ok and bool(np.array_equal(e2e_out.array, e2e_in.array))
    after, err1 = probe_ceiling()
    if err0 or err1 or not before[0] or not after[0]:
        ceiling = {"value": None, "unit": "NTT/s", "error": err0 or err1}
    else:
        both, h2d_alone, d2h_alone, piped_all, piped_equal = [
            (x + y) / 2 for x, y in zip(before[:5], after[:5])]
        per_gbps = 1e9 / (256 * 1024)
        ceiling = {
            "value": both * per_gbps, "unit": "NTT/s",
            "aggregate_both_directions_each_GBps": both,
            "aggregate_h2d_alone_GBps": h2d_alone,
            "aggregate_d2h_alone_GBps": d2h_alone,
            "before_and_after_the_e2e_run": [before[0] * per_gbps,
                                             after[0] * per_gbps],
            "how": "tools/pcie_probe.py: cudaMemcpyAsync of pinned buffers "
                   "on every rank at once, no library (N = 1: 4 copies of 128 "
                   "MiB; N > 1: 32 MiB copies for 0.3 s per measurement, so "
                   "that every rank copies against all the others throughout "
                   "-- a fixed amount per rank lets the better-connected GPUs "
                   "finish early and flatters the sum); one NTT moves 256 KiB "
                   "each way while both directions are busy; mean of a probe "
                   "before and a probe after the e2e segments",
            # the same copies as the e2e path issues (slice size, slices per
            # step, each download behind its upload, two buffer sets), without
            # the library and without kernels
            "pipelined": {
                "value": piped_all * per_gbps, "unit": "NTT/s",
                "aggregate_each_GBps": piped_all,
                "before_and_after_the_e2e_run": [before[3] * per_gbps,
                                                 after[3] * per_gbps],
                "slice_mib": chunk_elems * 8 / 2 ** 20,
                "slices_per_step": E2E_CHUNKS, "steps": e2e_steps},
            # N x the slowest rank's pipelined rate: the limb shards are equal
            # and the e2e clock stops with the last rank, so a box whose GPUs
            # get unequal shares of the host fabric (this pool at N = 8: 8.4
            # GB/s each way for GPUs 0-3, 11.9 for GPUs 4-7) caps the job here
            "equal_shards": {"value": piped_equal * per_gbps, "unit": "NTT/s",
                             "aggregate_each_GBps": piped_equal},
        }

    # the weak-scaling figure: every rank runs the whole one-GPU workload
    weak = None
    if world > 1:
        wtables = [vk.NttTables(N, q, params.find_psi(N, q), ctx=ctx)
                   for q in primes]
        wdata = ctx.vector(POLYS * N, zero=False)
        wwork = ctx.vector(POLYS * N, zero=False)
        ctx.forward_transform_rns(wdata, wdata, wtables, BATCH)   # any residues

        def step_weak():
            ctx.forward_transform_rns(wdata, wwork, wtables, BATCH)
            ctx.inverse_transform_rns(wwork, wwork, wtables, BATCH)

        ms_weak, _ = timed(step_weak, args.steps, max(args.warmup, 3))
        weak = {"value": world * 2 * POLYS / (ms_weak / args.steps * 1e-3),
                "unit": "NTT/s", "scaling": "weak",
                "ms_per_step": ms_weak / args.steps,
                "workload": "every GPU runs all 32 limbs x batch 16"}
        for v in (wdata, wwork):
            v.destroy()
        for t in wtables:
            t.destroy()

    # the same trip through the reference's 18 calls only (one vector per
    # polynomial, pageable host memory): copy_from_host, forward_transform,
    # inverse_transform, map (read the result), unmap.  A bounded sample:
    # LEGACY_POLYS polynomials over this rank's limbs, driven from Python ...
    legacy_polys = 64
    legacy_vecs = [ctx.vector(N, zero=False) for _ in range(legacy_polys)]
    legacy_in = np.array(host_in.array[:legacy_polys * N])   # pageable copy
    legacy_out = np.empty_like(legacy_in)

    def step_legacy():
        for p, v in enumerate(legacy_vecs):
            v.copy_from_host(legacy_in[p * N:(p + 1) * N])
        for p, v in enumerate(legacy_vecs):
            ctx.forward_transform(v, v, tables[p % own])
        for p, v in enumerate(legacy_vecs):
            ctx.inverse_transform(v, v, tables[p % own])
        for p, v in enumerate(legacy_vecs):
            v.read_into(legacy_out[p * N:(p + 1) * N])   # map, copy, unmap

    step_legacy()
    ctx.sync()
    barrier()
    t_legacy = time.perf_counter()
    legacy_steps = 5
    for _ in range(legacy_steps):
        step_legacy()
    ctx.sync()
    t_legacy = max_over_ranks((time.perf_counter() - t_legacy) / legacy_steps)
    ok = ok and bool(np.array_equal(legacy_out, legacy_in))
    legacy_value = world * 2 * legacy_polys / t_legacy
    # ... and from C, as the reference's own callers are (examples/api_e2e.c:
    # the same five calls per vector, no interpreter between them)
    legacy_c = None
    api_e2e = os.path.join(ROOT, "build", "bin", "api_e2e")
    if rank == 0 and os.path.exists(api_e2e):
        try:
            res = subprocess.run(
                [api_e2e, str(LOG2N), str(legacy_polys), "8"],
                env=dict(os.environ, VKHEL_DEVICE=str(local_rank)),
                capture_output=True, text=True, timeout=120)
            for text in res.stdout.splitlines():
                if text.startswith("{"):
                    legacy_c = json.loads(text)
        except (OSError, ValueError, subprocess.SubprocessError):
            legacy_c = None

    peaks = ctx.probe_int_peaks() if rank == 0 else None
    all_ok = max_over_ranks(0.0 if ok else 1.0) == 0.0
    checked_total = int(sum_over_ranks(checked))

    ntts_per_step = 2 * POLYS                       # whole job, every N
    value = ntts_per_step / (ms_per_step * 1e-3)
    sus_value = ntts_per_step / (ms_sus / sus_steps * 1e-3)
    e2e_value = ntts_per_step / (ms_e2e / e2e_steps * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peaks()
        peak *= world
        algo_bytes = 16 * N * ntts_per_step
        achieved = algo_bytes / (ms_per_step * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": "NTT/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": workload_config(world),
            "checks": {
                "all_ranks_exact": all_ok,
                "parity_checked_polys": checked_total,
                "what": "forward transform of one polynomial per limb of "
                        "every shard bit-exact against the CPU oracle; round "
                        "trip of every shard (resident, e2e and reference-API "
                        "paths) returns the input",
                "host_numa_binding": numa,
            },
            "sustained": {
                "value": sus_value, "unit": "NTT/s", "steps": sus_steps,
                "seconds": ms_sus * 1e-3,
                "ms_per_step": ms_sus / sus_steps, "clocks": sus_clocks},
            "e2e": {"value": e2e_value, "unit": "NTT/s",
                    "h2d_bytes_per_step": POLYS * N * 8,
                    "d2h_bytes_per_step": POLYS * N * 8,
                    "steps": e2e_steps,
                    "segments": [ntts_per_step / (m / e2e_steps * 1e-3)
                                 for m in e2e_segments],
                    "value_is": "median of three segments of `steps` steps",
                    "frac_of_ceiling": (e2e_value / ceiling["value"]
                                        if ceiling.get("value") else None),
                    "frac_of_pipelined_ceiling": (
                        e2e_value / ceiling["pipelined"]["value"]
                        if ceiling.get("pipelined") else None),
                    "frac_of_equal_shards_ceiling": (
                        e2e_value / ceiling["equal_shards"]["value"]
                        if ceiling.get("equal_shards") else None),
                    "sharding": ("one GPU: all limbs, all batch entries"
                                 if world == 1 else
                                 "pairs of ranks (slowest with fastest measured "
                                 "host <-> device rate) share the limbs of "
                                 "their two shards and split the 16 batch "
                                 "entries in proportion to their rates"),
                    "batch_entries_per_rank": e2e_batches,
                    "rates_GBps_per_rank": rates,
                    "path": "per slice of whole batch entries (up to 4 slices "
                            "per step): vkhel_vector_upload "
                            "(pinned) -> forward_transform_rns -> "
                            "inverse_transform_rns -> vkhel_vector_download;"
                            " copies on the context's H2D/D2H streams "
                            "overlap with each other and with compute"},
            "e2e_ceiling": ceiling,
            "e2e_reference_api": {
                "value": (legacy_c["ntt_per_s"]
                          if legacy_c and legacy_c.get("round_trip_exact")
                          and world == 1 else legacy_value),
                "unit": "NTT/s",
                "sample": "%d polynomials per rank, one vkhel_vector each"
                          % legacy_polys,
                "from_c": legacy_c, "from_python": legacy_value,
                "path": "the reference's 18 entry points only, pageable host "
                        "memory: vkhel_vector_copy_from_host -> "
                        "vkhel_vector_forward_transform -> "
                        "vkhel_vector_inverse_transform -> vkhel_vector_map "
                        "+ copy out + vkhel_vector_unmap (which writes the "
                        "vector back, as the reference does); host wall "
                        "clock; value = the C caller (examples/api_e2e.c) at "
                        "N = 1, the Python loop over all ranks otherwise"},
            # whole job: every rank launches the same kernels on its shard
            "gpu_launches": launches * world,
            # forward transforms (warm-up and timed steps of this rank) whose
            # results only the in-place inverse transform after them read, and
            # which therefore stored residues in [0,3q) instead of [0,q)
            # (vector.cu, held forward transform; $VKHEL_LAZY_FORWARD=0: none).
            # `checks` reads the forward transform through the API: canonical.
            "lazy_forward_stores": lazy_stores,
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic if world == 1 else None,
                "traffic_source": traffic_src + " (two passes per "
                                  "transform, the intermediate stays in L2: "
                                  "limb slices on two streams; N = 1)",
                "peak_source": peak_src + (" x %d GPUs" % world
                                           if world > 1 else ""),
                "kernel": "whole step (forward + inverse NTT, all passes)",
                "algorithmic_bytes_per_step": algo_bytes,
                "forward_ms": ms_fwd / args.steps,
                "inverse_ms": ms_inv / args.steps,
                "forward_GBps": 16 * N * POLYS / (ms_fwd / args.steps * 1e-3) / 1e9,
                "inverse_GBps": 16 * N * POLYS / (ms_inv / args.steps * 1e-3) / 1e9,
            },
        }
        if weak:
            line["weak"] = weak
        # the binding roofline: integer (fmaheavy-pipe) butterfly rate per GPU
        bfly_per_step = ntts_per_step * (N // 2) * LOG2N
        bfly_rate = bfly_per_step / (ms_per_step * 1e-3) / 1e9 / world
        line["issue_roofline"] = issue_roofline(
            peaks, bfly_rate, clocks.get("sm_mhz") if clocks else None)
        line["issue_roofline"]["per"] = "GPU"
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(12.0)
        emit_line(line)

    timer.destroy()
    for v in [data, work] + slices[0] + slices[1] + legacy_vecs:
        v.destroy()
    for t in tables:
        t.destroy()
    host_in.free()
    host_out.free()
    e2e_in.free()
    e2e_out.free()
    for t in extra_tables.values():
        t.destroy()
    ctx.destroy()
    if dist is not None:
        dist.destroy_process_group()
    if not all_ok:
        raise SystemExit("bench.py: parity / round trip mismatch -- result invalid")


def emit_line(line):
    """the one JSON line, on the process's real stdout"""
    os.write(REAL_STDOUT, (json.dumps(line) + "\n").encode())


REAL_STDOUT = 1


def main():
    # Libraries print on stdout behind our back (the context's device banner,
    # NCCL's version line at communicator creation, buffered until exit): keep
    # the real stdout for the one JSON line and point fd 1 at stderr for
    # everything else.
    global REAL_STDOUT
    sys.stdout.flush()
    REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
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
