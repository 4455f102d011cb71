#!/usr/bin/env python3
"""BASELINE configs[2], [3] and [4] across the GPUs of one box, one process per
GPU (run under torchrun for N > 1, plain python for N = 1):

  configs[2]  n = 2^16, 32 RNS limbs x batch 16, LIMB-sharded: rank r owns a
              contiguous range of 32/N limbs (and uploads only their tables),
              forward + inverse;
  configs[3]  polynomial product c = INTT(NTT(a) (*) NTT(b)), n = 2^16, batch
              1024, batch-sharded 1024/N per GPU, fused kernel;
  configs[4]  degree sweep n = 2^10 .. 2^17 at 2^27 coefficients in total
              (strong scaling: 2^27 / N per GPU), forward + inverse.

No collective on the data path: torch.distributed carries the
barrier and the max-over-ranks of the CUDA-event times only.  Every rank checks
its shard (round trip; product against the separate transforms on a sample).
One JSON object per measurement on rank 0, whole-job throughput."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkhel_b200 as vk  # noqa: E402
from vkhel_b200 import params, shard  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    device = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        device = torch.device("cuda", local_rank)
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if dist is not None:
            dist.barrier()

    def slowest(ms):
        return shard.max_over_ranks(ms, dist, device)

    def all_ok(flag):
        return slowest(0.0 if flag else 1.0) == 0.0

    ctx = vk.Context(local_rank)
    timer = ctx.timer()

    def time_ms(fn, iters, warmup=3):
        for _ in range(warmup):
            fn()
        ctx.sync()
        barrier()
        timer.start()
        for _ in range(iters):
            fn()
        timer.stop()
        ms = timer.elapsed_ms() / iters
        barrier()
        return slowest(ms)

    def emit(obj):
        if rank == 0:
            obj["n_gpus"] = world
            print(json.dumps(obj), flush=True)

    rng = np.random.default_rng(rank)

    # ---- configs[2]: limb-sharded RNS transform -------------------------------------
    n, limbs, batch = 1 << 16, 32, 16
    lo, hi = shard.limb_shard(limbs, world, rank)
    primes = params.ntt_primes(limbs)[lo:hi]
    tabs = [vk.NttTables(n, q, params.find_psi(n, q)) for q in primes]
    own = hi - lo
    host = np.concatenate([
        rng.integers(0, 1 << 62, n, dtype=np.uint64) % np.uint64(primes[p % own])
        for p in range(batch * own)])
    a = ctx.from_host(host)
    w = ctx.vector(host.size, zero=False)

    def step_rns():
        ctx.forward_transform_rns(a, w, tabs, batch)
        ctx.inverse_transform_rns(w, w, tabs, batch)

    ms = time_ms(step_rns, 50)
    ok = all_ok(np.array_equal(w.to_host(), host))
    emit({"config": "configs[2]: n=2^16, 32 limbs x batch 16, limb-sharded "
          "(%d limbs per GPU), forward+inverse" % own, "ms_per_step": ms,
          "ntt_per_s": 2 * limbs * batch / ms * 1e3, "round_trip_exact": ok})
    if dist is not None:
        # NOT on the hot path: the caller asks for the whole result on every
        # GPU.  NCCL all-gather over NVLink straight out of / into the
        # library's device vectors (zero-copy views), reported on its own.
        import torch
        ctx.forward_transform_rns(a, w, tabs, batch)
        full = ctx.vector(world * host.size, zero=False)
        ctx.sync()
        src = torch.as_tensor(w, device=device)
        dst = torch.as_tensor(full, device=device)
        ev0, ev1 = torch.cuda.Event(True), torch.cuda.Event(True)
        for _ in range(3):
            dist.all_gather_into_tensor(dst, src)
        torch.cuda.synchronize()
        barrier()
        ev0.record()
        for _ in range(10):
            dist.all_gather_into_tensor(dst, src)
        ev1.record()
        torch.cuda.synchronize()
        gms = slowest(ev0.elapsed_time(ev1) / 10)
        got = full.to_host()[rank * host.size:(rank + 1) * host.size]
        ok = all_ok(np.array_equal(got, w.to_host()))
        emit({"config": "configs[2] result gather (optional, off the hot "
              "path): NCCL all-gather of the limb shards, rank-major",
              "ms": gms, "bytes_per_gpu": 8 * host.size,
              "bytes_gathered": 8 * host.size * world,
              "algbw_GBps": 8 * host.size * world / gms / 1e6,
              "own_shard_intact": ok})
        full.destroy()
    if "--only-config2" in sys.argv:
        ctx.destroy()
        if dist is not None:
            dist.destroy_process_group()
        return
    a.destroy()
    w.destroy()
    for t in tabs:
        t.destroy()

    # ---- configs[3]: batch-sharded polynomial product -------------------------------
    q = params.P0
    total_batch = 1024
    b0, b1 = shard.batch_shard(total_batch, world, rank)
    mine = b1 - b0
    t16 = vk.NttTables(n, q, params.find_psi(n, q))
    ha = rng.integers(0, 1 << 62, mine * n, dtype=np.uint64) % np.uint64(q)
    hb = rng.integers(0, 1 << 62, mine * n, dtype=np.uint64) % np.uint64(q)
    va, vb = ctx.from_host(ha), ctx.from_host(hb)
    vc = ctx.vector(mine * n, zero=False)
    ms = time_ms(lambda: ctx.polymul_rns(va, vb, vc, [t16], mine), 10)
    # the fused product against forward, forward, elemmul, inverse on 2 polys
    sa, sb = ctx.from_host(ha[:2 * n]), ctx.from_host(hb[:2 * n])
    ctx.forward_transform_batch(sa, sa, t16, 2)
    ctx.forward_transform_batch(sb, sb, t16, 2)
    ctx.elemmul(sa, sb, sa, q)
    ctx.inverse_transform_batch(sa, sa, t16, 2)
    ok = all_ok(np.array_equal(vc.to_host()[:2 * n], sa.to_host()))
    emit({"config": "configs[3]: polymul n=2^16, batch 1024, batch-sharded "
          "(%d per GPU)" % mine, "ms": ms,
          "polymul_per_s": total_batch / ms * 1e3,
          "matches_separate_calls": ok})
    for v in (va, vb, vc, sa, sb):
        v.destroy()
    t16.destroy()

    # ---- configs[4]: degree sweep at 2^27 coefficients in total ---------------------
    total = (1 << 27) // world
    host = rng.integers(0, 1 << 62, total, dtype=np.uint64) % np.uint64(q)
    a = ctx.from_host(host)
    w = ctx.vector(total, zero=False)
    for log2n in range(10, 18):
        nn = 1 << log2n
        polys = total >> log2n
        t = vk.NttTables(nn, q, params.find_psi(nn, q))
        fwd = time_ms(lambda: ctx.forward_transform_batch(a, w, t, polys), 10)
        inv = time_ms(lambda: ctx.inverse_transform_batch(w, w, t, polys), 10)
        ctx.forward_transform_batch(a, w, t, polys)
        ctx.inverse_transform_batch(w, w, t, polys)
        ok = all_ok(np.array_equal(w.to_host(), host))
        bfly = world * polys * (nn // 2) * log2n
        emit({"config": "configs[4]: sweep, 2^27 coefficients in total",
              "log2n": log2n, "polys_per_gpu": polys, "fwd_ms": fwd,
              "inv_ms": inv, "fwd_ntt_per_s": world * polys / fwd * 1e3,
              "inv_ntt_per_s": world * polys / inv * 1e3,
              "fwd_Gbfly_per_s": bfly / fwd / 1e6,
              "inv_Gbfly_per_s": bfly / inv / 1e6, "round_trip_exact": ok})
        t.destroy()
    a.destroy()
    w.destroy()
    ctx.destroy()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
