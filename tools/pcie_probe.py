#!/usr/bin/env python3
"""Host <-> device copy bandwidth WITHOUT the library: plain cudaHostAlloc /
cudaMalloc / cudaMemcpyAsync through cuda-python, each direction alone and both
at once, on one GPU or on N GPUs of the box at the same time.

This is the ceiling of bench.py's `e2e` figure (one step moves 512 KiB in and
512 KiB out per two NTTs): if the aggregate does not grow with the number of
GPUs that copy concurrently, the box's host <-> device fabric is the limit and
not the library's transfer path.

    python tools/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 \\
        --master-addr 127.0.0.1 --master-port 29511 tools/pcie_probe.py

Prints one JSON line on rank 0; bench.py calls measure() for its `e2e_ceiling`.
"""
import json
import os
import sys

MIB = 1 << 20


def _check(res):
    err, rest = res[0], res[1:]
    if int(err) != 0:
        raise RuntimeError("CUDA runtime error %s" % err)
    return rest[0] if len(rest) == 1 else rest


def measure(device, mib=256, reps=6, barrier=None, write_combined=False,
            seconds=None):
    """GB/s of pinned-host <-> device copies on `device`: h2d alone, d2h alone,
    both at once (per direction).  `barrier()` (optional) is called right before
    each timed region so that several ranks copy concurrently.

    `seconds` (None: `reps` copies): copy for that long instead of a fixed
    number of times.  With several ranks a fixed amount of work flatters the
    aggregate -- the GPUs with the larger share of the host fabric finish early
    and the others then speed up -- while a fixed time keeps every rank copying
    against all the others for the whole measurement: the steady-state figure,
    which is what a job that keeps all GPUs busy can get."""
    from cuda.bindings import runtime as rt
    _check(rt.cudaSetDevice(device))
    nbytes = mib * MIB
    up_flags = rt.cudaHostAllocWriteCombined if write_combined \
        else rt.cudaHostAllocDefault
    hin = _check(rt.cudaHostAlloc(nbytes, up_flags))
    hout = _check(rt.cudaHostAlloc(nbytes, rt.cudaHostAllocDefault))
    da = _check(rt.cudaMalloc(nbytes))
    db = _check(rt.cudaMalloc(nbytes))
    s_up = _check(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
    s_down = _check(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
    e0, e1, ej = (_check(rt.cudaEventCreate()) for _ in range(3))
    h2d = rt.cudaMemcpyKind.cudaMemcpyHostToDevice
    d2h = rt.cudaMemcpyKind.cudaMemcpyDeviceToHost
    _check(rt.cudaMemsetAsync(da, 1, nbytes, s_up))
    _check(rt.cudaMemsetAsync(db, 2, nbytes, s_up))
    _check(rt.cudaStreamSynchronize(s_up))

    def once(up, down):
        if up:
            _check(rt.cudaMemcpyAsync(da, hin, nbytes, h2d, s_up))
        if down:
            _check(rt.cudaMemcpyAsync(hout, db, nbytes, d2h, s_down))

    def timed(up, down):
        once(up, down)
        _check(rt.cudaDeviceSynchronize())
        if barrier is not None:
            barrier()
        if seconds is not None:
            # one copy (pair) at a time until the time is up; both directions
            # of a pair run concurrently and the pair ends with the slower one
            import time
            t0 = time.perf_counter()
            done = 0
            while True:
                once(up, down)
                _check(rt.cudaStreamSynchronize(s_up))
                _check(rt.cudaStreamSynchronize(s_down))
                done += 1
                elapsed = time.perf_counter() - t0
                if elapsed >= seconds:
                    return nbytes * done / elapsed / 1e9
        # both streams start behind e0; e1 follows the end of both
        _check(rt.cudaEventRecord(e0, s_up))
        _check(rt.cudaStreamWaitEvent(s_down, e0, 0))
        for _ in range(reps):
            once(up, down)
        _check(rt.cudaEventRecord(ej, s_down))
        _check(rt.cudaStreamWaitEvent(s_up, ej, 0))
        _check(rt.cudaEventRecord(e1, s_up))
        _check(rt.cudaEventSynchronize(e1))
        ms = _check(rt.cudaEventElapsedTime(e0, e1))
        return nbytes * reps / (ms * 1e-3) / 1e9

    out = {
        "device": device, "mib_per_copy": mib, "reps": reps,
        "h2d_alone_GBps": timed(True, False),
        "d2h_alone_GBps": timed(False, True),
        "both_each_GBps": timed(True, True),
        "write_combined_upload_buffer": bool(write_combined),
    }
    for ev in (e0, e1, ej):
        _check(rt.cudaEventDestroy(ev))
    _check(rt.cudaStreamDestroy(s_up))
    _check(rt.cudaStreamDestroy(s_down))
    _check(rt.cudaFree(da))
    _check(rt.cudaFree(db))
    _check(rt.cudaFreeHost(hin))
    _check(rt.cudaFreeHost(hout))
    return out


def measure_pipelined(device, chunk_bytes, chunks=4, steps=20, barrier=None,
                      seconds=None):
    """GB/s per direction of the copy pattern of bench.py's e2e path WITHOUT
    the library and without kernels: per step `chunks` slices of `chunk_bytes`,
    each uploaded into its own device buffer and downloaded again as soon as
    its upload has ended; two sets of device buffers used alternately, an
    upload into a buffer waits for the download that last read it.  The figure
    a pipeline with these dependencies can reach, as opposed to two independent
    streams of 128 MiB copies (measure).  `seconds`: steps for that long
    instead of `steps` of them (see measure)."""
    from cuda.bindings import runtime as rt
    _check(rt.cudaSetDevice(device))
    total = chunk_bytes * chunks
    hin = _check(rt.cudaHostAlloc(total, rt.cudaHostAllocDefault))
    hout = _check(rt.cudaHostAlloc(total, rt.cudaHostAllocDefault))
    dev = [[_check(rt.cudaMalloc(chunk_bytes)) for _ in range(chunks)]
           for _ in range(2)]
    s_up = _check(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
    s_down = _check(rt.cudaStreamCreateWithFlags(rt.cudaStreamNonBlocking))
    up_done = [[_check(rt.cudaEventCreateWithFlags(rt.cudaEventDisableTiming))
                for _ in range(chunks)] for _ in range(2)]
    down_done = [[_check(rt.cudaEventCreateWithFlags(rt.cudaEventDisableTiming))
                  for _ in range(chunks)] for _ in range(2)]
    e0, e1, ej = (_check(rt.cudaEventCreate()) for _ in range(3))
    h2d = rt.cudaMemcpyKind.cudaMemcpyHostToDevice
    d2h = rt.cudaMemcpyKind.cudaMemcpyDeviceToHost

    def step(k, first):
        cur = k & 1
        for c in range(chunks):
            if not first:
                _check(rt.cudaStreamWaitEvent(s_up, down_done[cur][c], 0))
            _check(rt.cudaMemcpyAsync(dev[cur][c], hin + c * chunk_bytes,
                                      chunk_bytes, h2d, s_up))
            _check(rt.cudaEventRecord(up_done[cur][c], s_up))
            _check(rt.cudaStreamWaitEvent(s_down, up_done[cur][c], 0))
            _check(rt.cudaMemcpyAsync(hout + c * chunk_bytes, dev[cur][c],
                                      chunk_bytes, d2h, s_down))
            _check(rt.cudaEventRecord(down_done[cur][c], s_down))

    for k in range(2):                       # warm-up, fills both sets
        step(k, True)
    _check(rt.cudaDeviceSynchronize())
    if barrier is not None:
        barrier()
    _check(rt.cudaEventRecord(e0, s_up))
    _check(rt.cudaStreamWaitEvent(s_down, e0, 0))
    if seconds is None:
        for k in range(steps):
            step(k, False)
    else:
        # keep two steps enqueued (as the double-buffered e2e loop does) and
        # stop when the time is up
        import time
        t0 = time.perf_counter()
        steps = 0
        while True:
            step(steps, False)
            steps += 1
            if steps >= 2:
                # wait for the step before the one just enqueued
                _check(rt.cudaEventSynchronize(
                    down_done[(steps - 2) & 1][chunks - 1]))
                if time.perf_counter() - t0 >= seconds:
                    break
    _check(rt.cudaEventRecord(ej, s_down))
    _check(rt.cudaStreamWaitEvent(s_up, ej, 0))
    _check(rt.cudaEventRecord(e1, s_up))
    _check(rt.cudaEventSynchronize(e1))
    ms = _check(rt.cudaEventElapsedTime(e0, e1))
    for ev in [e0, e1, ej] + sum(up_done, []) + sum(down_done, []):
        _check(rt.cudaEventDestroy(ev))
    _check(rt.cudaStreamDestroy(s_up))
    _check(rt.cudaStreamDestroy(s_down))
    for buf in sum(dev, []):
        _check(rt.cudaFree(buf))
    _check(rt.cudaFreeHost(hin))
    _check(rt.cudaFreeHost(hout))
    return {"pipelined_each_GBps": total * steps / (ms * 1e-3) / 1e9,
            "chunk_mib": chunk_bytes / MIB, "chunks_per_step": chunks,
            "steps": steps}


def summarise(per_rank):
    """aggregate over the ranks that copied concurrently; the e2e ceiling in
    NTT/s at n = 2^16: 512 KiB in and 512 KiB out per NTT pair, i.e. 256 KiB
    each way per NTT while both directions are active"""
    total = {k: sum(r[k] for r in per_rank)
             for k in ("h2d_alone_GBps", "d2h_alone_GBps", "both_each_GBps",
                       "pipelined_each_GBps") if k in per_rank[0]}
    total["n_gpus"] = len(per_rank)
    total["e2e_ceiling_ntt_per_s"] = total["both_each_GBps"] * 1e9 / (256 * 1024)
    return total


def main():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wc = "--write-combined" in sys.argv
    barrier, dist = None, None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        barrier = dist.barrier
    secs = 0.4 if world > 1 else None
    mine = measure(local_rank, barrier=barrier, write_combined=wc,
                   mib=32 if secs else 256, seconds=secs)
    # the e2e path's slices at this number of GPUs (bench.py: 32/world limbs x
    # 4 batch entries of 512 KiB per slice, 4 slices per step)
    mine.update(measure_pipelined(local_rank, (32 // world) * 4 * 512 * 1024,
                                  barrier=barrier, seconds=secs))
    mine["seconds_per_measurement"] = secs
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        dist.destroy_process_group()
    else:
        gathered = [mine]
    if rank == 0:
        print(json.dumps({"aggregate": summarise(gathered),
                          "per_rank": gathered}))


if __name__ == "__main__":
    main()
