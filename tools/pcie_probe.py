#!/usr/bin/env python3
"""Host <-> device copy bandwidth through the library's own transfer calls
(vkhel_vector_upload / vkhel_vector_download on the context's copy streams):
each direction alone and both at once.  This is the ceiling of bench.py's
`e2e` figure: one step moves 256 MiB in and 256 MiB out for 1024 NTTs.

    python tools/pcie_probe.py          # prints one JSON line
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import vkhel_b200 as vk  # noqa: E402

MIB = 1 << 20


def main():
    ctx = vk.Context(0)
    count = 256 * MIB // 8
    hin, hout = vk.host_alloc(count), vk.host_alloc(count)
    hin.array[:] = 1
    a, b = ctx.vector(count, zero=False), ctx.vector(count, zero=False)
    timer = ctx.timer()
    reps = 8

    def timed(fn):
        fn()
        ctx.sync()
        timer.start()
        for _ in range(reps):
            fn()
        timer.stop()
        return timer.elapsed_ms() / reps

    ms_up = timed(lambda: a.upload(hin))
    ms_down = timed(lambda: b.download(hout))

    def both():
        a.upload(hin)
        b.download(hout)

    ms_both = timed(both)
    gb = count * 8 / 1e9
    print(json.dumps({
        "bytes_per_copy": count * 8,
        "h2d_alone_GBps": gb / (ms_up * 1e-3),
        "d2h_alone_GBps": gb / (ms_down * 1e-3),
        "both_each_GBps": gb / (ms_both * 1e-3),
        "e2e_ceiling_ntt_per_s": 1024 / (ms_both * 1e-3),
    }))
    timer.destroy()
    a.destroy(), b.destroy()
    hin.free(), hout.free()
    ctx.destroy()


if __name__ == "__main__":
    main()
