#!/usr/bin/env python3
"""DRAM traffic of one bench step from an ncu launch list taken with
    ncu --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        -c <N> --csv --log-file <csv> python bench.py --steps 2 --warmup 3 --no-cpu
(one pass per kernel, no replay, caches left alone: with the sliced two-pass
transforms what the second pass reads is still in L2, and flushing before
every kernel -- ncu's default -- would hide exactly that).

The first launches of bench.py are its clock spin-up: back-to-back resident
steps, i.e. the timed workload itself.  `--launches-per-step L --skip S --steps
K` averages K steps of L launches after S launches.

    python tools/ncu_traffic.py gpurun_out/traffic.csv --launches-per-step 32 \
        > profiles/r01_ntt_traffic.json
"""
import argparse
import collections
import csv
import json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--launches-per-step", type=int, required=True)
    ap.add_argument("--skip", type=int, default=None)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--algorithmic-bytes", type=int, default=1 << 30)
    args = ap.parse_args()
    launches = collections.OrderedDict()
    with open(args.csv) as f:
        for row in csv.reader(f):
            if len(row) < 15 or not row[0].isdigit():
                continue
            ent = launches.setdefault(int(row[0]), {"name": row[4]})
            value = float(row[14].replace(",", ""))
            unit = row[13]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
                     "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
            ent[row[12]] = value * scale
    ids = sorted(launches)
    lps = args.launches_per_step
    skip = args.skip if args.skip is not None else 2 * lps
    sel = ids[skip:skip + lps * args.steps]
    assert len(sel) == lps * args.steps, "launch list too short"
    per_kernel = collections.OrderedDict()
    total_r = total_w = total_ns = 0.0
    for i in sel:
        ent = launches[i]
        name = ent["name"].split("(")[0].replace("void ", "")
        k = per_kernel.setdefault(name, {"launches": 0, "read": 0.0,
                                         "write": 0.0, "ns": 0.0})
        k["launches"] += 1
        k["read"] += ent.get("dram__bytes_read.sum", 0.0)
        k["write"] += ent.get("dram__bytes_write.sum", 0.0)
        k["ns"] += ent.get("gpu__time_duration.sum", 0.0)
        total_r += ent.get("dram__bytes_read.sum", 0.0)
        total_w += ent.get("dram__bytes_write.sum", 0.0)
        total_ns += ent.get("gpu__time_duration.sum", 0.0)
    steps = args.steps
    out = {
        "source": args.csv,
        "how": "ncu --cache-control none, one pass per kernel; averaged over "
               "%d steps of %d launches after %d launches" % (steps, lps, skip),
        "bytes_per_step": int((total_r + total_w) / steps),
        "read_bytes_per_step": int(total_r / steps),
        "write_bytes_per_step": int(total_w / steps),
        "algorithmic_bytes_per_step": args.algorithmic_bytes,
        "traffic_over_algorithmic": (total_r + total_w) / steps
        / args.algorithmic_bytes,
        "serialized_us_per_step": total_ns / steps / 1e3,
        "per_kernel_per_step": {
            name: {"launches": k["launches"] / steps,
                   "read_MB": k["read"] / steps / 1e6,
                   "write_MB": k["write"] / steps / 1e6,
                   "us": k["ns"] / steps / 1e3,
                   "share_of_step": k["ns"] / total_ns}
            for name, k in per_kernel.items()},
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
