#!/usr/bin/env python
"""Operator micro-benchmarks on one B200 for the other BASELINE.json configurations
(the reference's bench_sort / bench_partitioning mains, src/bench/sortbench.c:85-202 and
partitioningbench.c:63-210, and the Zipf join): device-resident data, CUDA events on the
library's stream, best of `--reps`.  Prints one JSON object.  Not the driver's bench (bench.py)."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib


def timed(fn, reps):
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.smjdev_event_create(C.byref(e0))
    lib.smjdev_event_create(C.byref(e1))
    best = None
    for _ in range(reps):
        lib.smjdev_sync()
        lib.smjdev_event_record(e0)
        fn()
        lib.smjdev_event_record(e1)
        ms = C.c_float()
        smj.check(lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms)))
        best = ms.value if best is None else min(best, ms.value)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    n, out = args.n, {"n": args.n}
    smj.check(lib.smjdev_init(0))
    src, a, b = smj.DeviceBuffer(n), smj.DeviceBuffer(n), smj.DeviceBuffer(n + 8 * (1 << 14))
    smj.check(lib.smjdev_gen_unique(src.ptr, n, 0, n, 12345))
    res = C.c_void_p()

    for name, key_only in (("sort_whole_word", 0), ("sort_key_only", 1)):
        def run():
            smj.check(lib.smjdev_memcpy_d2d(a.ptr, src.ptr, n * 8))
            smj.check(lib.smjdev_sort(a.ptr, b.ptr, n, key_only, 1, C.byref(res)))
        copy_ms = timed(lambda: smj.check(lib.smjdev_memcpy_d2d(a.ptr, src.ptr, n * 8)), args.reps)
        ms = timed(run, args.reps) - copy_ms
        out[name] = {"ms": ms, "M_items_per_s": n / ms / 1e3, "GBps_at_16B_per_item": 16 * n / ms / 1e6}

    hist, offs = smj.DeviceBuffer(1 << 14), smj.DeviceBuffer(1 << 14)
    out["partition_relation_optimized"] = {}
    for nbits in range(4, 15):
        ms = timed(lambda: smj.check(lib.smjdev_partition(src.ptr, b.ptr, a.ptr, n, nbits, 0, 1, 1, hist.ptr, offs.ptr)),
                   args.reps)
        out["partition_relation_optimized"][str(nbits)] = {"ms": ms, "M_tuples_per_s": n / ms / 1e3,
                                                          "GBps_at_16B_per_tuple": 16 * n / ms / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
