#!/usr/bin/env python
"""Device-resident time of the three join algorithms (m-way, m-pass, mpsm) with the per-kernel
breakdown: python tools/time_algos.py [tuples per relation] [runs]"""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128_000_000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
smj.check(lib.smjdev_init(0))
dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, 12345))
smj.check(lib.smjdev_gen_unique(ds.ptr, n, 0, n, 54321))
e0, e1 = C.c_void_p(), C.c_void_p()
lib.smjdev_event_create(C.byref(e0)); lib.smjdev_event_create(C.byref(e1))
for name, algo in (("m-way", smj.ALGO_MWAY), ("m-pass", smj.ALGO_MPASS), ("mpsm", smj.ALGO_MPSM)):
    st = smj.smj_join_stats_t()
    best = 1e9
    for rep in range(3):
        lib.smjdev_sync(); lib.smjdev_event_record(e0)
        smj.check(lib.smjdev_join(algo, dr.ptr, n, ds.ptr, n, runs, C.byref(st)))
        lib.smjdev_event_record(e1)
        ms = C.c_float(); lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms))
        if rep:
            best = min(best, ms.value)
    lib.smjdev_profile_begin()
    smj.check(lib.smjdev_join(algo, dr.ptr, n, ds.ptr, n, runs, C.byref(st)))
    buf = C.create_string_buffer(1 << 16); lib.smjdev_profile_end(buf, len(buf))
    prof = {}
    for line in buf.value.decode().strip().splitlines():
        nm, cnt, t, items = line.rsplit(";", 3)
        prof[nm.strip("() ")[:28] + " x" + cnt] = round(float(t), 3)
    print(f"{name:7s} {runs} runs: {best:.3f} ms = {2 * n / best / 1e3:.0f} M tuples/s, matches {st.matches}, launches {st.launches}")
    print("   ", prof)
