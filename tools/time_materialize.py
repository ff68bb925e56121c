#!/usr/bin/env python
"""Device-resident time of the materialising join (smjdev_join_materialize): python tools/time_materialize.py [n]"""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128_000_000
smj.check(lib.smjdev_init(0))
dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, 12345))
smj.check(lib.smjdev_gen_unique(ds.ptr, n, 0, n, 54321))
out = smj.DeviceBuffer(n)
cnt = C.c_uint64(0)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.smjdev_event_create(C.byref(e0)); lib.smjdev_event_create(C.byref(e1))
best = 1e9
for rep in range(3):
    lib.smjdev_sync(); lib.smjdev_event_record(e0)
    smj.check(lib.smjdev_join_materialize(dr.ptr, n, ds.ptr, n, 0, out.ptr, n, C.byref(cnt)))
    lib.smjdev_event_record(e1)
    ms = C.c_float(); lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms))
    if rep:
        best = min(best, ms.value)
lib.smjdev_profile_begin()
smj.check(lib.smjdev_join_materialize(dr.ptr, n, ds.ptr, n, 0, out.ptr, n, C.byref(cnt)))
buf = C.create_string_buffer(1 << 16); lib.smjdev_profile_end(buf, len(buf))
prof = {}
for line in buf.value.decode().strip().splitlines():
    nm, c, t, items = line.rsplit(";", 3)
    prof[nm.strip("() ")[:26] + " x" + c] = round(float(t), 3)
print(f"materialising join {n} x {n}: {best:.3f} ms, {cnt.value} tuples emitted")
print("   ", prof)
