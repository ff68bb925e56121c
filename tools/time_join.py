#!/usr/bin/env python
"""Quick timing of the device-resident m-way join (bench.py's workload) with the per-kernel
breakdown: python tools/time_join.py [tuples per relation] [steps]"""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = smj.lib
smj.check(lib.smjdev_init(0), "init")
dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, 12345))
smj.check(lib.smjdev_gen_unique(ds.ptr, n, 0, n, 54321))
for _ in range(3):
    st = smj.device_join(smj.ALGO_MWAY, dr, ds)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.smjdev_event_create(C.byref(e0)); lib.smjdev_event_create(C.byref(e1))
lib.smjdev_sync(); lib.smjdev_event_record(e0)
for _ in range(steps):
    st = smj.device_join(smj.ALGO_MWAY, dr, ds)
lib.smjdev_event_record(e1)
ms = C.c_float(); lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms))
lib.smjdev_profile_begin()
for _ in range(steps):
    smj.device_join(smj.ALGO_MWAY, dr, ds)
buf = C.create_string_buffer(1 << 16); lib.smjdev_profile_end(buf, len(buf))
prof = {}
for line in buf.value.decode().strip().splitlines():
    name, cnt, t, items = line.rsplit(";", 3)
    prof[name.strip("() ")[:26]] = round(float(t) / steps, 3)
print(f"ms/step {ms.value / steps:.3f} matches {st.matches} path {st.path} repeats {st.repeats}", prof)
