#!/usr/bin/env python
"""Device-resident time of the KEY_8B (16-byte tuple) join and sort: python tools/time_key8b.py [n]"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
smj.check(lib.smjdev_init(0))
rng = np.random.default_rng(1)
def rel(seed):
    r = np.empty((n, 2), dtype=np.int64)
    r[:, 1] = np.random.default_rng(seed).permutation(n) + 1   # key
    r[:, 0] = np.arange(n)                                      # payload
    return r
r, s = rel(1), rel(2)
dr, ds = smj.DeviceBuffer(2 * n), smj.DeviceBuffer(2 * n)
smj.check(lib.smjdev_memcpy_h2d(dr.ptr, r.ctypes.data, r.nbytes)); smj.check(lib.smjdev_memcpy_h2d(ds.ptr, s.ctypes.data, s.nbytes))
lib.smjdev_sync()
e0, e1 = C.c_void_p(), C.c_void_p()
lib.smjdev_event_create(C.byref(e0)); lib.smjdev_event_create(C.byref(e1))
st = smj.smj_join_stats_t()
lib.smjdev_join_key8b.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(smj.smj_join_stats_t)]
best = 1e9
for rep in range(3):
    lib.smjdev_sync(); lib.smjdev_event_record(e0)
    smj.check(lib.smjdev_join_key8b(dr.ptr, n, ds.ptr, n, 0, C.byref(st)))
    lib.smjdev_event_record(e1)
    ms = C.c_float(); lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms))
    if rep: best = min(best, ms.value)
lib.smjdev_profile_begin()
smj.check(lib.smjdev_join_key8b(dr.ptr, n, ds.ptr, n, 0, C.byref(st)))
buf = C.create_string_buffer(1 << 16); lib.smjdev_profile_end(buf, len(buf))
prof = {}
for line in buf.value.decode().strip().splitlines():
    nm, c, t, items = line.rsplit(";", 3)
    prof[nm.strip("() ")[:26] + " x" + c] = round(float(t), 3)
print(f"KEY_8B join {n} x {n} 16-byte tuples: {best:.3f} ms = {2 * n / best / 1e3:.0f} M tuples/s, matches {st.matches}")
print("   ", prof)
