#!/usr/bin/env python
"""Timing of the merge-join scan (smjdev_merge_join) on sorted relations:
python tools/time_mj.py [tuples per relation] [S multiple]"""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128_000_000
mult = int(sys.argv[2]) if len(sys.argv) > 2 else 1
smj.check(lib.smjdev_init(0))
r, s = smj.DeviceBuffer(n), smj.DeviceBuffer(n * mult)
a, b = smj.DeviceBuffer(n * mult), smj.DeviceBuffer(n * mult)
smj.check(lib.smjdev_gen_unique(r.ptr, n, 0, n, 12345))
if mult == 1:
    smj.check(lib.smjdev_gen_unique(s.ptr, n, 0, n, 54321))
else:
    smj.check(lib.smjdev_gen_fk(s.ptr, 0, n * mult, n, 54321))
res = C.c_void_p()
sr, ss = smj.DeviceBuffer(n), smj.DeviceBuffer(n * mult)
for x, dst, m in ((r, sr, n), (s, ss, n * mult)):
    smj.check(lib.smjdev_memcpy_d2d(a.ptr, x.ptr, m * 8))
    smj.check(lib.smjdev_sort(a.ptr, b.ptr, m, 0, 1, C.byref(res)))
    smj.check(lib.smjdev_memcpy_d2d(dst.ptr, res.value, m * 8))
cnt, cks = C.c_uint64(0), C.c_uint64(0)
e0, e1 = C.c_void_p(), C.c_void_p()
lib.smjdev_event_create(C.byref(e0)); lib.smjdev_event_create(C.byref(e1))
best = 1e9
prof = os.environ.get("SMJ_TOOL_PROFILE")
if prof:
    cuda = C.CDLL("libcuda.so.1")
    for _ in range(2):
        smj.check(lib.smjdev_merge_join(sr.ptr, n, ss.ptr, n * mult, C.byref(cnt), C.byref(cks)))
    cuda.cuProfilerStart()
    smj.check(lib.smjdev_merge_join(sr.ptr, n, ss.ptr, n * mult, C.byref(cnt), C.byref(cks)))
    cuda.cuProfilerStop()
for _ in range(5):
    lib.smjdev_sync(); lib.smjdev_event_record(e0)
    smj.check(lib.smjdev_merge_join(sr.ptr, n, ss.ptr, n * mult, C.byref(cnt), C.byref(cks)))
    lib.smjdev_event_record(e1)
    ms = C.c_float(); lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms)); best = min(best, ms.value)
st = smj.device_join(smj.ALGO_MWAY, r, s)
ok = (st.matches, st.checksum) == (cnt.value, cks.value)
tot = n * (1 + mult)
print(f"merge_join {n} x {n * mult}: {best:.3f} ms = {tot * 8 / best / 1e6:.0f} GB/s; count {cnt.value} "
      f"checksum {cks.value:#x} equal to the m-way join's: {ok}")
assert ok
