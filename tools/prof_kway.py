#!/usr/bin/env python
"""One multiway merge of 128 sorted runs (tputbench's merge phase, scaled) for Nsight Compute:
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_kw \
        -o gpurun_out/prof_kway python tools/prof_kway.py"""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32_000_000
lib = smj.lib
smj.check(lib.smjdev_init(0), "init")
runs, per = 128, n // 128
src, runbuf, t1, t2 = smj.DeviceBuffer(n), smj.DeviceBuffer(per * runs), smj.DeviceBuffer(per), smj.DeviceBuffer(per)
out, tmp = smj.DeviceBuffer(per * runs), smj.DeviceBuffer(per * runs)
smj.check(lib.smjdev_gen_unique(src.ptr, n, 0, n, 12345))
res = C.c_void_p()
for i in range(runs):
    smj.check(lib.smjdev_memcpy_d2d(t1.ptr, src.ptr + 8 * per * i, per * 8))
    smj.check(lib.smjdev_sort(t1.ptr, t2.ptr, per, 0, 1, C.byref(res)))
    smj.check(lib.smjdev_memcpy_d2d(runbuf.ptr + 8 * per * i, res.value, per * 8))
ptrs = (C.c_void_p * runs)(*[runbuf.ptr + 8 * per * i for i in range(runs)])
lens = (C.c_uint64 * runs)(*[per] * runs)
smj.check(lib.smjdev_multiway_merge(out.ptr, tmp.ptr, ptrs, lens, runs, 0))
lib.smjdev_sync()
cuda = C.CDLL("libcuda.so.1")
cuda.cuProfilerStart()
lib.smjdev_profile_begin()
smj.check(lib.smjdev_multiway_merge(out.ptr, tmp.ptr, ptrs, lens, runs, 0))
buf = C.create_string_buffer(1 << 16)
lib.smjdev_profile_end(buf, len(buf))
cuda.cuProfilerStop()
stats = (C.c_uint64 * 3)()
smj.check(lib.smjdev_check_order(out.ptr, per * runs, 0, stats))
print("descents", stats[0])
print(buf.value.decode())
