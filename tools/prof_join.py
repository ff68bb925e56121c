#!/usr/bin/env python
"""One m-way join of 128M x 128M (bench.py's workload) for Nsight Compute: warm-up joins run
outside the profiled region (cuProfilerStart/Stop around the last one), so

    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -o gpurun_out/prof python tools/prof_join.py

captures exactly the kernels of one step."""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128_000_000
lib = smj.lib
smj.check(lib.smjdev_init(0), "init")
dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, 12345))
smj.check(lib.smjdev_gen_unique(ds.ptr, n, 0, n, 54321))
for _ in range(3):
    st = smj.device_join(smj.ALGO_MWAY, dr, ds)
cuda = C.CDLL("libcuda.so.1")
cuda.cuProfilerStart()
st = smj.device_join(smj.ALGO_MWAY, dr, ds)
cuda.cuProfilerStop()
print("matches", st.matches, "ms", st.ms_total, "path", st.path, "repeats", st.repeats, "launches", st.launches)
