#!/usr/bin/env python
"""End-to-end time of sortmergejoin_multiway() on ordinary malloc'ed (pageable) relations, the way the
reference's main.c passes them: python tools/time_pageable.py [n]   (env: SMJ_STAGE_THREADS,
SMJ_STAGE_CHUNK_KB)"""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128_000_000
smj.check(lib.smjdev_init(0))
dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, 12345))
smj.check(lib.smjdev_gen_unique(ds.ptr, n, 0, n, 54321))
r, s = dr.to_host(), ds.to_host()
dr.free()
ds.free()
best = None
for it in range(5):
    t0 = time.perf_counter()
    total, st = smj.join_host("multiway", r, s, nthreads=1)
    dt = time.perf_counter() - t0
    assert total == n
    if it:
        best = dt if best is None else min(best, dt)
print(f"pageable e2e: {best * 1e3:.2f} ms ({2 * n / best / 1e6:.0f} M tuples/s), H2D {st.ms_h2d:.2f} ms; "
      f"threads={os.environ.get('SMJ_STAGE_THREADS', 'default')} chunk_kb={os.environ.get('SMJ_STAGE_CHUNK_KB', '8192')}")
