#!/usr/bin/env python
"""How fast is the drop-in path when the caller's relations live in ordinary malloc'ed (pageable)
host memory, like the reference's own main.c:516-528 allocates them?  Prints ms per join and the
H2D share, next to the same join from pinned memory."""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128_000_000
smj.check(lib.smjdev_init(0))
dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, 12345))
smj.check(lib.smjdev_gen_unique(ds.ptr, n, 0, n, 54321))
r, s = dr.to_host(), ds.to_host()  # numpy arrays: pageable memory
for kind in ("pageable", "pinned"):
    if kind == "pinned":
        hr, hs = lib.smjdev_host_alloc(n * 8), lib.smjdev_host_alloc(n * 8)
        C.memmove(hr, r.ctypes.data, n * 8)
        C.memmove(hs, s.ctypes.data, n * 8)
        rr, ss = smj.relation_t(hr, n), smj.relation_t(hs, n)
    else:
        rr, ss = smj.rel(r), smj.rel(s)
    cfg = smj.joinconfig_t(1, 128, 0, 0, 0, smj.NUMA_NEXT)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    best, h2d = None, None
    for it in range(4):
        t0 = time.perf_counter()
        res = lib.sortmergejoin_multiway(C.byref(rr), C.byref(ss), C.byref(cfg))
        dt = time.perf_counter() - t0
        assert res.contents.totalresults == n
        libc.free(C.cast(res.contents.resultlist, C.c_void_p))
        libc.free(C.cast(res, C.c_void_p))
        st = smj.smj_join_stats_t()
        lib.smj_last_join_stats(C.byref(st))
        if it and (best is None or dt < best):
            best, h2d = dt, st.ms_h2d
    print(f"{kind:9s}: {best * 1e3:8.2f} ms per join ({2 * n / best / 1e6:9.1f} M tuples/s), H2D {h2d:7.2f} ms "
          f"= {2 * n * 8 / h2d / 1e6:6.1f} GB/s")
