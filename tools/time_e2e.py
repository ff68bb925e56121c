#!/usr/bin/env python
"""End-to-end timing of sortmergejoin_multiway() on pinned host relations (bench.py's e2e leg)."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200"); lib = smj.lib
n = 128_000_000
smj.check(lib.smjdev_init(0))
dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
lib.smjdev_gen_unique(dr.ptr, n, 0, n, 12345); lib.smjdev_gen_unique(ds.ptr, n, 0, n, 54321)
hr, hs = lib.smjdev_host_alloc(n * 8), lib.smjdev_host_alloc(n * 8)
lib.smjdev_memcpy_d2h(hr, dr.ptr, n * 8); lib.smjdev_memcpy_d2h(hs, ds.ptr, n * 8); lib.smjdev_sync()
cfg = smj.joinconfig_t(1, 128, 0, 0, 20 * 1024 * 1024, smj.NUMA_NEXT)
relr, rels = smj.relation_t(hr, n), smj.relation_t(hs, n)
libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
ts, h2d = [], []
for i in range(7):
    t0 = time.perf_counter()
    res = lib.sortmergejoin_multiway(C.byref(relr), C.byref(rels), C.byref(cfg))
    ts.append(time.perf_counter() - t0)
    assert res.contents.totalresults == n
    st = smj.smj_join_stats_t(); lib.smj_last_join_stats(C.byref(st)); h2d.append(st.ms_h2d)
    libc.free(C.cast(res.contents.resultlist, C.c_void_p)); libc.free(C.cast(res, C.c_void_p))
print(f"chunk {os.environ.get('SMJ_UPLOAD_CHUNK_MB', '128')} MiB: e2e {min(ts[2:]) * 1e3:.2f} ms (mean {sum(ts[2:]) / len(ts[2:]) * 1e3:.2f}), h2d {min(h2d[2:]):.2f} ms")
