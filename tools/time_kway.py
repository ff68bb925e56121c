#!/usr/bin/env python
"""Timing of smjdev_multiway_merge (k sorted runs of n/k items, chunks of a random permutation) by path:
python tools/time_kway.py [total items] [k ...]"""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128_000_000
ks = [int(x) for x in sys.argv[2:]] or [128]
smj.check(lib.smjdev_init(0))
r = smj.DeviceBuffer(n)
smj.check(lib.smjdev_gen_unique(r.ptr, n, 0, n, 12345))
a, b = smj.DeviceBuffer(n + 1024), smj.DeviceBuffer(n + 1024)
res = C.c_void_p()
smj.check(lib.smjdev_memcpy_d2d(a.ptr, r.ptr, n * 8))
smj.check(lib.smjdev_sort(a.ptr, b.ptr, n, 0, 1, C.byref(res)))
want = smj.DeviceBuffer(n)
smj.check(lib.smjdev_memcpy_d2d(want.ptr, res.value, n * 8))
ref = want.to_host() if n <= 1 << 28 else None
e0, e1 = C.c_void_p(), C.c_void_p()
lib.smjdev_event_create(C.byref(e0)); lib.smjdev_event_create(C.byref(e1))
for k in ks:
    per = n // k
    runbuf, t1, t2 = smj.DeviceBuffer(per * k), smj.DeviceBuffer(per), smj.DeviceBuffer(per)
    for i in range(k):
        smj.check(lib.smjdev_memcpy_d2d(t1.ptr, r.ptr + 8 * per * i, per * 8))
        smj.check(lib.smjdev_sort(t1.ptr, t2.ptr, per, 0, 1, C.byref(res)))
        smj.check(lib.smjdev_memcpy_d2d(runbuf.ptr + 8 * per * i, res.value, per * 8))
    ptrs = (C.c_void_p * k)(*[runbuf.ptr + 8 * per * i for i in range(k)])
    lens = (C.c_uint64 * k)(*[per] * k)
    line = f"k={k:4d} runs x {per} items:"
    for how in ("default", "kway-sorttile", "sort", "rounds"):
        os.environ.pop("SMJ_MERGE_PATH", None); os.environ.pop("SMJ_KWAY_TILE", None)
        if how == "kway-sorttile":
            os.environ["SMJ_KWAY_TILE"] = "sort"
        elif how != "default":
            os.environ["SMJ_MERGE_PATH"] = how
        best = 1e9
        for rep in range(4):
            lib.smjdev_sync(); lib.smjdev_event_record(e0)
            smj.check(lib.smjdev_multiway_merge(a.ptr, b.ptr, ptrs, lens, k, 0))
            lib.smjdev_event_record(e1)
            ms = C.c_float(); lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms))
            if rep:
                best = min(best, ms.value)
        line += f"  {how} {best:.3f} ms ({per * k * 16 / best / 1e6:.0f} GB/s)"
        if how == "default" and per * k == n and ref is not None:
            import numpy as np
            assert np.array_equal(a.to_host(n, 0), ref), "merged output differs from the sorted relation"
            line += " [checked]"
    print(line, flush=True)
    if os.environ.get("SMJ_TOOL_PROFILE"):
        os.environ.pop("SMJ_MERGE_PATH", None); os.environ.pop("SMJ_KWAY_TILE", None)
        lib.smjdev_profile_begin()
        smj.check(lib.smjdev_multiway_merge(a.ptr, b.ptr, ptrs, lens, k, 0))
        buf = C.create_string_buffer(1 << 16); lib.smjdev_profile_end(buf, len(buf))
        print(buf.value.decode())
    for x in (runbuf, t1, t2):
        x.free()
