#!/usr/bin/env python
"""Timing of smjdev_sort (whole-word order) on bench_sort's data (keys = permutation of 1..n,
payload = row number): python tools/time_sort.py [n] ; SMJ_SORT_PATH=lsd forces the LSD sort."""
import ctypes as C
import importlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
smj.check(lib.smjdev_init(0))
src, buf, tmp = smj.DeviceBuffer(n), smj.DeviceBuffer(n), smj.DeviceBuffer(n)
smj.check(lib.smjdev_gen_unique(src.ptr, n, 0, n, 12345))
res = C.c_void_p()
e0, e1 = C.c_void_p(), C.c_void_p()
lib.smjdev_event_create(C.byref(e0)); lib.smjdev_event_create(C.byref(e1))
best = None
for i in range(4):
    smj.check(lib.smjdev_memcpy_d2d(buf.ptr, src.ptr, n * 8))
    lib.smjdev_sync(); lib.smjdev_event_record(e0)
    smj.check(lib.smjdev_sort(buf.ptr, tmp.ptr, n, 0, 1, C.byref(res)))
    lib.smjdev_event_record(e1)
    ms = C.c_float(); smj.check(lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms)))
    if i: best = ms.value if best is None else min(best, ms.value)
st = (C.c_uint64 * 3)()
smj.check(lib.smjdev_check_order(res.value, n, 0, st))
print(f"n {n} sort {best:.3f} ms = {n / best / 1e3:.0f} M items/s, {16 * n / best / 1e6:.0f} GB/s at 16 B/item; descents {st[0]}")
lib.smjdev_profile_begin()
smj.check(lib.smjdev_memcpy_d2d(buf.ptr, src.ptr, n * 8))
smj.check(lib.smjdev_sort(buf.ptr, tmp.ptr, n, 0, 1, C.byref(res)))
pb = C.create_string_buffer(1 << 16)
lib.smjdev_profile_end(pb, len(pb))
for line in pb.value.decode().strip().splitlines():
    name, cnt, t, items = line.rsplit(";", 3)
    print(f"   {name.strip('() ')[:40]:42s} x{int(cnt):<3d} {float(t):9.3f} ms")
