#!/usr/bin/env python
"""Per-phase throughput of the join pipeline and of the copy paths on one B200: the GPU
counterpart of the reference's tputbench / numabench (src/bench/tputbench.c:249-1170), in its own
convention: TPUT = (|R| + |S|) / t in M tuples/s, MEMTPUT = TPUT x 2 x sizeof(tuple_t) (read +
write of every tuple, tputbench.c:321-325) in GB/s.  Phases as tputbench_thread runs them:
partition (partition_relation_optimized, 7 radix bits), sort (avxsort of the whole relation),
merge (multiway merge of 128 sorted runs), join (merge_join of the sorted relations).
Prints one JSON object.  Not the driver's bench (bench.py)."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
lib = smj.lib


def timed(fn, reps):
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.smjdev_event_create(C.byref(e0))
    lib.smjdev_event_create(C.byref(e1))
    best = None
    for _ in range(reps + 1):  # the first repetition warms the scratch pool
        lib.smjdev_sync()
        lib.smjdev_event_record(e0)
        fn()
        lib.smjdev_event_record(e1)
        ms = C.c_float()
        smj.check(lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms)))
        best = ms.value if best is None else min(best, ms.value)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128_000_000, help="tuples per relation")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    n = args.n
    smj.check(lib.smjdev_init(0))
    r, s = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
    a, b = smj.DeviceBuffer(n + 8 * 128), smj.DeviceBuffer(n + 8 * 128)
    smj.check(lib.smjdev_gen_unique(r.ptr, n, 0, n, 12345))
    smj.check(lib.smjdev_gen_unique(s.ptr, n, 0, n, 54321))
    hist, offs = smj.DeviceBuffer(128), smj.DeviceBuffer(128)
    res = C.c_void_p()
    out = {"tuples_per_relation": n, "tuple_bytes": 8, "phases": {}, "copies": {}}

    def phase(name, ms, tuples, bytes_per_tuple=16):
        out["phases"][name] = {"ms": round(ms, 3), "TPUT_Mtuples_s": round(tuples / ms / 1e3, 1),
                               "MEMTPUT_GBs": round(bytes_per_tuple * tuples / ms / 1e6, 1)}

    # partition: both relations, 7 radix bits, padded layout (tputbench.c:302-308)
    ms = timed(lambda: [smj.check(lib.smjdev_partition(x.ptr, a.ptr, b.ptr, n, 7, 0, 1, 1, hist.ptr, offs.ptr))
                        for x in (r, s)], args.reps)
    phase("partition_7bits", ms, 2 * n)
    # sort: whole relation, whole-word order (avxsort_tuples), input copied fresh every time
    copy_ms = timed(lambda: [smj.check(lib.smjdev_memcpy_d2d(a.ptr, x.ptr, n * 8)) for x in (r, s)], args.reps)

    def sort_both():
        for x in (r, s):
            smj.check(lib.smjdev_memcpy_d2d(a.ptr, x.ptr, n * 8))
            smj.check(lib.smjdev_sort(a.ptr, b.ptr, n, 0, 1, C.byref(res)))
    phase("sort", timed(sort_both, args.reps) - copy_ms, 2 * n)
    # sorted copies of R and S for the merge and join phases
    sr, ss = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
    for x, dst in ((r, sr), (s, ss)):
        smj.check(lib.smjdev_memcpy_d2d(a.ptr, x.ptr, n * 8))
        smj.check(lib.smjdev_sort(a.ptr, b.ptr, n, 0, 1, C.byref(res)))
        smj.check(lib.smjdev_memcpy_d2d(dst.ptr, res.value, n * 8))
    # merge: 128 sorted runs = the 128 chunks of the (unsorted) relation R, each sorted on its own,
    # like the runs the reference's threads hand to multiwaymerge_phase
    runs = 128
    per = n // runs
    runbuf, t1, t2 = smj.DeviceBuffer(per * runs), smj.DeviceBuffer(per), smj.DeviceBuffer(per)
    for i in range(runs):
        smj.check(lib.smjdev_memcpy_d2d(t1.ptr, r.ptr + 8 * per * i, per * 8))
        smj.check(lib.smjdev_sort(t1.ptr, t2.ptr, per, 0, 1, C.byref(res)))
        smj.check(lib.smjdev_memcpy_d2d(runbuf.ptr + 8 * per * i, res.value, per * 8))
    ptrs = (C.c_void_p * runs)(*[runbuf.ptr + 8 * per * i for i in range(runs)])
    lens = (C.c_uint64 * runs)(*[per] * runs)
    ms = timed(lambda: smj.check(lib.smjdev_multiway_merge(a.ptr, b.ptr, ptrs, lens, runs, 0)), args.reps)
    phase("multiway_merge_128_runs", ms, per * runs)  # default: one-pass k-way merge (csrc/smj_kway.cu)
    for how in ("sort", "rounds"):  # range partition of the union (MSD sort) / log2(k) pairwise rounds
        os.environ["SMJ_MERGE_PATH"] = how
        ms = timed(lambda: smj.check(lib.smjdev_multiway_merge(a.ptr, b.ptr, ptrs, lens, runs, 0)), args.reps)
        phase("multiway_merge_128_runs_" + how, ms, per * runs)
    os.environ.pop("SMJ_MERGE_PATH", None)
    # join: merge-path scan of the sorted relations
    cnt, cks = C.c_uint64(0), C.c_uint64(0)
    ms = timed(lambda: smj.check(lib.smjdev_merge_join(sr.ptr, n, ss.ptr, n, C.byref(cnt), C.byref(cks))), args.reps)
    assert cnt.value == n
    phase("merge_join_scan", ms, 2 * n, bytes_per_tuple=8)
    # the whole m-way join (what bench.py reports)
    st = smj.smj_join_stats_t()
    ms = timed(lambda: smj.check(lib.smjdev_join(smj.ALGO_MWAY, r.ptr, n, s.ptr, n, 1, C.byref(st))), args.reps)
    phase("mway_join_total", ms, 2 * n, bytes_per_tuple=40)

    # copy paths (numabench's memcpy matrix becomes: HBM, PCIe pinned / pageable)
    nb = n * 8
    ms = timed(lambda: smj.check(lib.smjdev_memcpy_d2d(a.ptr, r.ptr, nb)), args.reps)
    out["copies"]["hbm_d2d_GBs_read_plus_write"] = round(2 * nb / ms / 1e6, 1)
    hp = lib.smjdev_host_alloc(nb)
    ms = timed(lambda: smj.check(lib.smjdev_memcpy_h2d(a.ptr, hp, nb)), args.reps)
    out["copies"]["pcie_h2d_pinned_GBs"] = round(nb / ms / 1e6, 1)
    ms = timed(lambda: (smj.check(lib.smjdev_memcpy_d2h(hp, a.ptr, nb)), lib.smjdev_sync()), args.reps)
    out["copies"]["pcie_d2h_pinned_GBs"] = round(nb / ms / 1e6, 1)
    lib.smjdev_host_free(hp)
    pg = np.empty(n, dtype=np.uint64)
    pg[:] = 1
    ms = timed(lambda: smj.check(lib.smjdev_upload(a.ptr, pg.ctypes.data, nb)), args.reps)
    out["copies"]["pcie_h2d_pageable_staged_GBs"] = round(nb / ms / 1e6, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
