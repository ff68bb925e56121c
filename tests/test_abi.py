"""CPU-only checks of the drop-in boundary: the library loads, exports every symbol that
include/*.h declares, the ABI structs have the reference's layout, the host-side logic that needs
no GPU behaves like the reference, and compute calls FAIL LOUDLY when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in ("smj_b200.h", "smj_device.h", "smj_b200_key8b.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"^#define.*$", "", txt, flags=re.M)  # (the KEY_8B name-mapping macros)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", txt):
            if m.group(1) not in ("defined", "sizeof", "ALIGN_NUMTUPLES", "CACHELINEPADDING"):
                names.add(m.group(1))
    return names


def test_every_declared_symbol_is_exported(smj):
    out = subprocess.run(["nm", "-D", "--defined-only", smj.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    declared = declared_functions()
    assert len(declared) > 60
    missing = sorted(declared - exported)
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    # the Python signature table is the third copy of the contract
    assert declared == set(smj.SIGNATURES), sorted(declared ^ set(smj.SIGNATURES))


def test_struct_layout_matches_reference(smj):
    # reference src/types.h:51-98 compiled for x86-64
    assert C.sizeof(smj.tuple_t) == 8 and smj.tuple_t.payload.offset == 0 and smj.tuple_t.key.offset == 4
    assert C.sizeof(smj.relation_t) == 16 and smj.relation_t.num_tuples.offset == 8
    assert C.sizeof(smj.threadresult_t) == 24
    assert C.sizeof(smj.result_t) == 24 and smj.result_t.nthreads.offset == 16
    assert C.sizeof(smj.joinconfig_t) == 24
    # key<<32 | payload: what every kernel sorts on
    t = smj.tuple_t(payload=7, key=3)
    assert np.frombuffer(bytes(t), dtype=np.uint64)[0] == (3 << 32) | 7


def test_placement_matches_oracle(smj, orc):
    rng = np.random.default_rng(0)
    sizes = list(range(0, 300)) + [16383, 16384, 16385, 32684, 81920, 1000003, 4194381, 12345678] + \
        [int(x) for x in rng.integers(300, 1 << 26, size=300)]
    for n in sizes:
        assert smj.lib.smj_avxsort_placement(n, 0) == orc.placement(n, 0), n
    for n in [1, 5000, 2621440, 2621441, 3 * 2621440, 3 * 2621440 + 1, 4 * 2621440, 5 * 2621440 + 7,
              9 * 2621440 + 1]:
        assert smj.lib.smj_avxsort_placement(n, 1) == orc.placement(n, 1), n


def test_placement_matches_reference(smj, ref):
    import oracle_lib as ol
    rng = np.random.default_rng(1)
    for n in [1, 100, 255, 256, 300, 4095, 16384, 20000, 32684, 81920, 131072, 400003]:
        _, placement = ref.sort(ol.rand_items(rng, n, 1 << 26))
        assert smj.lib.smj_avxsort_placement(n, 0) == placement, n


def test_numa_shuffle_and_helpers(smj):
    lib = smj.lib
    lib.numa_shuffle_init(smj.NUMA_NEXT, 8)
    assert [lib.get_numa_shuffle_strategy(3, i, 8) for i in range(8)] == [(3 + i) % 8 for i in range(8)]
    lib.numa_shuffle_init(smj.NUMA_RING, 8)
    assert sorted(lib.get_numa_shuffle_strategy(2, i, 8) for i in range(8)) == list(range(8))
    lib.numa_shuffle_init(smj.NUMA_RANDOM, 8)
    assert sorted(lib.get_numa_shuffle_strategy(0, i, 8) for i in range(8)) == list(range(8))
    a = ((np.arange(1, 100, dtype=np.uint64)) << np.uint64(32)).view(np.int64)
    assert lib.is_sorted_helper(a.ctypes.data, a.size) == 1
    b = a[::-1].copy()
    assert lib.is_sorted_helper(b.ctypes.data, b.size) == 0
    # avxsort_int32 is an empty stub in the reference (src/avxsort/avxsort.c:247-250)
    x, y = np.array([3, 1, 2], dtype=np.int32), np.zeros(3, dtype=np.int32)
    px, py = C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data)
    lib.avxsort_int32(C.byref(px), C.byref(py), 3)
    assert list(x) == [3, 1, 2] and px.value == x.ctypes.data


def test_power_of_two_check_needs_no_gpu(smj, capfd):
    # reference src/joins/sortmergejoin_multiway.c:54-57: message + NULL before any work
    r = np.array([(1 << 32) | 1], dtype=np.uint64)
    total, _ = smj.join_host("multiway", r, r, nthreads=6)
    assert total is None
    assert "power of 2" in capfd.readouterr().out


def test_no_cpu_fallback(smj, capfd):
    if smj.have_gpu():
        pytest.skip("a GPU is present: the failure path cannot be exercised")
    assert smj.lib.smjdev_init(-1) != 0
    assert b"no CUDA device" in smj.lib.smjdev_last_error()
    items = np.arange(10, 0, -1, dtype=np.uint64) << np.uint64(32)
    a, b = items.copy(), np.zeros(10, dtype=np.uint64)
    pa, pb = C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data)
    smj.lib.avxsort_tuples(C.byref(pa), C.byref(pb), 10)  # SMJ_NO_EXIT is set by conftest
    assert "[ERROR]" in capfd.readouterr().out
    assert np.array_equal(a, items), "nothing may be computed on the CPU"
    total, _ = smj.join_host("multiway", items, items, nthreads=2)
    assert total is None


def test_p2p_layout():
    """offsets of the fused partition + exchange: slices land in source-rank order, no overlap."""
    import importlib
    dj = importlib.import_module("avx-sort-merge-joins_b200.distributed")
    rng = np.random.default_rng(3)
    W = 4
    counts = rng.integers(0, 1000, size=(W, 2, W))  # [src][rel][dst]
    layouts = [dj.p2p_layout(counts, g) for g in range(W)]
    for rel in range(2):
        for dst in range(W):
            pos = 0
            for src in range(W):
                assert layouts[src][0][rel][dst] == pos
                pos += counts[src][rel][dst]
            assert layouts[dst][1][rel] == pos


def test_msd_layout_and_plan():
    """bucket ownership / write offsets of the exchange-as-first-partition-pass (pure host logic)."""
    import importlib
    dj = importlib.import_module("avx-sort-merge-joins_b200.distributed")
    rng = np.random.default_rng(11)
    W = 4
    hist = rng.integers(0, 5000, size=(W, 2, 512))
    hist[:, :, 400:] = 0                       # unused buckets at the top of the key range
    hist[:, 1, 17] += 500000                   # a heavy S bucket
    lay = [dj.msd_layout(hist, g) for g in range(W)]
    owner = lay[0][0]
    assert all(np.array_equal(owner, l[0]) for l in lay) and np.all(np.diff(owner) >= 0)
    assert set(owner.tolist()) <= set(range(W))
    tot = hist.sum(axis=0)
    for rel in range(2):
        for g in range(W):
            pos = 0
            for b in np.nonzero(owner == g)[0]:
                for src in range(W):            # slices of a bucket in source-rank order, buckets back to back
                    assert lay[src][1][rel][b] == pos
                    pos += hist[src][rel][b]
                assert lay[g][2][rel][b] == tot[rel][b]
            assert lay[g][2][rel].sum() == pos
        assert sum(l[2][rel].sum() for l in lay) == tot[rel].sum()
    # plan: keys 1..128M -> 27 varying bits, top digit at item bit 32+26-8
    items_or, items_and = (128_000_000 | 0x7FFFFFF) << 32, 0
    assert dj.msd_plan(items_or, items_and, items_or, items_and) == (0, 26, 50)
    assert dj.msd_plan(5 << 32, 5 << 32, 5 << 32, 5 << 32) is None
    assert dj.msd_plan_holds(items_or, 0, items_or, 0, 0, 26)
    assert not dj.msd_plan_holds(items_or | (1 << 62), 0, items_or, 0, 0, 26)


def test_msd_layout_dma_simulated_exchange():
    """Layout of the copy-engine exchange (pure host logic): scatter every rank's items bucket-major
    into a send buffer, apply the planned copies, and check that every owner's receive buffer is
    exactly covered by its segments and that each segment holds one bucket's items."""
    import importlib
    dj = importlib.import_module("avx-sort-merge-joins_b200.distributed")
    rng = np.random.default_rng(0)
    for W in (1, 2, 3, 8):
        hist = rng.integers(0, 50, size=(W, 2, 512))
        hist[:, :, 100:140] = 0
        hist[:, 1, 17] += 5000  # a heavy S bucket
        if W == 8:
            hist[:, :, 300:] = 0  # some ranks may own nothing at the top
        lays = [dj.msd_layout_dma(hist, g) for g in range(W)]
        assert all(np.array_equal(lays[0]["owner"], l["owner"]) for l in lays)
        for rel in range(2):
            send = [np.concatenate([np.full(hist[g][rel][b], b) for b in range(512)]) for g in range(W)]
            recv = [np.full(lays[p]["recv"][rel], -1) for p in range(W)]
            for g in range(W):
                lay = lays[g]
                lo, hi = lay["first"][g], lay["first"][g + 1]
                start = lay["send_off"][rel][lo] if lo < 512 else 0
                blk = send[g][start:start + hist[g][rel][lo:hi].sum()]
                recv[g][lay["self_off"][rel]:lay["self_off"][rel] + blk.size] = blk
                for (crel, p, src, cnt, dst) in lay["copies"]:
                    if crel == rel:
                        assert p != g and cnt > 0
                        recv[p][dst:dst + cnt] = send[g][src:src + cnt]
            for p in range(W):
                assert (recv[p] >= 0).all()
                st, ct, bk = lays[p]["segs"][rel]
                seen = np.zeros(recv[p].size, bool)
                for a, c, b in zip(st, ct, bk):
                    assert lays[p]["owner"][b] == p and (recv[p][a:a + c] == b).all() and not seen[a:a + c].any()
                    seen[a:a + c] = True
                assert seen.all() and ct.sum() == hist[:, rel, lays[p]["owner"] == p].sum()
