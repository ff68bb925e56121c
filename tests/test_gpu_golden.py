"""The CUDA path (through the reference-facing C entry points, host buffers) against the frozen
reference outputs in tests/golden/ref_golden.npz.  Needs a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))
G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden_r2.npz"))


def test_sort_golden_gpu(smj):
    for n in G["sort_sizes"]:
        n = int(n)
        a, b = G[f"sort_in_{n}"].copy(), np.zeros(n, dtype=np.uint64)
        pa, pb = C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data)
        smj.lib.avxsort_tuples(C.byref(pa), C.byref(pb), n)
        placement = int(G[f"sort_place_{n}"])
        assert pb.value == (b.ctypes.data if placement else a.ctypes.data)
        assert np.array_equal(b if placement else a, G[f"sort_out_{n}"])


def test_merge_golden_gpu(smj):
    a, b = G["merge_a"].copy(), G["merge_b"].copy()
    out = np.zeros(a.size + b.size, dtype=np.uint64)
    assert smj.lib.avx_merge_tuples(a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size, b.size) == out.size
    assert np.array_equal(out, G["merge_avx"])
    a, b = G["merge_ak"].copy(), G["merge_bk"].copy()
    assert smj.lib.scalar_merge_tuples(a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size, b.size) == out.size
    assert np.array_equal(out, G["merge_scalar"])
    runs = [r.copy() for r in np.split(G["mway_in"], np.cumsum(G["mway_lens"])[:-1])]
    rels = (smj.relation_t * len(runs))(*[smj.relation_t(r.ctypes.data if r.size else None, r.size) for r in runs])
    pp = (C.POINTER(smj.relation_t) * len(runs))(*[C.pointer(rels[i]) for i in range(len(runs))])
    out = np.zeros(G["mway_out"].size, dtype=np.uint64)
    assert smj.lib.avx_multiway_merge(out.ctypes.data, pp, len(runs), None, 0) == out.size
    assert np.array_equal(out, G["mway_out"])


@pytest.mark.parametrize("nbits,shift", [(4, 0), (7, 3), (10, 1)])
def test_partition_golden_gpu(smj, nbits, shift):
    items = G["part_in"].copy()
    fan = 1 << nbits
    for name, tag in (("partition_relation", "plain"), ("partition_relation_optimized", "opt"),
                      ("partition_relation_optimized_V2", "v2")):
        out = np.zeros(items.size + 8 * fan + 8, dtype=np.uint64)
        parts = (smj.relation_t * fan)()
        pp = (C.POINTER(smj.relation_t) * fan)(*[C.pointer(parts[i]) for i in range(fan)])
        rin, rout = smj.rel(items), smj.relation_t(out.ctypes.data, items.size)
        getattr(smj.lib, name)(pp, C.byref(rin), C.byref(rout), nbits, shift)
        hist = np.array([parts[i].num_tuples for i in range(fan)], dtype=np.uint64)
        offs = np.array([(parts[i].tuples - out.ctypes.data) // 8 for i in range(fan)], dtype=np.uint64)
        assert np.array_equal(hist, G[f"part_{tag}_{nbits}_{shift}_hist"])
        assert np.array_equal(offs, G[f"part_{tag}_{nbits}_{shift}_offs"])
        packed = np.concatenate([out[int(o):int(o) + int(c)] for o, c in zip(offs, hist)])
        assert np.array_equal(packed, G[f"part_{tag}_{nbits}_{shift}_data"])


def test_join_golden_gpu(smj):
    r, s = G["mj_r"].copy(), G["mj_s"].copy()
    assert smj.lib.merge_join(r.ctypes.data, s.ctypes.data, r.size, s.size, None) == int(G["mj_count"])
    r, s = G["mji_r"].copy(), G["mji_s"].copy()
    assert smj.lib.merge_join_interpolation(r.ctypes.data, s.ctypes.data, r.size, s.size, None) == int(G["mji_count"])
    R, S, D1, D2 = G["join_R"].copy(), G["join_S"].copy(), G["join_D1"].copy(), G["join_D2"].copy()
    for algo, nthreads in (("multiway", 2), ("multiway", 4), ("multipass", 4), ("mpsm", 4)):
        assert smj.join_host(algo, R, S, nthreads=nthreads)[0] == int(G["join_RS_mway_2"])
        assert smj.join_host(algo, D1, D2, nthreads=nthreads)[0] == int(G["join_DD_mway_2"])


def test_materialize_golden_gpu(smj):
    """sortmergejoin_multiway_materialize (host buffers, the reference's JOIN_MATERIALIZE contract)
    against the tuples the compiled reference emitted (tests/golden/make_golden_r2.py)."""
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for name in G2["mat_names"]:
        r, s = G2[f"mat_{name}_r"].copy(), G2[f"mat_{name}_s"].copy()
        exp = G2[f"mat_{name}_sorted"]
        cfg = smj.joinconfig_t(4, 128, 0, 0, 20 * 1024 * 1024, smj.NUMA_NEXT)
        res = smj.lib.sortmergejoin_multiway_materialize(C.byref(smj.rel(r)), C.byref(smj.rel(s)), C.byref(cfg))
        assert res and res.contents.totalresults == exp.size, name
        tr = res.contents.resultlist[0]
        got = (np.ctypeslib.as_array(C.cast(tr.results, C.POINTER(C.c_uint64)), shape=(exp.size,)).copy()
               if exp.size else np.zeros(0, dtype=np.uint64))
        assert np.array_equal(np.sort(got), exp), name
        if tr.results:
            libc.free(C.cast(tr.results, C.c_void_p))
        libc.free(C.cast(res.contents.resultlist, C.c_void_p))
        libc.free(C.cast(res, C.c_void_p))
