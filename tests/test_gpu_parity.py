"""Parity of the CUDA path (through the C-ABI of libsmj_b200.so) with the CPU oracle on the same
seeded inputs.  Bit-exact everywhere: this is integer work.  Needs a GPU (pytest -m gpu)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

SORT_SIZES = [1, 2, 15, 16, 100, 255, 256, 6143, 6144, 6145, 16383, 16384, 16385, 100003, 1000003]


def _inputs(rng, n, kind):
    if kind == "unique":
        return ol.make_unique(rng, n)
    if kind == "dups":
        return ol.rand_items(rng, n, 7)
    if kind == "equal":
        return np.full(n, (5 << 32) | 9, dtype=np.uint64)
    if kind == "presorted":
        return np.sort(ol.rand_items(rng, n, 1 << 20))
    if kind == "reverse":
        return np.sort(ol.rand_items(rng, n, 1 << 20))[::-1].copy()
    if kind == "wide":  # all 32 payload bits and 31 key bits in use
        return ol.rand_items(rng, n, (1 << 31) - 1)
    raise ValueError(kind)


# ------------------------------------------------------------------------------------ sort
def _sort_device(smj, items, key_only=0):
    n = items.size
    buf, tmp = smj.DeviceBuffer.from_host(items), smj.DeviceBuffer(n)
    res = C.c_void_p()
    smj.check(smj.lib.smjdev_sort(buf.ptr, tmp.ptr, n, key_only, 1, C.byref(res)), "sort")
    smj.check(smj.lib.smjdev_sync())
    assert res.value in (buf.ptr, tmp.ptr)
    got = (buf if res.value == buf.ptr else tmp).to_host()
    buf.free()
    tmp.free()
    return got


def test_device_sort_msd_path(smj, orc, monkeypatch):
    """Whole-word sorts of >= 2^21 items take the MSD path (two histogram-free partition passes +
    in-bucket counting sort, ties ordered by payload) when it applies and the LSD sort otherwise;
    bit-identical to the oracle (== the reference's avxsort order) either way."""
    rng = np.random.default_rng(21)
    n = 3_000_017
    lo32 = np.uint64(0xFFFFFFFF)
    cases = {
        "unique": ol.make_unique(rng, n),                               # keys a permutation: no ties
        "row_ids": orc.gen_unique(n, 12345),                            # bench_sort's data
        "short_ties": ol.rand_items(rng, n, 1 << 21),                   # ~1.4 items per key: runs ordered in place
        "wide": ol.rand_items(rng, n, (1 << 31) - 1),                   # 31 key bits: 13 bits left per bucket
        "long_ties": ol.rand_items(rng, n, 1 << 10) | (np.uint64(1 << 30) << np.uint64(32)),  # runs > 64: LSD
        "narrow": ol.rand_items(rng, n, 1 << 12),                       # 12 key bits: LSD
        "sorted": np.sort(ol.make_unique(rng, n)),
        "reverse": np.sort(ol.make_unique(rng, n))[::-1].copy(),
    }
    skew = ol.make_unique(rng, n)
    skew[: n // 2] = (skew[: n // 2] & lo32) | (np.uint64(77) << np.uint64(32))  # half the items share a key
    cases["skew"] = skew
    for name, items in cases.items():
        exp = orc.sort(items)
        monkeypatch.setenv("SMJ_SORT_PATH", "msd")  # the MSD path from 2^21 items on (default: 40 M)
        assert np.array_equal(_sort_device(smj, items), exp), name
        monkeypatch.setenv("SMJ_SORT_PATH", "lsd")
        assert np.array_equal(_sort_device(smj, items), exp), name + " (lsd)"
        monkeypatch.delenv("SMJ_SORT_PATH")
    # the optimistic passes' repair paths under the sort
    monkeypatch.setenv("SMJ_SORT_PATH", "msd")
    for hook in ("shrink2", "nocoop"):
        monkeypatch.setenv("SMJ_MSD_TEST", hook)
        assert np.array_equal(_sort_device(smj, cases["unique"]), orc.sort(cases["unique"])), hook
        monkeypatch.delenv("SMJ_MSD_TEST")
    monkeypatch.delenv("SMJ_SORT_PATH")


@pytest.mark.parametrize("n", SORT_SIZES)
@pytest.mark.parametrize("algo", [1, 0])
def test_device_sort(smj, orc, n, algo):
    rng = np.random.default_rng(n)
    for kind in ("unique", "dups", "equal", "presorted", "reverse", "wide"):
        items = _inputs(rng, n, kind)
        buf, tmp = smj.DeviceBuffer.from_host(items), smj.DeviceBuffer(n)
        res = C.c_void_p()
        smj.check(smj.lib.smjdev_sort(buf.ptr, tmp.ptr, n, 0, algo, C.byref(res)), "sort")
        smj.check(smj.lib.smjdev_sync())
        assert res.value in (buf.ptr, tmp.ptr)
        got = (buf if res.value == buf.ptr else tmp).to_host()
        assert np.array_equal(got, orc.sort(items)), (kind, n)
        # key-only variant: stable on the key
        buf2, tmp2 = smj.DeviceBuffer.from_host(items), smj.DeviceBuffer(n)
        smj.check(smj.lib.smjdev_sort(buf2.ptr, tmp2.ptr, n, 1, algo, C.byref(res)), "sort")
        smj.check(smj.lib.smjdev_sync())
        got = (buf2 if res.value == buf2.ptr else tmp2).to_host()
        assert np.array_equal(got, orc.sort_key_stable(items)), (kind, n)
        for b in (buf, tmp, buf2, tmp2):
            b.free()


def test_key_sort_with_sampled_bit_plan(smj, orc, monkeypatch):
    """Key-only sorts of >= 2^22 items plan their radix passes from a SAMPLE of the key bits and
    verify the plan against the exact OR / AND the histogram pass computes on the way.  Same
    (stable) result as the oracle when the sample is right, when it misses a varying bit (one item
    with a high key bit at a place the sample does not look: the plan is redone from the exact
    bits), when all keys are equal, and with the sampling switched off."""
    rng = np.random.default_rng(4242)
    n = (1 << 22) + 12345
    plain = ol.rand_items(rng, n, 1 << 20)
    missed = plain.copy()
    missed[1] |= np.uint64(1 << 29) << np.uint64(32)       # index 1 is not a multiple of the sample stride
    missed[n - 3] |= np.uint64(1 << 27) << np.uint64(32)
    equal = (plain & np.uint64(0xFFFFFFFF)) | (np.uint64(99) << np.uint64(32))
    one_off = equal.copy()
    one_off[7] = (np.uint64(5) << np.uint64(32)) | np.uint64(1)  # "all keys equal" according to the sample only
    for name, items in (("plain", plain), ("missed", missed), ("equal", equal), ("one_off", one_off)):
        exp = orc.sort_key_stable(items)
        for off in (False, True):
            if off:
                monkeypatch.setenv("SMJ_NO_SAMPLED_BITS", "1")
            buf, tmp = smj.DeviceBuffer.from_host(items), smj.DeviceBuffer(n)
            res = C.c_void_p()
            smj.check(smj.lib.smjdev_sort(buf.ptr, tmp.ptr, n, 1, 0, C.byref(res)), "sort")
            smj.check(smj.lib.smjdev_sync())
            got = (buf if res.value == buf.ptr else tmp).to_host()
            assert np.array_equal(got, exp), (name, off)
            buf.free()
            tmp.free()
            monkeypatch.delenv("SMJ_NO_SAMPLED_BITS", raising=False)


@pytest.mark.parametrize("n", [0, 1, 100, 255, 16384, 32684, 81920, 1000003])
def test_avxsort_tuples_host_api(smj, orc, n):
    """avxsort_tuples with host buffers: sorted contents, pointer-swap contract and the buffer the
    reference would have chosen (reference src/avxsort/avxsort.c:117-118, 212-226)."""
    rng = np.random.default_rng(n + 1)
    items = ol.rand_items(rng, n, 1 << 27)
    for name, mw in (("avxsort_tuples", 0), ("avxsort_int64", 0), ("avxsortmultiway_tuples", 1)):
        a, b = items.copy(), np.zeros(max(n, 1), dtype=np.uint64)
        pa, pb = C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data)
        getattr(smj.lib, name)(C.byref(pa), C.byref(pb), n)
        if n == 0:
            assert (pa.value, pb.value) == (a.ctypes.data, b.ctypes.data)
            continue
        placement = orc.placement(n, mw)
        assert pb.value == (b.ctypes.data if placement else a.ctypes.data), name
        assert pa.value == (a.ctypes.data if placement else b.ctypes.data), name
        res = b if placement else a
        assert np.array_equal(res[:n], orc.sort(items)), name


def test_scalarsort_host_api(smj, orc):
    rng = np.random.default_rng(5)
    n = 50001
    items = ol.rand_items(rng, n, 1000)
    a, b = items.copy(), np.zeros(n, dtype=np.uint64)
    pa, pb = C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data)
    smj.lib.scalarsort_tuples(C.byref(pa), C.byref(pb), n)
    assert pb.value == a.ctypes.data and pa.value == b.ctypes.data  # sorted in place, swapped
    assert np.array_equal(a >> np.uint64(32), np.sort(items >> np.uint64(32)))
    assert np.array_equal(np.sort(a), np.sort(items))
    v = rng.integers(-2**31, 2**31 - 1, size=n, dtype=np.int64).astype(np.int32)
    w, z = v.copy(), np.zeros(n, dtype=np.int32)
    pa, pb = C.c_void_p(w.ctypes.data), C.c_void_p(z.ctypes.data)
    smj.lib.scalarsort_int32(C.byref(pa), C.byref(pb), n)
    assert pb.value == w.ctypes.data
    assert np.array_equal(w, np.sort(v))


# ------------------------------------------------------------------------------------ partition
@pytest.mark.parametrize("nbits,shift", [(2, 0), (4, 0), (7, 3), (8, 0), (9, 1), (11, 0), (14, 2),
                                         (15, 0)])
def test_partition_relation_host_api(smj, orc, nbits, shift):
    """partition_relation / _optimized / _V2 vs the oracle: counts, starts (8-tuple aligned for the
    optimized variants) and exact per-partition contents, i.e. the scatter is stable
    (reference tests/check_partitioning.c:73-90)."""
    rng = np.random.default_rng(nbits * 100 + shift)
    n = 200003
    items = ol.rand_items(rng, n, 1 << 22)
    fan = 1 << nbits
    for name, padded in (("partition_relation", False), ("partition_relation_optimized", True),
                         ("partition_relation_optimized_V2", True)):
        out = np.zeros(n + 8 * fan + 8, dtype=np.uint64)
        parts = (smj.relation_t * fan)()
        pp = (C.POINTER(smj.relation_t) * fan)(*[C.pointer(parts[i]) for i in range(fan)])
        rin, rout = smj.rel(items), smj.relation_t(out.ctypes.data, n)
        getattr(smj.lib, name)(pp, C.byref(rin), C.byref(rout), nbits, shift)
        eout, ehist, eoffs, _ = orc.partition(items, nbits, shift, padded)
        hist = np.array([parts[i].num_tuples for i in range(fan)], dtype=np.uint64)
        offs = np.array([(parts[i].tuples - out.ctypes.data) // 8 for i in range(fan)], dtype=np.uint64)
        assert np.array_equal(hist, ehist), name
        assert np.array_equal(offs, eoffs), name
        if padded:
            assert not np.any(offs % np.uint64(8)), name
        for p in range(fan):
            lo, c = int(offs[p]), int(hist[p])
            assert np.array_equal(out[lo:lo + c], eout[lo:lo + c]), (name, p)


def test_radix_cluster_and_histogram(smj, orc):
    rng = np.random.default_rng(11)
    n = 70001
    items = ol.rand_items(rng, n, 1 << 20)
    out = np.zeros(n, dtype=np.uint64)
    hist = np.zeros(64, dtype=np.int32)
    rin, rout = smj.rel(items), smj.relation_t(out.ctypes.data, n)
    smj.lib.radix_cluster(C.byref(rout), C.byref(rin), hist.ctypes.data_as(C.POINTER(C.c_int32)), 2, 6)
    eout, ehist, _, _ = orc.partition(items, 6, 2, False)
    assert np.array_equal(hist.astype(np.uint64), ehist)
    assert np.array_equal(out, eout[:n])
    out2 = np.zeros(n, dtype=np.uint64)
    rout2 = smj.relation_t(out2.ctypes.data, n)
    smj.lib.histogram_memcpy_bench(None, C.byref(rin), C.byref(rout2), 10)
    assert np.array_equal(out2, items)


def test_partition_empty_and_tiny(smj, orc):
    for n in (0, 1, 7):
        items = ol.rand_items(np.random.default_rng(n), n, 100)
        fan = 16
        out = np.zeros(n + 8 * fan + 8, dtype=np.uint64)
        parts = (smj.relation_t * fan)()
        pp = (C.POINTER(smj.relation_t) * fan)(*[C.pointer(parts[i]) for i in range(fan)])
        rin, rout = smj.relation_t(items.ctypes.data if n else None, n), smj.relation_t(out.ctypes.data, n)
        smj.lib.partition_relation_optimized(pp, C.byref(rin), C.byref(rout), 4, 0)
        _, ehist, eoffs, _ = orc.partition(items, 4, 0, True)
        assert [parts[i].num_tuples for i in range(fan)] == list(ehist)


# ------------------------------------------------------------------------------------ merge
@pytest.mark.parametrize("na,nb", [(0, 5), (5, 0), (1, 1), (100, 37), (4096, 4096), (4095, 4097),
                                   (100000, 333), (250000, 250001)])
def test_merge_host_api(smj, orc, na, nb):
    rng = np.random.default_rng(na * 31 + nb)
    a = np.sort(ol.rand_items(rng, na, 50))
    b = np.sort(ol.rand_items(rng, nb, 50))
    out = np.zeros(na + nb, dtype=np.uint64)
    for name in ("avx_merge_tuples", "avx_merge_int64", "scalar_merge_int64"):
        out[:] = 0
        k = getattr(smj.lib, name)(a.ctypes.data, b.ctypes.data, out.ctypes.data, na, nb)
        assert k == na + nb
        assert np.array_equal(out, orc.merge(a, b, 0)), name
    ak, bk = orc.sort_key_stable(a), orc.sort_key_stable(b)
    k = smj.lib.scalar_merge_tuples(ak.ctypes.data, bk.ctypes.data, out.ctypes.data, na, nb)
    assert k == na + nb
    assert np.array_equal(out, orc.merge(ak, bk, 1))


@pytest.mark.parametrize("nruns", [1, 2, 3, 4, 8, 64, 100])
def test_multiway_merge_host_api(smj, orc, nruns):
    rng = np.random.default_rng(nruns)
    runs = [np.sort(ol.rand_items(rng, int(rng.integers(0, 9000)), 1 << 16)) for _ in range(nruns)]
    total = sum(r.size for r in runs)
    out = np.zeros(total + 1, dtype=np.uint64)
    rels = (smj.relation_t * nruns)(*[smj.relation_t(r.ctypes.data if r.size else None, r.size) for r in runs])
    pp = (C.POINTER(smj.relation_t) * nruns)(*[C.pointer(rels[i]) for i in range(nruns)])
    k = smj.lib.avx_multiway_merge(out.ctypes.data, pp, nruns, None, 0)
    assert k == total
    assert np.array_equal(out[:total], orc.multiway_merge(runs))
    assert all(rels[i].num_tuples == 0 for i in range(nruns))  # inputs are consumed
    # scalar twin: key-only order; the reference only guarantees key-sortedness + the multiset
    rels = (smj.relation_t * nruns)(*[smj.relation_t(r.ctypes.data if r.size else None, r.size) for r in runs])
    pp = (C.POINTER(smj.relation_t) * nruns)(*[C.pointer(rels[i]) for i in range(nruns)])
    out[:] = 0
    k = smj.lib.scalar_multiway_merge(out.ctypes.data, pp, nruns, None, 0)
    assert k == total
    assert np.all(np.diff((out[:total] >> np.uint64(32)).astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(out[:total]), orc.multiway_merge(runs))


# ------------------------------------------------------------------------------------ merge join
@pytest.mark.parametrize("nr,ns,kmax", [(1000, 1000, 1000), (5000, 20000, 300), (100, 100, 1),
                                        (1, 1, 5), (3000, 10, 50), (200000, 300000, 150000),
                                        (50000, 50000, 3), (4096, 4096, 1 << 30), (100000, 1, 7)])
def test_merge_join(smj, orc, nr, ns, kmax):
    rng = np.random.default_rng(nr + ns + kmax)
    r = orc.sort(ol.rand_items(rng, nr, kmax))
    s = orc.sort(ol.rand_items(rng, ns, kmax))
    exp = orc.merge_join_runs(r, s)
    if nr * ns <= 4_000_000:
        assert orc.merge_join(r, s) == exp  # literal restatement agrees with the run formula
    dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
    cnt, cks = C.c_uint64(), C.c_uint64()
    smj.check(smj.lib.smjdev_merge_join(dr.ptr, nr, ds.ptr, ns, C.byref(cnt), C.byref(cks)))
    assert (cnt.value, cks.value) == exp
    assert smj.lib.merge_join(r.ctypes.data, s.ctypes.data, nr, ns, None) == exp[0]
    dr.free()
    ds.free()


@pytest.mark.parametrize("nr,ns,kmax", [(6_000_000, 5_000_000, 4_000_000), (5_000_000, 4_500_000, 700)])
def test_merge_join_many_tiles(smj, orc, nr, ns, kmax):
    """More than 2048 tiles of the merge-path scan: the boundary records are stitched in two levels
    (k_mj_fixup with a grid, then one CTA over the open records of its threads).  Mostly unique
    keys (about every second tile boundary cuts a key) and 700 heavy keys (runs that span whole
    tiles), count and checksum against the oracle's run formula."""
    rng = np.random.default_rng(nr + kmax)
    r = orc.sort(ol.rand_items(rng, nr, kmax))
    s = orc.sort(ol.rand_items(rng, ns, kmax))
    exp = orc.merge_join_runs(r, s)
    dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
    cnt, cks = C.c_uint64(), C.c_uint64()
    smj.check(smj.lib.smjdev_merge_join(dr.ptr, nr, ds.ptr, ns, C.byref(cnt), C.byref(cks)))
    assert (cnt.value, cks.value) == exp
    dr.free()
    ds.free()


def test_merge_join_interpolation(smj, orc):
    rng = np.random.default_rng(9)
    r = orc.sort(ol.make_unique(rng, 40000))
    s = orc.sort(ol.rand_items(rng, 90000, 40000))
    assert smj.lib.merge_join_interpolation(r.ctypes.data, s.ctypes.data, r.size, s.size, None) == \
        orc.merge_join_interpolation(r, s)
    # duplicates in R: the reference then counts every matching S tuple once
    r2 = orc.sort(ol.rand_items(rng, 40000, 1000))
    s2 = orc.sort(ol.rand_items(rng, 50000, 2000))
    assert smj.lib.merge_join_interpolation(r2.ctypes.data, s2.ctypes.data, r2.size, s2.size, None) == \
        orc.merge_join_interpolation(r2, s2)


# ------------------------------------------------------------------------------------ joins
def _join_cases(orc):
    rng = np.random.default_rng(2024)
    n = 1 << 18
    yield "unique", orc.gen_unique(n, 12345), orc.gen_unique(n, 54321)
    yield "fk", orc.gen_unique(n // 4, 12345), orc.gen_fk(n, n // 4, 99)
    yield "dups_both", ol.rand_items(rng, n, n // 8), ol.rand_items(rng, n, n // 8)
    yield "all_equal", np.full(3000, (7 << 32) | 1, dtype=np.uint64), np.full(5000, (7 << 32) | 2, dtype=np.uint64)
    yield "disjoint", ol.rand_items(rng, 10000, 1000), ol.rand_items(rng, 10000, 1000) + np.uint64(5000 << 32)
    yield "ragged", orc.gen_unique(100003, 1), orc.gen_fk(7, 100003, 2)
    zk = ol.zipf_keys(rng, n, n // 4, 1.0)
    yield "zipf", orc.gen_unique(n // 4, 5), (zk << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    yield "empty_r", np.zeros(0, dtype=np.uint64), orc.gen_unique(1000, 3)
    yield "empty_s", orc.gen_unique(1000, 3), np.zeros(0, dtype=np.uint64)


@pytest.mark.parametrize("algo", ["multiway", "multipass", "mpsm"])
def test_join_host_api(smj, orc, algo):
    """sortmergejoin_* called like src/main.c:603 with host relations: cardinality and checksum
    equal to the oracle's (brute force == the reference's m-way pipeline restated)."""
    for name, r, s in _join_cases(orc):
        exp = orc.join_bruteforce(r, s)
        r0, s0 = r.copy(), s.copy()
        total, st = smj.join_host(algo, r, s, nthreads=4, fanout=128)
        assert total == exp[0], (algo, name)
        assert (st.matches, st.checksum) == exp, (algo, name)
        assert np.array_equal(r, r0) and np.array_equal(s, s0), "inputs must stay unmodified"
        if algo == "multiway" and name in ("unique", "fk", "dups_both"):
            assert orc.join_mway(r, s, 4, 128) == exp


def test_join_error_behaviour(smj, orc, capfd):
    r, s = orc.gen_unique(1000, 1), orc.gen_unique(1000, 2)
    total, _ = smj.join_host("multiway", r, s, nthreads=3)  # not a power of two
    assert total is None
    assert "power of 2" in capfd.readouterr().out


def test_join_device_resident(smj, orc):
    n = 1 << 20
    r, s = orc.gen_unique(n, 12345), orc.gen_fk(2 * n, n, 7)
    exp = orc.join_bruteforce(r, s)
    dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
    for algo, nruns in ((smj.ALGO_MWAY, 1), (smj.ALGO_MPASS, 8), (smj.ALGO_MPASS, 5), (smj.ALGO_MPSM, 8)):
        st = smj.device_join(algo, dr, ds, nruns)
        assert (st.matches, st.checksum) == exp, algo
        assert st.launches > 0
    assert np.array_equal(dr.to_host(), r) and np.array_equal(ds.to_host(), s)


# ------------------------------------------------------------------------------------ generators
def test_generators_match_oracle(smj, orc):
    for n in (1, 2, 1000, 100003, 1 << 20):
        d = smj.DeviceBuffer(n)
        smj.check(smj.lib.smjdev_gen_unique(d.ptr, n, 0, n, 12345))
        got = d.to_host()
        assert np.array_equal(got, orc.gen_unique(n, 12345))
        assert np.array_equal(np.sort(got >> np.uint64(32)), np.arange(1, n + 1, dtype=np.uint64))
        smj.check(smj.lib.smjdev_gen_fk(d.ptr, 0, n, max(n // 3, 1), 77))
        assert np.array_equal(d.to_host(), orc.gen_fk(n, max(n // 3, 1), 77))
        d.free()


@pytest.mark.parametrize("alphabet,theta", [(1, 0.9), (1000, 0.5), (100003, 1.0), (65536, 0.75), (300, 1.05),
                                            (70001, 0.0), (1 << 20, 0.6)])
def test_gen_zipf_matches_oracle(smj, orc, alphabet, theta):
    """smjdev_zipf_lut / smjdev_gen_zipf (genzipf.c:29-159 on the device) bit for bit against the
    oracle's restatement: cumulative table, drawn keys, row ranges, shared table."""
    n = 200_003
    lut = smj.DeviceBuffer(alphabet)
    tmp = smj.DeviceBuffer(alphabet // 256 + alphabet // 65536 + 8)
    smj.check(smj.lib.smjdev_zipf_lut(lut.ptr, alphabet, theta, tmp.ptr))
    got_lut = lut.to_host(dtype=np.float64)
    assert np.array_equal(got_lut.view(np.uint64), orc.zipf_lut(alphabet, theta).view(np.uint64))
    d = smj.DeviceBuffer(n)
    smj.check(smj.lib.smjdev_gen_zipf(d.ptr, 0, n, alphabet, theta, 54321, 12345, None))
    exp = orc.gen_zipf(n, alphabet, theta, 54321, 12345)
    assert np.array_equal(d.to_host(), exp)
    # a row range with a prepared table gives the same rows
    smj.check(smj.lib.smjdev_gen_zipf(d.ptr, 1000, 5000, alphabet, theta, 54321, 12345, lut.ptr))
    assert np.array_equal(d.to_host(5000), exp[1000:6000])
    keys = exp >> np.uint64(32)
    assert keys.min() >= 1 and keys.max() <= alphabet
    for b in (lut, tmp, d):
        b.free()


def test_gen_zipf_joins_like_the_reference_workload(smj, orc):
    """config 4 in small: R = permutation of 1..A, S = Zipf over R's keys: every S tuple finds its
    partner (count = |S|), checksum as the oracle's."""
    a, n = 1 << 20, (1 << 21) + 11
    dr, ds = smj.DeviceBuffer(a), smj.DeviceBuffer(n)
    smj.check(smj.lib.smjdev_gen_unique(dr.ptr, a, 0, a, 12345))
    smj.check(smj.lib.smjdev_gen_zipf(ds.ptr, 0, n, a, 0.9, 54321, 777, None))
    st = smj.device_join(smj.ALGO_MWAY, dr, ds)
    exp = orc.join_bruteforce(orc.gen_unique(a, 12345), orc.gen_zipf(n, a, 0.9, 54321, 777))
    assert (st.matches, st.checksum) == exp and st.matches == n
    dr.free()
    ds.free()


def test_range_partition(smj, orc):
    rng = np.random.default_rng(4)
    n = 300007
    items = ol.rand_items(rng, n, 1 << 20)
    for nparts in (1, 2, 8):
        sp = np.array([(i + 1) * (1 << 20) // nparts for i in range(nparts - 1)], dtype=np.uint32)
        d_in, d_out, d_cnt = smj.DeviceBuffer.from_host(items), smj.DeviceBuffer(n), smj.DeviceBuffer(nparts)
        smj.check(smj.lib.smjdev_range_partition(d_in.ptr, d_out.ptr, n,
                                                 sp.ctypes.data_as(C.POINTER(C.c_uint32)), nparts, d_cnt.ptr))
        cnt, out = d_cnt.to_host(), d_out.to_host()
        dest = np.searchsorted(sp, (items >> np.uint64(32)).astype(np.uint32), side="right")
        assert np.array_equal(cnt, np.bincount(dest, minlength=nparts).astype(np.uint64))
        assert np.array_equal(out, items[np.argsort(dest, kind="stable")])  # stable


# ------------------------------------------------------------------------------------ m-way fast path
def _device_mway(smj, r, s):
    dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
    st = smj.device_join(smj.ALGO_MWAY, dr, ds)
    dr.free()
    ds.free()
    return st.matches, st.checksum


@pytest.mark.parametrize("key_bits", [19, 23, 28, 29, 31])
def test_mway_fast_path_key_domains(smj, orc, key_bits):
    """The MSD fast path (>= 2^20 tuples, 19..31 varying key bits) against the oracle, including the
    CTA-per-bucket variant for domains wider than 2^28."""
    rng = np.random.default_rng(key_bits)
    nr, ns = 1_300_003, 2_700_001
    kmax = (1 << key_bits) - 1
    keys = np.unique(rng.integers(1, kmax + 1, size=int(nr * 1.2), dtype=np.uint64))[:nr]
    rng.shuffle(keys)
    keys[0], keys[1] = np.uint64(1), np.uint64(kmax)  # make the whole domain vary
    r = ol.keys_ok((keys << np.uint64(32)) | np.arange(keys.size, dtype=np.uint64))
    s = ol.keys_ok((rng.choice(keys, size=ns) << np.uint64(32)) | rng.integers(0, 1 << 32, size=ns, dtype=np.uint64))
    assert _device_mway(smj, r, s) == orc.join_bruteforce(r, s)


def test_mway_fast_path_heavy_hitters(smj, orc):
    rng = np.random.default_rng(77)
    n = 1 << 22
    r = orc.gen_unique(n // 4, 5)
    zk = ol.zipf_keys(rng, n, n // 4, 1.0)  # hottest key: ~ n / 14 tuples -> cooperative bucket join
    s = (zk << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    assert np.bincount(zk.astype(np.int64)).max() > (1 << 16)
    exp = orc.join_bruteforce(r, s)
    assert exp[0] == n
    assert _device_mway(smj, r, s) == exp
    # a key that is heavy on BOTH sides is left to the general sort + scan path
    hot = np.full(100_000, (12345 << 32) | 7, dtype=np.uint64)
    r2, s2 = np.concatenate([r, hot]), np.concatenate([s, hot, hot])
    assert _device_mway(smj, r2, s2) == orc.join_bruteforce(r2, s2)


def test_mway_fast_path_wrong_sample_guess(smj, orc):
    """Two outliers far outside the sampled key range: the sampled plan must be rejected by the
    exact OR/AND check and the general path must take over."""
    n = 1 << 21
    r, s = orc.gen_unique(n, 12345), orc.gen_unique(n, 54321)
    r, s = r.copy(), s.copy()
    r[n // 3 + 1] = np.uint64(((1 << 30) + 5) << 32 | 1)
    s[n // 5 + 7] = np.uint64(((1 << 30) + 5) << 32 | 2)
    assert _device_mway(smj, r, s) == orc.join_bruteforce(r, s)


def _device_mway_stats(smj, r, s):
    dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
    st = smj.device_join(smj.ALGO_MWAY, dr, ds)
    dr.free()
    ds.free()
    return st


def test_mpsm_one_pass_probe_and_its_fallbacks(smj, orc, monkeypatch):
    """P-MPSM joins all sorted runs of S against the sorted R in ONE pass over R (csrc/smj_mpsm.cu:
    R tiles that end on key boundaries, the pieces of every run looked up by binary search).  Same
    count and checksum as the oracle and as one merge-path scan per run (SMJ_MPSM_PER_RUN=1) for:
    a primary-key R, short runs of duplicates in R (walked inside a tile), an R run longer than a
    tile (flag -> per-run scans), a hot key of S that would land on one CTA (flag -> per-run scans),
    S keys below / above R's key range, and more runs than items in a piece."""
    rng = np.random.default_rng(77)
    n = 400_000
    pk = orc.gen_unique(n, 3)
    cases = {
        "pk_fk": (pk, orc.gen_fk(3 * n + 11, n, 4)),
        "dups_r": (ol.rand_items(rng, n, n // 6), ol.rand_items(rng, 2 * n, n // 6)),
        "long_run_r": (np.concatenate([pk[: n // 2], np.full(6000, (np.uint64(77) << np.uint64(32)) | np.uint64(9))]),
                       orc.gen_fk(n, n, 5)),
        "outside": (pk + (np.uint64(1000) << np.uint64(32)), ol.rand_items(rng, n, n + 3000)),
    }
    hot = orc.gen_fk(1_300_000, n, 6)
    hot[: 1_200_000] = (hot[: 1_200_000] & np.uint64(0xFFFFFFFF)) | (np.uint64(4242) << np.uint64(32))
    cases["hot_s"] = (pk, hot)
    for name, (r, s) in cases.items():
        exp = orc.join_bruteforce(r, s)
        for nruns in (4, 64):
            for per_run in (False, True):
                if per_run:
                    monkeypatch.setenv("SMJ_MPSM_PER_RUN", "1")
                dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
                st = smj.device_join(smj.ALGO_MPSM, dr, ds, nruns)
                dr.free()
                ds.free()
                monkeypatch.delenv("SMJ_MPSM_PER_RUN", raising=False)
                assert (st.matches, st.checksum) == exp, (name, nruns, per_run)


def test_mway_optimistic_partition_and_its_exact_repeats(smj, orc, monkeypatch):
    """The histogram-free partition passes (sampled bin capacities, fixed second-level slots):
    same count and checksum as the exact passes and as the oracle; when a region overflows (forced
    here by under-sizing the regions, SMJ_MSD_TEST) the pass is repeated exactly and the result
    does not change."""
    n = 3_000_017
    r, s = orc.gen_unique(n, 12345), orc.gen_fk(2 * n + 5, n, 7)
    exp = orc.join_bruteforce(r, s)
    st = _device_mway_stats(smj, r, s)
    assert (st.matches, st.checksum) == exp
    # only the partly filled first-level bucket at the top of the dense key domain (one per
    # relation) may need its second level repeated
    assert st.path == smj.PATH_MSD_OPT and st.repeats <= 2, st.repeats
    monkeypatch.setenv("SMJ_MSD_MODE", "exact")
    st = _device_mway_stats(smj, r, s)
    assert (st.matches, st.checksum, st.path, st.repeats) == (*exp, smj.PATH_MSD, 0)
    monkeypatch.delenv("SMJ_MSD_MODE")
    monkeypatch.setenv("SMJ_MSD_TEST", "nocoop")  # no cooperative launch: sampled bits + exact passes
    st = _device_mway_stats(smj, r, s)
    assert (st.matches, st.checksum, st.path, st.repeats) == (*exp, smj.PATH_MSD, 0)
    for hook in ("shrink1", "shrink2"):
        monkeypatch.setenv("SMJ_MSD_TEST", hook)
        st = _device_mway_stats(smj, r, s)
        assert (st.matches, st.checksum) == exp, hook
        # shrink1: both relations overflow at level 1 and are repartitioned exactly by the host;
        # shrink2: every first-level bucket of R (unique keys: equal slots; S has duplicate keys
        # and takes the exact second level anyway) is repartitioned exactly on the device
        assert st.path == smj.PATH_MSD_OPT, hook
        assert st.repeats == 2 if hook == "shrink1" else st.repeats > 300, (hook, st.repeats)
    monkeypatch.delenv("SMJ_MSD_TEST")


def test_mway_optimistic_partition_input_orders(smj, orc):
    """Sampled capacities must hold for inputs that are not in random order: sorted, reverse
    sorted, sorted blocks, and a relation too small to be sampled sparsely."""
    n = 2_500_003
    base = orc.gen_unique(n, 99)
    s = orc.gen_fk(n + 11, n, 5)
    exp = orc.join_bruteforce(base, s)
    blocks = np.concatenate([np.sort(c) for c in np.array_split(base, 37)])
    for name, r in (("sorted", np.sort(base)), ("reverse", np.sort(base)[::-1].copy()), ("blocks", blocks)):
        st = _device_mway_stats(smj, r, s)
        assert (st.matches, st.checksum) == exp, name
        assert st.path == smj.PATH_MSD_OPT, name
    # second-level skew that no single heavy key explains: S keys concentrated on one key range
    rng = np.random.default_rng(5)
    hot = rng.integers(1000, 1000 + 4096, size=n, dtype=np.uint64)
    s2 = np.concatenate([s, (hot << np.uint64(32)) | np.arange(n, dtype=np.uint64)])
    rng.shuffle(s2)
    st = _device_mway_stats(smj, base, s2)
    assert (st.matches, st.checksum) == orc.join_bruteforce(base, s2)


@pytest.mark.parametrize("seed", range(10))
def test_mway_fast_path_fuzz(smj, orc, seed):
    """Randomly shaped relations through the optimistic m-way path against the oracle: sizes around
    the 2^20 threshold, 19..31 key bits, unique / duplicated / Zipf / clustered / partly sorted keys.
    Whatever the passes guess, count and checksum must be the oracle's."""
    rng = np.random.default_rng(1000 + seed)
    bits = int(rng.integers(19, 32))
    kmax = (1 << bits) - 1
    nr = int(rng.integers(1 << 19, 3 << 20))
    ns = int(rng.integers(1 << 19, 5 << 20))

    def keys(n, kind):
        if kind == "unique" and min(kmax, 8 * n) >= n:
            k = rng.permutation(np.arange(1, min(kmax, 8 * n) + 1, dtype=np.uint64))[:n]
        elif kind == "uniform":
            k = rng.integers(1, kmax + 1, size=n, dtype=np.uint64)
        elif kind == "zipf":
            k = ol.zipf_keys(rng, n, min(kmax, n), float(rng.uniform(0.3, 1.1)))
        elif kind == "clustered":  # most keys inside a few narrow ranges
            centres = rng.integers(1, kmax + 1, size=5, dtype=np.uint64)
            k = (centres[rng.integers(0, 5, size=n)] + rng.integers(0, 4096, size=n, dtype=np.uint64)) % np.uint64(kmax) + np.uint64(1)
        else:  # sorted blocks
            k = np.sort(rng.integers(1, kmax + 1, size=n, dtype=np.uint64))
        k[0], k[-1] = np.uint64(1), np.uint64(kmax)  # the whole key domain varies in both relations
        return k

    kr = keys(nr, ["unique", "uniform", "sorted"][seed % 3])
    ks = keys(ns, ["uniform", "zipf", "clustered", "unique", "sorted"][seed % 5])
    r = ol.keys_ok((kr << np.uint64(32)) | rng.integers(0, 1 << 32, size=nr, dtype=np.uint64))
    s = ol.keys_ok((ks << np.uint64(32)) | np.arange(ns, dtype=np.uint64))
    st = _device_mway_stats(smj, r, s)
    assert (st.matches, st.checksum) == orc.join_bruteforce(r, s), (seed, bits, nr, ns, st.path, st.repeats)


def test_join_pinned_host_relations_overlap_upload(smj, orc, monkeypatch):
    """Pinned host relations go to smjdev_join as they are: uploaded in 64 MiB chunks on the copy
    stream while the first partitioning pass consumes the chunks (several chunks per relation
    here).  Same result as the device-resident join and the oracle, for every algorithm, through
    the device layer and through sortmergejoin_multiway()."""
    nr, ns = 17_000_001, 9_000_001
    r, s = orc.gen_unique(nr, 12345), orc.gen_fk(ns, nr, 7)
    exp = orc.join_bruteforce(r, s)
    assert exp[0] == ns
    lib = smj.lib
    hr, hs = lib.smjdev_host_alloc(nr * 8), lib.smjdev_host_alloc(ns * 8)
    assert hr and hs
    C.memmove(hr, r.ctypes.data, nr * 8)
    C.memmove(hs, s.ctypes.data, ns * 8)
    try:
        for algo, nruns in ((smj.ALGO_MWAY, 1), (smj.ALGO_MPASS, 4), (smj.ALGO_MPSM, 2)):
            st = smj.smj_join_stats_t()
            smj.check(lib.smjdev_join(algo, hr, nr, hs, ns, nruns, C.byref(st)), "smjdev_join")
            assert (st.matches, st.checksum) == exp, algo
            assert st.ms_h2d > 0
            if algo == smj.ALGO_MWAY:
                assert st.path == smj.PATH_MSD_OPT
        # small pinned relations (below the fast path's threshold: general path behind the upload)
        st = smj.smj_join_stats_t()
        smj.check(lib.smjdev_join(smj.ALGO_MWAY, hr, 5000, hs, 3000, 1, C.byref(st)), "smjdev_join")
        assert (st.matches, st.checksum) == orc.join_bruteforce(r[:5000], s[:3000]) and st.path == smj.PATH_GENERAL
        # mixed: R on the device, S pinned; and the exact passes (the upload is awaited first)
        dr = smj.DeviceBuffer.from_host(r)
        st = smj.smj_join_stats_t()
        smj.check(lib.smjdev_join(smj.ALGO_MWAY, dr.ptr, nr, hs, ns, 1, C.byref(st)), "smjdev_join")
        assert (st.matches, st.checksum) == exp
        monkeypatch.setenv("SMJ_MSD_MODE", "exact")
        smj.check(lib.smjdev_join(smj.ALGO_MWAY, hr, nr, hs, ns, 1, C.byref(st)), "smjdev_join")
        assert (st.matches, st.checksum, st.path) == (*exp, smj.PATH_MSD)
        monkeypatch.delenv("SMJ_MSD_MODE")
        monkeypatch.setenv("SMJ_MSD_TEST", "shrink1")  # overflow -> exact repeat from the uploaded copy
        smj.check(lib.smjdev_join(smj.ALGO_MWAY, hr, nr, hs, ns, 1, C.byref(st)), "smjdev_join")
        assert (st.matches, st.checksum, st.repeats) == (*exp, 2)
        monkeypatch.delenv("SMJ_MSD_TEST")
        dr.free()
        # the reference-facing entry point with the same pinned buffers
        cfg = smj.joinconfig_t(1, 128, 0, 0, 20 * 1024 * 1024, smj.NUMA_NEXT)
        relr, rels = smj.relation_t(hr, nr), smj.relation_t(hs, ns)
        res = lib.sortmergejoin_multiway(C.byref(relr), C.byref(rels), C.byref(cfg))
        assert res and res.contents.totalresults == ns
        hst = smj.smj_join_stats_t()
        lib.smj_last_join_stats(C.byref(hst))
        assert (hst.matches, hst.checksum) == exp and hst.ms_h2d > 0
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(C.cast(res.contents.resultlist, C.c_void_p))
        libc.free(C.cast(res, C.c_void_p))
        assert np.array_equal(np.ctypeslib.as_array(C.cast(hr, C.POINTER(C.c_uint64)), shape=(nr,)), r)
    finally:
        lib.smjdev_host_free(hr)
        lib.smjdev_host_free(hs)


def test_multiway_merge_many_runs_via_partition(smj, orc, monkeypatch):
    """Eight or more runs in whole-word order with a large total are merged by range-partitioning
    their union (the MSD sort) instead of log2(k) merge rounds: same output as the oracle's merge,
    with adjacent and with scattered runs, and the rounds still run when the shortcut is off."""
    rng = np.random.default_rng(8)
    nruns, per = 16, 150_001
    runs = [np.sort(ol.rand_items(rng, per + 13 * i, 1 << 24)) for i in range(nruns)]
    exp = orc.multiway_merge(runs)
    total = exp.size
    bufs = [smj.DeviceBuffer.from_host(x) for x in runs]
    out, tmp = smj.DeviceBuffer(total), smj.DeviceBuffer(total)
    ptrs = (C.c_void_p * nruns)(*[b.ptr for b in bufs])
    lens = (C.c_uint64 * nruns)(*[x.size for x in runs])
    monkeypatch.setenv("SMJ_MERGE_PATH", "sort")
    for path in ("msd", "lsd"):  # lsd: the sort declines, the merge rounds run
        monkeypatch.setenv("SMJ_SORT_PATH", path)
        smj.check(smj.lib.smjdev_multiway_merge(out.ptr, tmp.ptr, ptrs, lens, nruns, 0), "multiway_merge")
        assert np.array_equal(out.to_host(), exp), path
    monkeypatch.setenv("SMJ_SORT_PATH", "msd")
    # runs that lie back to back in one buffer are sorted where they are (no gathering)
    cat = smj.DeviceBuffer.from_host(np.concatenate(runs))
    offs = np.concatenate([[0], np.cumsum([x.size for x in runs])])
    ptrs2 = (C.c_void_p * nruns)(*[cat.ptr + 8 * int(o) for o in offs[:-1]])
    smj.check(smj.lib.smjdev_multiway_merge(out.ptr, tmp.ptr, ptrs2, lens, nruns, 0), "multiway_merge")
    assert np.array_equal(out.to_host(), exp)
    assert np.array_equal(cat.to_host(), np.concatenate(runs)), "inputs must stay unmodified"
    monkeypatch.setenv("SMJ_MERGE_PATH", "rounds")
    smj.check(smj.lib.smjdev_multiway_merge(out.ptr, tmp.ptr, ptrs, lens, nruns, 0), "multiway_merge")
    assert np.array_equal(out.to_host(), exp)
    for b in bufs + [out, tmp, cat]:
        b.free()


@pytest.mark.parametrize("nruns", [3, 17, 128, 256, 300])
def test_kway_merge_one_pass(smj, orc, nruns, monkeypatch):
    """smjdev_multiway_merge takes the one-pass k-way merge (csrc/smj_kway.cu, the counterpart of
    avx_multiway_merge, src/merge/avx_multiwaymerge.c:199-338): value-range tiles from regular
    samples, gathered (three gather shapes: k <= 8, k < 64, k >= 64), put in order in shared memory
    by the interpolating counting sort -- or, for tiles whose values are not evenly spread (the
    "disjoint" shape) and under SMJ_KWAY_TILE=sort, by the comparison sort -- and written once.
    Same output as the oracle for runs of very different lengths, empty runs, heavy duplicates
    (tiles overflow -> another path takes over) and more than 256 runs (declined)."""
    monkeypatch.setenv("SMJ_MERGE_PATH", "kway")
    rng = np.random.default_rng(nruns)
    lens = [int(x) for x in rng.integers(0, 2 * 700_000 // nruns, size=nruns)]
    lens[0], lens[-1] = 0, 400_000 // nruns * 3 + 5
    shapes = {
        "uniform": [np.sort(ol.rand_items(rng, n, 1 << 27)) for n in lens],
        "narrow": [np.sort(ol.rand_items(rng, n, 50, payload_bits=4)) for n in lens],   # piles of equal words
        "disjoint": [np.sort(ol.rand_items(rng, n, 1 << 16) + (np.uint64(i) << np.uint64(50))) for i, n in enumerate(lens)],
    }
    for name, runs in shapes.items():
        exp = orc.multiway_merge(runs)
        bufs = [smj.DeviceBuffer.from_host(x) for x in runs]
        out, tmp = smj.DeviceBuffer(exp.size), smj.DeviceBuffer(exp.size)
        ptrs = (C.c_void_p * nruns)(*[b.ptr for b in bufs])
        ln = (C.c_uint64 * nruns)(*[x.size for x in runs])
        for tile in ("count", "sort"):
            monkeypatch.setenv("SMJ_KWAY_TILE", tile)
            smj.check(smj.lib.smjdev_memset(out.ptr, 0, exp.size * 8), "memset")
            smj.check(smj.lib.smjdev_multiway_merge(out.ptr, tmp.ptr, ptrs, ln, nruns, 0), name)
            assert np.array_equal(out.to_host(), exp), (name, tile)
        monkeypatch.delenv("SMJ_KWAY_TILE")
        for b, x in zip(bufs, runs):
            assert np.array_equal(b.to_host(), x), "inputs must stay unmodified"
        for b in bufs + [out, tmp]:
            b.free()


# ------------------------------------------------------------------------------------ materialisation
def _materialize(smj, r, s, layout):
    dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
    cnt = C.c_uint64(0)
    smj.check(smj.lib.smjdev_join_materialize(dr.ptr, r.size, ds.ptr, s.size, layout, None, 0, C.byref(cnt)))
    out = smj.DeviceBuffer(max(cnt.value, 1))
    cnt2 = C.c_uint64(0)
    smj.check(smj.lib.smjdev_join_materialize(dr.ptr, r.size, ds.ptr, s.size, layout, out.ptr, cnt.value,
                                              C.byref(cnt2)))
    assert cnt2.value == cnt.value
    res = out.to_host(cnt.value, 0) if cnt.value else np.zeros(0, dtype=np.uint64)
    for b in (dr, ds, out):
        b.free()
    return res


@pytest.mark.parametrize("layout", [0, 1])
def test_join_materialize(smj, orc, layout, monkeypatch):
    """smjdev_join_materialize against the oracle's restatement of the JOIN_MATERIALIZE loop
    (joincommon.c:256-278): same multiset of emitted tuples, for unique / foreign-key / duplicate /
    heavy-run / disjoint / empty inputs."""
    for name, r, s in _join_cases(orc):
        if name == "all_equal":
            r, s = r[:300], s[:500]
        exp = orc.join_materialize(r, s, layout)
        got = _materialize(smj, r, s, layout)
        assert got.size == exp.size == orc.join_bruteforce(r, s)[0], name
        assert np.array_equal(np.sort(got), np.sort(exp)), name
    # long R runs (duplicates in R beyond the shared-memory staging) and a larger foreign-key join
    rng = np.random.default_rng(3)
    r = ol.rand_items(rng, 200_000, 40)
    s = ol.rand_items(rng, 3_000, 50)
    assert np.array_equal(np.sort(_materialize(smj, r, s, layout)), np.sort(orc.join_materialize(r, s, layout)))
    r, s = orc.gen_unique(1 << 20, 5), orc.gen_fk((1 << 21) + 7, 1 << 20, 9)
    monkeypatch.setenv("SMJ_SORT_PATH", "msd")  # S (2^21 items) is sorted MSD-first, R LSD
    got = _materialize(smj, r, s, layout)
    monkeypatch.delenv("SMJ_SORT_PATH")
    assert got.size == s.size
    assert np.array_equal(np.sort(got), np.sort(orc.join_materialize(r, s, layout)))


def test_join_materialize_host_api(smj, orc):
    """sortmergejoin_multiway_materialize: resultlist[0].results holds totalresults tuples
    <S.key, S.payload>, malloc()ed, freed by the caller; partial output when the buffer is small."""
    r, s = orc.gen_unique(50_000, 1), orc.gen_fk(120_001, 50_000, 2)
    cfg = smj.joinconfig_t(2, 128, 0, 0, 20 * 1024 * 1024, smj.NUMA_NEXT)
    res = smj.lib.sortmergejoin_multiway_materialize(C.byref(smj.rel(r)), C.byref(smj.rel(s)), C.byref(cfg))
    assert res and res.contents.totalresults == s.size and res.contents.nthreads == 2
    tr = res.contents.resultlist[0]
    assert tr.nresults == s.size
    got = np.ctypeslib.as_array(C.cast(tr.results, C.POINTER(C.c_uint64)), shape=(s.size,)).copy()
    assert np.array_equal(np.sort(got), np.sort(s))  # every S tuple has exactly one partner
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(C.cast(tr.results, C.c_void_p))
    libc.free(C.cast(res.contents.resultlist, C.c_void_p))
    libc.free(C.cast(res, C.c_void_p))
    # capacity smaller than the result: the count is complete, the output is cut
    dr, ds, out = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s), smj.DeviceBuffer(1000)
    cnt = C.c_uint64(0)
    smj.check(smj.lib.smjdev_join_materialize(dr.ptr, r.size, ds.ptr, s.size, 0, out.ptr, 1000, C.byref(cnt)))
    assert cnt.value == s.size
    part = out.to_host()
    assert np.all(np.isin(part, s))
    for b in (dr, ds, out):
        b.free()


# ------------------------------------------------------------------------------------ host memory paths
def test_pageable_relations_use_the_staging_pipeline(smj, orc):
    """Relations in ordinary malloc'ed memory above the staging threshold (32 MiB): upload through
    the threaded bounce-buffer pipeline, sort results downloaded the same way."""
    nr, ns = 4_500_000, 6_000_001
    r, s = orc.gen_unique(nr, 7), orc.gen_fk(ns, nr, 8)
    assert r.nbytes > 32 << 20 and s.nbytes > 32 << 20
    total, st = smj.join_host("multiway", r, s, nthreads=2)
    assert total == ns
    exp = orc.join_bruteforce(r, s)
    assert (st.matches, st.checksum) == exp
    a, b = s.copy(), np.zeros(ns, dtype=np.uint64)
    pa, pb = C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data)
    smj.lib.avxsort_tuples(C.byref(pa), C.byref(pb), ns)
    res = b if pb.value == b.ctypes.data else a
    assert np.array_equal(res, np.sort(s))


def _cli(smj):
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(smj.LIB_PATH), "sortmergejoins")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.dirname(smj.LIB_PATH), exe], check=True, capture_output=True)
    env = dict(os.environ)
    env.pop("SMJ_QUIET", None)
    return exe, env


def test_reference_style_cli(smj):
    """The sortmergejoins command line driver (src/main.c) on top of the library: short options,
    the reference's long options (src/main.c:716-745) and its data set switches."""
    import subprocess
    exe, env = _cli(smj)
    for algo, extra in (("m-way", []), ("m-pass", ["-n", "4"]), ("mpsm", ["-n", "4"]), ("m-way", ["-d"]),
                        ("m-way", ["-s", "3000000"])):
        out = subprocess.run([exe, "-a", algo, "-r", "1000000", "-s", "1000000"] + extra, capture_output=True,
                             text=True, env=env, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        n_s = 3000000 if "-s" in extra else 1000000
        assert f"Results = {n_s}. DONE." in out.stdout, out.stdout
        assert "TUPLES-PER-SECOND" in out.stderr  # the reference's timing line, src/joins/joincommon.c:214-227
    # the call the reference's scripts make (scripts/tput-scalability.sh:32,53)
    for extra, expect in ((["--scalarmerge", "--scalarsort"], 1500000),
                          (["--skew=0.9", "--s-size=4000000"], 4000000),          # every Zipf key is a key of R
                          (["--partfanout=256", "--numastrategy=RING", "--mwaybufsize=1048576", "--verbose"], 1500000),
                          (["--device-resident", "--r-seed=7", "--s-seed=8", "--no-numa-localize"], 1500000)):
        out = subprocess.run([exe, "--algo=m-way", "--nthreads=4", "--r-size=1500000", "--s-size=1500000"] + extra,
                             capture_output=True, text=True, env=env, timeout=300)
        assert out.returncode == 0 and f"Results = {expect}. DONE." in out.stdout, out.stdout + out.stderr
    # --non-unique / --full-range: the result is whatever the key multisets give; it must be the
    # same from all three algorithms (and --full-range S holds R's rows, so every S tuple matches)
    for flag in ("--non-unique", "--full-range"):
        res = set()
        for algo in ("m-way", "m-pass", "mpsm"):
            out = subprocess.run([exe, f"--algo={algo}", "--nthreads=2", "--r-size=1200000", "--s-size=2500000", flag],
                                 capture_output=True, text=True, env=env, timeout=300)
            assert out.returncode == 0, out.stdout + out.stderr
            res.add([ln for ln in out.stdout.splitlines() if "Results =" in ln][0])
        assert len(res) == 1, res
        if flag == "--full-range":
            assert int(res.pop().split("=")[1].split(".")[0]) >= 2500000
    bad = subprocess.run([exe, "-n", "3", "-r", "1000", "-s", "1000"], capture_output=True, text=True, env=env)
    assert "power of 2" in bad.stdout and "Results" not in bad.stdout  # src/joins/sortmergejoin_multiway.c:54-57
    bad = subprocess.run([exe, "--partfanout=100"], capture_output=True, text=True, env=env)
    assert "must be a power of 2" in bad.stdout
    bad = subprocess.run([exe, "--nthreads=256", "--partfanout=128"], capture_output=True, text=True, env=env)
    assert "must be >= num-threads" in bad.stdout
    bad = subprocess.run([exe, "--algo=nope"], capture_output=True, text=True, env=env)
    assert "does not exist" in bad.stdout


def test_experiment_script_runs_unmodified_options(smj, tmp_path):
    """scripts/tput-scalability.sh (the reference's sweep: --algo= --nthreads= --r-size= --s-size=
    [--scalarmerge --scalarsort], stderr scraped into the results table) against this build."""
    import os
    import subprocess
    exe, env = _cli(smj)
    root = os.path.dirname(os.path.dirname(smj.LIB_PATH))
    env.update(BIN=exe, RESULTS=str(tmp_path), NTUPLES="1000000", THREADS="4 1", REPS="1", OUTPUTFILE="tput-test")
    out = subprocess.run(["bash", os.path.join(root, "scripts", "tput-scalability.sh")], env=env, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    rows = [ln.split() for ln in open(tmp_path / "tput-test.txt") if ln and ln[0] not in "#="]
    rows = [r for r in rows if r]
    assert len(rows) == 3 * 2 * 2  # algorithms x thread counts x {avx, scalar}
    for r in rows:
        # <algo> <nthreads> <rep> five stage times, then NUM-TUPLES = n TOTAL-TIME-USECS = t TUPLES-PER-SECOND = x
        assert r[0].split("-scalar")[0] in ("m-pass", "m-way", "mpsm") and r[1] in ("4", "1"), r
        assert "NUM-TUPLES" in r and r[r.index("NUM-TUPLES") + 2] == "1000000", r
        assert float(r[r.index("TUPLES-PER-SECOND") + 2]) > 0
    log = open(tmp_path / "tput-test.log").read()
    assert log.count("Results = 1000000. DONE.") == 12


def test_dropin_reference_main(smj):
    """The reference's OWN driver (src/main.c, compiled unmodified by `make -C oracle dropin` with its
    generators and cpu mapping) linked against libsmj_b200.so: its joins run on the GPU."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(smj.LIB_PATH))
    exe = os.path.join(root, "oracle", "_ref", "sortmergejoins_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/sortmergejoins_dropin not built (needs /root/reference: make -C oracle dropin)")
    env = dict(os.environ)
    env.pop("SMJ_QUIET", None)
    for args, expect in ((["--algo=m-way", "--nthreads=4", "--r-size=2000000", "--s-size=2000000"], 2000000),
                         (["--algo=m-pass", "--nthreads=2", "--r-size=1000000", "--s-size=3000000"], 3000000),
                         (["--algo=mpsm", "--nthreads=2", "--r-size=1000000", "--s-size=1000000"], 1000000),
                         (["--algo=m-way", "--nthreads=2", "--r-size=1000000", "--s-size=2000000", "--skew=0.8"], 2000000)):
        out = subprocess.run([exe] + args, capture_output=True, text=True, env=env, timeout=600, cwd=root)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert f"Results = {expect}. DONE." in out.stdout, out.stdout[-2000:]
        assert "TOTAL-TIME-USECS" in out.stderr


def test_remaining_exports(smj, orc, capfd):
    """Entry points no other test calls: sortmergejoin_initrun (joincommon.h:59-61), the _modulo /
    _bitand twins of the scalar multiway merge (scalar_multiwaymerge.h:47-78), avxsortmultiway_int64,
    scalarsort_int64, avx_merge_int64 / scalar_merge_int64, print_timing, is_sorted_helper, check_sorted."""
    rng = np.random.default_rng(77)
    r, s = orc.gen_unique(300_000, 3), orc.gen_fk(700_001, 300_000, 4)
    cfg = smj.joinconfig_t(4, 128, 1, 1, 1 << 20, smj.NUMA_RING)  # SCALARSORT / SCALARMERGE set: same result
    res = smj.lib.sortmergejoin_initrun(C.byref(smj.rel(r)), C.byref(smj.rel(s)), C.byref(cfg), None)
    assert res and res.contents.totalresults == s.size and res.contents.nthreads == 4
    st = smj.smj_join_stats_t()
    assert smj.lib.smj_last_join_stats(C.byref(st)) == 0
    assert (st.matches, st.checksum) == orc.join_bruteforce(r, s)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(C.cast(res.contents.resultlist, C.c_void_p))
    libc.free(C.cast(res, C.c_void_p))
    # scalar multiway merge twins: key-sorted, same multiset, inputs consumed
    runs = [np.sort(ol.rand_items(rng, int(rng.integers(1, 5000)), 1 << 14)) for _ in range(8)]
    total = sum(x.size for x in runs)
    for name in ("scalar_multiway_merge_modulo", "scalar_multiway_merge_bitand"):
        rels = (smj.relation_t * 8)(*[smj.relation_t(x.ctypes.data, x.size) for x in runs])
        pp = (C.POINTER(smj.relation_t) * 8)(*[C.pointer(rels[i]) for i in range(8)])
        out = np.zeros(total, dtype=np.uint64)
        assert getattr(smj.lib, name)(out.ctypes.data, pp, 8, None, 0) == total
        assert np.all(np.diff((out >> np.uint64(32)).astype(np.int64)) >= 0), name
        assert np.array_equal(np.sort(out), orc.multiway_merge(runs)), name
        assert all(rels[i].num_tuples == 0 for i in range(8))
    # int64 twins of the sorts and merges
    n = 70_001
    items = ol.rand_items(rng, n, 1 << 27)
    for name, mw in (("avxsortmultiway_int64", 1), ("scalarsort_int64", None)):
        a, b = items.copy(), np.zeros(n, dtype=np.uint64)
        pa, pb = C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data)
        getattr(smj.lib, name)(C.byref(pa), C.byref(pb), n)
        placement = 0 if mw is None else orc.placement(n, mw)  # scalarsort sorts in place and swaps
        assert pb.value == (b.ctypes.data if placement else a.ctypes.data), name
        assert np.array_equal((b if placement else a), orc.sort(items)), name
    a, b = np.sort(ol.rand_items(rng, 5000, 99)), np.sort(ol.rand_items(rng, 3001, 99))
    out = np.zeros(8001, dtype=np.uint64)
    for name in ("avx_merge_int64", "scalar_merge_int64"):
        out[:] = 0
        assert getattr(smj.lib, name)(a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size, b.size) == 8001
        assert np.array_equal(out, orc.merge(a, b, 0)), name
    # host-side helpers
    srt = orc.sort(items)
    assert smj.lib.is_sorted_helper(srt.ctypes.data, n) == 1
    assert smj.lib.is_sorted_helper(items.ctypes.data, n) == 0
    capfd.readouterr()
    smj.lib.check_sorted(srt.ctypes.data, items.ctypes.data, n, n, 3)
    C.CDLL(None).fflush(None)
    txt = capfd.readouterr().out
    assert "3-thread -> R is sorted" in txt and "3-thread -> S is NOT sorted" in txt


def test_gen_fk_from_and_nan_keys(smj, orc):
    """smjdev_gen_fk_from (create_relation_fk_from_pk, generator.c:458-488) and the avoid_NaN rule
    of smjdev_gen_fk for full-range keys (generator.c:107-121) against the oracle."""
    nr, ns = 100_003, 350_001
    r = orc.gen_fk(nr, (1 << 31) - 1, 5)
    assert not np.any((r & np.uint64(0x7FF << 52)) == np.uint64(0x7FF << 52))
    dr, ds = smj.DeviceBuffer(nr), smj.DeviceBuffer(ns)
    smj.check(smj.lib.smjdev_gen_fk(dr.ptr, 0, nr, (1 << 31) - 1, 5))
    assert np.array_equal(dr.to_host(), r)
    smj.check(smj.lib.smjdev_gen_fk_from(ds.ptr, ns, 0, ns, dr.ptr, nr, 9))
    s = ds.to_host()
    assert np.array_equal(s, orc.gen_fk_from(ns, r, 9))
    vals, cnt = np.unique(s, return_counts=True)
    assert np.array_equal(vals, np.unique(r)) and cnt.min() >= ns // nr  # every row of R 3 or 4 times
    dr.free()
    ds.free()


@pytest.mark.parametrize("n", [2, 31, 32, 33, 511, 1000, 4097, 8192, 16383, 16384])
def test_small_sort_is_one_launch(smj, orc, n, monkeypatch):
    """Inputs up to the reference's 16384-item in-cache block (src/avxsort/avxsort.c:17,
    avxsort_core.h:1213-1274, 2056-2214) are sorted by ONE kernel launch (register sorting network +
    merge-path rounds in shared memory), whole-word and stable key-only; the radix passes give the
    same result."""
    rng = np.random.default_rng(n)
    for kind in ("unique", "dups", "equal", "reverse", "wide"):
        items = _inputs(rng, n, kind)
        for key_only, exp in ((0, orc.sort(items)), (1, orc.sort_key_stable(items))):
            buf, tmp = smj.DeviceBuffer.from_host(items), smj.DeviceBuffer(n)
            res = C.c_void_p()
            before = smj.lib.smjdev_launch_count()
            smj.check(smj.lib.smjdev_sort(buf.ptr, tmp.ptr, n, key_only, 1, C.byref(res)), "sort")
            assert smj.lib.smjdev_launch_count() - before == 1
            smj.check(smj.lib.smjdev_sync())
            got = (buf if res.value == buf.ptr else tmp).to_host()
            assert np.array_equal(got, exp), (kind, key_only)
            monkeypatch.setenv("SMJ_NO_SMALL_SORT", "1")
            buf2, tmp2 = smj.DeviceBuffer.from_host(items), smj.DeviceBuffer(n)
            smj.check(smj.lib.smjdev_sort(buf2.ptr, tmp2.ptr, n, key_only, 1, C.byref(res)), "sort")
            smj.check(smj.lib.smjdev_sync())
            assert np.array_equal((buf2 if res.value == buf2.ptr else tmp2).to_host(), exp), (kind, key_only, "radix")
            monkeypatch.delenv("SMJ_NO_SMALL_SORT")
            for b in (buf, tmp, buf2, tmp2):
                b.free()


def test_run_sorts_share_one_bit_probe(smj, orc):
    """m-pass / mpsm with many runs: the varying key bits are probed once per relation, tiny runs go
    through the one-CTA sort; same {count, checksum} as the oracle for 1 .. 64 runs."""
    rng = np.random.default_rng(12)
    r, s = ol.rand_items(rng, 300_001, 40_000), ol.rand_items(rng, 500_003, 40_000)
    exp = orc.join_bruteforce(r, s)
    dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
    for algo in (smj.ALGO_MPASS, smj.ALGO_MPSM):
        for runs in (1, 2, 8, 64):
            st = smj.device_join(algo, dr, ds, nruns=runs)
            assert (st.matches, st.checksum) == exp, (algo, runs)
    dr.free()
    ds.free()


def test_tputbench_tool(smj):
    """tools/tputbench.py, the counterpart of the reference's tputbench / numabench
    (src/bench/tputbench.c:249-1170: per-phase throughput in its TPUT / MEMTPUT convention + the copy
    paths), runs as a program and reports every phase."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(smj.LIB_PATH))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "tputbench.py"), "--n", "4000000", "--reps", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    for phase in ("partition_7bits", "sort", "multiway_merge_128_runs", "multiway_merge_128_runs_rounds",
                  "multiway_merge_128_runs_sort",
                  "merge_join_scan", "mway_join_total"):
        assert res["phases"][phase]["TPUT_Mtuples_s"] > 0 and res["phases"][phase]["MEMTPUT_GBs"] > 0, phase
    for path in ("hbm_d2d_GBs_read_plus_write", "pcie_h2d_pinned_GBs", "pcie_d2h_pinned_GBs", "pcie_h2d_pageable_staged_GBs"):
        assert res["copies"][path] > 0, path


# ------------------------------------------------------------------------------------ KEY_8B (16-byte tuples)
def _key8b_cases(rng):
    n = 200_003
    uniq = np.empty((n, 2), dtype=np.int64)
    uniq[:, 1] = rng.permutation(n) + 1
    uniq[:, 0] = np.arange(n)
    yield "fk", uniq, ol.rand_tuples16(rng, 3 * n + 1, 1, n)
    yield "dups", ol.rand_tuples16(rng, n, 1, n // 50), ol.rand_tuples16(rng, n // 2, 1, n // 50)
    yield "wide", ol.rand_tuples16(rng, n, (1 << 40), (1 << 40) + 3000), ol.rand_tuples16(rng, n, (1 << 40), (1 << 40) + 3000)
    yield "signed", ol.rand_tuples16(rng, n, -5000, 5000), ol.rand_tuples16(rng, n, -5000, 5000)
    yield "high_only", (ol.rand_tuples16(rng, n, 0, 900) * np.array([1, 1 << 32])), (ol.rand_tuples16(rng, n, 0, 900) * np.array([1, 1 << 32]))
    yield "all_equal", ol.rand_tuples16(rng, 3000, 7, 7), ol.rand_tuples16(rng, 2000, 7, 7)
    yield "small", ol.rand_tuples16(rng, 1, 3, 3), ol.rand_tuples16(rng, 5, 1, 5)
    yield "disjoint", ol.rand_tuples16(rng, n, 1, 1000), ol.rand_tuples16(rng, n, 2000, 3000)


def test_key8b_join_and_sort(smj, orc):
    """16-byte tuples (the reference's KEY_8B build, src/types.h:23-29) through the _key8b entry
    points with host buffers: count and checksum of all three joins and of merge_join against the
    oracle (pinned against the reference's KEY_8B build in test_oracle_vs_ref.py), the stable key
    sort, signed keys and keys beyond 32 bits."""
    rng = np.random.default_rng(88)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for name, r, s in _key8b_cases(rng):
        r, s = np.ascontiguousarray(r, dtype=np.int64), np.ascontiguousarray(s, dtype=np.int64)
        exp = orc.join16(r, s)
        cfg = smj.joinconfig_t(4, 128, 1, 1, 1 << 20, smj.NUMA_NEXT)
        for fn in ("sortmergejoin_multiway_key8b", "sortmergejoin_multipass_key8b", "sortmergejoin_mpsm_key8b"):
            relr, rels = smj.relation_t(r.ctypes.data, r.shape[0]), smj.relation_t(s.ctypes.data, s.shape[0])
            res = getattr(smj.lib, fn)(C.byref(relr), C.byref(rels), C.byref(cfg))
            assert res and res.contents.totalresults == exp[0], (name, fn)
            st = smj.smj_join_stats_t()
            assert smj.lib.smj_last_join_stats(C.byref(st)) == 0 and (st.matches, st.checksum) == exp, (name, fn)
            libc.free(C.cast(res.contents.resultlist, C.c_void_p))
            libc.free(C.cast(res, C.c_void_p))
        # scalarsort_tuples: the input buffer ends up key-sorted (stable here), pointers swapped
        a, b = s.copy(), np.zeros_like(s)
        pa, pb = C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data)
        smj.lib.scalarsort_tuples_key8b(C.byref(pa), C.byref(pb), s.shape[0])
        assert pb.value == a.ctypes.data and pa.value == b.ctypes.data
        assert np.array_equal(a, orc.sort16(s)), name
        sr = orc.sort16(r)
        assert smj.lib.merge_join_key8b(sr.ctypes.data, a.ctypes.data, r.shape[0], s.shape[0], None) == exp[0], name
    # device-resident relations
    r, s = ol.rand_tuples16(rng, 1 << 20, 1, 1 << 19), ol.rand_tuples16(rng, (1 << 21) + 3, 1, 1 << 19)
    dr, ds = smj.DeviceBuffer(2 * r.shape[0]), smj.DeviceBuffer(2 * s.shape[0])
    smj.check(smj.lib.smjdev_memcpy_h2d(dr.ptr, r.ctypes.data, r.nbytes))
    smj.check(smj.lib.smjdev_memcpy_h2d(ds.ptr, s.ctypes.data, s.nbytes))
    st = smj.smj_join_stats_t()
    smj.check(smj.lib.smjdev_join_key8b(dr.ptr, r.shape[0], ds.ptr, s.shape[0], 0, C.byref(st)))
    assert (st.matches, st.checksum) == orc.join16(r, s)
    dr.free()
    ds.free()


def test_scatter_async_staging_variant(smj, orc, monkeypatch):
    """SMJ_SCATTER_ASYNC=1: persistent scatter CTAs whose next tile is copied into the staging buffer
    with cp.async behind the output loop.  Same join results (full and ragged tiles, both levels,
    the overflow -> repair paths) as the oracle."""
    monkeypatch.setenv("SMJ_SCATTER_ASYNC", "1")
    rng = np.random.default_rng(99)
    for n_r, n_s in ((1 << 21, (1 << 21) + 12345), (3_000_017, 2_500_001), (1 << 20, 1 << 20)):
        r, s = orc.gen_unique(n_r, 5), orc.gen_fk(n_s, n_r, 6)
        st = _device_mway_stats(smj, r, s)
        assert (st.matches, st.checksum) == orc.join_bruteforce(r, s)
        assert st.path == smj.PATH_MSD_OPT
    r, s = orc.gen_unique(1 << 22, 7), orc.gen_unique(1 << 22, 8)
    for hook in ("shrink1", "shrink2"):
        monkeypatch.setenv("SMJ_MSD_TEST", hook)
        st = _device_mway_stats(smj, r, s)
        assert (st.matches, st.checksum) == orc.join_bruteforce(r, s), hook
    monkeypatch.delenv("SMJ_MSD_TEST")
    r = ol.rand_items(rng, 3_000_000, (1 << 28) - 1)
    s = ol.rand_items(rng, 3_000_000, (1 << 28) - 1)
    st = _device_mway_stats(smj, r, s)
    assert (st.matches, st.checksum) == orc.join_bruteforce(r, s)


def test_scatter_bulk_store_variant(smj, orc, monkeypatch):
    """SMJ_SCATTER_BULK=1: every bin's run leaves the staging buffer as one bulk copy (TMA,
    cp.async.bulk shared -> global) plus at most one plain store at either end -- the variant the
    NVLink exchange pass of the join group uses by default.  Same join results (full and ragged
    tiles, both levels, skipped runs of the overflow -> repair paths) as the oracle."""
    monkeypatch.setenv("SMJ_SCATTER_BULK", "1")
    rng = np.random.default_rng(99)
    for n_r, n_s in ((1 << 21, (1 << 21) + 12345), (3_000_017, 2_500_001), (1 << 20, 1 << 20)):
        r, s = orc.gen_unique(n_r, 5), orc.gen_fk(n_s, n_r, 6)
        st = _device_mway_stats(smj, r, s)
        assert (st.matches, st.checksum) == orc.join_bruteforce(r, s)
        assert st.path == smj.PATH_MSD_OPT
    r, s = orc.gen_unique(1 << 22, 7), orc.gen_unique(1 << 22, 8)
    for hook in ("shrink1", "shrink2"):
        monkeypatch.setenv("SMJ_MSD_TEST", hook)
        st = _device_mway_stats(smj, r, s)
        assert (st.matches, st.checksum) == orc.join_bruteforce(r, s), hook
    monkeypatch.delenv("SMJ_MSD_TEST")
    r = ol.rand_items(rng, 3_000_000, (1 << 28) - 1)
    s = ol.rand_items(rng, 3_000_000, (1 << 28) - 1)
    st = _device_mway_stats(smj, r, s)
    assert (st.matches, st.checksum) == orc.join_bruteforce(r, s)
