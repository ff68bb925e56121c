"""Pins the C oracle (oracle/smj_oracle.c) against the UNMODIFIED reference compiled from
/root/reference (oracle/_ref/libsmj_ref.so).  CPU only.  The reference has no golden vectors and
its own tests use time() seeds, so this side-by-side run IS the pin; a frozen
subset of it is committed under tests/golden/ (see test_oracle_golden.py)."""
import numpy as np
import pytest

import oracle_lib as ol

SIZES = [1, 2, 15, 16, 17, 100, 255, 256, 257, 1000, 4096, 16383, 16384, 16385, 32684, 81920,
         100003]


@pytest.mark.parametrize("n", SIZES)
def test_sort_matches_reference(orc, ref, n):
    rng = np.random.default_rng(n)
    for kind in ("unique", "dups", "equal", "presorted", "reverse"):
        if kind == "unique":
            items = ol.rand_items(rng, n, 1 << 27)
        elif kind == "dups":
            items = ol.rand_items(rng, n, 7)
        elif kind == "equal":
            items = np.full(n, (5 << 32) | 9, dtype=np.uint64)
        elif kind == "presorted":
            items = np.sort(ol.rand_items(rng, n, 1 << 20))
        else:
            items = np.sort(ol.rand_items(rng, n, 1 << 20))[::-1].copy()
        got, placement = ref.sort(items)
        assert np.array_equal(orc.sort(items), got), kind
        assert np.array_equal(np.sort(items), got), kind  # == ascending uint64 order
        assert orc.placement(n) == placement, kind


@pytest.mark.parametrize("n", [5000, 2621440 + 17, 3 * 2621440 + 5, 4 * 2621440 + 100])
def test_sortmultiway_matches_reference(orc, ref, n):
    rng = np.random.default_rng(n)
    items = ol.rand_items(rng, n, 1 << 27)
    got, placement = ref.sort(items, "avxsortmultiway_tuples")
    assert np.array_equal(np.sort(items), got)
    assert orc.placement(n, 1) == placement


def test_placement_sweep(orc, ref):
    rng = np.random.default_rng(7)
    sizes = list(range(1, 70)) + [int(x) for x in rng.integers(70, 600000, size=60)]
    for n in sizes:
        items = ol.rand_items(rng, n, 1 << 26)
        _, placement = ref.sort(items)
        assert orc.placement(n) == placement, n


@pytest.mark.parametrize("na,nb", [(0, 5), (5, 0), (1, 1), (100, 37), (4096, 4096), (10000, 333)])
def test_merge_matches_reference(orc, ref, na, nb):
    rng = np.random.default_rng(na * 31 + nb)
    a = np.sort(ol.rand_items(rng, na, 50))
    b = np.sort(ol.rand_items(rng, nb, 50))
    assert np.array_equal(orc.merge(a, b, 0), ref.merge(a, b, "avx_merge_tuples"))
    # scalar variant: key-only compare, B first on ties -- compare on key-sorted inputs
    ak, bk = orc.sort_key_stable(a), orc.sort_key_stable(b)
    assert np.array_equal(orc.merge(ak, bk, 1), ref.merge(ak, bk, "scalar_merge_tuples"))


@pytest.mark.parametrize("nruns", [4, 8, 64])
def test_multiway_merge_matches_reference(orc, ref, nruns):
    rng = np.random.default_rng(nruns)
    runs = [np.sort(ol.rand_items(rng, int(rng.integers(0, 5000)), 1 << 16)) for _ in range(nruns)]
    got, k = ref.multiway_merge(runs)
    exp = orc.multiway_merge(runs)
    assert k == exp.size
    assert np.array_equal(exp, got)


@pytest.mark.parametrize("nbits,shift", [(2, 0), (4, 0), (7, 3), (8, 0), (11, 0), (14, 2)])
def test_partition_matches_reference(orc, ref, nbits, shift):
    rng = np.random.default_rng(nbits * 100 + shift)
    n = 50000
    items = ol.rand_items(rng, n, 1 << 22)
    for which, padded in (("partition_relation", False), ("partition_relation_optimized", True),
                          ("partition_relation_optimized_V2", True)):
        rout, rhist, roffs = ref.partition(items, nbits, shift, which)
        out, hist, offs, used = orc.partition(items, nbits, shift, padded)
        assert np.array_equal(hist, rhist), which
        assert np.array_equal(offs, roffs), which
        for p in range(1 << nbits):
            lo, c = int(offs[p]), int(hist[p])
            assert np.array_equal(out[lo:lo + c], rout[lo:lo + c]), (which, p)


def test_merge_join_matches_reference(orc, ref):
    rng = np.random.default_rng(3)
    for (nr, ns, kmax) in [(1000, 1000, 1000), (5000, 20000, 300), (100, 100, 1), (1, 1, 5),
                           (3000, 10, 50)]:
        r = orc.sort(ol.rand_items(rng, nr, kmax))
        s = orc.sort(ol.rand_items(rng, ns, kmax))
        cnt = ref.merge_join(r, s)
        assert orc.merge_join(r, s)[0] == cnt
        assert orc.merge_join_runs(r, s) == orc.merge_join(r, s)
    # interpolation variant assumes unique R keys that bracket S
    r = orc.sort(ol.make_unique(rng, 4000))
    s = orc.sort(ol.rand_items(rng, 9000, 4000))
    assert orc.merge_join_interpolation(r, s) == ref.merge_join(r, s, "merge_join_interpolation")


@pytest.mark.parametrize("nthreads,fanout", [(1, 16), (2, 128), (4, 128), (8, 32)])
def test_join_matches_reference(orc, ref, nthreads, fanout):
    rng = np.random.default_rng(nthreads)
    n = 1 << 17
    r = orc.gen_unique(n, 12345)
    s = orc.gen_unique(n, 54321)
    assert ref.join(r, s, "sortmergejoin_multiway", nthreads, fanout) == n
    assert orc.join_mway(r, s, nthreads, fanout)[0] == n
    assert orc.join_bruteforce(r, s) == orc.join_mway(r, s, nthreads, fanout)
    # foreign keys with duplicates on the S side, both algorithms of the reference
    s2 = orc.gen_fk(3 * n, n, 99)
    exp = orc.join_bruteforce(r, s2)
    assert exp[0] == 3 * n
    assert ref.join(r, s2, "sortmergejoin_multiway", nthreads, fanout) == exp[0]
    if nthreads > 1:  # the reference's m-pass frees interior pointers when run with one thread
        assert ref.join(r, s2, "sortmergejoin_multipass", nthreads, fanout) == exp[0]
    assert orc.join_mway(r, s2, nthreads, fanout) == exp
    # duplicates on both sides
    r3 = ol.rand_items(rng, n, n // 4)
    s3 = ol.rand_items(rng, n, n // 4)
    exp = orc.join_bruteforce(r3, s3)
    assert ref.join(r3, s3, "sortmergejoin_multiway", nthreads, fanout) == exp[0]
    assert orc.join_mway(r3, s3, nthreads, fanout) == exp


# ------------------------------------------------------------------------------------ Zipf generator
@pytest.mark.parametrize("n,alphabet,theta", [(5000, 1000, 0.5), (60000, 50000, 1.0), (30000, 77777, 0.75),
                                             (4000, 64, 1.05), (1000, 1, 0.9)])
def test_zipf_literal_restatement_matches_reference(orc, ref, n, alphabet, theta):
    """orc_gen_zipf_ref (genzipf.c:29-159 restated with glibc rand() and pow()) reproduces the
    compiled reference's gen_zipf bit for bit under the same srand()."""
    got, alpha, lut = orc.gen_zipf_ref(n, alphabet, theta, 54321)
    assert np.array_equal(got, ref.gen_zipf(n, alphabet, theta, 54321))
    assert np.array_equal(np.sort(alpha), np.arange(1, alphabet + 1, dtype=np.uint32))
    assert lut[-1] == 1.0 and np.all(np.diff(lut) > 0)


@pytest.mark.parametrize("alphabet,theta", [(1000, 0.5), (100003, 1.0), (65536, 0.75), (300, 1.05), (70000, 0.0)])
def test_zipf_reproducible_table_close_to_reference(orc, alphabet, theta):
    """The device generator's table (IEEE-only i^-theta, blocked sums; orc_zipf_lut restates it)
    against the reference's pow()-based running sum: 1e-13 relative."""
    _, _, lut_ref = orc.gen_zipf_ref(1, alphabet, theta, 1)
    lut = orc.zipf_lut(alphabet, theta)
    assert lut[-1] == 1.0
    assert np.max(np.abs(lut - lut_ref) / lut_ref) < 1e-13


@pytest.mark.parametrize("theta", [0.5, 1.0])
def test_zipf_device_variant_has_the_reference_distribution(orc, ref, theta):
    """Rank frequencies of the counter-based variant (what smjdev_gen_zipf generates) against the
    compiled reference's gen_zipf: both are samples of the same table, so the two-sample chi-square
    statistic over the ranks (tail pooled) stays near its degrees of freedom."""
    n, alphabet = 400_000, 2000
    a = ref.gen_zipf(n, alphabet, theta, 54321)
    _, alpha_ref, _ = orc.gen_zipf_ref(1, alphabet, theta, 54321)  # the alphabet the reference drew from
    rank_of = np.empty(alphabet + 1, dtype=np.int64)
    rank_of[alpha_ref] = np.arange(alphabet)
    ranks_a = rank_of[(a >> np.uint64(32)).astype(np.int64)]
    b = orc.gen_zipf(n, alphabet, theta, 777, 12345)
    alpha_dev = (orc.gen_unique(alphabet, 12345) >> np.uint64(32)).astype(np.int64)
    rank_dev = np.empty(alphabet + 1, dtype=np.int64)
    rank_dev[alpha_dev] = np.arange(alphabet)
    ranks_b = rank_dev[(b >> np.uint64(32)).astype(np.int64)]
    assert np.array_equal(b & np.uint64(0xFFFFFFFF), np.arange(n, dtype=np.uint64))  # payload = row
    ca, cb = np.bincount(ranks_a, minlength=alphabet), np.bincount(ranks_b, minlength=alphabet)
    keep = (ca + cb) >= 40  # pool the sparse tail
    ca = np.append(ca[keep], ca[~keep].sum()).astype(np.float64)
    cb = np.append(cb[keep], cb[~keep].sum()).astype(np.float64)
    ok = (ca + cb) > 0
    chi2 = float(np.sum((ca[ok] - cb[ok]) ** 2 / (ca[ok] + cb[ok])))
    dof = int(ok.sum()) - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)


# ------------------------------------------------------------------------------------ materialisation
@pytest.mark.parametrize("nthreads", [1, 2, 4])
def test_join_materialize_matches_reference(orc, refmat, nthreads):
    """The tuples the reference's JOIN_MATERIALIZE build emits (unmodified sources + the
    tuple_buffer.h shim): same multiset as orc_join_materialize; single-threaded also checked per
    partition order (sorted by key inside every emitted run)."""
    rng = np.random.default_rng(100 + nthreads)
    n = 1 << 13
    for r, s in ((ol.make_unique(rng, n), ol.rand_items(rng, 3 * n, n)),
                 (ol.rand_items(rng, n, n // 8), ol.rand_items(rng, 2 * n, n // 8)),
                 (ol.rand_items(rng, n, 1 << 20), ol.rand_items(rng, n, 1 << 20))):
        total, outs = refmat.join_materialize(r, s, "sortmergejoin_multiway", nthreads, 128)
        got = np.concatenate(outs)
        exp = orc.join_materialize(r, s, 0)
        assert total == got.size == exp.size == orc.join_bruteforce(r, s)[0]
        assert np.array_equal(np.sort(got), np.sort(exp))


# ------------------------------------------------------------------------------------ KEY_8B build
def test_key8b_join_and_sort_match_reference(orc, refk8):
    """The oracle's 16-byte-tuple restatement (orc_sort16 / orc_join16) against the reference
    compiled with -DKEY_8B (src/types.h:23-29), run with scalar sort and merge as src/main.c:871-877
    demands: join cardinality for unique / foreign-key / duplicate keys, keys beyond 32 bits, the key
    order of scalarsort_tuples, merge_join on sorted inputs."""
    rng = np.random.default_rng(816)
    n = 1 << 13
    uniq = np.empty((n, 2), dtype=np.int64)
    uniq[:, 1] = rng.permutation(n) + 1
    uniq[:, 0] = np.arange(n)
    cases = {
        "fk": (uniq, ol.rand_tuples16(rng, 3 * n, 1, n)),
        "dups": (ol.rand_tuples16(rng, n, 1, n // 8), ol.rand_tuples16(rng, 2 * n, 1, n // 8)),
        "wide": (ol.rand_tuples16(rng, n, 1 << 33, (1 << 33) + 2000), ol.rand_tuples16(rng, n, 1 << 33, (1 << 33) + 2000)),
    }
    for name, (r, s) in cases.items():
        exp = orc.join16(r, s)[0]
        for which, nt in (("sortmergejoin_multiway", 1), ("sortmergejoin_multiway", 4), ("sortmergejoin_multipass", 2)):
            if name == "wide" and nt > 1:
                continue  # the reference's partition hash takes the low key bits: all tuples in few partitions
            assert refk8.join(r, s, which, nt) == exp, (name, which, nt)
        sr, ss = orc.sort16(r), orc.sort16(s)
        assert refk8.merge_join(sr, ss) == exp, name
        got = refk8.sort(s)
        assert np.array_equal(got[:, 1], ss[:, 1]), name            # same key order
        assert np.array_equal(np.sort(got.view([("p", "i8"), ("k", "i8")]).reshape(-1), order=["k", "p"]),
                              np.sort(ss.view([("p", "i8"), ("k", "i8")]).reshape(-1), order=["k", "p"])), name
