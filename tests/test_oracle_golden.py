"""The C oracle against the frozen reference outputs in tests/golden/ref_golden.npz (generated from
the unmodified reference by tests/golden/make_golden.py).  CPU only; runs where /root/reference
does not exist."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))
G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden_r2.npz"))


def test_sort_golden(orc):
    for n in G["sort_sizes"]:
        assert np.array_equal(orc.sort(G[f"sort_in_{n}"]), G[f"sort_out_{n}"])
        assert orc.placement(int(n)) == int(G[f"sort_place_{n}"])
    for n, v in zip(G["place_n"], G["place_v"]):
        assert orc.placement(int(n)) == int(v), n


def test_merge_golden(orc):
    assert np.array_equal(orc.merge(G["merge_a"], G["merge_b"], 0), G["merge_avx"])
    assert np.array_equal(orc.merge(G["merge_ak"], G["merge_bk"], 1), G["merge_scalar"])
    runs = np.split(G["mway_in"], np.cumsum(G["mway_lens"])[:-1])
    assert np.array_equal(orc.multiway_merge(runs), G["mway_out"])


@pytest.mark.parametrize("nbits,shift", [(4, 0), (7, 3), (10, 1)])
def test_partition_golden(orc, nbits, shift):
    for tag, padded in (("plain", False), ("opt", True), ("v2", True)):
        out, hist, offs, _ = orc.partition(G["part_in"], nbits, shift, padded)
        assert np.array_equal(hist, G[f"part_{tag}_{nbits}_{shift}_hist"])
        assert np.array_equal(offs, G[f"part_{tag}_{nbits}_{shift}_offs"])
        packed = np.concatenate([out[int(o):int(o) + int(c)] for o, c in zip(offs, hist)])
        assert np.array_equal(packed, G[f"part_{tag}_{nbits}_{shift}_data"])


def test_join_golden(orc):
    assert orc.merge_join(G["mj_r"], G["mj_s"])[0] == int(G["mj_count"])
    assert orc.merge_join_runs(G["mj_r"], G["mj_s"]) == orc.merge_join(G["mj_r"], G["mj_s"])
    assert orc.merge_join_interpolation(G["mji_r"], G["mji_s"]) == int(G["mji_count"])
    R, S, D1, D2 = G["join_R"], G["join_S"], G["join_D1"], G["join_D2"]
    assert orc.join_bruteforce(R, S)[0] == int(G["join_RS_mway_2"]) == int(G["join_RS_mway_4"]) == \
        int(G["join_RS_mpass_4"])
    assert orc.join_mway(R, S, 2, 128) == orc.join_mway(R, S, 4, 64) == orc.join_bruteforce(R, S)
    assert orc.join_bruteforce(D1, D2)[0] == int(G["join_DD_mway_2"]) == int(G["join_DD_mpass_2"])
    assert orc.join_mway(D1, D2, 2, 128) == orc.join_bruteforce(D1, D2)


def test_oracle_join_materialize_against_numpy():
    """orc_join_materialize (joincommon.c:256-278 restated): the emitted tuples are the S tuples,
    once per matching R tuple; the row-id layout pairs every matching (R.payload, S.payload)."""
    import oracle_lib as ol
    orc = ol.oracle()
    rng = np.random.default_rng(11)
    r = ol.rand_items(rng, 3000, 400)
    s = ol.rand_items(rng, 5000, 500)
    cnt, _ = orc.join_bruteforce(r, s)
    kr, ks = (r >> np.uint64(32)).astype(np.int64), (s >> np.uint64(32)).astype(np.int64)
    mult = np.bincount(kr, minlength=512)
    expect = np.sort(np.repeat(s, mult[ks]))
    got = orc.join_materialize(r, s, 0)
    assert got.size == cnt == int(mult[ks].sum())
    assert np.array_equal(np.sort(got), expect)
    pairs = orc.join_materialize(r, s, 1)
    lo32 = np.uint64(0xFFFFFFFF)
    brute = np.sort(np.array([((a & lo32) << np.uint64(32)) | (b & lo32)
                              for a in r[:300] for b in s[(s >> np.uint64(32)) == (a >> np.uint64(32))]], dtype=np.uint64))
    sub = orc.join_materialize(r[:300], s, 1)
    assert np.array_equal(np.sort(sub), brute)
    assert pairs.size == cnt
    assert orc.join_materialize(r[:0], s, 0).size == 0 and orc.join_materialize(r, s[:0], 1).size == 0


def test_zipf_golden(orc):
    """orc_gen_zipf_ref against frozen outputs of the compiled reference's gen_zipf under
    srand(54321) (tests/golden/make_golden_r2.py).  Depends on glibc's rand(): skipped elsewhere."""
    import platform
    if platform.libc_ver()[0] != "glibc":
        pytest.skip("the frozen vectors were drawn from glibc's rand()")
    for i, (n, a, th) in enumerate(G2["zipf_cases"]):
        got, _, _ = orc.gen_zipf_ref(int(n), int(a), float(th), 54321)
        assert np.array_equal(got, G2[f"zipf_{i}"]), (n, a, th)


def test_materialize_golden(orc):
    """orc_join_materialize against the tuples the reference's JOIN_MATERIALIZE build emitted
    (frozen as sorted multisets; the order is per thread and per partition)."""
    for name in G2["mat_names"]:
        r, s = G2[f"mat_{name}_r"], G2[f"mat_{name}_s"]
        exp = G2[f"mat_{name}_sorted"]
        got = orc.join_materialize(r, s, 0)
        assert got.size == exp.size == orc.join_bruteforce(r, s)[0], name
        assert np.array_equal(np.sort(got), exp), name
        assert np.array_equal(np.sort(G2[f"mat_{name}_order1"]), exp), name
