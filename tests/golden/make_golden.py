"""Generates tests/golden/ref_golden.npz from the UNMODIFIED reference (oracle/_ref/libsmj_ref.so,
compiled from /root/reference by `make -C oracle ref`).  Run in the dev container:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own (all of its tests use time() seeds);
these frozen input/output pairs are what travels to machines without /root/reference.  Inputs come
from numpy's PCG64 with fixed seeds and are stored next to the outputs, so nothing depends on a
generator version.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def main():
    ref = ol.reference()
    assert ref is not None, "build the reference first: make -C oracle ref"
    rng = np.random.default_rng(20261018)
    g = {}

    # --- sort (avxsort_tuples): inputs of several shapes, output + which buffer it landed in
    sort_sizes = [1, 15, 16, 100, 255, 256, 1000, 4096, 5000, 16384, 20001]
    for i, n in enumerate(sort_sizes):
        kind = ["unique", "dups", "wide"][i % 3]
        items = {"unique": ol.make_unique(rng, n), "dups": ol.rand_items(rng, n, 9),
                 "wide": ol.rand_items(rng, n, (1 << 31) - 1)}[kind]
        out, placement = ref.sort(items)
        g[f"sort_in_{n}"], g[f"sort_out_{n}"], g[f"sort_place_{n}"] = items, out, np.int64(placement)
    g["sort_sizes"] = np.array(sort_sizes)
    place_n = np.array(list(range(1, 40)) + [255, 256, 257, 511, 512, 8191, 8192, 16383, 16384, 16385,
                                             32684, 49152, 81920, 100003, 262144, 300000], dtype=np.int64)
    g["place_n"] = place_n
    g["place_v"] = np.array([ref.sort(ol.rand_items(rng, int(n), 1 << 20))[1] for n in place_n], dtype=np.int64)

    # --- merge
    a, b = np.sort(ol.rand_items(rng, 3000, 40)), np.sort(ol.rand_items(rng, 1777, 40))
    g["merge_a"], g["merge_b"] = a, b
    g["merge_avx"] = ref.merge(a, b, "avx_merge_tuples")
    ak, bk = ol.oracle().sort_key_stable(a), ol.oracle().sort_key_stable(b)
    g["merge_ak"], g["merge_bk"] = ak, bk
    g["merge_scalar"] = ref.merge(ak, bk, "scalar_merge_tuples")
    runs = [np.sort(ol.rand_items(rng, int(rng.integers(0, 900)), 1 << 12)) for _ in range(8)]
    g["mway_lens"] = np.array([r.size for r in runs])
    g["mway_in"] = np.concatenate(runs)
    g["mway_out"] = ref.multiway_merge(runs)[0]

    # --- partition (all three variants), n = 6000
    items = ol.rand_items(rng, 6000, 1 << 18)
    g["part_in"] = items
    for nbits, shift in ((4, 0), (7, 3), (10, 1)):
        for which, tag in (("partition_relation", "plain"), ("partition_relation_optimized", "opt"),
                           ("partition_relation_optimized_V2", "v2")):
            out, hist, offs = ref.partition(items, nbits, shift, which)
            used = int(offs[-1] + hist[-1])
            # padding gaps hold unspecified bytes: store only the partitions themselves
            packed = np.concatenate([out[int(o):int(o) + int(c)] for o, c in zip(offs, hist)])
            g[f"part_{tag}_{nbits}_{shift}_hist"], g[f"part_{tag}_{nbits}_{shift}_offs"] = hist, offs
            g[f"part_{tag}_{nbits}_{shift}_data"] = packed
            assert used <= out.size

    # --- merge join + whole joins
    r, s = np.sort(ol.rand_items(rng, 5000, 700)), np.sort(ol.rand_items(rng, 9000, 700))
    g["mj_r"], g["mj_s"], g["mj_count"] = r, s, np.int64(ref.merge_join(r, s))
    ru = np.sort(ol.make_unique(rng, 3000))
    g["mji_r"], g["mji_s"] = ru, np.sort(ol.rand_items(rng, 7000, 3000))
    g["mji_count"] = np.int64(ref.merge_join(g["mji_r"], g["mji_s"], "merge_join_interpolation"))
    n = 1 << 13
    R = ol.make_unique(rng, n)
    S = ol.rand_items(rng, 3 * n, n)
    D1, D2 = ol.rand_items(rng, n, n // 8), ol.rand_items(rng, n, n // 8)
    g["join_R"], g["join_S"], g["join_D1"], g["join_D2"] = R, S, D1, D2
    g["join_RS_mway_2"] = np.int64(ref.join(R, S, "sortmergejoin_multiway", 2, 128))
    g["join_RS_mway_4"] = np.int64(ref.join(R, S, "sortmergejoin_multiway", 4, 64))
    g["join_RS_mpass_4"] = np.int64(ref.join(R, S, "sortmergejoin_multipass", 4, 128))
    g["join_DD_mway_2"] = np.int64(ref.join(D1, D2, "sortmergejoin_multiway", 2, 128))
    g["join_DD_mpass_2"] = np.int64(ref.join(D1, D2, "sortmergejoin_multipass", 2, 128))
    path = os.path.join(HERE, "ref_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(g), "arrays")


if __name__ == "__main__":
    main()
