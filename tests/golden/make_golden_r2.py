"""Generates tests/golden/ref_golden_r2.npz from the UNMODIFIED reference: round-2 additions to
ref_golden.npz (which stays as it is, so that its stored random inputs do not change).

    python tests/golden/make_golden_r2.py          # in the dev container, after `make -C oracle ref`

* zipf_*: outputs of the compiled reference's gen_zipf (src/datagen/genzipf.c:98-159) under
  srand(54321) -- keys, for several (n, alphabet, theta).
* mat_*: the tuples the reference's JOIN_MATERIALIZE build emits (src/joins/joincommon.c:256-278),
  from oracle/_ref/libsmj_ref_mat.so = the unmodified sources + oracle/shim/tuple_buffer.h for the
  header the reference lacks; stored as sorted multisets per case (the per-thread order is an
  artefact of the thread count), plus the emission ORDER of the single-threaded run.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def mat_cases(rng):
    n = 1 << 12
    yield "fk", ol.make_unique(rng, n), ol.rand_items(rng, 3 * n, n)
    yield "dups", ol.rand_items(rng, n, n // 8), ol.rand_items(rng, 2 * n, n // 8)
    yield "heavy", ol.rand_items(rng, 2000, 12), ol.rand_items(rng, 1500, 16)
    yield "disjoint", ol.rand_items(rng, 1000, 500), ol.rand_items(rng, 1000, 500) + (np.uint64(600) << np.uint64(32))
    yield "ragged", ol.rand_items(rng, 4097, 999), ol.rand_items(rng, 12345, 1200)


def main():
    ref, refm = ol.reference(), ol.reference_mat()
    assert ref is not None and refm is not None, "build the reference first: make -C oracle ref"
    rng = np.random.default_rng(20261019)
    g = {}
    zipf = [(5000, 1000, 0.5), (20000, 4096, 1.0), (30000, 77777, 0.75), (10000, 100, 1.05)]
    g["zipf_cases"] = np.array(zipf, dtype=np.float64)
    for i, (n, a, th) in enumerate(zipf):
        g[f"zipf_{i}"] = ref.gen_zipf(n, a, th, 54321)
    names = []
    for name, r, s in mat_cases(rng):
        names.append(name)
        g[f"mat_{name}_r"], g[f"mat_{name}_s"] = r, s
        tot1, outs1 = refm.join_materialize(r, s, "sortmergejoin_multiway", 1, 128)
        # Several threads only where the reference is sound: a dozen distinct keys crashes its
        # multiway merge (skew is outside what it supports, src/main.c:270-276), and a last chunk
        # that crosses a power of two gives its threads different partition shifts, so that matches
        # are LOST (sortmergejoin_multiway.c:352,372; 4097 x 12345 tuples: 31477 results on two
        # threads, 41927 on one -- SURVEY.md appendix A.2).
        nt = 4 if name in ("fk", "dups", "disjoint") else 1
        tot4, outs4 = refm.join_materialize(r, s, "sortmergejoin_multiway", nt, 128)
        a1, a4 = np.concatenate(outs1), np.concatenate(outs4)
        assert tot1 == tot4 == a1.size == a4.size
        assert np.array_equal(np.sort(a1), np.sort(a4))
        g[f"mat_{name}_sorted"] = np.sort(a1)
        g[f"mat_{name}_order1"] = a1
    g["mat_names"] = np.array(names)
    path = os.path.join(HERE, "ref_golden_r2.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(g), "arrays")


if __name__ == "__main__":
    main()
