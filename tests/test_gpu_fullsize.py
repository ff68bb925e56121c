"""BASELINE.json's full-size configurations on one B200, checked through size-independent
properties (the arrays never leave HBM): sortedness + multiset fingerprints, partition membership
+ stability, closed-form join cardinalities and agreement of independent join pipelines.
Sizes can be scaled down with SMJ_FULLSIZE_SCALE (e.g. 0.1) on smaller GPUs."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SCALE = float(os.environ.get("SMJ_FULLSIZE_SCALE", "1"))


def _stats(smj, fn, *args, n=3):
    out = (C.c_uint64 * n)()
    smj.check(fn(*args, out), fn.__name__)
    return [int(v) for v in out]


def test_config2_sort_one_billion_items(smj):
    """avxsort standalone sort of 1B packed key|payload items (bench_sort, sortbench.c:85-202):
    keys = permutation of 1..n, payload = row number."""
    n = int(1_000_000_000 * SCALE)
    buf, tmp = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
    smj.check(smj.lib.smjdev_gen_unique(buf.ptr, n, 0, n, 12345), "gen")
    inv0, sum0, xor0 = _stats(smj, smj.lib.smjdev_check_order, buf.ptr, n, 0)
    assert inv0 > n // 3  # a random permutation has ~n/2 descents
    res = C.c_void_p()
    smj.check(smj.lib.smjdev_sort(buf.ptr, tmp.ptr, n, 0, 1, C.byref(res)), "sort")
    assert res.value in (buf.ptr, tmp.ptr)
    inv1, sum1, xor1 = _stats(smj, smj.lib.smjdev_check_order, res.value, n, 0)
    assert inv1 == 0, "output is not sorted"
    assert (sum1, xor1) == (sum0, xor0), "output is not a permutation of the input"
    out = buf if res.value == buf.ptr else tmp
    assert int(out.to_host(1, 0)[0]) >> 32 == 1 and int(out.to_host(1, n - 1)[0]) >> 32 == n
    # sorting sorted data is the identity (idempotence), through the same entry point
    other = tmp if out is buf else buf
    smj.check(smj.lib.smjdev_sort(out.ptr, other.ptr, n, 0, 1, C.byref(res)), "sort")
    inv2, sum2, xor2 = _stats(smj, smj.lib.smjdev_check_order, res.value, n, 0)
    assert (inv2, sum2, xor2) == (0, sum0, xor0)
    buf.free()
    tmp.free()


@pytest.mark.parametrize("nbits", [4, 6, 8, 9, 10, 12, 14])
def test_config3_partition_fanout_sweep(smj, nbits):
    """radix partitioning fan-out sweep on 1B tuples, shift 0, padded layout
    (bench_partitioning, partitioningbench.c:96-181 with partition_relation_optimized)."""
    n = int(1_000_000_000 * SCALE)
    fan = 1 << nbits
    src, out, tmp = smj.DeviceBuffer(n), smj.DeviceBuffer(n + 8 * fan), smj.DeviceBuffer(n if nbits > 9 else 1)
    hist, offs = smj.DeviceBuffer(fan), smj.DeviceBuffer(fan)
    smj.check(smj.lib.smjdev_gen_unique(src.ptr, n, 0, n, 12345), "gen")
    _, sum0, _ = _stats(smj, smj.lib.smjdev_check_order, src.ptr, n, 0)
    smj.check(smj.lib.smjdev_partition(src.ptr, out.ptr, tmp.ptr, n, nbits, 0, 1, 1, hist.ptr, offs.ptr), "partition")
    h, o = hist.to_host(), offs.to_host()
    assert int(h.sum()) == n
    assert not np.any(o % np.uint64(8)), "partition starts must be multiples of 8 tuples"
    assert np.array_equal(o[1:], np.cumsum((h + np.uint64(7)) & ~np.uint64(7))[:-1])
    # keys are a permutation of 1..n: bucket sizes are known in closed form
    expect = np.bincount((np.arange(n % fan, dtype=np.int64)), minlength=fan)[:fan] * 0 + n // fan
    expect[: n % fan] += 1
    assert np.array_equal(h.astype(np.int64), expect)
    used = int(o[-1] + h[-1])
    bad, unstable, items, sum1 = _stats(smj, smj.lib.smjdev_check_partition, out.ptr, offs.ptr, hist.ptr,
                                        nbits, 0, 1, used, n=4)
    assert (bad, unstable, items, sum1) == (0, 0, n, sum0)
    for b in (src, out, tmp, hist, offs):
        b.free()


@pytest.mark.parametrize("theta", [0.5, 1.0])
def test_config4_zipf_foreign_key_join(smj, theta):
    """m-way join with Zipf-skewed S: |R| = 128M unique keys, |S| = 1.28B foreign keys drawn with
    Zipf(theta) from a permuted alphabet of R's keys (genzipf.c:29-159).  Every S tuple has exactly
    one partner: count == |S|; the checksum of the m-way pipeline must equal the one of the
    independent mpsm pipeline (R sorted, S in 4 separately sorted runs)."""
    import torch
    nr, ns = int(128_000_000 * SCALE), int(1_280_000_000 * SCALE)
    dev = torch.device("cuda", smj.lib.smjdev_current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(54321)
    lut = torch.cumsum(torch.arange(1, nr + 1, dtype=torch.float64, device=dev).pow_(-theta), 0)
    lut /= lut[-1].clone()
    alphabet = torch.randperm(nr, device=dev, generator=g) + 1
    r = smj.DeviceBuffer(nr)
    smj.check(smj.lib.smjdev_gen_unique(r.ptr, nr, 0, nr, 12345), "gen")
    s = torch.empty(ns, dtype=torch.int64, device=dev)
    chunk = 1 << 27
    for lo in range(0, ns, chunk):
        hi = min(ns, lo + chunk)
        u = torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g)
        pos = torch.searchsorted(lut, u).clamp_(max=nr - 1)
        s[lo:hi] = (alphabet[pos] << 32) | torch.arange(lo, hi, dtype=torch.int64, device=dev).bitwise_and_(0xFFFFFFFF)
    del lut, alphabet
    torch.cuda.synchronize()
    hot = int(torch.bincount((s[: 1 << 24] >> 32)).max())
    st, st2 = smj.smj_join_stats_t(), smj.smj_join_stats_t()
    smj.check(smj.lib.smjdev_join(smj.ALGO_MWAY, r.ptr, nr, s.data_ptr(), ns, 1, C.byref(st)), "m-way")
    assert st.matches == ns, (theta, hot)
    smj.check(smj.lib.smjdev_join(smj.ALGO_MPSM, r.ptr, nr, s.data_ptr(), ns, 4, C.byref(st2)), "mpsm")
    assert (st2.matches, st2.checksum) == (st.matches, st.checksum)
    print(f"zipf {theta}: hottest key {hot} of {1 << 24} sampled, m-way {st.ms_total:.1f} ms, mpsm {st2.ms_total:.1f} ms")
    r.free()
