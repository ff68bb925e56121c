"""The join group (csrc/smj_mgpu.cu): the m-way join whose first partitioning pass is the exchange,
driven from the device.  world = 1 runs on any GPU box (same kernels, mailboxes and layout, the one
rank owns every bucket); the multi-GPU cases need two GPUs and are skipped otherwise."""
import ctypes as C
import threading

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _cases(orc):
    rng = np.random.default_rng(31)
    n = 2_000_000  # (not a power of two: a sample that misses the one key with the top bit set is a legitimate repeat)
    lo32 = np.uint64(0xFFFFFFFF)
    yield "unique", orc.gen_unique(n, 12345), orc.gen_unique(n, 54321), True
    yield "fk", orc.gen_unique(n, 1), orc.gen_fk(3 * n + 5, n, 2), True
    yield "zipf", orc.gen_unique(n, 12345), orc.gen_zipf(2 * n, n, 0.6, 54321, 12345), True
    yield "wide_keys", ol.rand_items(rng, n, (1 << 30) - 1), ol.rand_items(rng, n, (1 << 30) - 1), None
    yield "narrow_keys", ol.rand_items(rng, n, 1 << 12), ol.rand_items(rng, n // 2, 1 << 12), False   # no MSD plan
    heavy = orc.gen_unique(n, 3)
    heavy[: n // 3] = (heavy[: n // 3] & lo32) | (np.uint64(77) << np.uint64(32))                      # heavy on both sides
    yield "heavy_both", heavy, heavy.copy(), False
    yield "small", orc.gen_unique(5000, 4), orc.gen_fk(7000, 5000, 5), False
    yield "empty_s", orc.gen_unique(n, 6), np.zeros(0, dtype=np.uint64), False


def _group(smj, rank, world, cap_r, cap_s):
    g = C.c_void_p()
    smj.check(smj.lib.smjdev_group_create(rank, world, cap_r, cap_s, C.byref(g)), "group_create")
    h = C.create_string_buffer(smj.GROUP_HANDLE_BYTES)
    smj.check(smj.lib.smjdev_group_export(g, h), "group_export")
    return g, h.raw


def test_group_world1(smj, orc):
    smj.check(smj.lib.smjdev_init(0))
    for name, r, s, fast in _cases(orc):
        cap = max(r.size, s.size) * 5 // 4 + 65536
        g, h = _group(smj, 0, 1, cap, cap)
        smj.check(smj.lib.smjdev_group_connect(g, h), "connect")
        dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
        st, sent = smj.smj_join_stats_t(), C.c_uint64(7)
        for _ in range(2):  # twice: epochs, buffers and mailboxes are reused
            smj.check(smj.lib.smjdev_group_join(g, dr.ptr, r.size, ds.ptr, s.size, C.byref(st), C.byref(sent)), name)
            assert (st.matches, st.checksum) == orc.join_bruteforce(r, s), name
            assert sent.value == 0
            if fast is True:
                assert st.path == smj.PATH_MSD_OPT and st.repeats == 0, name
            if fast is False:
                assert st.repeats == 1, name
        smj.lib.smjdev_group_destroy(g)
        dr.free()
        ds.free()


def test_group_world1_receive_buffer_too_small(smj, orc):
    """The layout does not fit the announced capacity -> exact repeat -> which does not fit either:
    the join fails loudly instead of writing past the buffer."""
    smj.check(smj.lib.smjdev_init(0))
    r, s = orc.gen_unique(1 << 21, 1), orc.gen_unique(1 << 21, 2)
    g, h = _group(smj, 0, 1, 1 << 20, 1 << 20)
    smj.check(smj.lib.smjdev_group_connect(g, h), "connect")
    dr, ds = smj.DeviceBuffer.from_host(r), smj.DeviceBuffer.from_host(s)
    st = smj.smj_join_stats_t()
    rc = smj.lib.smjdev_group_join(g, dr.ptr, r.size, ds.ptr, s.size, C.byref(st), None)
    assert rc != 0 and b"buffer" in smj.lib.smjdev_last_error()
    smj.lib.smjdev_group_destroy(g)


def _two_gpus(smj):
    if smj.lib.smjdev_device_count() < 2:
        pytest.skip("needs two GPUs")


@pytest.mark.parametrize("world", [2, 4])
def test_group_threads(smj, orc, world):
    """One process, one host thread per GPU (peer access): every rank gets the global result."""
    if smj.lib.smjdev_device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    for name, r, s, fast in _cases(orc):
        cap = max(r.size, s.size) + 65536
        groups, handles, res, errs = [None] * world, [None] * world, [None] * world, []
        bar = threading.Barrier(world)

        def work(rank):
            try:
                smj.check(smj.lib.smjdev_init(rank))
                groups[rank], handles[rank] = _group(smj, rank, world, cap, cap)
                bar.wait()
                smj.check(smj.lib.smjdev_group_connect(groups[rank], b"".join(handles)), "connect")
                bar.wait()
                sl = lambda a: a[(a.size * rank) // world:(a.size * (rank + 1)) // world]
                dr, ds = smj.DeviceBuffer.from_host(sl(r)), smj.DeviceBuffer.from_host(sl(s))
                st, sent = smj.smj_join_stats_t(), C.c_uint64(0)
                for _ in range(2):
                    smj.check(smj.lib.smjdev_group_join(groups[rank], dr.ptr, dr.n, ds.ptr, ds.n, C.byref(st),
                                                        C.byref(sent)), name)
                res[rank] = (st.matches, st.checksum, st.path, st.repeats, sent.value)
                bar.wait()
                smj.lib.smjdev_group_destroy(groups[rank])
            except Exception as exc:  # noqa: BLE001
                errs.append((rank, repr(exc)))
                bar.abort()

        th = [threading.Thread(target=work, args=(k,)) for k in range(world)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, (name, errs)
        exp = orc.join_bruteforce(r, s)
        for rank in range(world):
            assert res[rank][:2] == exp, (name, rank, res[rank], exp)
            if fast is True:
                assert res[rank][2] == smj.PATH_MSD_OPT and res[rank][3] == 0, (name, res[rank])
                assert res[rank][4] > 0  # items did cross to other GPUs
            if fast is False:
                assert res[rank][3] == 1, (name, res[rank])
    smj.check(smj.lib.smjdev_init(0))


def test_sortmergejoin_multiway_over_gpus(smj, orc, monkeypatch):
    """SMJ_NGPUS=2: sortmergejoin_multiway() itself fans out over the GPUs (host buffers in, global
    result out), like the reference fans out over threads inside the call (joincommon.c:105-173)."""
    _two_gpus(smj)
    monkeypatch.setenv("SMJ_NGPUS", "2")
    for name, r, s, _ in _cases(orc):
        total, st = smj.join_host("multiway", r, s, nthreads=4)
        exp = orc.join_bruteforce(r, s)
        assert total == exp[0] and (st.matches, st.checksum) == exp, name
    smj.lib.smjdev_join_multi_release()
    monkeypatch.delenv("SMJ_NGPUS")
    smj.check(smj.lib.smjdev_init(0))
