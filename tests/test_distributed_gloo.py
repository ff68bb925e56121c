"""world_size-2 (and 3) CPU test of the multi-GPU orchestration (splitters, ragged all-to-all,
reduction) in avx-sort-merge-joins_b200/distributed.py over the gloo backend.  The per-rank engine
is a TEST stand-in built on the CPU oracle; on GPUs the engine is CudaEngine (same interface)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import importlib, os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["SMJ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SMJ_ROOT"], "tests"))
    import oracle_lib as ol
    dj_mod = importlib.import_module("avx-sort-merge-joins_b200.distributed")

    class OracleEngine:  # test stand-in with the interface of CudaEngine
        def __init__(self): self.o = ol.oracle()
        def empty(self, n): return torch.empty(max(int(n), 1), dtype=torch.int64)
        def range_partition(self, src, n, splitters, nparts):
            items = src[:n].numpy().view(np.uint64)
            dest = np.searchsorted(np.asarray(splitters, dtype=np.uint32), (items >> np.uint64(32)).astype(np.uint32), side="right")
            order = np.argsort(dest, kind="stable")
            out = self.empty(n); out[:n] = torch.from_numpy(items[order].view(np.int64))
            return out, torch.from_numpy(np.bincount(dest, minlength=nparts).astype(np.int64))
        def join(self, r, nr, s, ns):
            return self.o.join_bruteforce(r[:nr].numpy().view(np.uint64), s[:ns].numpy().view(np.uint64))
        def sample_keys(self, t, n, k):
            if n == 0: return torch.empty(0, dtype=torch.int64)
            return (t[:n:max(n // k, 1)] >> 32).contiguous()
        def result_tensor(self, count, checksum):
            return torch.from_numpy(np.array([count, checksum], dtype=np.uint64).view(np.int64).copy())

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    o = ol.oracle()
    n = 60000
    R = o.gen_unique(n, 12345)
    rng = np.random.default_rng(5)
    zk = ol.zipf_keys(rng, 3 * n, n, 0.9)
    cases = {"unique": (R, o.gen_unique(n, 54321), "uniform"),
             "fk": (R, o.gen_fk(2 * n + 7, n, 9), "uniform"),
             "zipf": (R, (zk << np.uint64(32)) | np.arange(3 * n, dtype=np.uint64), "sampled"),
             "empty_s": (R, np.zeros(0, dtype=np.uint64), "uniform")}
    eng = OracleEngine(); dj = dj_mod.DistributedJoin(eng, rank, world)
    for name, (r, s, how) in cases.items():
        expect = o.join_bruteforce(r, s)
        def mine(a):
            lo, hi = (a.size * rank) // world, (a.size * (rank + 1)) // world
            t = eng.empty(hi - lo); t[:hi - lo] = torch.from_numpy(a[lo:hi].view(np.int64).copy()); return t, hi - lo
        (lr, nr), (ls, ns) = mine(r), mine(s)
        sp = dj_mod.uniform_splitters(n, world) if how == "uniform" else dj.sampled_splitters(lr, nr, ls, ns, 512)
        assert len(sp) == world - 1 and all(sp[i] <= sp[i + 1] for i in range(len(sp) - 1))
        cnt, cks, sent = dj.step(lr, nr, ls, ns, sp)
        assert (cnt, cks) == expect, (name, rank, (cnt, cks), expect)
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''')


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_join_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SMJ_ROOT=ROOT, OMP_NUM_THREADS="1")
    port = 29600 + world
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
                          str(port), str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == world


WORKER_MSD = textwrap.dedent('''
    import importlib, os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["SMJ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SMJ_ROOT"], "tests"))
    import oracle_lib as ol
    dj = importlib.import_module("avx-sort-merge-joins_b200.distributed")

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    o = ol.oracle()
    n = 1 << 20                      # 20 varying key bits: the smallest domain the MSD plan takes
    R, S = o.gen_unique(n, 12345), o.gen_fk(n + n // 3, n, 9)
    expect = o.join_bruteforce(R, S)
    sl = lambda a: a[(a.size * rank) // world:(a.size * (rank + 1)) // world]
    loc = [sl(R), sl(S)]

    def gather(vec):                 # all-gather of an int64 vector, as the CUDA path does
        t = torch.from_numpy(np.ascontiguousarray(vec, dtype=np.int64))
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return np.stack([x.numpy() for x in out])

    # plan from the OR / AND of all ranks' items (CUDA: smjdev_sample_key_bits + exact bits)
    bits = gather(np.array([np.bitwise_or.reduce(loc[0]), np.bitwise_and.reduce(loc[0]),
                            np.bitwise_or.reduce(loc[1]), np.bitwise_and.reduce(loc[1])], dtype=np.uint64).view(np.int64)).view(np.uint64)
    fold = lambda c, op: int(op.reduce(bits[:, c]))
    lo, hi, shift1 = dj.msd_plan(fold(0, np.bitwise_or), fold(1, np.bitwise_and), fold(2, np.bitwise_or), fold(3, np.bitwise_and))
    assert dj.msd_plan_holds(fold(0, np.bitwise_or), fold(1, np.bitwise_and), fold(2, np.bitwise_or), fold(3, np.bitwise_and), lo, hi)
    digit = lambda a: ((a >> np.uint64(shift1)) & np.uint64(511)).astype(np.int64)
    hist = gather(np.concatenate([np.bincount(digit(loc[0]), minlength=512), np.bincount(digit(loc[1]), minlength=512)])).reshape(world, 2, 512)

    def run(plan_of_rank, recv_sizes):
        """gloo has no one-sided stores: ranks take turns, rank g sends while the others receive"""
        bufs = [np.full(recv_sizes[rel], np.iinfo(np.int64).min, dtype=np.int64) for rel in range(2)]
        for g in range(world):
            for rel in range(2):
                for p in range(world):
                    if g == rank:
                        pieces = plan_of_rank[rel][p]
                        if p == rank:
                            for off, items in pieces:
                                bufs[rel][off:off + items.size] = items.view(np.int64)
                        else:
                            dist.send(torch.tensor([len(pieces)], dtype=torch.int64), dst=p)
                            for off, items in pieces:
                                dist.send(torch.tensor([off, items.size], dtype=torch.int64), dst=p)
                                if items.size:
                                    dist.send(torch.from_numpy(items.view(np.int64).copy()), dst=p)
                    elif p == rank:
                        k = torch.zeros(1, dtype=torch.int64); dist.recv(k, src=g)
                        for _ in range(int(k)):
                            hdr = torch.zeros(2, dtype=torch.int64); dist.recv(hdr, src=g)
                            off, cnt = int(hdr[0]), int(hdr[1])
                            if cnt:
                                t = torch.zeros(cnt, dtype=torch.int64); dist.recv(t, src=g)
                                bufs[rel][off:off + cnt] = t.numpy()
        return bufs

    def finish(bufs, name):
        assert all((b != np.iinfo(np.int64).min).all() for b in bufs), name
        cnt, cks = o.join_bruteforce(bufs[0].view(np.uint64), bufs[1].view(np.uint64))
        res = torch.from_numpy(np.array([cnt, cks], dtype=np.uint64).view(np.int64).copy())
        dist.all_reduce(res)
        got = tuple(int(x) for x in res.numpy().view(np.uint64))
        assert got == expect, (name, rank, got, expect)

    # ---- copy-engine exchange: bucket-major send buffer, one contiguous range per owner
    lay = dj.msd_layout_dma(hist, rank)
    send = [a[np.argsort(digit(a), kind="stable")] for a in loc]
    plan = [[[] for _ in range(world)] for _ in range(2)]
    first = lay["first"]
    for rel in range(2):
        own = send[rel][int(lay["send_off"][rel][first[rank]]) if first[rank] < 512 else 0:][:int(hist[rank][rel][first[rank]:first[rank + 1]].sum())]
        plan[rel][rank].append((lay["self_off"][rel], own))
    for (rel, p, src, cnt, dst) in lay["copies"]:
        plan[rel][p].append((dst, send[rel][src:src + cnt]))
    bufs = run(plan, lay["recv"])
    for rel in range(2):                      # every segment holds exactly one bucket's items
        st, ct, bk = lay["segs"][rel]
        for a, c, b in zip(st, ct, bk):
            assert (digit(bufs[rel][a:a + c].view(np.uint64)) == b).all()
        assert ct.sum() == bufs[rel].size
    finish(bufs, "dma")

    # ---- fused-store exchange: every rank writes its slice of bucket b at offs[rel][b] in the owner's buffer
    owner, offs, mine = dj.msd_layout(hist, rank)
    plan = [[[] for _ in range(world)] for _ in range(2)]
    for rel in range(2):
        d = digit(loc[rel])
        for b in np.nonzero(hist[rank][rel])[0]:
            plan[rel][int(owner[b])].append((int(offs[rel][b]), loc[rel][d == b]))
    bufs = run(plan, [int(mine[rel].sum()) for rel in range(2)])
    for rel in range(2):                      # buckets back to back, in bucket order
        d = digit(bufs[rel].view(np.uint64))
        assert (np.diff(d) >= 0).all() and np.array_equal(np.bincount(d, minlength=512), mine[rel])
    finish(bufs, "msd")
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''')


@pytest.mark.parametrize("world", [2, 3])
def test_msd_exchanges_gloo(tmp_path, world):
    """The two MSD exchanges of the multi-GPU join (fused peer stores: msd_layout; copy engines:
    msd_layout_dma) across real processes: every rank plans from the all-gathered histograms,
    'stores' / 'copies' become gloo sends at the planned offsets, the receive buffers must be
    exactly covered, bucket-pure where the layout says so, and the joins of what arrived must add
    up to the global join."""
    script = tmp_path / "worker_msd.py"
    script.write_text(WORKER_MSD)
    env = dict(os.environ, SMJ_ROOT=ROOT, OMP_NUM_THREADS="1")
    port = 29610 + world
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
                          str(port), str(script)], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == world


WORKER_GROUP = textwrap.dedent('''
    import ctypes as C, importlib, os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, os.environ["SMJ_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SMJ_ROOT"], "tests"))
    import oracle_lib as ol
    smj = importlib.import_module("avx-sort-merge-joins_b200")
    dj = importlib.import_module("avx-sort-merge-joins_b200.distributed")

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    o = ol.oracle()
    n = 1 << 20
    R, S = o.gen_unique(n, 12345), o.gen_fk(n + n // 3, n, 9)
    expect = o.join_bruteforce(R, S)
    sl = lambda a: a[(a.size * rank) // world:(a.size * (rank + 1)) // world]
    loc = [sl(R), sl(S)]
    POSTW = 16 + 2 * 512

    def make_post(caps, every=4):
        """what k_msd_sample + k_xg_post_sample publish: here the sample is every 4th item"""
        p = np.zeros(POSTW, dtype=np.uint64)
        samp = [a[::every] for a in loc]
        p[0], p[1] = np.bitwise_or.reduce(samp[0]), np.bitwise_and.reduce(samp[0])
        p[2], p[3] = np.bitwise_or.reduce(samp[1]), np.bitwise_and.reduce(samp[1])
        plan = dj.msd_plan(int(p[0]), int(p[1]), int(p[2]), int(p[3]))
        shift1 = plan[2] if plan else 0
        p[8], p[9], p[10], p[11], p[12], p[13], p[14] = samp[0].size, samp[1].size, loc[0].size, loc[1].size, shift1, caps[0], caps[1]
        if shift1:
            for rel in range(2):
                d = ((samp[rel] >> np.uint64(shift1)) & np.uint64(511)).astype(np.int64)
                p[16 + 512 * rel:16 + 512 * (rel + 1)] = np.bincount(d, minlength=512)
        return p, shift1

    def gather(vec):
        t = torch.from_numpy(np.ascontiguousarray(vec).view(np.int64).copy())
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return np.stack([x.numpy() for x in out]).view(np.uint64)

    def layout(posts, shift1):
        owner = np.zeros(512, dtype=np.uint32); start = np.zeros(1024, dtype=np.uint64); cap = np.zeros(1024, dtype=np.uint64)
        seg = np.zeros(2 * 512 * world, dtype=np.uint64); need = np.zeros(2 * world, dtype=np.uint64)
        u64p = C.POINTER(C.c_uint64)
        st = smj.lib.smjdev_group_layout_host(world, rank, np.ascontiguousarray(posts).ctypes.data_as(u64p), shift1,
                                              owner.ctypes.data_as(C.POINTER(C.c_uint32)), start.ctypes.data_as(u64p),
                                              cap.ctypes.data_as(u64p), seg.ctypes.data_as(u64p), need.ctypes.data_as(u64p))
        return st, owner, start.reshape(2, 512), cap.reshape(2, 512), seg.reshape(2, 512, world), need.reshape(world, 2)

    capacity = int(1.4 * (n + n // 3) / world) + 65536
    post, shift1 = make_post((capacity, capacity))
    posts = gather(post)
    status, owner, start, cap, seg, need = layout(posts, shift1)
    assert status == 0, status
    assert (np.diff(owner.astype(np.int64)) >= 0).all() and owner.max() == world - 1 and owner.min() == 0
    assert (need <= capacity).all()
    # every rank computed the same layout: its own regions are the (b, rank) entries of the segment table
    assert np.array_equal(start, seg[:, :, rank])
    # regions of one owner are disjoint and back to back in (bucket, source) order
    for rel in range(2):
        for p in range(world):
            bs = np.nonzero(owner == p)[0]
            flat = seg[rel, bs, :].reshape(-1).astype(np.int64)
            assert flat[0] == 0 and (np.diff(flat) >= 0).all() and flat[-1] <= need[p][rel]
    # ---- replay the exchange: rank g 'stores' its items of bucket b at start[rel][b] of the owner's buffer
    digit = lambda a: ((a >> np.uint64(shift1)) & np.uint64(511)).astype(np.int64)
    SENT = np.iinfo(np.int64).min
    bufs = [np.full(int(need[rank][rel]), SENT, dtype=np.int64) for rel in range(2)]
    for g in range(world):
        for rel in range(2):
            for p in range(world):
                if g == rank:
                    d = digit(loc[rel])
                    pieces = [(int(start[rel][b]), loc[rel][d == b]) for b in np.nonzero(owner == p)[0]]
                    for off, items in pieces:
                        assert items.size <= int(cap[rel][int(digit(items[:1])[0])]) if items.size else True  # fits its region
                    if p == rank:
                        for off, items in pieces:
                            bufs[rel][off:off + items.size] = items.view(np.int64)
                    else:
                        cat = np.concatenate([it for _, it in pieces]) if pieces else np.zeros(0, dtype=np.uint64)
                        hdr = np.array([x for off, it in pieces for x in (off, it.size)], dtype=np.int64)
                        dist.send(torch.tensor([hdr.size, cat.size], dtype=torch.int64), dst=p)
                        if hdr.size: dist.send(torch.from_numpy(hdr), dst=p)
                        if cat.size: dist.send(torch.from_numpy(cat.view(np.int64).copy()), dst=p)
                elif p == rank:
                    k = torch.zeros(2, dtype=torch.int64); dist.recv(k, src=g)
                    hdr = torch.zeros(int(k[0]), dtype=torch.int64); cat = torch.zeros(int(k[1]), dtype=torch.int64)
                    if int(k[0]): dist.recv(hdr, src=g)
                    if int(k[1]): dist.recv(cat, src=g)
                    pos = 0
                    for off, cnt in hdr.numpy().reshape(-1, 2):
                        assert (bufs[rel][off:off + cnt] == SENT).all()       # nobody else wrote here
                        bufs[rel][off:off + cnt] = cat.numpy()[pos:pos + cnt]; pos += cnt
    got = [b[b != SENT].view(np.uint64) for b in bufs]
    for rel in range(2):  # what arrived belongs to this rank's buckets
        assert (owner[digit(got[rel])] == rank).all()
    cnt, cks = o.join_bruteforce(got[0], got[1])
    res = torch.from_numpy(np.array([cnt, cks], dtype=np.uint64).view(np.int64).copy())
    dist.all_reduce(res)
    assert tuple(int(x) for x in res.numpy().view(np.uint64)) == expect
    # ---- verdicts that must be the same on every rank
    small, _ = make_post((1000, 1000))                       # buffers too small for the layout
    st2 = layout(gather(small), shift1)[0]
    assert st2 & 2, st2
    odd = post.copy(); odd[12] = shift1 + (1 if rank == 0 else 0)   # rank 0 derived another digit
    st3 = layout(gather(odd), int(odd[12]))[0]
    assert st3 & 1, st3
    try:                                                      # ADVICE r1: capacity checks are global
        dj.check_capacity(np.array([[10, 10]] * (world - 1) + [[10, 99]]), np.full((world, 2), 50))
        raise SystemExit("check_capacity did not raise")
    except RuntimeError as exc:
        assert f"rank {world - 1}" in str(exc)
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''')


@pytest.mark.parametrize("world", [2, 3])
def test_group_layout_gloo(tmp_path, world):
    """The join group's layout rule (csrc/smj_mgpu.cu: k_xg_layout, here through its host twin
    smjdev_group_layout_host) across real processes: every rank publishes its sample, all compute
    the layout from the same posts, the scatter into per-(bucket, source) regions is replayed with
    gloo sends; regions must be disjoint, what arrives must be exactly the owner's buckets, the
    per-rank joins must add up to the global join, and capacity / plan disagreements must give the
    same verdict on every rank."""
    script = tmp_path / "worker_group.py"
    script.write_text(WORKER_GROUP)
    env = dict(os.environ, SMJ_ROOT=ROOT, OMP_NUM_THREADS="1")
    port = 29620 + world
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
                          str(port), str(script)], capture_output=True, text=True, env=env, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("ok") == world
