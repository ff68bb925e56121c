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
