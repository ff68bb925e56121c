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
