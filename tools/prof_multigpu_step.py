"""Diagnostic: host-side timeline of one DistributedJoin.step_msd on rank 0 (collectives bracketed by\ndevice synchronisations).  torchrun --nproc-per-node N tools/prof_multigpu_step.py"""
import os, sys, time, importlib, ctypes as C
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
smj = importlib.import_module("avx-sort-merge-joins_b200")
dj_mod = importlib.import_module("avx-sort-merge-joins_b200.distributed")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
eng = dj_mod.CudaEngine(smj, dev); dj = dj_mod.DistributedJoin(eng, rank, world)
n = 128_000_000
first, count = (n*rank)//world, (n*(rank+1))//world - (n*rank)//world
r = eng.gen_unique(n, first, count, 12345); s = eng.gen_unique(n, first, count, 54321)
px = dj_mod.P2PExchange(eng, rank, world, int(count*1.25)+65536, int(count*1.25)+65536)
# monkeypatch timing: wrap dist collectives and lib calls
marks = []
def T(label):
    torch.cuda.synchronize(); marks.append((label, time.perf_counter()))
orig_ag, orig_ar = dist.all_gather, dist.all_reduce
def ag(*a, **k):
    T("pre_allgather"); x = orig_ag(*a, **k); T("allgather"); return x
def ar(*a, **k):
    T("pre_allreduce"); x = orig_ar(*a, **k); T("allreduce"); return x
for i in range(6):
    if i == 5:
        dist.all_gather, dist.all_reduce = ag, ar
        dj_mod.dist.all_gather, dj_mod.dist.all_reduce = ag, ar
        torch.cuda.synchronize(); dist.barrier(); marks.append(("start", time.perf_counter()))
    out = dj.step_msd(r, count, s, count, px)
    if i == 5: T("end")
if rank == 0:
    t0 = marks[0][1]
    prev = t0
    for lab, t in marks[1:]:
        print(f"{lab:16s} +{(t-prev)*1e3:7.3f} ms   (t={(t-t0)*1e3:7.3f})"); prev = t
    print(out)
px.close(); dist.destroy_process_group()
