#!/usr/bin/env python
"""NVLink throughput of SM stores into a peer GPU's memory (2 ranks): contiguous copy with 8 or 16
bytes per lane, and copies broken into runs of 16 / 64 / 4096 items with gaps, as the scatter
kernels produce them.  torchrun --nproc-per-node 2 tools/p2p_store_bw.py"""
import ctypes as C
import importlib
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
dj = importlib.import_module("avx-sort-merge-joins_b200.distributed")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
eng = dj.CudaEngine(smj, dev)
n = 32 * 1024 * 1024  # 256 MB
px = dj.P2PExchange(eng, rank, world, 2 * n + 65536, 16)
src = torch.arange(n, dtype=torch.int64, device=dev)
peer = px.ptrs[0][(rank + 1) % world]
lib = smj.lib
for width, run, gap in ((8, 1 << 30, 0), (16, 1 << 30, 0), (8, 4096, 64), (8, 64, 16), (8, 16, 4), (16, 16, 4), (8, 16, 0)):
    for target, name in ((peer, "peer"), (px.local[0], "local")):
        ms = []
        for _ in range(4):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            smj.check(lib.smjdev_debug_copy(src.data_ptr(), C.c_void_p(target), n, width, run, gap))
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        if rank == 0:
            print(f"width {width:2d} run {run if run < (1 << 30) else 'inf':>5} gap {gap:3d} -> {name:5s}: {n * 8 / min(ms[1:]) / 1e6:7.1f} GB/s")
px.close()
dist.destroy_process_group()
