#!/usr/bin/env python
"""NVLink throughput of the first-level scatter kernel storing into a PEER GPU's memory (2 ranks),
by run length: 512 bins x 8192-item tiles (128-byte runs, two CTAs per SM), 256 bins (256-byte
runs), 16384-item tiles (one CTA per SM).  All 2^k bins point into the peer's buffer.
torchrun --nproc-per-node 2 tools/p2p_scatter_bw.py"""
import ctypes as C
import importlib
import os
import sys

import numpy as np
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
n = 64_000_000
px = dj.P2PExchange(eng, rank, world, int(n * 1.6) + 65536, 16)
src = eng.gen_unique(128_000_000, rank * n, n, 12345)   # keys of 1..128M: 27 bits, top digit at bit 18+32
lib = smj.lib
shift1 = 32 + 26 - 8
for bits, big, al in ((9, 0, 1), (8, 0, 1), (9, 1, 1), (8, 1, 1), (9, 1, 4), (9, 1, 16), (8, 1, 4), (8, 1, 16), (9, 0, 16)):
    os.environ["SMJ_SCATTER_TO_BITS"], os.environ["SMJ_SCATTER_TO_BIG"] = str(bits), str(big)
    os.environ["SMJ_SCATTER_TO_ALIGN"] = str(al)
    nb = 1 << bits
    sh = shift1 if bits == 9 else shift1 + 1
    for target, name in ((px.ptrs[0][(rank + 1) % world], "peer"), (px.local[0], "local")):
        # bins laid out back to back with 30 % slack each
        cap = int(n * 1.6) // nb
        ptrs = (C.c_void_p * 512)(*([target] * 512))
        offs = (C.c_uint64 * 512)(*[min(i, nb - 1) * cap for i in range(512)])
        ms = []
        for _ in range(4):
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            smj.check(lib.smjdev_msd_scatter(src.data_ptr(), n, sh, ptrs, offs))
            e1.record(); torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        if rank == 0:
            print(f"align {al:2d} bins {nb:3d} tile {'16384' if big else ' 8192'} -> {name:5s}: {min(ms[1:]):6.3f} ms  {n * 8 / min(ms[1:]) / 1e6:7.1f} GB/s", flush=True)
px.close()
dist.destroy_process_group()
