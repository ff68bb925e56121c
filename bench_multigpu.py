"""N-GPU leg of bench.py (launched by torchrun, one rank per GPU): the 128M x 128M m-way join with
the relations spread evenly over the ranks.  A step = plan + first-level histograms (one small
collective), the first MSD partitioning pass as the exchange (the scatter kernel stores into the
bucket owners' buffers over NVLink, or copy engines move contiguous bucket ranges), second pass +
bucket join per rank, all-reduce of {count, checksum} (avx-sort-merge-joins_b200/distributed.py).
Strong scaling: the total work is the same at every N."""
from __future__ import annotations

import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

import bench as B


def run(args, smj):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dj_mod = importlib.import_module("avx-sort-merge-joins_b200.distributed")
    eng = dj_mod.CudaEngine(smj, dev)
    dj = dj_mod.DistributedJoin(eng, rank, world)
    lib = smj.lib

    n = args.size or (2_000_000_000 if args.config == 5 else B.N1)
    first, count = (n * rank) // world, (n * (rank + 1)) // world - (n * rank) // world
    r = eng.gen_unique(n, first, count, B.SEED_R)
    s = eng.gen_unique(n, first, count, B.SEED_S)
    splitters = dj_mod.uniform_splitters(n, world)  # keys are a permutation of 1..n

    # SMJ_EXCHANGE = msd (the pass-A scatter kernel stores into the owners' buffers) | dma (copy
    # engines) | p2p (range partition fused with the exchange) | nccl (all-to-all).  Measured on
    # B200 (128M x 128M): N=2 msd 3.11-3.17 ms, dma 3.28 ms; N=8 msd 1.91 ms, dma 2.50 ms.
    # auto: dma when a rank sends at least 1 GB per step (config 5: 12.7 ms against 17.0 ms), else msd
    # group (default): the library's device-driven exchange, csrc/smj_mgpu.cu
    mode = os.environ.get("SMJ_EXCHANGE", "group")
    if mode == "auto":
        mode = "dma" if 2 * count * 8 * (world - 1) / world >= 1e9 else "msd"
    use_p2p = mode not in ("nccl", "group")
    px, gj = None, None
    cap = int(count * 1.25) + 65536
    if use_p2p:
        px = dj_mod.P2PExchange(eng, rank, world, cap, cap)
    if mode == "group":
        gj = dj_mod.GroupJoin(eng, rank, world, cap, cap)
    path = [None]

    def one_step(rr, ss):
        if mode == "group":
            c_, k_, sent_, st_ = gj.step(rr, count, ss, count)
            path[0] = (int(st_.path), int(st_.repeats))
            return c_, k_, sent_
        if mode == "dma":
            return dj.step_msd_dma(rr, count, ss, count, px)
        if mode == "msd":
            return dj.step_msd(rr, count, ss, count, px)
        if use_p2p:
            return dj.step_p2p(rr, count, ss, count, splitters, px)
        return dj.step(rr, count, ss, count, splitters)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        cnt, cks, sent = one_step(r, s)
    assert cnt == n, f"rank {rank}: distributed join returned {cnt}, expected {n}"

    sampler = B.ClockSampler(local) if rank == 0 else None
    launches0 = lib.smjdev_launch_count()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        cnt, cks, sent = one_step(r, s)
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([lib.smjdev_launch_count() - launches0], dtype=torch.int64, device=dev)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_step = float(ms.item()) / args.steps
    # how much of a step is kernels of this library (rank 0, CUDA events around every launch)
    lib.smjdev_profile_begin()
    for _ in range(args.steps):
        one_step(r, s)
    pbuf = C.create_string_buffer(1 << 16)
    lib.smjdev_profile_end(pbuf, len(pbuf))
    prof = B.parse_profile(pbuf.value.decode())
    kernel_ms = sum(p["ms"] for p in prof.values()) / args.steps
    value = 2 * n / (ms_step * 1e-3) / 1e6
    assert cnt == n
    # the CPU oracle's full-size result for this workload (bench.py: ORACLE_RESULT), when it is known
    if n in B.ORACLE_RESULT:
        assert (cnt, cks) == B.ORACLE_RESULT[n], f"count / checksum {(cnt, cks)} differ from the oracle's {B.ORACLE_RESULT[n]}"

    # ---- end to end: each rank's slices start in pinned host memory
    if os.environ.get("SMJ_BENCH_NO_E2E"):  # very large --size: the pinned copies do not fit the host
        e2e_steps, e2e_s = 0, float("nan")
    else:
        e2e_steps, e2e_s = _e2e(args, r, s, count, dev, one_step, sync_all, n)
    # the sampler runs through the per-kernel pass and the end-to-end steps as well: the timed
    # region alone (a few milliseconds) is shorter than nvidia-smi's sampling period
    clocks = sampler.stop() if sampler else None
    line = _line(args, rank, world, n, count, mode, use_p2p, ms_step, value, sent, kernel_ms, prof, clocks,
                 e2e_s, e2e_steps, launches, cnt, cks)
    if rank == 0 and line is not None and path[0] is not None:
        line["config"]["group_path"] = {"path": path[0][0], "exact_repeats": path[0][1]}
    if px is not None:
        px.close()
    if gj is not None:
        gj.close()
    dist.barrier()
    dist.destroy_process_group()
    return line


def _bind_to_gpu_numa_node(local: int) -> str:
    """Pinned host slices should live on the NUMA node the GPU hangs off: eight ranks that all pin
    memory on node 0 (torchrun's default affinity) share one memory controller and one socket's PCIe
    root.  Best effort, before the buffers are allocated: (1) move this process to the cpus of the
    GPU's node if the cpuset allows it (first touch then places the pages); (2) in any case ask the
    kernel to PREFER that node for this process's new pages (set_mempolicy(MPOL_PREFERRED): works
    without cpus on the node, falls back to any node when it is not allowed).  Returns what happened
    (reported in the bench line)."""
    try:
        prop = torch.cuda.get_device_properties(local)
        path = f"/sys/bus/pci/devices/{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0/numa_node"
        node = int(open(path).read().strip())
    except Exception as exc:  # noqa: BLE001
        return f"gpu numa node unknown ({type(exc).__name__})"
    if node < 0:
        return "gpu numa node unknown (-1)"
    what = [f"gpu on node {node}"]
    try:
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus)
        what.append(f"cpus bound ({len(cpus)})")
    except Exception as exc:  # noqa: BLE001
        what.append(f"cpus not bound ({type(exc).__name__})")
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238  # x86-64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), 1024)
        what.append("memory preferred on node" if rc == 0 else f"set_mempolicy failed (errno {ctypes.get_errno()})")
    except Exception as exc:  # noqa: BLE001
        what.append(f"set_mempolicy unavailable ({type(exc).__name__})")
    return ", ".join(what)


def _e2e(args, r, s, count, dev, one_step, sync_all, n):
    numa = _bind_to_gpu_numa_node(dev.index)
    if os.environ.get("SMJ_VERBOSE"):
        print(f"[bench] rank {dist.get_rank()}: {numa}", file=sys.stderr)
    hr = torch.empty(count, dtype=torch.int64).pin_memory()
    hs = torch.empty(count, dtype=torch.int64).pin_memory()
    hr.copy_(r[:count])
    hs.copy_(s[:count])
    sync_all()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dr = hr.to(dev, non_blocking=True)
        ds = hs.to(dev, non_blocking=True)
        c2, _, _ = one_step(dr, ds)
        assert c2 == n
    sync_all()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    return e2e_steps, float(e2e_s.item())


def _roofline(prof, steps, sent, n, world, ms_step, peak, peak_src, link_gbs):
    """The limiting kernel of the N-GPU step is the first partitioning pass, which IS the exchange:
    its roofline is NVLink (bytes that must leave the GPU / measured peer-copy bandwidth, 770 GB/s
    per direction per GPU, B200_PROFILING.md), not HBM."""
    top = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
    hbm_frac = 40.0 * 2 * n / (ms_step * 1e-3) / 1e9 / world / peak
    out = {"bound": "nvlink", "kernel": top, "peak": 770.0, "unit": "GB/s", "traffic": None,
           "peak_source": "measured peer copy, 770 GB/s per direction per GPU (B200_PROFILING.md; nominal 900)",
           "note": "achieved = item bytes this GPU stores into other GPUs per step / duration of the exchange kernel "
                   "(rank 0, CUDA events, serialised profiling pass); whole_step_hbm_frac = 40 B/tuple over the step "
                   "per GPU against the measured HBM peak",
           "whole_step_hbm_frac": hbm_frac, "hbm_peak": peak, "hbm_peak_source": peak_src,
           "nvlink_gbs_sent_per_gpu_over_step": link_gbs}
    if top and "onesweep_unstable" in top and prof[top]["ms"]:
        ms = prof[top]["ms"] / steps
        out.update(achieved=sent / (ms * 1e-3) / 1e9, frac=sent / (ms * 1e-3) / 1e9 / 770.0,
                   exchange_kernel_ms_per_step=ms, share_of_step=ms / (sum(p["ms"] for p in prof.values()) / steps))
    else:  # an exchange mode whose biggest kernel is not the scatter: report the step against HBM
        out.update(bound="hbm", achieved=hbm_frac * peak, peak=peak, frac=hbm_frac, peak_source=peak_src)
    return out


def _line(args, rank, world, n, count, mode, use_p2p, ms_step, value, sent, kernel_ms, prof, clocks, e2e_s,
          e2e_steps, launches, cnt, cks):
    line = None
    if rank == 0:
        peak, peak_src = B.measured_peak()
        link_gbs = sent / (ms_step * 1e-3) / 1e9  # exchange bytes leaving this GPU / whole step
        line = {
            "metric": B.METRICS[5 if n == 2_000_000_000 else 1][0], "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": B.workload(5 if n == 2_000_000_000 else 1, n),
                       "distribution": f"relations spread evenly over {world} GPUs",
                       "parallelism": (f"join group over {world} GPUs (csrc/smj_mgpu.cu): samples / counts / results through "
                                       f"mailboxes in peer memory, first MSD partitioning pass stores into the bucket owners' "
                                       f"buffers over NVLink (CUDA IPC), no collective and no host round trip between sample "
                                       f"and result" if mode == "group" else
                                       f"first MSD partitioning pass into a bucket-major send buffer, contiguous bucket "
                                       f"ranges copied to their owners over NVLink by the copy engines (overlapping "
                                       f"the scatter of the other relation) over {world} GPUs" if mode == "dma" else
                                       f"first MSD partitioning pass fused with the exchange (the scatter kernel stores "
                                       f"into the bucket owners' buffers over NVLink, CUDA IPC) over {world} GPUs"
                                       if mode == "msd" else
                                       f"range partition fused with the exchange (scatter kernel stores into peer receive "
                                       f"buffers over NVLink, CUDA IPC) over {world} GPUs" if use_p2p else
                                       f"range partition by key + NCCL all-to-all over {world} GPUs")
                                      + ", local m-way join, all-reduce of count/checksum",
                       "l2": "per-rank inputs (2 x %.2f GB) are larger than L2" % (count * 8 / 1e9),
                       "timing": "CUDA events on the stream shared with NCCL, max over ranks",
                       "exchange_bytes_sent_per_rank": sent,
                       "rank0_kernel_ms_per_step": kernel_ms,
                       "rank0_kernels": {k: round(v["ms"] / args.steps, 4) for k, v in prof.items()}},
            "clocks": clocks,
            "e2e": ({"value": 2 * n / e2e_s / 1e6, "unit": B.UNIT, "h2d_bytes_per_step": 2 * n * 8,
                     "d2h_bytes_per_step": 16 * world, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                     "api": "GroupJoin.step / DistributedJoin.step on per-rank pinned host slices"} if e2e_steps else
                    {"value": None, "unit": B.UNIT, "skipped": "SMJ_BENCH_NO_E2E"}),
            "gpu_launches": int(launches.item()),
            "roofline": _roofline(prof, args.steps, sent, n, world, ms_step, peak, peak_src, link_gbs),
            "cpu_baseline": None, "matches": cnt, "checksum": cks,
        }
    return line
