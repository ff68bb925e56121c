"""Multi-GPU sort-merge join: one process per GPU, torch.distributed (NCCL over NVLink) for the
plumbing, the library's CUDA kernels for everything that touches tuples.

This replaces the reference's cross-socket step: there, thread t owns partitions [t*F/T, (t+1)*F/T)
and its multiway merge PULLS the sorted runs of those partitions out of every other thread's
memory in NUMA-shuffle order (src/joins/sortmergejoin_multiway.c:492-517,
src/util/numa_shuffle.c:56-85).  Here rank g owns a key range; every rank range-partitions its
slice of R and S by the same splitters (smjdev_range_partition: one stable scatter pass whose
output is already laid out as per-destination contiguous send regions, so there is no pack
kernel), the slices are exchanged with one all-to-all per relation, each rank joins what it
received with the single-GPU pipeline and the counts / checksums are all-reduced.

    exchange-then-sort instead of the reference's sort-then-merge: the same bytes cross the
    links, but no G-way merge is needed afterwards.

The per-rank engine is pluggable so that the orchestration (splitters, count exchange, ragged
all-to-all, reduction) can be tested on CPU with the gloo backend; the product engine is
CudaEngine.

GroupJoin (the default of bench.py --gpus N) is the library's own device-driven exchange
(csrc/smj_mgpu.cu).  DistributedJoin keeps the four host-orchestrated exchanges of round 1, from
the simplest to the most fused, as references and fallbacks (SMJ_EXCHANGE=nccl|p2p|msd|dma):
  step          range partition into send regions + NCCL all-to-all (the reference design)
  step_p2p      range partition whose scatter kernel stores straight into the peers' receive
                buffers (CUDA IPC over NVLink), then the single-GPU join on what arrived
  step_msd      the m-way join's FIRST partitioning pass is the exchange: every item is stored
                into the buffer of the GPU that owns its key bucket; the receiver only runs the
                second pass and the bucket join (default for small relations)
  step_msd_dma  the same first pass into a local bucket-major send buffer, contiguous bucket
                ranges moved by the copy engines, second pass over the received segments
                (default when a rank sends >= 1 GB per step)
All of them return the same {count, checksum} as one GPU does.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class CudaEngine:
    """Local engine backed by libsmj_b200.so; relations are torch.int64 CUDA tensors holding the
    8-byte items (key<<32 | payload)."""

    def __init__(self, smj, device: torch.device):
        self.smj, self.lib, self.device = smj, smj.lib, device
        torch.cuda.set_device(device)
        smj.check(self.lib.smjdev_init(device.index if device.index is not None else 0), "smjdev_init")
        # run on torch's current stream so that NCCL collectives and our kernels are ordered
        smj.check(self.lib.smjdev_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def empty(self, n: int) -> torch.Tensor:
        return torch.empty(max(int(n), 1), dtype=torch.int64, device=self.device)

    def to_local(self, host: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(host.view(np.int64)).to(self.device, non_blocking=True)

    def gen_unique(self, n_total: int, first: int, count: int, seed: int) -> torch.Tensor:
        t = self.empty(count)
        self.smj.check(self.lib.smjdev_gen_unique(t.data_ptr(), n_total, first, count, seed), "gen")
        return t

    def range_partition(self, src: torch.Tensor, n: int, splitters: np.ndarray, nparts: int):
        out, counts = self.empty(n), torch.zeros(nparts, dtype=torch.int64, device=self.device)
        sp = np.ascontiguousarray(splitters, dtype=np.uint32)
        self.smj.check(self.lib.smjdev_range_partition(
            src.data_ptr(), out.data_ptr(), n, sp.ctypes.data_as(C.POINTER(C.c_uint32)), nparts,
            counts.data_ptr()), "range_partition")
        return out, counts

    def join(self, r: torch.Tensor, nr: int, s: torch.Tensor, ns: int):
        st = self.smj.smj_join_stats_t()
        self.smj.check(self.lib.smjdev_join(self.smj.ALGO_MWAY, r.data_ptr(), nr, s.data_ptr(), ns, 1,
                                            C.byref(st)), "smjdev_join")
        return st.matches, st.checksum

    def sample_keys(self, t: torch.Tensor, n: int, k: int) -> torch.Tensor:
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device=self.device)
        stride = max(n // k, 1)
        return (t[:n:stride] >> 32).contiguous()

    def result_tensor(self, count: int, checksum: int) -> torch.Tensor:
        # two's complement view: the checksum is a uint64, all-reduce sums wrap the same way
        vals = np.array([count, checksum], dtype=np.uint64).view(np.int64)
        return torch.from_numpy(vals).to(self.device)


def p2p_layout(counts: np.ndarray, rank: int):
    """counts[src][rel][dst] = items rank `src` sends to rank `dst` for relation `rel` (the
    all-gathered count matrix).  Returns (offs, recv): offs[rel][dst] = where this rank's slice
    starts in dst's receive buffer of relation rel (slices are laid out in source-rank order),
    recv[rel] = number of items this rank receives."""
    counts = np.asarray(counts, dtype=np.int64)
    offs = counts[:rank].sum(axis=0)          # [rel][dst]: everything lower ranks send there
    recv = counts[:, :, rank].sum(axis=0)     # [rel]
    return offs, recv


def msd_plan(or_r: int, and_r: int, or_s: int, and_s: int):
    """Key digits of the m-way fast path from the OR / AND of all items of R and S (the same rule
    as msd_plan() in csrc/smj_join.cu): returns (lo_bit, hi_bit, shift1) or None when the key
    domain does not suit it (fewer than 19 or more than 31 varying key bits)."""
    v = (((or_r | or_s) >> 32) & ~((and_r & and_s) >> 32)) & 0xFFFFFFFF
    if v == 0:
        return None
    hi, lo = v.bit_length() - 1, (v & -v).bit_length() - 1
    if not 19 <= hi - lo + 1 <= 31:
        return None
    return lo, hi, 32 + hi - 8


def msd_plan_holds(or_r: int, and_r: int, or_s: int, and_s: int, lo: int, hi: int) -> bool:
    """No key bit outside [lo, hi] may vary (exact OR / AND of all items of all ranks)."""
    v = (((or_r | or_s) >> 32) & ~((and_r & and_s) >> 32)) & 0xFFFFFFFF
    rng = ((1 << (hi + 1)) - 1) & ~((1 << lo) - 1)
    return (v & ~rng) == 0


def msd_layout(hist: np.ndarray, rank: int):
    """hist[src][rel][512] = first-level bucket sizes of every rank's slice.  Buckets are given to
    ranks in contiguous runs of roughly equal total size.  Returns
      owner[512]        rank that owns each bucket
      offs[rel][512]    where THIS rank's items of a bucket start inside the owner's buffer
                        (buckets back to back, inside a bucket the slices in source-rank order)
      mine[rel][512]    sizes of the buckets this rank owns (0 for the others)"""
    hist = np.asarray(hist, dtype=np.int64)
    world = hist.shape[0]
    tot = hist.sum(axis=0)                      # [rel][512]
    load = tot.sum(axis=0)                      # [512]
    prefix = np.cumsum(load) - load
    owner = np.minimum(((2 * prefix + load) * world) // max(2 * int(load.sum()), 1), world - 1)
    owner = np.maximum.accumulate(owner)        # contiguous, non-decreasing
    first = np.searchsorted(owner, np.arange(world), side="left")   # first bucket of every rank
    csum = np.cumsum(tot, axis=1) - tot         # [rel][512] exclusive prefix over all buckets
    start_of_owner = np.concatenate([csum, tot.sum(axis=1, keepdims=True)], axis=1)[:, np.minimum(first, 512)]
    base = csum - start_of_owner[:, owner]      # bucket start inside its owner's buffer
    offs = base + hist[:rank].sum(axis=0)
    mine = np.where(owner[None, :] == rank, tot, 0)
    return owner, offs, mine


def msd_layout_dma(hist: np.ndarray, rank: int):
    """Layout of the copy-engine exchange.  hist[src][rel][512] as in msd_layout; buckets are given
    to ranks in the same contiguous runs.  Every rank scatters its slice bucket-major into a LOCAL
    send buffer, so the buckets of one owner are one contiguous range there, and copies that range
    into the owner's receive buffer, which is therefore laid out source-major: for every source
    rank one block holding its items of the owner's buckets, bucket by bucket.  Returns a dict:
      owner[512], first[W + 1]       bucket b belongs to rank owner[b]; rank p owns first[p]..first[p+1]-1
      send_off[rel][512]             where bucket b starts in THIS rank's send buffer
      self_off[rel]                  where this rank's own block starts in its own receive buffer
      copies                         [(rel, dst rank, first item in the send buffer, items, item offset
                                       in dst's receive buffer)] for every other rank
      segs[rel] = (start, count, bucket) arrays: the segments of this rank's receive buffer
      recv[rel]                      items this rank receives (its own included)"""
    hist = np.asarray(hist, dtype=np.int64)
    world = hist.shape[0]
    tot = hist.sum(axis=0)
    load = tot.sum(axis=0)
    prefix = np.cumsum(load) - load
    owner = np.minimum(((2 * prefix + load) * world) // max(2 * int(load.sum()), 1), world - 1)
    owner = np.maximum.accumulate(owner)
    first = np.searchsorted(owner, np.arange(world + 1), side="left")
    first[world] = 512
    csum = np.concatenate([np.zeros((world, 2, 1), dtype=np.int64), np.cumsum(hist, axis=2)], axis=2)
    # blk[p][g][rel] = items source g sends to owner p; blockstart = exclusive prefix over g
    blk = np.stack([csum[:, :, first[p + 1]] - csum[:, :, first[p]] for p in range(world)])
    blockstart = np.cumsum(blk, axis=1) - blk
    send_off = csum[rank][:, :512]
    copies = [(rel, p, int(send_off[rel][first[p]]) if first[p] < 512 else 0, int(blk[p][rank][rel]),
               int(blockstart[p][rank][rel]))
              for rel in range(2) for p in range(world) if p != rank and blk[p][rank][rel] > 0]
    lo, hi = int(first[rank]), int(first[rank + 1])
    segs = []
    for rel in range(2):
        st, ct, bk = [], [], []
        for g in range(world):
            c = hist[g][rel][lo:hi]
            inner = np.cumsum(c) - c
            st.append(blockstart[rank][g][rel] + inner)
            ct.append(c)
            bk.append(np.arange(lo, hi, dtype=np.int64))
        st, ct, bk = (np.concatenate(x) if x else np.zeros(0, dtype=np.int64) for x in (st, ct, bk))
        keep = ct > 0
        segs.append((st[keep], ct[keep], bk[keep]))
    self_off = [int(blockstart[rank][rank][rel]) for rel in range(2)]
    recv = [int(blk[rank][:, rel].sum()) for rel in range(2)]
    return {"owner": owner, "first": first, "send_off": send_off, "self_off": self_off, "copies": copies,
            "segs": segs, "recv": recv}


class P2PExchange:
    """Receive buffers for R and S on every GPU, mapped into every process with CUDA IPC.  With it
    DistributedJoin.step_p2p partitions and exchanges in ONE kernel per relation: the scatter
    stores each item directly into the receive buffer of its destination GPU over NVLink."""

    def __init__(self, engine: "CudaEngine", rank: int, world: int, capacity_r: int, capacity_s: int):
        self.e, self.rank, self.world = engine, rank, world
        lib, smj = engine.lib, engine.smj
        self.cap = [int(capacity_r), int(capacity_s)]
        self.local, handles = [], []
        for cap in self.cap:
            p = C.c_void_p()
            smj.check(lib.smjdev_ipc_alloc(max(cap, 1) * 8, C.byref(p)), "ipc_alloc")
            h = C.create_string_buffer(64)
            smj.check(lib.smjdev_ipc_export(p, h), "ipc_export")
            self.local.append(p.value)
            handles.append(np.frombuffer(h.raw, dtype=np.uint8).copy())
        mine = torch.from_numpy(np.concatenate(handles)).to(engine.device)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        self.ptrs = [[0] * world for _ in range(2)]  # [rel][rank] -> pointer usable in THIS process
        self.opened = []
        for g in range(world):
            raw = allh[g].cpu().numpy().tobytes()
            for rel in range(2):
                if g == rank:
                    self.ptrs[rel][g] = self.local[rel]
                else:
                    q = C.c_void_p()
                    smj.check(lib.smjdev_ipc_open(raw[64 * rel:64 * rel + 64], C.byref(q)), "ipc_open")
                    self.ptrs[rel][g] = q.value
                    self.opened.append(q.value)
        self.ptr_table = np.array(self.ptrs, dtype=np.uint64)  # [rel][rank], for vectorised lookups
        # capacities of ALL ranks: the "does it fit" decision must be the same on every rank
        capt = torch.tensor(self.cap, dtype=torch.int64, device=engine.device)
        allc = [torch.empty_like(capt) for _ in range(world)]
        dist.all_gather(allc, capt)
        self.all_caps = np.stack([c.cpu().numpy() for c in allc])  # [rank][rel]
        dist.barrier()

    def check_capacity(self, recv_all) -> None:
        check_capacity(recv_all, self.all_caps)

    def close(self):
        lib = self.e.lib
        dist.barrier()
        for q in self.opened:
            lib.smjdev_ipc_close(q)
        dist.barrier()
        for p in self.local:
            lib.smjdev_ipc_free(p)
        self.opened, self.local = [], []


def check_capacity(recv_all, all_caps) -> None:
    """recv_all[rank][rel] = items every rank would receive, all_caps[rank][rel] = what its buffers
    hold.  Every rank evaluates this on the SAME all-gathered numbers BEFORE any scatter or copy is
    enqueued, so either nobody writes or everybody does: no rank can store past the end of a peer's
    buffer while that peer has already given up (and nobody is left waiting in a collective)."""
    recv_all, all_caps = np.asarray(recv_all, dtype=np.int64), np.asarray(all_caps, dtype=np.int64)
    bad = np.argwhere(recv_all > all_caps)
    if bad.size:
        p, rel = (int(x) for x in bad[0])
        raise RuntimeError(f"receive buffer of rank {p} too small for relation {'RS'[rel]}: "
                           f"{int(recv_all[p][rel])} > {int(all_caps[p][rel])} (same verdict on every rank)")


class GroupJoin:
    """The library's join group (csrc/smj_mgpu.cu, smjdev_group_*) with one process per GPU:
    torch.distributed only carries the 256-byte handles at set-up; a join is ONE C call per rank,
    with no collective and no host round trip between the sample and the result (samples, counts
    and results travel through mailboxes in peer memory, the first partitioning pass stores into
    the owners' buffers over NVLink)."""

    def __init__(self, engine: "CudaEngine", rank: int, world: int, capacity_r: int, capacity_s: int):
        self.e, self.rank, self.world = engine, rank, world
        lib, smj = engine.lib, engine.smj
        self.group = C.c_void_p()
        smj.check(lib.smjdev_group_create(rank, world, int(capacity_r), int(capacity_s), C.byref(self.group)),
                  "group_create")
        h = C.create_string_buffer(smj.GROUP_HANDLE_BYTES)
        smj.check(lib.smjdev_group_export(self.group, h), "group_export")
        mine = torch.from_numpy(np.frombuffer(h.raw, dtype=np.uint8).copy()).to(engine.device)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)  # also: every rank has been created
        blob = b"".join(t.cpu().numpy().tobytes() for t in allh)
        smj.check(lib.smjdev_group_connect(self.group, blob), "group_connect")
        dist.barrier()

    def step(self, r, nr: int, s, ns: int):
        """-> (global count, global checksum, item bytes this rank sent, stats)"""
        smj, lib = self.e.smj, self.e.lib
        st, sent = smj.smj_join_stats_t(), C.c_uint64(0)
        smj.check(lib.smjdev_group_join(self.group, r.data_ptr(), nr, s.data_ptr(), ns, C.byref(st), C.byref(sent)),
                  "group_join")
        return int(st.matches), int(st.checksum), int(sent.value), st

    def close(self):
        dist.barrier()
        self.e.lib.smjdev_group_destroy(self.group)
        self.group = None
        dist.barrier()


def uniform_splitters(maxkey: int, nparts: int) -> np.ndarray:
    """Equal-width key ranges for dense keys 1..maxkey: dest(key) = #splitters <= key."""
    return np.array([(i * maxkey) // nparts + 1 for i in range(1, nparts)], dtype=np.uint32)


class DistributedJoin:
    def __init__(self, engine, rank: int, world: int):
        self.e, self.rank, self.world = engine, rank, world

    # ---- splitters: the same on every rank, for R and S alike
    def sampled_splitters(self, r, nr: int, s, ns: int, per_rank: int = 4096) -> np.ndarray:
        """Quantiles of a sample of the keys of both relations gathered from all ranks (skewed
        keys: SURVEY.md 8e).  Heavy hitters end up on one rank: the result stays exact, the load
        becomes uneven."""
        mine = torch.cat([self.e.sample_keys(r, nr, per_rank), self.e.sample_keys(s, ns, per_rank)])
        size = torch.tensor([mine.numel()], dtype=torch.int64, device=mine.device)
        sizes = [torch.zeros_like(size) for _ in range(self.world)]
        dist.all_gather(sizes, size)
        cap = int(max(int(x.item()) for x in sizes))
        padded = torch.full((cap,), -1, dtype=torch.int64, device=mine.device)
        padded[:mine.numel()] = mine
        allp = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(allp, padded)
        keys = torch.cat(allp)
        keys = torch.sort(keys[keys >= 0]).values.cpu().numpy()
        if keys.size == 0:
            return np.zeros(self.world - 1, dtype=np.uint32)
        q = [keys[min(keys.size - 1, (i * keys.size) // self.world)] for i in range(1, self.world)]
        return np.asarray(q, dtype=np.uint32)

    # ---- ragged all-to-all of one partitioned relation
    def exchange(self, part: torch.Tensor, counts: torch.Tensor):
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts)
        send = [int(x) for x in counts.cpu().tolist()]  # the collective needs host-side sizes
        recv = [int(x) for x in recv_counts.cpu().tolist()]
        n_recv = sum(recv)
        out = self.e.empty(n_recv)
        dist.all_to_all_single(out[:n_recv], part[:sum(send)], output_split_sizes=recv,
                               input_split_sizes=send)
        return out, n_recv, send, recv

    def step_p2p(self, r, nr: int, s, ns: int, splitters: np.ndarray, px: "P2PExchange"):
        """Same result as step(), but partition + exchange are fused: count per destination,
        all-gather the counts, then one scatter kernel per relation writes straight into the peers'
        receive buffers (no send buffer, no NCCL data collective)."""
        e, lib, smj, W = self.e, self.e.lib, self.e.smj, self.world
        sp = np.ascontiguousarray(splitters, dtype=np.uint32)
        spp = sp.ctypes.data_as(C.POINTER(C.c_uint32))
        cnt = torch.zeros(2, W, dtype=torch.int64, device=e.device)
        smj.check(lib.smjdev_range_count(r.data_ptr(), nr, spp, W, cnt[0].data_ptr()), "range_count")
        smj.check(lib.smjdev_range_count(s.data_ptr(), ns, spp, W, cnt[1].data_ptr()), "range_count")
        allc = [torch.empty_like(cnt) for _ in range(W)]
        dist.all_gather(allc, cnt)
        counts = torch.stack(allc).cpu().numpy()  # [src][rel][dst]; the copy synchronises
        offs, recv = p2p_layout(counts, self.rank)
        px.check_capacity(counts.sum(axis=0).T)  # [dst][rel]: every rank checks EVERY receiver
        for rel, (t, n) in enumerate(((r, nr), (s, ns))):
            dptr = (C.c_void_p * W)(*px.ptrs[rel])
            doff = (C.c_uint64 * W)(*[int(x) for x in offs[rel]])
            smj.check(lib.smjdev_range_scatter(t.data_ptr(), n, spp, W, dptr, doff), "range_scatter")
        # Device-side barrier: the tiny all-reduce is enqueued behind this rank's scatter kernels and
        # completes only when every rank has reached it, i.e. when all stores into this GPU's
        # receive buffers have been issued and retired.  No host synchronisation.
        dist.all_reduce(self._token(e.device), op=dist.ReduceOp.SUM)
        st = smj.smj_join_stats_t()
        smj.check(lib.smjdev_join(smj.ALGO_MWAY, px.local[0], int(recv[0]), px.local[1], int(recv[1]), 1,
                                  C.byref(st)), "smjdev_join")
        res = e.result_tensor(st.matches, st.checksum)
        dist.all_reduce(res, op=dist.ReduceOp.SUM)  # also keeps anyone from scattering into a buffer in use
        out = res.cpu().numpy().view(np.uint64)
        sent = 8 * int(counts[self.rank].sum() - counts[self.rank, :, self.rank].sum())
        return int(out[0]), int(out[1]), sent

    def step_msd(self, r, nr: int, s, ns: int, px: "P2PExchange"):
        """The m-way join with its first partitioning pass AS the exchange: every rank histograms
        its slices by the top 9 varying key bits, the ranks agree on bucket owners, and each rank's
        pass-A scatter kernel stores its items straight into the owners' buffers over NVLink
        (smjdev_msd_scatter).  Each rank then holds exactly the input of pass B and finishes with
        smjdev_join_level1.  Compared with step_p2p this saves one full scatter pass and one
        histogram pass per relation.  Falls back to step_p2p when the key domain does not suit the
        MSD plan."""
        e, lib, smj, W = self.e, self.e.lib, self.e.smj, self.world
        dev = e.device
        # ---- 1 + 2. plan, first-level histograms and exact OR/AND of all ranks (one collective)
        pre = self._msd_plan_and_hist(r, nr, s, ns)
        if pre is None:
            return self.step_p2p(r, nr, s, ns, self._fallback, px)
        lo, hi, shift1, hist = pre
        owner, offs, mine_counts = msd_layout(hist, self.rank)
        tot = np.asarray(hist, dtype=np.int64).sum(axis=0)  # [rel][512]
        px.check_capacity(np.array([[int(tot[rel][owner == p].sum()) for rel in range(2)] for p in range(W)]))
        # ---- 3. pass A of the partition == the exchange
        for rel, (t, n) in enumerate(((r, nr), (s, ns))):
            dptr = np.ascontiguousarray(px.ptr_table[rel][owner])          # 512 destination pointers
            doff = np.ascontiguousarray(offs[rel], dtype=np.uint64)
            smj.check(lib.smjdev_msd_scatter(t.data_ptr(), n, shift1, dptr.ctypes.data_as(C.POINTER(C.c_void_p)),
                                             doff.ctypes.data_as(C.POINTER(C.c_uint64))), "msd_scatter")
        dist.all_reduce(self._token(dev), op=dist.ReduceOp.SUM)  # device-side barrier (see step_p2p)
        # ---- 4. pass B + bucket join on what this rank received
        mc = np.ascontiguousarray(mine_counts, dtype=np.uint64)
        st = smj.smj_join_stats_t()
        smj.check(lib.smjdev_join_level1(px.local[0], mc[0].ctypes.data_as(C.POINTER(C.c_uint64)), px.local[1],
                                         mc[1].ctypes.data_as(C.POINTER(C.c_uint64)), shift1, lo, C.byref(st)),
                  "join_level1")
        res = e.result_tensor(st.matches, st.checksum)
        dist.all_reduce(res, op=dist.ReduceOp.SUM)
        out = res.cpu().numpy().view(np.uint64)
        sent = 8 * int(hist[self.rank][:, owner != self.rank].sum())
        return int(out[0]), int(out[1]), sent

    def step_msd_dma(self, r, nr: int, s, ns: int, px: "P2PExchange"):
        """The m-way join across GPUs with a COPY-ENGINE exchange: pass A of the MSD partition
        scatters every rank's slice bucket-major into a local send buffer, so the buckets of one
        owner are one contiguous range; that range is copied into the owner's receive buffer with
        one cudaMemcpyAsync over NVLink on the copy stream, and every rank finishes the buckets it
        owns from the segments it received (smjdev_join_level1_seg).  Contiguous copies cross
        NVLink at ~700 GB/s per direction where the scattered 128-byte stores of step_msd reach
        ~350.  The local scatters need only the LOCAL histograms, on the device: they are enqueued
        behind the all-gather of the histograms and run while the host waits for it and computes
        the layout."""
        e, lib, smj, W = self.e, self.e.lib, self.e.smj, self.world
        dev = e.device
        sends = [e.empty(nr), e.empty(ns)]

        def scatter_local(loc, shift1):
            smj.check(lib.smjdev_msd_scatter_local(r.data_ptr(), nr, shift1, loc[0:512].data_ptr(),
                                                   sends[0].data_ptr()), "msd_scatter_local")
            smj.check(lib.smjdev_msd_scatter_local(s.data_ptr(), ns, shift1, loc[512:1024].data_ptr(),
                                                   sends[1].data_ptr()), "msd_scatter_local")

        pre = self._msd_plan_and_hist(r, nr, s, ns, overlap=scatter_local)
        if pre is None:
            return self.step_p2p(r, nr, s, ns, self._fallback, px)
        lo, hi, shift1, hist = pre
        lay = msd_layout_dma(hist, self.rank)
        first = lay["first"]
        tot = np.asarray(hist, dtype=np.int64).sum(axis=0)
        px.check_capacity(np.array([[int(tot[rel][first[p]:first[p + 1]].sum()) for rel in range(2)] for p in range(W)]))
        smj.check(lib.smjdev_exchange_begin(), "exchange_begin")
        # rank g sends to g+1, g+2, ... (mod W): at any moment every receiver has one sender
        order = sorted(lay["copies"], key=lambda c: (c[0], (c[1] - self.rank) % W))
        for rel in range(2):
            for (crel, p, src, cnt, dst) in order:
                if crel == rel:
                    smj.check(lib.smjdev_exchange_copy(C.c_void_p(px.ptrs[rel][p] + 8 * dst),
                                                       C.c_void_p(sends[rel].data_ptr() + 8 * src), 8 * cnt),
                              "exchange_copy")
            own_n = int(hist[self.rank][rel][first[self.rank]:first[self.rank + 1]].sum())
            if own_n:  # this rank's own buckets: a local copy
                src = int(lay["send_off"][rel][first[self.rank]])
                smj.check(lib.smjdev_exchange_copy(C.c_void_p(px.local[rel] + 8 * lay["self_off"][rel]),
                                                   C.c_void_p(sends[rel].data_ptr() + 8 * src), 8 * own_n),
                          "exchange_copy")
        smj.check(lib.smjdev_exchange_end(), "exchange_end")
        dist.all_reduce(self._token(dev), op=dist.ReduceOp.SUM)  # every rank's copies have landed
        st = smj.smj_join_stats_t()
        (rs, rc, rb), (ss, sc_, sb) = lay["segs"]
        if rs.size and ss.size:
            rseg = np.ascontiguousarray(np.concatenate([rs, rc, rb]), dtype=np.uint64)
            sseg = np.ascontiguousarray(np.concatenate([ss, sc_, sb]), dtype=np.uint64)
            smj.check(lib.smjdev_join_level1_seg(px.local[0], rs.size, rseg.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                 px.local[1], ss.size, sseg.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                 shift1, lo, C.byref(st)), "join_level1_seg")
        res = e.result_tensor(st.matches, st.checksum)
        dist.all_reduce(res, op=dist.ReduceOp.SUM)  # also: nobody copies into a buffer still in use
        out = res.cpu().numpy().view(np.uint64)
        del sends
        sent = 8 * sum(c[3] for c in lay["copies"])
        return int(out[0]), int(out[1]), sent

    def _msd_plan_and_hist(self, r, nr: int, s, ns: int, overlap=None):
        """Steps 1-2 of the MSD exchanges in ONE collective: every rank derives the digit plan from
        the sampled OR/AND of its OWN slices, histograms them by that plan's first digit and
        all-gathers {histograms, exact OR/AND, sampled OR/AND, sizes, the digit it used}.  When all
        ranks used the digit of the global plan (the usual case: every slice spans the key domain)
        the histograms are the ones needed; otherwise they are taken again with the global digit
        (a second collective).  overlap(loc, shift1): called after the all-gather has been ENQUEUED
        and before the host waits for it, with the device tensor of the local histograms (R at
        0:512, S at 512:1024); called again if the digit turns out to be another one.
        Returns (lo, hi, shift1, hist[W][2][512]) or None (then self._fallback holds splitters)."""
        e, lib, smj, W = self.e, self.e.lib, self.e.smj, self.world
        dev = e.device
        b4 = (C.c_uint64 * 4)()
        smj.check(lib.smjdev_sample_key_bits(r.data_ptr(), nr, s.data_ptr(), ns, b4), "sample_key_bits")
        local = msd_plan(int(b4[0]), int(b4[1]), int(b4[2]), int(b4[3])) if nr and ns else None
        guess = local[2] if local is not None else 0

        def gather(shift1):
            """histograms by the digit at shift1 (0: none) + the bookkeeping words, all-gathered"""
            tail = np.array([0, -1, 0, -1] + [int(x) for x in np.array(list(b4), dtype=np.uint64).view(np.int64)]
                            + [nr, ns, shift1], dtype=np.int64)
            loc = torch.zeros(1024 + tail.size, dtype=torch.int64, device=dev)
            loc[1024:] = torch.from_numpy(tail).to(dev, non_blocking=True)
            if shift1:
                smj.check(lib.smjdev_msd_hist(r.data_ptr(), nr, shift1, loc[0:512].data_ptr(),
                                              loc[1024:1026].data_ptr()), "msd_hist")
                smj.check(lib.smjdev_msd_hist(s.data_ptr(), ns, shift1, loc[512:1024].data_ptr(),
                                              loc[1026:1028].data_ptr()), "msd_hist")
            alll = torch.empty(W, loc.numel(), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(alll, loc)
            if getattr(self, "_pin", None) is None or self._pin.shape != alll.shape:
                self._pin = torch.empty(alll.shape, dtype=torch.int64).pin_memory()
            self._pin.copy_(alll, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if overlap is not None and shift1:
                overlap(loc, shift1)  # work that needs only local data runs while the host waits
            ev.synchronize()
            return self._pin.numpy().copy()

        a = gather(guess)
        m = a[:, 1028:1034].view(np.uint64)  # sampled OR_R AND_R OR_S AND_S, nr, ns of every rank
        self._fallback = self._fallback_splitters(m)
        fold = lambda col, op: int(op.reduce(m[:, col]))
        sizes_ok = int(m[:, 4].sum() + m[:, 5].sum()) >= (1 << 20) and m[:, 4].sum() > 0 and m[:, 5].sum() > 0
        plan = msd_plan(fold(0, np.bitwise_or), fold(1, np.bitwise_and), fold(2, np.bitwise_or),
                        fold(3, np.bitwise_and)) if sizes_ok else None
        if plan is None:
            return None
        lo, hi, shift1 = plan
        if not np.all(a[:, 1034] == shift1):  # some rank guessed another digit: once more, all alike
            a = gather(shift1)
        ex = a[:, 1024:1028].view(np.uint64)
        if not msd_plan_holds(int(np.bitwise_or.reduce(ex[:, 0])), int(np.bitwise_and.reduce(ex[:, 1])),
                              int(np.bitwise_or.reduce(ex[:, 2])), int(np.bitwise_and.reduce(ex[:, 3])), lo, hi):
            return None
        return lo, hi, shift1, a[:, :1024].reshape(W, 2, 512)

    def _fallback_splitters(self, m: np.ndarray) -> np.ndarray:
        """Equal-width splitters over the sampled key range (used when step_msd cannot apply)."""
        vor = int(np.bitwise_or.reduce(m[:, 0]) | np.bitwise_or.reduce(m[:, 2])) >> 32
        kmax = (1 << max(vor.bit_length(), 1)) - 1
        return uniform_splitters(kmax, self.world)

    def _token(self, device):
        if getattr(self, "_tok", None) is None:
            self._tok = torch.zeros(1, dtype=torch.int32, device=device)
        return self._tok

    def step(self, r, nr: int, s, ns: int, splitters: np.ndarray):
        """One complete distributed join of the local slices; returns the GLOBAL (count, checksum)
        and the number of item bytes this rank sent to other ranks."""
        pr, cr = self.e.range_partition(r, nr, splitters, self.world)
        ps, cs = self.e.range_partition(s, ns, splitters, self.world)
        rr, nrr, send_r, _ = self.exchange(pr, cr)
        ss, nss, send_s, _ = self.exchange(ps, cs)
        cnt, cks = self.e.join(rr, nrr, ss, nss)
        res = self.e.result_tensor(cnt, cks)
        dist.all_reduce(res, op=dist.ReduceOp.SUM)
        out = res.cpu().numpy().view(np.uint64)
        sent = 8 * (sum(send_r) - send_r[self.rank] + sum(send_s) - send_s[self.rank])
        return int(out[0]), int(out[1]), sent
