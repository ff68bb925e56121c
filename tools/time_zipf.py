#!/usr/bin/env python
"""Per-kernel timing of the m-way join on a Zipf-skewed foreign-key workload (BASELINE config 4,
scaled): python tools/time_zipf.py [theta] [|R|] [|S|]"""
import ctypes as C
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("SMJ_QUIET", "1")
smj = importlib.import_module("avx-sort-merge-joins_b200")
theta = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 128_000_000
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 1_280_000_000
lib = smj.lib
smj.check(lib.smjdev_init(0), "init")
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(54321)
lut = torch.cumsum(torch.arange(1, nr + 1, dtype=torch.float64, device=dev).pow_(-theta), 0)
lut /= lut[-1].clone()
alphabet = torch.randperm(nr, device=dev, generator=g) + 1
r = smj.DeviceBuffer(nr)
smj.check(lib.smjdev_gen_unique(r.ptr, nr, 0, nr, 12345), "gen")
s = torch.empty(ns, dtype=torch.int64, device=dev)
chunk = 1 << 27
for lo in range(0, ns, chunk):
    hi = min(ns, lo + chunk)
    u = torch.rand(hi - lo, dtype=torch.float64, device=dev, generator=g)
    pos = torch.searchsorted(lut, u).clamp_(max=nr - 1)
    s[lo:hi] = (alphabet[pos] << 32) | torch.arange(lo, hi, dtype=torch.int64, device=dev).bitwise_and_(0xFFFFFFFF)
del lut, alphabet
torch.cuda.synchronize()
st = smj.smj_join_stats_t()
for _ in range(2):
    smj.check(lib.smjdev_join(smj.ALGO_MWAY, r.ptr, nr, s.data_ptr(), ns, 1, C.byref(st)), "m-way")
lib.smjdev_profile_begin()
smj.check(lib.smjdev_join(smj.ALGO_MWAY, r.ptr, nr, s.data_ptr(), ns, 1, C.byref(st)), "m-way")
buf = C.create_string_buffer(1 << 16)
lib.smjdev_profile_end(buf, len(buf))
prof = {}
for line in buf.value.decode().strip().splitlines():
    name, cnt, t, items = line.rsplit(";", 3)
    prof[name.strip("() ")[:28]] = (int(cnt), round(float(t), 3))
print(f"theta {theta}: ms {st.ms_total:.2f} matches {st.matches} (|S| {ns}) path {st.path} repeats {st.repeats}")
for k, v in prof.items():
    print(f"   {k:30s} x{v[0]:<3d} {v[1]:9.3f} ms")
