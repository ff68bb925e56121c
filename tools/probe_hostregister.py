#!/usr/bin/env python
"""How fast can ordinary (pageable) memory be pinned in place on this host?  cudaHostRegister of a
1 GB numpy buffer: one call, 64 MiB pieces from one thread, 64 MiB pieces from 4 / 8 threads.
(Decides whether registering the caller's relations could replace the bounce-buffer staging.)"""
import ctypes as C
import threading
import time

import numpy as np
import torch

torch.cuda.init()
rt = C.CDLL("libcudart.so.12")
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
rt.cudaHostUnregister.argtypes = [C.c_void_p]
n = 1 << 30
buf = np.ones(n, dtype=np.uint8)
base = buf.ctypes.data
base = (base + 4095) & ~4095
n -= 8192
for rep in range(2):
    t0 = time.perf_counter(); rc = rt.cudaHostRegister(base, n, 0); t1 = time.perf_counter()
    rc2 = rt.cudaHostUnregister(base); t2 = time.perf_counter()
    print(f"one call: register {1e3 * (t1 - t0):.1f} ms ({n / (t1 - t0) / 1e9:.1f} GB/s) rc {rc}, unregister {1e3 * (t2 - t1):.1f} ms rc {rc2}")
piece = 64 << 20
offs = list(range(0, n - piece + 1, piece))
for nth in (1, 4, 8):
    def work(k):
        for o in offs[k::nth]:
            rt.cudaHostRegister(base + o, piece, 0)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k,)) for k in range(nth)]
    [t.start() for t in th]; [t.join() for t in th]
    t1 = time.perf_counter()
    for o in offs:
        rt.cudaHostUnregister(base + o)
    t2 = time.perf_counter()
    print(f"{nth} threads x 64 MiB pieces: register {1e3 * (t1 - t0):.1f} ms ({len(offs) * piece / (t1 - t0) / 1e9:.1f} GB/s), unregister {1e3 * (t2 - t1):.1f} ms")
