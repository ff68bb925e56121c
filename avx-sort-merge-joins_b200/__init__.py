"""avx-sort-merge-joins_b200 -- ctypes binding of libsmj_b200.so.

The product is the shared library (hand-written sm_100a kernels + C host entry points, see
include/smj_b200.h and include/smj_device.h).  This module only loads it and declares the C
signatures so that tests, bench.py and Python callers can reach the reference-compatible
entry points; there is no Python implementation of anything and no CPU fallback: if the
library is missing, importing fails; if no GPU is usable, every compute call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsmj_b200.so")


def build(verbose: bool = False) -> str:
    """Compile libsmj_b200.so in-tree with nvcc for sm_100a (see Makefile)."""
    res = subprocess.run(["make", "-C", _HERE, "-j8", LIB_PATH], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libsmj_b200.so failed")
    return LIB_PATH


# ----------------------------------------------------------------------------- ABI structs
class tuple_t(C.Structure):  # reference src/types.h:51-54
    _fields_ = [("payload", C.c_int32), ("key", C.c_int32)]


class relation_t(C.Structure):  # reference src/types.h:61-64
    _fields_ = [("tuples", C.c_void_p), ("num_tuples", C.c_uint64)]


class threadresult_t(C.Structure):  # reference src/types.h:68-72
    _fields_ = [("nresults", C.c_int64), ("results", C.c_void_p), ("threadid", C.c_uint32)]


class result_t(C.Structure):  # reference src/types.h:75-79
    _fields_ = [("totalresults", C.c_int64), ("resultlist", C.POINTER(threadresult_t)),
                ("nthreads", C.c_int)]


class joinconfig_t(C.Structure):  # reference src/types.h:91-98
    _fields_ = [("NTHREADS", C.c_int), ("PARTFANOUT", C.c_int), ("SCALARSORT", C.c_int),
                ("SCALARMERGE", C.c_int), ("MWAYMERGEBUFFERSIZE", C.c_int),
                ("NUMASTRATEGY", C.c_int)]


class smj_join_stats_t(C.Structure):  # include/smj_device.h
    _fields_ = [("matches", C.c_uint64), ("checksum", C.c_uint64), ("ms_h2d", C.c_float),
                ("ms_partition", C.c_float), ("ms_sort", C.c_float), ("ms_merge", C.c_float),
                ("ms_join", C.c_float), ("ms_total", C.c_float), ("launches", C.c_uint64),
                ("hbm_bytes", C.c_uint64), ("path", C.c_uint32), ("repeats", C.c_uint32)]


NUMA_RANDOM, NUMA_RING, NUMA_NEXT = 0, 1, 2
ALGO_MWAY, ALGO_MPASS, ALGO_MPSM = 0, 1, 2
PATH_GENERAL, PATH_MSD, PATH_MSD_OPT = 0, 1, 2
OUT_S_TUPLE, OUT_RID_PAIR = 0, 1
GROUP_HANDLE_BYTES = 256

_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p

# every exported symbol with its signature: (restype, argtypes).  tests/test_abi.py checks that
# this table, include/*.h and the .so agree.
SIGNATURES = {
    # --- device layer (include/smj_device.h)
    "smjdev_device_count": (C.c_int, []),
    "smjdev_init": (C.c_int, [C.c_int]),
    "smjdev_current_device": (C.c_int, []),
    "smjdev_set_stream": (C.c_int, [_vp]),
    "smjdev_use_own_stream": (C.c_int, []),
    "smjdev_get_stream": (_vp, []),
    "smjdev_sync": (C.c_int, []),
    "smjdev_last_error": (C.c_char_p, []),
    "smjdev_launch_count": (C.c_uint64, []),
    "smjdev_sm_count": (C.c_int, []),
    "smjdev_malloc": (_vp, [C.c_size_t]),
    "smjdev_free": (C.c_int, [_vp]),
    "smjdev_host_alloc": (_vp, [C.c_size_t]),
    "smjdev_host_free": (C.c_int, [_vp]),
    "smjdev_memcpy_h2d": (C.c_int, [_vp, _vp, C.c_size_t]),
    "smjdev_memcpy_d2h": (C.c_int, [_vp, _vp, C.c_size_t]),
    "smjdev_memcpy_d2d": (C.c_int, [_vp, _vp, C.c_size_t]),
    "smjdev_upload": (C.c_int, [_vp, _vp, C.c_size_t]),
    "smjdev_download": (C.c_int, [_vp, _vp, C.c_size_t]),
    "smjdev_memset": (C.c_int, [_vp, C.c_int, C.c_size_t]),
    "smjdev_pointer_kind": (C.c_int, [_vp]),
    "smjdev_event_create": (C.c_int, [C.POINTER(_vp)]),
    "smjdev_event_record": (C.c_int, [_vp]),
    "smjdev_event_elapsed_ms": (C.c_int, [_vp, _vp, C.POINTER(C.c_float)]),
    "smjdev_event_destroy": (C.c_int, [_vp]),
    "smjdev_profile_begin": (C.c_int, []),
    "smjdev_profile_end": (C.c_int, [C.c_char_p, C.c_size_t]),
    "smjdev_sort": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, C.c_int, C.POINTER(_vp)]),
    "smjdev_sort_int32": (C.c_int, [_vp, C.c_uint64]),
    "smjdev_partition": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                   C.c_int, _vp, _vp]),
    "smjdev_histogram": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _vp]),
    "smjdev_merge": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_int]),
    "smjdev_multiway_merge": (C.c_int, [_vp, _vp, C.POINTER(_vp), _u64p, C.c_uint32, C.c_int]),
    "smjdev_merge_join": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, _u64p, _u64p]),
    "smjdev_merge_join_ex": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, C.c_int, _u64p, _u64p]),
    "smjdev_merge_join_acc": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, _vp]),
    "smjdev_join": (C.c_int, [C.c_int, _vp, C.c_uint64, _vp, C.c_uint64, C.c_int,
                              C.POINTER(smj_join_stats_t)]),
    "smjdev_join_materialize": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, C.c_int, _vp, C.c_uint64, _u64p]),
    "smjdev_range_partition": (C.c_int, [_vp, _vp, C.c_uint64, C.POINTER(C.c_uint32), C.c_uint32,
                                         _vp]),
    "smjdev_sample_key_bits": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, _u64p]),
    "smjdev_msd_hist": (C.c_int, [_vp, C.c_uint64, C.c_uint32, _vp, _vp]),
    "smjdev_msd_scatter": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.POINTER(_vp), _u64p]),
    "smjdev_join_level1": (C.c_int, [_vp, _u64p, _vp, _u64p, C.c_uint32, C.c_uint32,
                                     C.POINTER(smj_join_stats_t)]),
    "smjdev_msd_scatter_local": (C.c_int, [_vp, C.c_uint64, C.c_uint32, _vp, _vp]),
    "smjdev_join_level1_seg": (C.c_int, [_vp, C.c_uint32, _u64p, _vp, C.c_uint32, _u64p, C.c_uint32, C.c_uint32,
                                         C.POINTER(smj_join_stats_t)]),
    "smjdev_debug_copy": (C.c_int, [_vp, _vp, C.c_uint64, C.c_int, C.c_uint32, C.c_uint32]),
    "smjdev_exchange_begin": (C.c_int, []),
    "smjdev_exchange_copy": (C.c_int, [_vp, _vp, C.c_size_t]),
    "smjdev_exchange_end": (C.c_int, []),
    "smjdev_range_count": (C.c_int, [_vp, C.c_uint64, C.POINTER(C.c_uint32), C.c_uint32, _vp]),
    "smjdev_range_scatter": (C.c_int, [_vp, C.c_uint64, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(_vp),
                                       _u64p]),
    "smjdev_ipc_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "smjdev_ipc_free": (C.c_int, [_vp]),
    "smjdev_ipc_export": (C.c_int, [_vp, C.c_char_p]),
    "smjdev_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "smjdev_ipc_close": (C.c_int, [_vp]),
    "smjdev_sort_key8b": (C.c_int, [_vp, _vp, C.c_uint64]),
    "smjdev_join_key8b": (C.c_int, [_vp, C.c_uint64, _vp, C.c_uint64, C.c_int, C.POINTER(smj_join_stats_t)]),
    "smjdev_group_create": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(_vp)]),
    "smjdev_group_export": (C.c_int, [_vp, C.c_char_p]),
    "smjdev_group_connect": (C.c_int, [_vp, C.c_char_p]),
    "smjdev_group_join": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_uint64, C.POINTER(smj_join_stats_t), _u64p]),
    "smjdev_group_destroy": (C.c_int, [_vp]),
    "smjdev_join_multi": (C.c_int, [C.c_int, _vp, C.c_uint64, _vp, C.c_uint64, C.POINTER(smj_join_stats_t)]),
    "smjdev_join_multi_release": (C.c_int, []),
    "smjdev_group_layout_host": (C.c_uint32, [C.c_int, C.c_int, _u64p, C.c_uint32, C.POINTER(C.c_uint32), _u64p, _u64p,
                                              _u64p, _u64p]),
    "smjdev_check_order": (C.c_int, [_vp, C.c_uint64, C.c_int, _u64p]),
    "smjdev_check_partition": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_int, C.c_uint64,
                                         _u64p]),
    "smjdev_gen_unique": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "smjdev_gen_fk": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]),
    "smjdev_gen_fk_from": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _vp, C.c_uint64, C.c_uint64]),
    "smjdev_zipf_lut": (C.c_int, [_vp, C.c_uint32, C.c_double, _vp]),
    "smjdev_gen_zipf": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, C.c_uint64, C.c_uint64,
                                  _vp]),
    # --- reference-facing entry points (include/smj_b200.h)
    "sortmergejoin_multiway": (C.POINTER(result_t), [C.POINTER(relation_t), C.POINTER(relation_t),
                                                     C.POINTER(joinconfig_t)]),
    "sortmergejoin_multipass": (C.POINTER(result_t), [C.POINTER(relation_t), C.POINTER(relation_t),
                                                      C.POINTER(joinconfig_t)]),
    "sortmergejoin_mpsm": (C.POINTER(result_t), [C.POINTER(relation_t), C.POINTER(relation_t),
                                                 C.POINTER(joinconfig_t)]),
    "sortmergejoin_multiway_materialize": (C.POINTER(result_t), [C.POINTER(relation_t), C.POINTER(relation_t),
                                                                 C.POINTER(joinconfig_t)]),
    "sortmergejoin_initrun": (C.POINTER(result_t), [C.POINTER(relation_t), C.POINTER(relation_t),
                                                    C.POINTER(joinconfig_t), _vp]),
    "sortmergejoin_multiway_key8b": (C.POINTER(result_t), [C.POINTER(relation_t), C.POINTER(relation_t),
                                                           C.POINTER(joinconfig_t)]),
    "sortmergejoin_multipass_key8b": (C.POINTER(result_t), [C.POINTER(relation_t), C.POINTER(relation_t),
                                                            C.POINTER(joinconfig_t)]),
    "sortmergejoin_mpsm_key8b": (C.POINTER(result_t), [C.POINTER(relation_t), C.POINTER(relation_t),
                                                       C.POINTER(joinconfig_t)]),
    "scalarsort_tuples_key8b": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "merge_join_key8b": (C.c_uint64, [_vp, _vp, C.c_uint64, C.c_uint64, _vp]),
    "smj_last_join_stats": (C.c_int, [C.POINTER(smj_join_stats_t)]),
    "merge_join": (C.c_uint64, [_vp, _vp, C.c_uint64, C.c_uint64, _vp]),
    "merge_join_interpolation": (C.c_uint64, [_vp, _vp, C.c_uint64, C.c_uint64, _vp]),
    "print_timing": (None, [C.c_uint64, _vp, _vp, _vp]),
    "is_sorted_helper": (C.c_int, [_vp, C.c_uint64]),
    "check_sorted": (None, [_vp, _vp, C.c_uint64, C.c_uint64, C.c_int]),
    "numa_shuffle_init": (None, [C.c_int, C.c_int]),
    "get_numa_shuffle_strategy": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "avxsort_tuples": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "avxsort_int64": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "avxsort_int32": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "avxsortmultiway_tuples": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "avxsortmultiway_int64": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "scalarsort_tuples": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "scalarsort_int64": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "scalarsort_int32": (None, [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]),
    "smj_avxsort_placement": (C.c_int, [C.c_uint64, C.c_int]),
    "radix_cluster": (None, [C.POINTER(relation_t), C.POINTER(relation_t), C.POINTER(C.c_int32),
                             C.c_int, C.c_int]),
    "partition_relation": (None, [C.POINTER(C.POINTER(relation_t)), C.POINTER(relation_t),
                                  C.POINTER(relation_t), C.c_int, C.c_int]),
    "partition_relation_optimized": (None, [C.POINTER(C.POINTER(relation_t)), C.POINTER(relation_t),
                                            C.POINTER(relation_t), C.c_uint32, C.c_uint32]),
    "partition_relation_optimized_V2": (None, [C.POINTER(C.POINTER(relation_t)),
                                               C.POINTER(relation_t), C.POINTER(relation_t),
                                               C.c_uint32, C.c_uint32]),
    "histogram_memcpy_bench": (None, [C.POINTER(C.POINTER(relation_t)), C.POINTER(relation_t),
                                      C.POINTER(relation_t), C.c_uint32]),
    "avx_merge_tuples": (C.c_uint64, [_vp, _vp, _vp, C.c_uint64, C.c_uint64]),
    "avx_merge_int64": (C.c_uint64, [_vp, _vp, _vp, C.c_uint64, C.c_uint64]),
    "scalar_merge_int64": (C.c_uint64, [_vp, _vp, _vp, C.c_uint64, C.c_uint64]),
    "scalar_merge_tuples": (C.c_uint64, [_vp, _vp, _vp, C.c_uint64, C.c_uint64]),
    "avx_multiway_merge": (C.c_uint64, [_vp, C.POINTER(C.POINTER(relation_t)), C.c_uint32, _vp,
                                        C.c_uint32]),
    "scalar_multiway_merge": (C.c_uint64, [_vp, C.POINTER(C.POINTER(relation_t)), C.c_uint32, _vp,
                                           C.c_uint32]),
    "scalar_multiway_merge_modulo": (C.c_uint64, [_vp, C.POINTER(C.POINTER(relation_t)),
                                                  C.c_uint32, _vp, C.c_uint32]),
    "scalar_multiway_merge_bitand": (C.c_uint64, [_vp, C.POINTER(C.POINTER(relation_t)),
                                                  C.c_uint32, _vp, C.c_uint32]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no Python or CPU fallback for the sort-merge join kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class SmjError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SmjError(f"{what} failed ({rc}): {lib.smjdev_last_error().decode()}")


def have_gpu() -> bool:
    return lib.smjdev_device_count() > 0


# ----------------------------------------------------------------------------- helpers
def as_items(a) -> np.ndarray:
    """View a relation (uint64 items or an array of (payload,key) int32 pairs) as uint64 items."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        return a
    if a.dtype == np.int64:
        return a.view(np.uint64)
    raise TypeError("relations are uint64 item arrays (key<<32 | payload)")


def make_items(keys, payloads) -> np.ndarray:
    keys = np.asarray(keys).astype(np.uint64)
    payloads = np.asarray(payloads).astype(np.uint64) & np.uint64(0xFFFFFFFF)
    return (keys << np.uint64(32)) | payloads


def rel(arr: np.ndarray) -> relation_t:
    """relation_t over a host numpy array (keeps no reference: keep `arr` alive)."""
    return relation_t(arr.ctypes.data, arr.size)


class DeviceBuffer:
    """n uint64 items in HBM, owned by the library's stream-ordered pool."""

    def __init__(self, n: int):
        self.n = int(n)
        self.ptr = lib.smjdev_malloc(max(self.n, 1) * 8)
        if not self.ptr:
            raise SmjError("smjdev_malloc: " + lib.smjdev_last_error().decode())

    @classmethod
    def from_host(cls, arr: np.ndarray) -> "DeviceBuffer":
        arr = as_items(arr)
        b = cls(arr.size)
        if arr.size:
            check(lib.smjdev_memcpy_h2d(b.ptr, arr.ctypes.data, arr.nbytes), "h2d")
            check(lib.smjdev_sync(), "sync")
        return b

    def to_host(self, n: int | None = None, offset: int = 0, dtype=np.uint64) -> np.ndarray:
        n = self.n - offset if n is None else n
        out = np.empty(n, dtype=dtype)
        if n:
            check(lib.smjdev_memcpy_d2h(out.ctypes.data, self.ptr + offset * out.itemsize,
                                        out.nbytes), "d2h")
            check(lib.smjdev_sync(), "sync")
        return out

    def free(self) -> None:
        if self.ptr:
            lib.smjdev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def device_join(algo: int, r: DeviceBuffer, s: DeviceBuffer, nruns: int = 1) -> smj_join_stats_t:
    st = smj_join_stats_t()
    check(lib.smjdev_join(algo, r.ptr, r.n, s.ptr, s.n, nruns, C.byref(st)), "smjdev_join")
    return st


def join_host(name: str, r: np.ndarray, s: np.ndarray, nthreads: int = 1, fanout: int = 128):
    """Call sortmergejoin_{multiway,multipass,mpsm} exactly like the reference's main.c:603 does,
    with host buffers.  Returns (totalresults, stats)."""
    r, s = as_items(r), as_items(s)
    cfg = joinconfig_t(nthreads, fanout, 0, 0, 20 * 1024 * 1024, NUMA_NEXT)
    rr, ss = rel(r), rel(s)
    fn = getattr(lib, "sortmergejoin_" + name)
    res = fn(C.byref(rr), C.byref(ss), C.byref(cfg))
    if not res:
        return None, None
    total = res.contents.totalresults
    libc = C.CDLL(None)
    libc.free.argtypes = [_vp]
    libc.free(C.cast(res.contents.resultlist, _vp))
    libc.free(C.cast(res, _vp))
    st = smj_join_stats_t()
    lib.smj_last_join_stats(C.byref(st))
    return total, st
