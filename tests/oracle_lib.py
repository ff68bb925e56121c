"""Test-side loaders for the CPU checker libraries (TEST INFRASTRUCTURE ONLY).

* oracle/libsmj_oracle.so  -- plain-C restatement of the reference path (oracle/smj_oracle.c)
* oracle/_ref/libsmj_ref.so -- the unmodified reference compiled by `make -C oracle ref`
                               (present in the dev container and shipped prebuilt to the GPU box)
Nothing in the product imports this module.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libsmj_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsmj_ref.so")

_u64p = C.POINTER(C.c_uint64)
_vp = C.c_void_p


def _p(a: np.ndarray):
    return a.ctypes.data_as(_vp)


@contextlib.contextmanager
def silence_stdout():
    """The reference fork prints a debug line per partition sort (avxsort_core.h:2258)."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        yield
    finally:
        C.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


@contextlib.contextmanager
def silence_all():
    sys.stdout.flush()
    sys.stderr.flush()
    saved = (os.dup(1), os.dup(2))
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    os.dup2(devnull, 2)
    try:
        yield
    finally:
        C.CDLL(None).fflush(None)
        os.dup2(saved[0], 1)
        os.dup2(saved[1], 2)
        os.close(saved[0])
        os.close(saved[1])
        os.close(devnull)


# --------------------------------------------------------------------------------- oracle
class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO) or (
                os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "smj_oracle.c"))):
            subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
        L = self.lib = C.CDLL(ORACLE_SO)
        L.orc_sort.argtypes = [_vp, C.c_uint64]
        L.orc_sort_key_stable.argtypes = [_vp, C.c_uint64]
        L.orc_avxsort_placement.argtypes = [C.c_uint64, C.c_int]
        L.orc_partition.argtypes = [_vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp]
        L.orc_partition.restype = C.c_uint64
        L.orc_merge.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_int]
        L.orc_merge.restype = C.c_uint64
        L.orc_multiway_merge.argtypes = [C.POINTER(_vp), _u64p, C.c_uint32, _vp]
        L.orc_multiway_merge.restype = C.c_uint64
        for f in (L.orc_merge_join, L.orc_merge_join_runs, L.orc_join_bruteforce):
            f.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64, _u64p, _u64p]
            f.restype = None
        L.orc_merge_join_interpolation.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64]
        L.orc_merge_join_interpolation.restype = C.c_uint64
        L.orc_join_mway.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64, C.c_int, C.c_int, _u64p, _u64p]
        L.orc_join_materialize.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64, C.c_int, _vp, C.c_uint64]
        L.orc_join_materialize.restype = C.c_uint64
        L.orc_gen_unique.argtypes = [_vp, C.c_uint64, C.c_uint64]
        L.orc_gen_fk.argtypes = [_vp, C.c_uint64, C.c_uint32, C.c_uint64]
        L.orc_sort16.argtypes = [_vp, C.c_uint64]
        L.orc_join16.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64, _u64p, _u64p]
        L.orc_gen_fk_from.argtypes = [_vp, C.c_uint64, _vp, C.c_uint64, C.c_uint64]
        L.orc_gen_zipf_ref.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_double, _vp, _vp]
        L.orc_zipf_weight.argtypes = [C.c_uint32, C.c_double]
        L.orc_zipf_weight.restype = C.c_double
        L.orc_zipf_lut.argtypes = [_vp, C.c_uint32, C.c_double]
        L.orc_gen_zipf.argtypes = [_vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_double, C.c_uint64,
                                   C.c_uint64, _vp]

    # sort
    def sort(self, items: np.ndarray) -> np.ndarray:
        out = np.array(items, dtype=np.uint64, copy=True)
        self.lib.orc_sort(_p(out), out.size)
        return out

    def sort_key_stable(self, items: np.ndarray) -> np.ndarray:
        out = np.array(items, dtype=np.uint64, copy=True)
        self.lib.orc_sort_key_stable(_p(out), out.size)
        return out

    def placement(self, n: int, multiway: int = 0) -> int:
        return self.lib.orc_avxsort_placement(n, multiway)

    # partition -> (out, hist, offs, used)
    def partition(self, items: np.ndarray, nbits: int, shift: int, padded: bool):
        items = np.ascontiguousarray(items, dtype=np.uint64)
        fan = 1 << nbits
        out = np.zeros(items.size + 8 * fan + 8, dtype=np.uint64)
        hist = np.zeros(fan, dtype=np.uint64)
        offs = np.zeros(fan, dtype=np.uint64)
        used = self.lib.orc_partition(_p(items), items.size, nbits, shift, int(padded), _p(out),
                                      _p(hist), _p(offs))
        return out, hist, offs, used

    def merge(self, a: np.ndarray, b: np.ndarray, mode: int = 0) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.empty(a.size + b.size, dtype=np.uint64)
        self.lib.orc_merge(_p(a), a.size, _p(b), b.size, _p(out), mode)
        return out

    def multiway_merge(self, runs) -> np.ndarray:
        runs = [np.ascontiguousarray(r, dtype=np.uint64) for r in runs]
        total = sum(r.size for r in runs)
        out = np.empty(total, dtype=np.uint64)
        ptrs = (_vp * len(runs))(*[r.ctypes.data for r in runs])
        lens = (C.c_uint64 * len(runs))(*[r.size for r in runs])
        self.lib.orc_multiway_merge(ptrs, lens, len(runs), _p(out))
        return out

    def _cc(self, fn, r, s):
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        cnt, cks = C.c_uint64(0), C.c_uint64(0)
        fn(_p(r), r.size, _p(s), s.size, C.byref(cnt), C.byref(cks))
        return cnt.value, cks.value

    def merge_join(self, r, s):
        return self._cc(self.lib.orc_merge_join, r, s)

    def merge_join_runs(self, r, s):
        return self._cc(self.lib.orc_merge_join_runs, r, s)

    def merge_join_interpolation(self, r, s) -> int:
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        return self.lib.orc_merge_join_interpolation(_p(r), r.size, _p(s), s.size)

    def join_bruteforce(self, r, s):
        return self._cc(self.lib.orc_join_bruteforce, r, s)

    def join_materialize(self, r, s, layout: int = 0) -> np.ndarray:
        """the join result written out (joincommon.c:256-278), in the reference loop's order"""
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        n = self.lib.orc_join_materialize(_p(r), r.size, _p(s), s.size, layout, None, 0)
        out = np.empty(n, dtype=np.uint64)
        got = self.lib.orc_join_materialize(_p(r), r.size, _p(s), s.size, layout, _p(out), n)
        assert got == n
        return out

    def join_mway(self, r, s, nthreads: int, fanout: int):
        r = np.ascontiguousarray(r, dtype=np.uint64)
        s = np.ascontiguousarray(s, dtype=np.uint64)
        cnt, cks = C.c_uint64(0), C.c_uint64(0)
        rc = self.lib.orc_join_mway(_p(r), r.size, _p(s), s.size, nthreads, fanout, C.byref(cnt),
                                    C.byref(cks))
        assert rc == 0
        return cnt.value, cks.value

    def gen_unique(self, n: int, seed: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        self.lib.orc_gen_unique(_p(out), n, seed)
        return out

    def gen_fk(self, n: int, maxkey: int, seed: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint64)
        self.lib.orc_gen_fk(_p(out), n, maxkey, seed)
        return out

    # KEY_8B: relations are int64 arrays of shape (n, 2) = {payload, key}
    def sort16(self, t: np.ndarray) -> np.ndarray:
        out = np.ascontiguousarray(t, dtype=np.int64).copy()
        self.lib.orc_sort16(_p(out), out.shape[0])
        return out

    def join16(self, r: np.ndarray, s: np.ndarray):
        r, s = np.ascontiguousarray(r, dtype=np.int64), np.ascontiguousarray(s, dtype=np.int64)
        cnt, cks = C.c_uint64(0), C.c_uint64(0)
        self.lib.orc_join16(_p(r), r.shape[0], _p(s), s.shape[0], C.byref(cnt), C.byref(cks))
        return cnt.value, cks.value

    def gen_fk_from(self, n: int, src: np.ndarray, seed: int) -> np.ndarray:
        src = np.ascontiguousarray(src, dtype=np.uint64)
        out = np.empty(n, dtype=np.uint64)
        self.lib.orc_gen_fk_from(_p(out), n, _p(src), src.size, seed)
        return out

    def gen_zipf_ref(self, n: int, alphabet: int, theta: float, srand_seed: int):
        """genzipf.c:98-159 literally, under srand(srand_seed).  Returns (items with payload = row
        index, alphabet, lut)."""
        out = np.arange(n, dtype=np.uint64)
        alpha = np.empty(alphabet, dtype=np.uint32)
        lut = np.empty(alphabet, dtype=np.float64)
        C.CDLL(None).srand(srand_seed)
        self.lib.orc_gen_zipf_ref(_p(out), n, alphabet, theta, _p(alpha), _p(lut))
        return out, alpha, lut

    def zipf_lut(self, alphabet: int, theta: float) -> np.ndarray:
        lut = np.empty(alphabet, dtype=np.float64)
        self.lib.orc_zipf_lut(_p(lut), alphabet, theta)
        return lut

    def gen_zipf(self, n: int, alphabet: int, theta: float, seed: int, alphabet_seed: int,
                 first: int = 0) -> np.ndarray:
        """what smjdev_gen_zipf generates (rows first .. first + n)"""
        out = np.empty(n, dtype=np.uint64)
        self.lib.orc_gen_zipf(_p(out), first, n, alphabet, theta, seed, alphabet_seed, None)
        return out


# --------------------------------------------------------------------------------- reference
class _Rel(C.Structure):
    _fields_ = [("tuples", _vp), ("num_tuples", C.c_uint64)]


class _Res(C.Structure):
    _fields_ = [("totalresults", C.c_int64), ("resultlist", _vp), ("nthreads", C.c_int)]


class _Cfg(C.Structure):
    _fields_ = [("NTHREADS", C.c_int), ("PARTFANOUT", C.c_int), ("SCALARSORT", C.c_int),
                ("SCALARMERGE", C.c_int), ("MWAYMERGEBUFFERSIZE", C.c_int), ("NUMASTRATEGY", C.c_int)]


def aligned_u64(n: int, align: int = 64) -> np.ndarray:
    """n uint64 in a buffer whose start is `align`-byte aligned (the reference picks its aligned
    AVX kernels only then, src/avxsort/avxsort.c:218-222)."""
    raw = np.zeros(n * 8 + align + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * 8].view(np.uint64)


class Reference:
    """The UNMODIFIED reference (C++-mangled symbols)."""
    SYMS = {
        "avxsort_tuples": "_Z14avxsort_tuplesPP7tuple_tS1_m",
        "avxsortmultiway_tuples": "_Z22avxsortmultiway_tuplesPP7tuple_tS1_m",
        "scalarsort_tuples": "_Z17scalarsort_tuplesPP7tuple_tS1_m",
        "avx_merge_tuples": "_Z16avx_merge_tuplesP7tuple_tS0_S0_mm",
        "scalar_merge_tuples": "_Z19scalar_merge_tuplesP7tuple_tS0_S0_mm",
        "avx_multiway_merge": "_Z18avx_multiway_mergeP7tuple_tPP10relation_tjS0_j",
        "scalar_multiway_merge": "_Z21scalar_multiway_mergeP7tuple_tPP10relation_tjS0_j",
        "partition_relation": "_Z18partition_relationPP10relation_tS0_S0_ii",
        "partition_relation_optimized": "_Z28partition_relation_optimizedPP10relation_tS0_S0_jj",
        "partition_relation_optimized_V2": "_Z31partition_relation_optimized_V2PP10relation_tS0_S0_jj",
        "merge_join": "_Z10merge_joinP7tuple_tS0_mmPv",
        "merge_join_interpolation": "_Z24merge_join_interpolationP7tuple_tS0_mmPv",
        "sortmergejoin_multiway": "_Z22sortmergejoin_multiwayP10relation_tS0_P12joinconfig_t",
        "sortmergejoin_multipass": "_Z23sortmergejoin_multipassP10relation_tS0_P12joinconfig_t",
        "numa_shuffle_init": "_Z17numa_shuffle_init15numa_strategy_ti",
        "cpu_mapping_init": "_Z16cpu_mapping_initv",
    }

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self):
        L = self.lib = C.CDLL(REF_SO)
        self.f = {k: getattr(L, v) for k, v in self.SYMS.items()}
        for k in ("avxsort_tuples", "avxsortmultiway_tuples", "scalarsort_tuples"):
            self.f[k].argtypes = [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]
            self.f[k].restype = None
        for k in ("avx_merge_tuples", "scalar_merge_tuples"):
            self.f[k].argtypes = [_vp, _vp, _vp, C.c_uint64, C.c_uint64]
            self.f[k].restype = C.c_uint64
        for k in ("avx_multiway_merge", "scalar_multiway_merge"):
            self.f[k].argtypes = [_vp, C.POINTER(C.POINTER(_Rel)), C.c_uint32, _vp, C.c_uint32]
            self.f[k].restype = C.c_uint64
        self.f["partition_relation"].argtypes = [C.POINTER(C.POINTER(_Rel)), C.POINTER(_Rel),
                                                 C.POINTER(_Rel), C.c_int, C.c_int]
        for k in ("partition_relation_optimized", "partition_relation_optimized_V2"):
            self.f[k].argtypes = [C.POINTER(C.POINTER(_Rel)), C.POINTER(_Rel), C.POINTER(_Rel),
                                  C.c_uint32, C.c_uint32]
        for k in ("merge_join", "merge_join_interpolation"):
            self.f[k].argtypes = [_vp, _vp, C.c_uint64, C.c_uint64, _vp]
            self.f[k].restype = C.c_uint64
        for k in ("sortmergejoin_multiway", "sortmergejoin_multipass"):
            self.f[k].argtypes = [C.POINTER(_Rel), C.POINTER(_Rel), C.POINTER(_Cfg)]
            self.f[k].restype = C.POINTER(_Res)
        self.f["numa_shuffle_init"].argtypes = [C.c_int, C.c_int]
        with silence_stdout():
            self.f["cpu_mapping_init"]()

    def gen_zipf(self, n: int, alphabet: int, theta: float, srand_seed: int) -> np.ndarray:
        """the compiled reference's gen_zipf (src/datagen/genzipf.c:98-159) under srand(srand_seed);
        payload = row index (the reference writes keys only)"""
        fn = self.lib._Z8gen_zipfjjdPP7tuple_t
        fn.argtypes = [C.c_uint, C.c_uint, C.c_double, C.POINTER(_vp)]
        fn.restype = _vp
        out = np.arange(n, dtype=np.uint64)
        po = _vp(out.ctypes.data)
        C.CDLL(None).srand(srand_seed)
        fn(n, alphabet, theta, C.byref(po))
        assert po.value == out.ctypes.data
        return out

    def sort(self, items: np.ndarray, which: str = "avxsort_tuples"):
        """returns (sorted array, placement) ; placement 1 = result in the output buffer"""
        n = items.size
        a, b = aligned_u64(max(n, 1)), aligned_u64(max(n, 1))
        a[:n] = items
        pa, pb = _vp(a.ctypes.data), _vp(b.ctypes.data)
        with silence_stdout():
            self.f[which](C.byref(pa), C.byref(pb), n)
        placement = 1 if pb.value == b.ctypes.data else 0
        res = b if placement else a
        assert pb.value in (a.ctypes.data, b.ctypes.data)
        return res[:n].copy(), placement

    def merge(self, a: np.ndarray, b: np.ndarray, which: str = "avx_merge_tuples") -> np.ndarray:
        # the AVX kernels may write up to 16 items into their inputs (avxsort_core.h:460-475)
        aa, bb = aligned_u64(a.size + 32), aligned_u64(b.size + 32)
        aa[:a.size], bb[:b.size] = a, b
        out = aligned_u64(a.size + b.size + 32)
        with silence_stdout():
            k = self.f[which](aa.ctypes.data, bb.ctypes.data, out.ctypes.data, a.size, b.size)
        assert k == a.size + b.size
        return out[:k].copy()

    def multiway_merge(self, runs, which: str = "avx_multiway_merge", bufntuples: int = 1 << 16):
        bufs = []
        rels = (_Rel * len(runs))()
        ptrs = (C.POINTER(_Rel) * len(runs))()
        total = 0
        for i, r in enumerate(runs):
            b = aligned_u64(r.size + 64)
            b[:r.size] = r
            bufs.append(b)
            rels[i] = _Rel(b.ctypes.data, r.size)
            ptrs[i] = C.pointer(rels[i])
            total += r.size
        out = aligned_u64(total + 64)
        fifo = aligned_u64(bufntuples + 64)
        with silence_stdout():
            k = self.f[which](out.ctypes.data, ptrs, len(runs), fifo.ctypes.data, bufntuples)
        return out[:k].copy(), k

    def partition(self, items: np.ndarray, nbits: int, shift: int, which: str):
        """returns (out buffer, hist, offs) with offs in items relative to the out buffer"""
        n, fan = items.size, 1 << nbits
        inp = aligned_u64(n + 8)
        inp[:n] = items
        out = aligned_u64(n + 8 * fan + 64)
        rin, rout = _Rel(inp.ctypes.data, n), _Rel(out.ctypes.data, n)
        parts = (_Rel * fan)()
        pp = (C.POINTER(_Rel) * fan)(*[C.pointer(parts[i]) for i in range(fan)])
        with silence_stdout():
            self.f[which](pp, C.byref(rin), C.byref(rout), nbits, shift)
        hist = np.array([parts[i].num_tuples for i in range(fan)], dtype=np.uint64)
        offs = np.array([(parts[i].tuples - out.ctypes.data) // 8 for i in range(fan)], dtype=np.uint64)
        return out, hist, offs

    def merge_join(self, r: np.ndarray, s: np.ndarray, which: str = "merge_join") -> int:
        rr, ss = aligned_u64(r.size + 8), aligned_u64(s.size + 8)
        rr[:r.size], ss[:s.size] = r, s
        return self.f[which](rr.ctypes.data, ss.ctypes.data, r.size, s.size, None)

    def join(self, r: np.ndarray, s: np.ndarray, which: str = "sortmergejoin_multiway",
             nthreads: int = 2, fanout: int = 128, mwaybuf: int = 1 << 22) -> int:
        return self.join_timed(self.join_prepare(r, s, nthreads, fanout), which, nthreads, fanout, mwaybuf)[0]

    @staticmethod
    def join_prepare(r: np.ndarray, s: np.ndarray, nthreads: int, fanout: int):
        """Copies both relations ONCE into 64-byte aligned buffers with the slack the reference's
        main.c gives them (RELATION_PADDING, src/params.h:55-57, src/main.c:517-519).  The joins leave
        their inputs unmodified, so the pair can be joined any number of times."""
        pad = nthreads * fanout * 8 + 64
        rr, ss = aligned_u64(r.size + pad), aligned_u64(s.size + pad)
        rr[:r.size], ss[:s.size] = r, s
        return rr, ss, r.size, s.size

    def join_timed(self, prepared, which: str = "sortmergejoin_multiway", nthreads: int = 2,
                   fanout: int = 128, mwaybuf: int = 1 << 22):
        """One join of a prepared pair.  Returns (totalresults, wall seconds around the bare call,
        seconds of the reference's OWN timed region or None): the reference starts its clock inside
        the threads behind the first barrier and prints TOTAL-TIME-USECS on stderr
        (src/joins/sortmergejoin_multiway.c:145-153, 310-314; src/joins/joincommon.c:195, 214-227);
        allocation of its temporaries is outside that region."""
        import re
        import tempfile
        import time
        rr, ss, nr, ns = prepared
        relr, rels = _Rel(rr.ctypes.data, nr), _Rel(ss.ctypes.data, ns)
        cfg = _Cfg(nthreads, fanout, 0, 0, mwaybuf, 2)
        self.f["numa_shuffle_init"](2, nthreads)
        sys.stdout.flush()
        sys.stderr.flush()
        saved = (os.dup(1), os.dup(2))
        devnull = os.open(os.devnull, os.O_WRONLY)
        with tempfile.TemporaryFile() as errf:
            os.dup2(devnull, 1)
            os.dup2(errf.fileno(), 2)
            try:
                t0 = time.perf_counter()
                res = self.f[which](C.byref(relr), C.byref(rels), C.byref(cfg))
                wall = time.perf_counter() - t0
            finally:
                C.CDLL(None).fflush(None)
                os.dup2(saved[0], 1)
                os.dup2(saved[1], 2)
                os.close(saved[0])
                os.close(saved[1])
                os.close(devnull)
            errf.seek(0)
            txt = errf.read().decode(errors="replace")
        m = re.search(r"TOTAL-TIME-USECS = ([0-9.]+)", txt)
        region = float(m.group(1)) * 1e-6 if m else None
        if not res:
            return -1, wall, region
        return res.contents.totalresults, wall, region


class _ThreadRes(C.Structure):
    _fields_ = [("nresults", C.c_int64), ("results", _vp), ("threadid", C.c_uint32)]


REF_MAT_SO = os.path.join(ORACLE_DIR, "_ref", "libsmj_ref_mat.so")


class ReferenceMat:
    """The UNMODIFIED reference compiled with -DJOIN_MATERIALIZE (oracle/Makefile `ref`), with
    oracle/shim/tuple_buffer.h standing in for the header its tree lacks: the joins then emit
    <S.key, S.payload> per matching pair into per-thread chained buffers
    (src/joins/joincommon.c:256-278, sortmergejoin_multiway.c:559-599)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_MAT_SO)

    def __init__(self):
        L = self.lib = C.CDLL(REF_MAT_SO)
        self.f = {k: getattr(L, Reference.SYMS[k]) for k in
                  ("sortmergejoin_multiway", "sortmergejoin_multipass", "numa_shuffle_init",
                   "cpu_mapping_init", "merge_join")}
        for k in ("sortmergejoin_multiway", "sortmergejoin_multipass"):
            self.f[k].argtypes = [C.POINTER(_Rel), C.POINTER(_Rel), C.POINTER(_Cfg)]
            self.f[k].restype = _vp
        self.f["merge_join"].argtypes = [_vp, _vp, C.c_uint64, C.c_uint64, _vp]
        self.f["merge_join"].restype = C.c_uint64
        self.f["numa_shuffle_init"].argtypes = [C.c_int, C.c_int]
        L.orc_shim_cb_count.argtypes = [_vp]
        L.orc_shim_cb_count.restype = C.c_uint64
        L.orc_shim_cb_copy.argtypes = [_vp, _vp, C.c_uint64]
        L.orc_shim_cb_copy.restype = C.c_uint64
        L.orc_shim_cb_free.argtypes = [_vp]
        with silence_stdout():
            self.f["cpu_mapping_init"]()

    def join_materialize(self, r: np.ndarray, s: np.ndarray, which: str = "sortmergejoin_multiway",
                         nthreads: int = 2, fanout: int = 128, mwaybuf: int = 1 << 22):
        """returns (totalresults, [emitted items of thread 0, thread 1, ...]) in emission order"""
        rr, ss, nr, ns = Reference.join_prepare(r, s, nthreads, fanout)
        relr, rels = _Rel(rr.ctypes.data, nr), _Rel(ss.ctypes.data, ns)
        cfg = _Cfg(nthreads, fanout, 0, 0, mwaybuf, 2)
        with silence_all():
            self.f["numa_shuffle_init"](2, nthreads)
            res = self.f[which](C.byref(relr), C.byref(rels), C.byref(cfg))
        assert res
        resv = _Res.from_address(res)
        lst = (_ThreadRes * nthreads).from_address(resv.resultlist)
        outs = []
        for t in range(nthreads):
            k = self.lib.orc_shim_cb_count(lst[t].results)
            assert k == lst[t].nresults
            out = np.empty(k, dtype=np.uint64)
            assert self.lib.orc_shim_cb_copy(lst[t].results, out.ctypes.data, k) == k
            self.lib.orc_shim_cb_free(lst[t].results)
            outs.append(out)
        libc = C.CDLL(None)
        libc.free.argtypes = [_vp]
        libc.free(resv.resultlist)
        total = resv.totalresults
        libc.free(res)
        return total, outs


REF_K8_SO = os.path.join(ORACLE_DIR, "_ref", "libsmj_ref_key8b.so")


class ReferenceKey8B:
    """The UNMODIFIED reference compiled with -DKEY_8B (oracle/Makefile `ref`): 16-byte tuples
    {int64 payload; int64 key}, scalar sort and merge only (src/main.c:871-877)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_K8_SO)

    def __init__(self):
        L = self.lib = C.CDLL(REF_K8_SO)
        self.f = {k: getattr(L, Reference.SYMS[k]) for k in
                  ("sortmergejoin_multiway", "sortmergejoin_multipass", "numa_shuffle_init", "cpu_mapping_init",
                   "scalarsort_tuples", "merge_join")}
        for k in ("sortmergejoin_multiway", "sortmergejoin_multipass"):
            self.f[k].argtypes = [C.POINTER(_Rel), C.POINTER(_Rel), C.POINTER(_Cfg)]
            self.f[k].restype = C.POINTER(_Res)
        self.f["scalarsort_tuples"].argtypes = [C.POINTER(_vp), C.POINTER(_vp), C.c_uint64]
        self.f["scalarsort_tuples"].restype = None
        self.f["merge_join"].argtypes = [_vp, _vp, C.c_uint64, C.c_uint64, _vp]
        self.f["merge_join"].restype = C.c_uint64
        self.f["numa_shuffle_init"].argtypes = [C.c_int, C.c_int]
        with silence_stdout():
            self.f["cpu_mapping_init"]()

    @staticmethod
    def _aligned(t: np.ndarray, extra_tuples: int) -> np.ndarray:
        n = t.shape[0]
        buf = aligned_u64(2 * (n + extra_tuples) + 16)
        buf[:2 * n] = np.ascontiguousarray(t, dtype=np.int64).reshape(-1).view(np.uint64)
        return buf

    def join(self, r: np.ndarray, s: np.ndarray, which: str = "sortmergejoin_multiway", nthreads: int = 2,
             fanout: int = 128) -> int:
        pad = nthreads * fanout * 8 + 64
        rr, ss = self._aligned(r, pad), self._aligned(s, pad)
        relr, rels = _Rel(rr.ctypes.data, r.shape[0]), _Rel(ss.ctypes.data, s.shape[0])
        cfg = _Cfg(nthreads, fanout, 1, 1, 1 << 22, 2)  # SCALARSORT, SCALARMERGE
        with silence_all():
            self.f["numa_shuffle_init"](2, nthreads)
            res = self.f[which](C.byref(relr), C.byref(rels), C.byref(cfg))
        return res.contents.totalresults if res else -1

    def sort(self, t: np.ndarray) -> np.ndarray:
        n = t.shape[0]
        a, b = self._aligned(t, 8), aligned_u64(2 * n + 32)
        pa, pb = _vp(a.ctypes.data), _vp(b.ctypes.data)
        self.f["scalarsort_tuples"](C.byref(pa), C.byref(pb), n)
        res = a if pb.value == a.ctypes.data else b
        return res[:2 * n].view(np.int64).reshape(n, 2).copy()

    def merge_join(self, r: np.ndarray, s: np.ndarray) -> int:
        rr, ss = self._aligned(r, 8), self._aligned(s, 8)
        return self.f["merge_join"](rr.ctypes.data, ss.ctypes.data, r.shape[0], s.shape[0], None)


def rand_tuples16(rng: np.random.Generator, n: int, key_lo: int, key_hi: int) -> np.ndarray:
    """(n, 2) int64 {payload, key}: keys uniform in [key_lo, key_hi], payloads any int64"""
    t = np.empty((n, 2), dtype=np.int64)
    t[:, 0] = rng.integers(-2**62, 2**62, size=n, dtype=np.int64)
    t[:, 1] = rng.integers(key_lo, key_hi + 1, size=n, dtype=np.int64)
    return t


_ORACLE = None
_REF = None
_REF_MAT = None
_REF_K8 = None


def reference_key8b() -> "ReferenceKey8B | None":
    global _REF_K8
    if _REF_K8 is None and ReferenceKey8B.available():
        _REF_K8 = ReferenceKey8B()
    return _REF_K8


def reference_mat() -> "ReferenceMat | None":
    global _REF_MAT
    if _REF_MAT is None and ReferenceMat.available():
        _REF_MAT = ReferenceMat()
    return _REF_MAT


def oracle() -> Oracle:
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE


def reference() -> Reference | None:
    global _REF
    if _REF is None and Reference.available():
        _REF = Reference()
    return _REF


# --------------------------------------------------------------------------------- inputs
def keys_ok(items: np.ndarray) -> np.ndarray:
    """Clear the NaN exponent pattern like src/datagen/generator.c:107-121 avoid_NaN does."""
    items = items.copy()
    nan = (items & np.uint64(0x7FF << 52)) == np.uint64(0x7FF << 52)
    items[nan] &= ~np.uint64(1 << 52)
    return items


def rand_items(rng: np.random.Generator, n: int, key_max: int, payload_bits: int = 32) -> np.ndarray:
    keys = rng.integers(1, key_max + 1, size=n, dtype=np.uint64)
    pay = rng.integers(0, 1 << payload_bits, size=n, dtype=np.uint64)
    return keys_ok((keys << np.uint64(32)) | pay)


def zipf_keys(rng: np.random.Generator, n: int, alphabet: int, theta: float) -> np.ndarray:
    """src/datagen/genzipf.c:29-159 restated with numpy: permuted alphabet 1..alphabet,
    cumulative LUT of 1/i^theta, binary search of a uniform variate."""
    alpha = rng.permutation(alphabet).astype(np.uint64) + np.uint64(1)
    lut = np.cumsum(1.0 / np.power(np.arange(1, alphabet + 1, dtype=np.float64), theta))
    lut /= lut[-1]
    pos = np.searchsorted(lut, rng.random(n), side="left")
    pos = np.minimum(pos, alphabet - 1)
    return alpha[pos]


def make_unique(rng: np.random.Generator, n: int) -> np.ndarray:
    """keys = random permutation of 1..n, payload = row index"""
    keys = rng.permutation(n).astype(np.uint64) + np.uint64(1)
    return (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
