#!/usr/bin/env python
"""bench.py -- the BASELINE.json workloads on B200, and the reference's CPU path beside them.

    python bench.py [--gpus N] [--steps K] [--warmup W]      # headline: config 1 (+ configs 2-4 summarised)
    python bench.py --config {1..5} ...                      # one configuration as the headline line
    python bench.py --impl reference [--config C] ...        # the reference's own CPU implementation
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

config 1  sortmergejoins -a m-way, |R| = |S| = 128 M 8-byte tuples, uniform unique keys      (join)
config 2  avxsort_tuples of 10^9 packed key|payload items (bench_sort)                        (sort)
config 3  partition_relation_optimized fan-out sweep 2^4 .. 2^14 on 10^9 tuples               (partition)
config 4  m-way join, |R| = 128 M, |S| = 1.28 G foreign keys, Zipf theta = 0.5 .. 1.0         (join, skew)
config 5  m-way join 2 G x 2 G over 8 GPUs                                                   (needs --gpus 8)

A step = one complete pass of the configuration's operator over its input.  value = whole-job
throughput with the inputs resident in HBM (CUDA events on the library's stream); e2e = the same
through the reference-facing C entry point with HOST buffers (H2D copies inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "M input tuples/s"
SEED_R, SEED_S = 12345, 54321  # reference src/main.c:474-475
N1 = 128_000_000

# {count, checksum} of the brute-force CPU oracle (oracle/smj_oracle.c: orc_join_bruteforce on
# orc_gen_unique(n, 12345) x orc_gen_unique(n, 54321)) at the FULL size of config 1, computed once
# in the dev container (83 s).  The GPU result must equal it bit for bit.
ORACLE_RESULT = {N1: (128_000_000, 13717379736529395287)}

METRICS = {
    1: ("join throughput, m-way sort-merge join 128M x 128M 8-byte tuples", UNIT),
    2: ("sort throughput, avxsort_tuples of 10^9 packed 64-bit key|payload items", "M items/s"),
    3: ("partition throughput, partition_relation_optimized fan-out sweep 2^4..2^14 on 10^9 tuples", "M tuples/s"),
    4: ("join throughput, m-way join 128M x 1.28G foreign keys, Zipf theta 0.5..1.0", UNIT),
    5: ("join throughput, m-way sort-merge join 2G x 2G 8-byte tuples over 8 GPUs", UNIT),
}
THETAS = (0.5, 0.6, 0.7, 0.8, 0.9, 1.0)


def workload(config: int, n: int) -> str:
    """the SAME string in both arms (`--impl ours` / `--impl reference`)"""
    if config == 1:
        return f"sortmergejoins -a m-way, |R|=|S|={n} 8-byte tuples, uniform unique keys (BASELINE.json configs[0])"
    if config == 2:
        return f"avxsort standalone sort of {n} packed 64-bit key|payload items (BASELINE.json configs[1])"
    if config == 3:
        return f"radix partitioning fanout sweep 2^4..2^14 on {n} tuples (BASELINE.json configs[2])"
    if config == 4:
        return (f"m-way join with Zipf-skewed S (theta=0.5..1.0, genzipf), |R|={n} |S|={10 * n} foreign-key "
                f"(BASELINE.json configs[3])")
    return f"scale-out m-way join |R|=|S|={n} over 8 GPUs (BASELINE.json configs[4])"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc, self.path = None, None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                for name, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------- reference arm
def pow2_le(x: int) -> int:
    p = 1
    while p * 2 <= x:
        p *= 2
    return p


class CpuReferenceJoin:
    """The reference's own sortmergejoin_multiway (oracle/_ref/libsmj_ref.so, the unmodified
    reference compiled by oracle/Makefile) on the host cores.  The relations are copied ONCE into
    aligned, padded buffers (what src/main.c:516-528 allocates); a run then times
      region: the reference's own timed region, TOTAL-TIME-USECS on its stderr (the clock starts
              inside the threads behind the first barrier and stops behind the last one,
              src/joins/sortmergejoin_multiway.c:145-153, 310-314; allocation of its temporaries
              is outside), and
      wall:   perf_counter around the bare sortmergejoin_multiway() call (allocation + threads)."""

    def __init__(self, r: np.ndarray, s: np.ndarray):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        self.ol = ol
        self.ref = ol.reference()
        self.r, self.s = r, s
        cpus = sorted(os.sched_getaffinity(0))
        self.cpus = cpus
        self.nthreads = min(pow2_le(len(cpus)), 64)
        self.fanout = max(128, self.nthreads)
        self.kind = "reference" if self.ref is not None else "port"
        self.prepared = self.ref.join_prepare(r, s, self.nthreads, self.fanout) if self.ref else None
        self.tmp = tempfile.TemporaryDirectory()
        # pin the reference's threads onto the cpus we may use (src/util/cpu_mapping.c:46-80)
        ncpu = os.sysconf("SC_NPROCESSORS_ONLN")
        if cpus != list(range(len(cpus))) and len(cpus) <= ncpu <= 1024:
            rest = [c for c in range(ncpu) if c not in cpus]
            with open(os.path.join(self.tmp.name, "cpu-mapping.txt"), "w") as f:
                f.write(" ".join(str(x) for x in [ncpu] + cpus + rest) + " 1\n")

    def run(self):
        """-> (matches, region seconds, wall seconds)"""
        if self.ref is None:  # the C port of the oracle: single thread, sort + scan
            t0 = time.perf_counter()
            cnt, _ = self.ol.oracle().join_bruteforce(self.r, self.s)
            dt = time.perf_counter() - t0
            return cnt, dt, dt
        cwd = os.getcwd()
        os.chdir(self.tmp.name)
        try:
            cnt, wall, region = self.ref.join_timed(self.prepared, "sortmergejoin_multiway", self.nthreads,
                                                    self.fanout, 32 * 1024 * 1024)
        finally:
            os.chdir(cwd)
        return cnt, (region if region else wall), wall

    @property
    def cores(self):
        return self.nthreads if self.ref is not None else 1


def host_relations(config: int, n: int):
    """config 1 / 4 inputs on the host from the oracle's generators (== the device generators, bit for bit)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    orc = ol.oracle()
    r = orc.gen_unique(n, SEED_R)
    if config == 4:
        return r, orc.gen_zipf(10 * n, n, 0.75, SEED_S, SEED_R)
    return r, orc.gen_unique(n, SEED_S)


def run_reference_arm(args) -> None:
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = args.config or 1
    metric, unit = METRICS[cfg]
    base = {"impl": "reference", "metric": metric, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "gpu_launches": 0}
    if cfg == 5:
        print(json.dumps({"impl": "reference", "unavailable": "config 5 (2G x 2G) needs ~100 GB of host memory and "
                          "int32 per-thread counts overflow in the reference (SURVEY.md 8d)"}))
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    if cfg in (1, 4):
        # config 1: the full workload.  config 4: a bounded sample (|R| = 12.8 M, |S| = 128 M, theta 0.75)
        n = args.size or (N1 if cfg == 1 else 12_800_000)
        r, s = host_relations(cfg, n)
        job = CpuReferenceJoin(r, s)
        reg, wal, cnt = [], [], -1
        for i in range(args.warmup + args.steps):
            cnt, region, wall = job.run()
            if i >= args.warmup:
                reg.append(region)
                wal.append(wall)
        assert cnt == s.size, f"reference join returned {cnt}, expected {s.size}"
        tuples = r.size + s.size
        sec, wsec = sum(reg) / len(reg), sum(wal) / len(wal)
        val, e2e = tuples / sec / 1e6, tuples / wsec / 1e6
        sample = "full workload, every step" if cfg == 1 else f"bounded sample |R|={n} |S|={10 * n} theta=0.75"
        line = dict(base, value=val, ms_per_step=sec * 1e3,
                    config={"workload": workload(cfg, args.size or (N1 if cfg == 1 else N1)), "threads": job.cores,
                            "sample": sample,
                            "timing": "value: the reference's own timed region (TOTAL-TIME-USECS, "
                                      "sortmergejoin_multiway.c:147,313); e2e: wall clock around the bare "
                                      "sortmergejoin_multiway() call; relations copied into aligned buffers once, "
                                      "outside both"},
                    cpu_baseline={"value": val, "unit": unit, "cores": job.cores, "kind": job.kind, "sample": sample},
                    e2e={"value": e2e, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                         "ms_per_step": wsec * 1e3})
        print(json.dumps(line))
        return
    # configs 2 / 3: the single-threaded operators (sortbench.c / partitioningbench.c call them on one
    # thread) on a bounded sample of 2^26 items
    ref = ol.reference()
    n = args.size or (1 << 26)
    items = ol.oracle().gen_unique(n, SEED_R)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if cfg == 2:
            if ref is not None:
                ref.sort(items)
            else:
                ol.oracle().sort(items)
        else:
            for nbits in range(4, 15):
                if ref is not None:
                    ref.partition(items, nbits, 0, "partition_relation_optimized")
                else:
                    ol.oracle().partition(items, nbits, 0, True)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    units = n if cfg == 2 else 11 * n
    val = units / sec / 1e6
    sample = f"bounded sample of {n} items, one thread (wall clock incl. the harness's buffer copies)"
    print(json.dumps(dict(base, value=val, ms_per_step=sec * 1e3,
                          config={"workload": workload(cfg, 1_000_000_000), "threads": 1, "sample": sample},
                          cpu_baseline={"value": val, "unit": unit, "cores": 1,
                                        "kind": "reference" if ref is not None else "port", "sample": sample},
                          e2e={"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})))


# ------------------------------------------------------------------------------------- helpers
def parse_profile(txt: str) -> dict:
    prof = {}
    for line in txt.strip().splitlines():
        name, cnt, ms, items = line.rsplit(";", 3)
        name = name.strip("() ")
        prof[name] = {"launches": int(cnt), "ms": float(ms), "items": int(items)}
    return prof


BYTES_PER_ITEM = {"k_onesweep": 16, "k_msd_scatter2": 16, "k_msd_hist2": 8, "k_msd_hist1": 8, "k_bucket_join": 8,
                  "k_mj_tile": 8, "k_hist_digits": 8, "k_bits_reduce": 8, "k_bucket_sort": 16, "k_hist_wide": 8,
                  "k_kway": 16, "k_mp_merge": 16, "k_small_sort": 16, "k_fused": 24}


class Timer:
    def __init__(self, smj):
        self.smj, self.lib = smj, smj.lib
        self.e0, self.e1 = C.c_void_p(), C.c_void_p()
        self.lib.smjdev_event_create(C.byref(self.e0))
        self.lib.smjdev_event_create(C.byref(self.e1))

    def time(self, fn, steps: int) -> float:
        """ms per step: CUDA events on the library's stream around `steps` calls"""
        self.lib.smjdev_sync()
        self.lib.smjdev_event_record(self.e0)
        for _ in range(steps):
            fn()
        self.lib.smjdev_event_record(self.e1)
        ms = C.c_float()
        self.smj.check(self.lib.smjdev_event_elapsed_ms(self.e0, self.e1, C.byref(ms)), "elapsed")
        return ms.value / steps


def profile_steps(smj, fn, steps: int) -> dict:
    lib = smj.lib
    lib.smjdev_profile_begin()
    for _ in range(steps):
        fn()
    buf = C.create_string_buffer(1 << 16)
    lib.smjdev_profile_end(buf, len(buf))
    return parse_profile(buf.value.decode())


def roofline_of(prof: dict, steps: int, whole_bytes: float, ms_step: float) -> dict:
    """dominant kernel: algorithmic bytes per launch / its event-timed duration, against the measured peak"""
    peak, peak_src = measured_peak()
    total_kernel_ms = sum(p["ms"] for p in prof.values()) or 1.0
    top = max(prof, key=lambda k: prof[k]["ms"])
    bpi = next((v for k, v in BYTES_PER_ITEM.items() if k in top), 16)
    achieved = bpi * prof[top]["items"] / (prof[top]["ms"] * 1e-3) / 1e9 if prof[top]["ms"] else 0.0
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        for kname, rec in sorted(tj.get("kernels", {}).items(), key=lambda kv: len(kv[0])):
            if kname in top and rec.get("items"):  # the longest matching name wins
                traffic = rec["dram_bytes_per_launch"] * (prof[top]["items"] / prof[top]["launches"]) / rec["items"]
                traffic_src = ("stored ncu capture (profiles/traffic.json: " + tj.get("_source", "round 1") +
                               "), scaled to this launch's items; NOT measured in this run")
    except Exception:
        pass
    return {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_item": bpi, "launches": prof[top]["launches"],
            "avg_launch_ms": prof[top]["ms"] / max(prof[top]["launches"], 1),
            "share_of_step": prof[top]["ms"] / total_kernel_ms,
            "whole_op_frac": (whole_bytes / (ms_step * 1e-3) / 1e9) / peak,
            "kernels": {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["launches"] / steps}
                        for k, v in prof.items()}}


# ------------------------------------------------------------------------------------- config 1
def config1(args, smj, full: bool) -> dict:
    lib = smj.lib
    n = args.size or N1
    expect = ORACLE_RESULT.get(n)
    dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
    smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, SEED_R), "gen R")
    smj.check(lib.smjdev_gen_unique(ds.ptr, n, 0, n, SEED_S), "gen S")
    smj.check(lib.smjdev_sync())

    def check(st):
        assert st.matches == n, f"join returned {st.matches}, expected {n}"
        if expect:
            assert (st.matches, st.checksum) == expect, f"join result {(st.matches, st.checksum)} != oracle {expect}"

    st = None
    for _ in range(args.warmup):
        st = smj.device_join(smj.ALGO_MWAY, dr, ds)
        check(st)
    tm = Timer(smj)
    sampler = ClockSampler(lib.smjdev_current_device())
    time.sleep(0.3)
    launches0 = lib.smjdev_launch_count()
    phases = {"partition": 0.0, "sort": 0.0, "merge": 0.0, "join": 0.0}

    def step():
        nonlocal st
        st = smj.device_join(smj.ALGO_MWAY, dr, ds)
        check(st)
        for k in phases:
            phases[k] += getattr(st, "ms_" + k)

    ms_step = tm.time(step, args.steps)
    launches = lib.smjdev_launch_count() - launches0
    value = 2 * n / (ms_step * 1e-3) / 1e6
    prof = profile_steps(smj, lambda: smj.device_join(smj.ALGO_MWAY, dr, ds), args.steps)
    roofline = roofline_of(prof, args.steps, 40.0 * 2 * n, ms_step)
    roofline["whole_join_frac"] = roofline["whole_op_frac"]
    path_names = {0: "general (LSD sort + merge-path scan)", 1: "msd (exact histograms)", 2: "msd_opt"}

    # ---- the literal sort-merge pipeline of north_star (LSD radix sort of R and S + merge-path scan)
    lsd = None
    try:
        os.environ["SMJ_MWAY_PATH"] = "lsd"
        for _ in range(2):
            st2 = smj.device_join(smj.ALGO_MWAY, dr, ds)
            check(st2)
        lsd_ms = tm.time(lambda: check(smj.device_join(smj.ALGO_MWAY, dr, ds)), max(2, min(args.steps, 5)))
        lprof = profile_steps(smj, lambda: smj.device_join(smj.ALGO_MWAY, dr, ds), 2)
        lsd = {"workload": "config 1 with SMJ_MWAY_PATH=lsd: key-only LSD radix sort of R and S + merge-path scan",
               "value": 2 * n / (lsd_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": lsd_ms,
               "matches": int(st2.matches), "checksum": int(st2.checksum),
               "roofline": roofline_of(lprof, 2, 40.0 * 2 * n, lsd_ms)}
    except Exception as exc:
        lsd = {"error": str(exc)}
    finally:
        os.environ.pop("SMJ_MWAY_PATH", None)

    # ---- end to end through sortmergejoin_multiway() with pinned host relations
    hr, hs = lib.smjdev_host_alloc(n * 8), lib.smjdev_host_alloc(n * 8)
    assert hr and hs, "pinned host allocation failed"
    smj.check(lib.smjdev_memcpy_d2h(hr, dr.ptr, n * 8))
    smj.check(lib.smjdev_memcpy_d2h(hs, ds.ptr, n * 8))
    smj.check(lib.smjdev_sync())
    cfg = smj.joinconfig_t(1, 128, 0, 0, 20 * 1024 * 1024, smj.NUMA_NEXT)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    h2d_ms = []

    def host_join(relr, rels):
        res = lib.sortmergejoin_multiway(C.byref(relr), C.byref(rels), C.byref(cfg))
        assert res, "sortmergejoin_multiway failed: " + lib.smjdev_last_error().decode()
        total = res.contents.totalresults
        hst = smj.smj_join_stats_t()
        lib.smj_last_join_stats(C.byref(hst))
        h2d_ms.append(hst.ms_h2d)
        libc.free(C.cast(res.contents.resultlist, C.c_void_p))
        libc.free(C.cast(res, C.c_void_p))
        assert total == n
        if expect:
            assert hst.checksum == expect[1]

    relr, rels = smj.relation_t(hr, n), smj.relation_t(hs, n)
    for _ in range(max(1, min(args.warmup, 2))):
        host_join(relr, rels)
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_join(relr, rels)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    pinned = {"value": 2 * n / e2e_s / 1e6, "unit": UNIT, "ms_per_step": e2e_s * 1e3,
              "h2d_ms_per_step": sum(h2d_ms[-e2e_steps:]) / e2e_steps}
    # the same call with ordinary malloc'ed (pageable) relations, which is what the reference's main.c
    # passes (src/main.c:516-528): this is the DROP-IN number and the headline e2e
    pageable = None
    try:
        r_pg = np.ctypeslib.as_array(C.cast(hr, C.POINTER(C.c_uint64)), shape=(n,)).copy()
        s_pg = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_uint64)), shape=(n,)).copy()
        relr_pg, rels_pg = smj.relation_t(r_pg.ctypes.data, n), smj.relation_t(s_pg.ctypes.data, n)
        host_join(relr_pg, rels_pg)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_join(relr_pg, rels_pg)
        pg_s = (time.perf_counter() - t0) / e2e_steps
        pageable = {"value": 2 * n / pg_s / 1e6, "unit": UNIT, "ms_per_step": pg_s * 1e3}
        del r_pg, s_pg
    except Exception as exc:
        pageable = {"value": None, "error": str(exc)}
    head = pageable if pageable.get("value") else pinned
    e2e = {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": 16,
           "ms_per_step": head["ms_per_step"], "steps": e2e_steps,
           "api": "sortmergejoin_multiway(relation_t*, relation_t*, joinconfig_t*) on ordinary malloc'ed "
                  "(pageable) host relations, as the reference's main.c passes them" if head is pageable else
                  "sortmergejoin_multiway(relation_t*, relation_t*, joinconfig_t*) on pinned host relations",
           "pageable_host_memory": pageable, "pinned_host_memory": pinned}
    clocks = sampler.stop()

    cpu = None
    if full and not args.no_cpu_baseline:
        r_h = np.ctypeslib.as_array(C.cast(hr, C.POINTER(C.c_uint64)), shape=(n,))
        s_h = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_uint64)), shape=(n,))
        try:
            job = CpuReferenceJoin(r_h, s_h)
            best_r, best_w = None, None
            for _ in range(2):
                cnt, region, wall = job.run()
                assert cnt == n
                best_r = region if best_r is None else min(best_r, region)
                best_w = wall if best_w is None else min(best_w, wall)
            cpu = {"value": 2 * n / best_r / 1e6, "unit": UNIT, "cores": job.cores, "kind": job.kind,
                   "sample": f"full workload ({n} x {n}), best of 2 runs: {best_r:.3f} s in the reference's own timed "
                             f"region (TOTAL-TIME-USECS), {best_w:.3f} s wall around the bare call",
                   "wall_value": 2 * n / best_w / 1e6}
            del job
        except Exception as exc:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "reference", "sample": f"failed: {exc}"}
    lib.smjdev_host_free(hr)
    lib.smjdev_host_free(hs)
    dr.free()
    ds.free()
    return {
        "metric": METRICS[1][0], "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload(1, n),
                   "keys": "pseudo-random permutations of 1..n, seeds 12345/54321, payload = row index",
                   "l2": "inputs (2 x %.2f GB) are larger than L2, no flush needed" % (n * 8 / 1e9),
                   "timing": "CUDA events on the library's stream around all steps",
                   "path": path_names.get(int(st.path), str(st.path)),
                   "result_check": ("count and checksum equal the CPU oracle's at full size " + str(expect)) if expect
                                   else "count == n",
                   "phases_ms": {k: v / args.steps for k, v in phases.items()}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "cpu_baseline": cpu, "matches": int(st.matches), "checksum": int(st.checksum),
        "configs": {"1-lsd": lsd},
    }


# ------------------------------------------------------------------------------------- config 2
def config2(args, smj, steps: int) -> dict:
    """avxsort_tuples(&in, &out, n) with DEVICE buffers (the entry point accepts them): keys =
    permutation of 1..n (create_relation_pk, sortbench.c:103-105), payload = row index."""
    lib = smj.lib
    n = args.size or 1_000_000_000
    src, a, b = smj.DeviceBuffer(n), smj.DeviceBuffer(n), smj.DeviceBuffer(n)
    smj.check(lib.smjdev_gen_unique(src.ptr, n, 0, n, SEED_R), "gen")
    before = (C.c_uint64 * 3)()
    smj.check(lib.smjdev_check_order(src.ptr, n, 0, before), "check")
    tm = Timer(smj)
    out_ptr = [0]

    def sort_once():
        smj.check(lib.smjdev_memcpy_d2d(a.ptr, src.ptr, n * 8))
        pa, pb = C.c_void_p(a.ptr), C.c_void_p(b.ptr)
        lib.avxsort_tuples(C.byref(pa), C.byref(pb), n)
        out_ptr[0] = pb.value

    for _ in range(2):
        sort_once()
    copy_ms = tm.time(lambda: smj.check(lib.smjdev_memcpy_d2d(a.ptr, src.ptr, n * 8)), steps)
    ms = tm.time(sort_once, steps) - copy_ms
    after = (C.c_uint64 * 3)()
    smj.check(lib.smjdev_check_order(out_ptr[0], n, 0, after), "check")
    assert after[0] == 0 and (after[1], after[2]) == (before[1], before[2]), "sort output is not a sorted permutation"
    prof = profile_steps(smj, sort_once, 1)
    prof = {k: v for k, v in prof.items()}
    line = {"workload": workload(2, n), "value": n / (ms * 1e-3) / 1e6, "unit": "M items/s", "ms_per_step": ms,
            "steps": steps, "check": "sorted, multiset fingerprint (sum, xor) unchanged",
            "roofline": roofline_of(prof, 1, 16.0 * n, ms)}
    for x in (src, a, b):
        x.free()
    return line


# ------------------------------------------------------------------------------------- config 3
def config3(args, smj, steps: int) -> dict:
    """partition_relation_optimized semantics (padded partitions, hash on key-1, shift 0,
    partitioningbench.c:96,113-116,152) for 4..14 radix bits on n tuples, device resident."""
    lib = smj.lib
    n = args.size or 1_000_000_000
    src, tmp, out = smj.DeviceBuffer(n), smj.DeviceBuffer(n), smj.DeviceBuffer(n + 8 * (1 << 14) + 8)
    hist, offs = smj.DeviceBuffer(1 << 14), smj.DeviceBuffer(1 << 14)
    smj.check(lib.smjdev_gen_unique(src.ptr, n, 0, n, SEED_R), "gen")
    tm = Timer(smj)
    sweep, total_ms = {}, 0.0
    top_prof = None
    for nbits in range(4, 15):
        run = lambda: smj.check(lib.smjdev_partition(src.ptr, out.ptr, tmp.ptr, n, nbits, 0, 1, 1, hist.ptr, offs.ptr))
        run()
        ms = tm.time(run, steps)
        h, o = hist.to_host(1 << nbits), offs.to_host(1 << nbits)
        used = int(o[-1] + h[-1])
        stats = (C.c_uint64 * 4)()
        smj.check(lib.smjdev_check_partition(out.ptr, offs.ptr, hist.ptr, nbits, 0, 1, used, stats), "check")
        assert stats[0] == 0 and stats[1] == 0 and stats[2] == n and int(h.sum()) == n and not np.any(o % 8), \
            f"partition check failed at {nbits} bits: {list(stats)}"
        sweep[str(nbits)] = {"ms": ms, "M_tuples_per_s": n / ms / 1e3, "hbm_frac_at_16B": 16 * n / ms / 1e6 / measured_peak()[0]}
        total_ms += ms
        if nbits == 7:
            top_prof = profile_steps(smj, run, 1)
    line = {"workload": workload(3, n), "value": 11 * n / (total_ms * 1e-3) / 1e6, "unit": "M tuples/s",
            "ms_per_step": total_ms, "steps": steps, "sweep": sweep,
            "check": "every partition: membership, 8-tuple aligned start, stable order, all tuples present",
            "roofline": roofline_of(top_prof, 1, 16.0 * n, sweep["7"]["ms"])}
    line["roofline"]["note"] = "dominant kernel and whole_op_frac are those of the 7-bit (fan-out 128) pass"
    for x in (src, tmp, out, hist, offs):
        x.free()
    return line


# ------------------------------------------------------------------------------------- config 4
def config4(args, smj, steps: int) -> dict:
    """R = permutation of 1..n, S = 10 n foreign keys drawn Zipf(theta) from R's keys with the
    reference's generator restated on the device (smjdev_gen_zipf, genzipf.c:29-159)."""
    lib = smj.lib
    n = args.size or N1
    ns = 10 * n
    dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(ns)
    lut, tmp = smj.DeviceBuffer(n), smj.DeviceBuffer(n // 256 + n // 65536 + 8)
    smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, SEED_R), "gen R")
    tm = Timer(smj)
    per, total_ms = {}, 0.0
    prof = None
    for theta in THETAS:
        smj.check(lib.smjdev_zipf_lut(lut.ptr, n, theta, tmp.ptr), "zipf lut")
        smj.check(lib.smjdev_gen_zipf(ds.ptr, 0, ns, n, theta, SEED_S, SEED_R + 1, lut.ptr), "gen zipf")
        st = smj.device_join(smj.ALGO_MWAY, dr, ds)
        assert st.matches == ns, f"theta {theta}: {st.matches} matches, expected |S| = {ns}"
        ms = tm.time(lambda: smj.device_join(smj.ALGO_MWAY, dr, ds), steps)
        per[f"{theta:.1f}"] = {"ms": ms, "M_tuples_per_s": (n + ns) / ms / 1e3, "matches": int(st.matches),
                               "checksum": int(st.checksum), "path": int(st.path), "repeats": int(st.repeats)}
        total_ms += ms
        if theta == THETAS[-1]:
            prof = profile_steps(smj, lambda: smj.device_join(smj.ALGO_MWAY, dr, ds), 1)
    line = {"workload": workload(4, n), "value": len(THETAS) * (n + ns) / (total_ms * 1e-3) / 1e6, "unit": UNIT,
            "ms_per_step": total_ms, "steps": steps, "thetas": per, "check": "matches == |S| for every theta",
            "roofline": roofline_of(prof, 1, 40.0 * (n + ns), per[f"{THETAS[-1]:.1f}"]["ms"])}
    line["roofline"]["note"] = "dominant kernel and whole_op_frac are those of theta = 1.0"
    for x in (dr, ds, lut, tmp):
        x.free()
    return line


def run_single_gpu(args, smj) -> dict:
    lib = smj.lib
    smj.check(lib.smjdev_init(int(os.environ.get("LOCAL_RANK", "0"))), "init")
    cfg = args.config
    if cfg in (None, 1):
        line = config1(args, smj, full=True)
        if cfg is None and not args.no_extra and not args.size:
            small = max(2, min(args.steps, 3))
            for key, fn in (("2", config2), ("3", config3), ("4", config4)):
                try:
                    line["configs"][key] = fn(args, smj, small)
                except Exception as exc:  # an extra line must never take the headline down
                    line["configs"][key] = {"error": f"{type(exc).__name__}: {exc}"}
                    try:
                        lib.smjdev_sync()
                    except Exception:
                        pass
        return line
    if cfg == 5:
        raise SystemExit("config 5 needs --gpus 8 (python bench.py --config 5 --gpus 8)")
    sampler = ClockSampler(lib.smjdev_current_device())
    launches0 = lib.smjdev_launch_count()
    sub = {2: config2, 3: config3, 4: config4}[cfg](args, smj, max(args.steps, 2))
    clocks = sampler.stop()
    metric, unit = METRICS[cfg]
    roofline = sub.pop("roofline")
    wl = sub.pop("workload")
    return {"metric": metric, "value": sub.pop("value"), "unit": unit, "n_gpus": 1, "steps": sub.pop("steps"),
            "warmup": args.warmup, "ms_per_step": sub.pop("ms_per_step"), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": dict({"workload": wl, "timing": "CUDA events on the library's stream",
                            "l2": "inputs are larger than L2, no flush needed"}, **sub),
            "clocks": clocks, "e2e": None, "gpu_launches": int(lib.smjdev_launch_count() - launches0),
            "roofline": roofline, "cpu_baseline": None}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=None, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configuration (1-based); default: 1 with 2-4 summarised under `configs`")
    ap.add_argument("--size", type=int, default=0, help="override the configuration's base size (tuples)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="config 1 only (skip the summaries of configs 2-4)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    os.environ.setdefault("SMJ_QUIET", "1")

    if args.impl == "reference":
        run_reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    smj = importlib.import_module("avx-sort-merge-joins_b200")  # fails loudly if the .so is missing
    if world == 1 and args.gpus > 1:
        # not started by torchrun: start one rank per GPU ourselves
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if world > 1:
        import bench_multigpu
        line = bench_multigpu.run(args, smj)
    else:
        line = run_single_gpu(args, smj)
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
