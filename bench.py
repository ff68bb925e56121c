#!/usr/bin/env python
"""bench.py -- m-way sort-merge join throughput, |R| = |S| = 128M 8-byte tuples (BASELINE.json
configs[0], the reference's `sortmergejoins -a m-way` default workload).

    python bench.py --gpus 1 --steps K --warmup W            # this repo, CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU m-way join
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A step = one complete join (range partition/sort of R and S, merge-join scan, count + checksum on
the host).  value = (|R|+|S|) / t with inputs resident in HBM; e2e = the same through the
reference-facing C entry point sortmergejoin_multiway(relation_t*, relation_t*, joinconfig_t*)
with pinned HOST buffers (H2D copy of both relations inside the timed region).
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

METRIC = "join throughput, m-way sort-merge join 128M x 128M 8-byte tuples"
UNIT = "M input tuples/s"
SEED_R, SEED_S = 12345, 54321  # reference src/main.c:474-475


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


def cpu_reference_join(r: np.ndarray, s: np.ndarray, runs: int):
    """Times the reference's own sortmergejoin_multiway (oracle/_ref/libsmj_ref.so, the unmodified
    reference compiled by oracle/Makefile) on the host cores; falls back to the C port of the oracle
    when the reference library is not there.  Returns (best_seconds, matches, info)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol

    cpus = sorted(os.sched_getaffinity(0))
    ref = ol.reference()
    if ref is None:
        orc = ol.oracle()
        t0 = time.perf_counter()
        cnt, _ = orc.join_bruteforce(r, s)
        return time.perf_counter() - t0, cnt, {"kind": "port", "cores": 1}
    nthreads = min(pow2_le(len(cpus)), 64)
    fanout = max(128, nthreads)
    best, cnt = None, -1
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        # pin the reference's threads onto the cpus we are allowed to use (src/util/cpu_mapping.c:46-80)
        ncpu = os.sysconf("SC_NPROCESSORS_ONLN")
        if cpus != list(range(len(cpus))) and len(cpus) <= ncpu <= 1024:
            rest = [c for c in range(ncpu) if c not in cpus]
            with open(os.path.join(tmp, "cpu-mapping.txt"), "w") as f:
                f.write(" ".join(str(x) for x in [ncpu] + cpus + rest) + " 1\n")
        os.chdir(tmp)
        try:
            for _ in range(runs):
                t0 = time.perf_counter()
                cnt = ref.join(r, s, "sortmergejoin_multiway", nthreads, fanout, 32 * 1024 * 1024)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
        finally:
            os.chdir(cwd)
    return best, cnt, {"kind": "reference", "cores": nthreads}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol

    n = args.size
    orc = ol.oracle()
    r, s = orc.gen_unique(n, SEED_R), orc.gen_unique(n, SEED_S)
    times, cnt, info = [], -1, {}
    for i in range(args.warmup + args.steps):
        t, cnt, info = cpu_reference_join(r, s, 1)
        if i >= args.warmup:
            times.append(t)
    assert cnt == n, f"reference join returned {cnt}, expected {n}"
    sec = sum(times) / len(times)
    val = 2 * n / sec / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"sortmergejoins -a m-way, |R|=|S|={n} 8-byte tuples, unique keys",
                   "threads": info.get("cores"), "timing": "wall clock around sortmergejoin_multiway"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": info.get("cores"), "kind": info.get("kind"),
                         "sample": "full workload, every step"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- our arm
def parse_profile(txt: str) -> dict:
    prof = {}
    for line in txt.strip().splitlines():
        name, cnt, ms, items = line.rsplit(";", 3)
        name = name.strip("() ")
        prof[name] = {"launches": int(cnt), "ms": float(ms), "items": int(items)}
    return prof


def run_single_gpu(args, smj) -> dict:
    lib = smj.lib
    n = args.size
    smj.check(lib.smjdev_init(int(os.environ.get("LOCAL_RANK", "0"))), "init")
    dr, ds = smj.DeviceBuffer(n), smj.DeviceBuffer(n)
    smj.check(lib.smjdev_gen_unique(dr.ptr, n, 0, n, SEED_R), "gen R")
    smj.check(lib.smjdev_gen_unique(ds.ptr, n, 0, n, SEED_S), "gen S")
    smj.check(lib.smjdev_sync())

    # ---- device-resident throughput
    st = None
    for _ in range(args.warmup):
        st = smj.device_join(smj.ALGO_MWAY, dr, ds)
    assert st is None or st.matches == n, f"join returned {st.matches}, expected {n}"
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.smjdev_event_create(C.byref(e0))
    lib.smjdev_event_create(C.byref(e1))
    sampler = ClockSampler(smj.lib.smjdev_current_device())
    time.sleep(0.3)
    launches0 = lib.smjdev_launch_count()
    lib.smjdev_sync()
    lib.smjdev_event_record(e0)
    phases = {"partition": 0.0, "sort": 0.0, "merge": 0.0, "join": 0.0}
    for _ in range(args.steps):
        st = smj.device_join(smj.ALGO_MWAY, dr, ds)
        assert st.matches == n
        phases["partition"] += st.ms_partition
        phases["sort"] += st.ms_sort
        phases["merge"] += st.ms_merge
        phases["join"] += st.ms_join
    lib.smjdev_event_record(e1)
    ms = C.c_float()
    smj.check(lib.smjdev_event_elapsed_ms(e0, e1, C.byref(ms)), "elapsed")
    launches = lib.smjdev_launch_count() - launches0
    ms_step = ms.value / args.steps
    value = 2 * n / (ms_step * 1e-3) / 1e6

    # ---- per-kernel durations (CUDA events on the launch stream), same steps again
    lib.smjdev_profile_begin()
    for _ in range(args.steps):
        smj.device_join(smj.ALGO_MWAY, dr, ds)
    buf = C.create_string_buffer(1 << 16)
    lib.smjdev_profile_end(buf, len(buf))
    prof = parse_profile(buf.value.decode())
    total_kernel_ms = sum(p["ms"] for p in prof.values()) or 1.0
    top = max(prof, key=lambda k: prof[k]["ms"])
    bytes_per_item = {"k_onesweep": 16, "k_msd_scatter2": 16, "k_msd_hist2": 8, "k_msd_hist1": 8,
                      "k_bucket_join": 8, "k_mj_tile": 8, "k_hist_digits": 8, "k_bits_reduce": 8}
    bpi = next((v for k, v in bytes_per_item.items() if k in top), 16)
    peak, peak_src = measured_peak()
    achieved = bpi * prof[top]["items"] / (prof[top]["ms"] * 1e-3) / 1e9 if prof[top]["ms"] else 0.0
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        for kname, rec in sorted(tj.get("kernels", {}).items(), key=lambda kv: len(kv[0])):
            if kname in top and rec.get("items"):  # the longest matching name wins
                traffic = rec["dram_bytes_per_launch"] * (prof[top]["items"] / prof[top]["launches"]) / rec["items"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_item": bpi, "launches": prof[top]["launches"],
                "avg_launch_ms": prof[top]["ms"] / max(prof[top]["launches"], 1),
                "share_of_step": prof[top]["ms"] / total_kernel_ms,
                "whole_join_frac": (40.0 * 2 * n / (ms_step * 1e-3) / 1e9) / peak,
                "kernels": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps}
                            for k, v in prof.items()}}

    # ---- end to end through sortmergejoin_multiway() with pinned host relations
    hr, hs = lib.smjdev_host_alloc(n * 8), lib.smjdev_host_alloc(n * 8)
    assert hr and hs, "pinned host allocation failed"
    smj.check(lib.smjdev_memcpy_d2h(hr, dr.ptr, n * 8))
    smj.check(lib.smjdev_memcpy_d2h(hs, ds.ptr, n * 8))
    smj.check(lib.smjdev_sync())
    cfg = smj.joinconfig_t(1, 128, 0, 0, 20 * 1024 * 1024, smj.NUMA_NEXT)
    relr, rels = smj.relation_t(hr, n), smj.relation_t(hs, n)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]

    h2d_ms = []

    def host_join():
        res = lib.sortmergejoin_multiway(C.byref(relr), C.byref(rels), C.byref(cfg))
        assert res, "sortmergejoin_multiway failed: " + lib.smjdev_last_error().decode()
        total = res.contents.totalresults
        hst = smj.smj_join_stats_t()
        lib.smj_last_join_stats(C.byref(hst))
        h2d_ms.append(hst.ms_h2d)
        libc.free(C.cast(res.contents.resultlist, C.c_void_p))
        libc.free(C.cast(res, C.c_void_p))
        return total

    for _ in range(max(1, min(args.warmup, 2))):
        assert host_join() == n
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert host_join() == n
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e = {"value": 2 * n / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8,
           "d2h_bytes_per_step": 16, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "h2d_ms_per_step": sum(h2d_ms[-e2e_steps:]) / e2e_steps,
           "api": "sortmergejoin_multiway(relation_t*, relation_t*, joinconfig_t*) with pinned host buffers"}

    # ---- the same call with ordinary malloc'ed (pageable) relations, as the reference's main.c
    #      allocates them: goes through the library's threaded staging pipeline
    try:
        r_pg = np.ctypeslib.as_array(C.cast(hr, C.POINTER(C.c_uint64)), shape=(n,)).copy()
        s_pg = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_uint64)), shape=(n,)).copy()
        relr_pg, rels_pg = smj.relation_t(r_pg.ctypes.data, n), smj.relation_t(s_pg.ctypes.data, n)

        def pageable_join():
            res = lib.sortmergejoin_multiway(C.byref(relr_pg), C.byref(rels_pg), C.byref(cfg))
            assert res and res.contents.totalresults == n
            libc.free(C.cast(res.contents.resultlist, C.c_void_p))
            libc.free(C.cast(res, C.c_void_p))

        pageable_join()
        t0 = time.perf_counter()
        for _ in range(2):
            pageable_join()
        pg_s = (time.perf_counter() - t0) / 2
        e2e["pageable_host_memory"] = {"value": 2 * n / pg_s / 1e6, "unit": UNIT, "ms_per_step": pg_s * 1e3}
        del r_pg, s_pg
    except Exception as exc:
        e2e["pageable_host_memory"] = {"value": None, "error": str(exc)}

    # the clock sampler ran through the device-resident steps, the per-kernel pass and the
    # end-to-end steps (the first timed region alone, a few milliseconds, is shorter than
    # nvidia-smi's sampling period); it stops before the CPU baseline loads the host
    clocks = sampler.stop()

    # ---- CPU baseline: the reference's own m-way join on this box's host cores
    cpu = None
    if not args.no_cpu_baseline:
        r_h = np.ctypeslib.as_array(C.cast(hr, C.POINTER(C.c_uint64)), shape=(n,))
        s_h = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_uint64)), shape=(n,))
        try:
            sec, cnt, info = cpu_reference_join(r_h, s_h, 2)
            assert cnt == n
            cpu = {"value": 2 * n / sec / 1e6, "unit": UNIT, "cores": info["cores"], "kind": info["kind"],
                   "sample": f"full workload ({n} x {n}), best of 2 runs, {sec:.3f} s"}
        except Exception as exc:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "reference", "sample": f"failed: {exc}"}
    lib.smjdev_host_free(hr)
    lib.smjdev_host_free(hs)

    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"sortmergejoins -a m-way, |R|=|S|={n} 8-byte tuples, uniform unique keys "
                               f"(BASELINE.json configs[0])",
                   "keys": "pseudo-random permutations of 1..n, seeds 12345/54321, payload = row index",
                   "l2": "inputs (2 x %.2f GB) are larger than L2, no flush needed" % (n * 8 / 1e9),
                   "timing": "CUDA events on the library's stream around all steps",
                   "phases_ms": {k: v / args.steps for k, v in phases.items()}},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "cpu_baseline": cpu, "matches": int(st.matches), "checksum": int(st.checksum),
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=128_000_000, help="tuples per relation")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
