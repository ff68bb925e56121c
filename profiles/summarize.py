#!/usr/bin/env python
"""Turns an Nsight Compute report (.ncu-rep, captured on the B200 box under gpurun with
`ncu --set full --clock-control none --import-source on`) into the text summaries committed in this
directory: headline metrics per launch, warp-stall breakdown, and stall samples per block of SASS
instructions (program order) so that the phases of a kernel can be told apart.

    python profiles/summarize.py gpurun_out/prof.ncu-rep [instructions-per-bucket [kernel-regex ...]] > profiles/rNN_x.txt
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def raw_page(rep):
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=== launch:", r[hdr.index("Kernel Name")][:110])
        print("    grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for w in WANT:
            if w in hdr:
                print(f"    {w:82s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
        st = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued"):
                try:
                    st.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        tot = sum(v for v, _ in st) or 1
        print("    warp stall samples:", ", ".join(f"{h} {100 * v / tot:.0f}%" for v, h in st[:8]))


def source_page(rep, bucket, kernel=None):
    """SASS stall samples of the first launch in the report, or of the first launch whose name
    matches the regex `kernel`."""
    rows, hdr, nk = [], None, 0
    extra = ["--kernel-name", "regex:" + kernel, "--launch-count", "1"] if kernel else []
    for r in csv.reader(ncu(rep, "--page", "source", "--csv", "--print-source", "sass", *extra).splitlines()):
        if r and r[0] == "Kernel Name":
            nk += 1
            if nk > 1:
                break
            print("=== SASS stall samples, first launch:", r[1][:100])
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) >= len(hdr) - 2:
            rows.append(r)
    if not hdr:
        return
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si] or 0) for r in rows) or 1
    print(f"    {len(rows)} SASS instructions, {tot} samples; buckets of {bucket} instructions in program order")
    for b in range(0, len(rows), bucket):
        chunk = rows[b:b + bucket]
        n = sum(int(r[si] or 0) for r in chunk)
        if n == 0:
            continue
        ex = sum(int(r[ie] or 0) for r in chunk)
        agg, ops = {}, {}
        for r in chunk:
            for i in stall:
                agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i] or 0)
            t = r[src].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        top = ", ".join(f"{k} {v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:3])
        topo = ", ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
        print(f"    [{b:5d},{b + bucket:5d}) {100 * n / tot:5.1f}% of samples, {ex / 1e6:7.1f}M warp-instr | {top} | {topo}")
    print("    hottest instructions:")
    for idx, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1][si] or 0))[:12]:
        print(f"      #{idx:5d} {100 * int(r[si] or 0) / tot:5.1f}%  {r[src].strip()[:90]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    raw_page(rep)
    bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    if len(sys.argv) > 3:  # regexes of the kernels whose SASS pages are wanted
        for k in sys.argv[3:]:
            source_page(rep, bucket, k)
    else:
        source_page(rep, bucket)
