#!/usr/bin/env python3
"""Summarises an .ncu-rep (ncu --set full --import-source on) into the text committed under profiles/:
headline metrics of the kernel plus stall samples / executed instructions aggregated per device function
and the hottest source lines.  Usage: ncu_summary.py report.ncu-rep [out.txt]"""
import bisect, csv, io, os, re, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout

raw = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_shared_ld.sum",
        "smsp__inst_executed_op_shared_st.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_dmma.sum", "smsp__inst_executed_pipe_tensor_op_dmma.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
for row in raw[2:]:
    print("=" * 100, file=out)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:90s} {row[i]} {units[i]}", file=out)

src = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass"))))
sections = [i for i, r in enumerate(src) if r and r[0] == "File Path"]
print("\nPer-function share of warp-stall samples and executed warp instructions (source-correlated)", file=out)
for si, start in enumerate(sections):
    path = src[start][1]
    end = sections[si + 1] if si + 1 < len(sections) else len(src)
    h = src[start + 2] if src[start + 1] and src[start + 1][0] == "Function Name" else src[start + 1]
    base = start + 3 if src[start + 1][0] == "Function Name" else start + 2
    if "# Samples" not in h:
        continue
    iS, iI = h.index("# Samples"), h.index("Instructions Executed")
    iW = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
    iE = h.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in h else None
    lines = []
    for r in src[base:end]:
        if r and r[0].isdigit():
            try:
                lines.append((int(r[0]), r[1], int(r[iS]), int(r[iI]), int(r[iW] or 0) if iW is not None else 0,
                              int(r[iE] or 0) if iE is not None else 0))
            except ValueError:
                pass
    if not lines:
        continue
    ts, ti = sum(l[2] for l in lines) or 1, sum(l[3] for l in lines) or 1
    text = open(path).read().split("\n") if os.path.exists(path) else []
    funcs = [(n, m.group(1)) for n, l in enumerate(text, 1)
             for m in [re.match(r"(?:LMH_HDN?|__global__|__device__(?: __forceinline__)?)\s+.*?(\w+)\(", l)] if m]
    agg = {}
    for ln, _, s, i, _w, _e in lines:
        k = bisect.bisect_right([f[0] for f in funcs], ln) - 1
        a = agg.setdefault(funcs[k][1] if k >= 0 else "?", [0, 0])
        a[0] += s; a[1] += i
    print(f"\n-- {os.path.relpath(path, ROOT)}: {ts} samples, {ti} warp instructions", file=out)
    for name, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"   {name:30s} samples {100 * s / ts:5.1f}%   instructions {100 * i / ti:5.1f}%", file=out)
    print("   hottest lines:", file=out)
    for ln, t, s, i, _w, _e in sorted(lines, key=lambda l: -l[2])[:25]:
        print(f"   {ln:5d} {100 * s / ts:5.1f}% {100 * i / ti:5.1f}%  {t.strip()[:100]}", file=out)
    tw = sum(l[4] for l in lines)
    if tw:
        te = sum(l[5] for l in lines)
        print(f"   shared-memory wavefronts: {tw} ({te} = {100 * te / tw:.1f}% excessive, i.e. bank conflicts); lines with the most:", file=out)
        print("    line  wavefronts  of which excessive (bank conflicts)", file=out)
        for ln, t, s, i, w_, e_ in sorted(lines, key=lambda l: -l[4])[:25]:
            print(f"   {ln:5d} {100 * w_ / tw:5.1f}% {100 * e_ / tw:5.1f}%  {t.strip()[:100]}", file=out)
