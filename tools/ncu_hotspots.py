#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics per kernel and the hottest source lines (share of executed warp
instructions and of stall samples).  Usage: python tools/ncu_hotspots.py report.ncu-rep [kernel-substring] [top-n]"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
           "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
           "launch__block_size", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def run(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        if want not in name:
            continue
        print(f"## {name}")
        for m in METRICS:
            if m in hdr:
                print(f"{m} = {row[hdr.index(m)]} {units[hdr.index(m)]}")
    text = run(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + (["-k", f"regex:{want}"] if want else []))
    kernel, fname, cols = None, None, None
    by_kernel = {}
    for row in csv.reader(io.StringIO(text)):
        if not row:
            continue
        if row[0] == "Function Name":
            kernel = row[1]
        elif row[0] == "File Path":
            fname = row[1].split("/")[-1]
        elif row[0] == "Line No":
            cols = row
        elif cols is not None and row[0].isdigit() and "Instructions Executed" in cols:
            ii, si = cols.index("Instructions Executed"), cols.index("# Samples")
            try:
                inst, samp = int(row[ii] or 0), int(row[si] or 0)
            except ValueError:
                continue
            if inst or samp:
                by_kernel.setdefault(kernel, []).append((inst, samp, fname, int(row[0]), row[1].strip()))
    for kernel, lines in by_kernel.items():
        tot_i, tot_s = sum(l[0] for l in lines) or 1, sum(l[1] for l in lines) or 1
        print(f"\n## hottest source lines of {kernel}: total warp instructions {tot_i}, samples {tot_s}")
        for inst, samp, fname, ln, src in sorted(lines, reverse=True)[:top]:
            print(f"{fname:22s}:{ln:4d} inst {100 * inst / tot_i:5.1f}%  samples {100 * samp / tot_s:5.1f}%  {src[:110]}")


if __name__ == "__main__":
    main()
