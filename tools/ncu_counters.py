#!/usr/bin/env python
"""Executed-work counters of every kernel of a bench step, from one `ncu --set full --import-source on` capture.

    python tools/ncu_counters.py report.ncu-rep --workload balmer_series --batch 4096 --out profiles/r2_balmer_series_counters.json

For each captured kernel the SASS source page gives the predicated-on THREAD instruction count of every opcode; this
script sums them by class (no hand-written flop model):
    fp32 flop  = 2 FFMA + FMUL + FADD + 2 (2 FFMA2 + FMUL2 + FADD2)      (packed .f32x2 forms carry two lanes)
    mufu ops   = MUFU                                                     (ex2, lg2, rcp, rsq, sqrt, sin, cos)
    xu ops     = sm__inst_executed_pipe_xu (busy fraction x cycles x 16 lanes/clk x SMs): MUFU plus the conversions
                 that execute on the same pipe
    fp64 flop  = 2 DFMA + DMUL + DADD
    tensor flop = UTCHMMA x 2 x 128 x N x 8                               (kind::tf32, M = 128, K = 8 per instruction)
and the raw page gives dram bytes, duration and the pipe utilisations.  bench.py divides by the thetas of the captured
launch (`--batch`) and by the kernel's live CUDA-event time to form `roofline.achieved`.
"""
import argparse
import collections
import csv
import io
import json
import re
import subprocess

STAGE = [("los_stage", "los"), ("chord_stage", "chord"), ("pix_contract", "contract"), ("if_contract", "contract"),
         ("pixel_fold", "pixel"), ("pixel_stage", "pixel"), ("finalize", "finalize")]
RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
       "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
       "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed"]
XU_LANES_PER_CLK_SM = 16  # one warp instruction per two cycles and SM: the rate the in-run MUFU micro-benchmark measures
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def stage_of(name):
    for key, stage in STAGE:
        if key in name:
            return stage
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--batch", type=int, required=True, help="thetas of the captured launches")
    ap.add_argument("--out", required=True)
    ap.add_argument("--command", default="")
    ap.add_argument("--sms", type=int, default=148, help="SMs of the GPU the capture ran on")
    args = ap.parse_args()

    raw = list(csv.reader(io.StringIO(ncu("-i", args.report, "--page", "raw", "--csv"))))
    hdr, units = raw[0], raw[1]
    kernels = {}
    for row in raw[2:]:
        name = row[hdr.index("Kernel Name")]
        stage = stage_of(name)
        if stage is None or stage in kernels:
            continue  # first capture of each stage
        k = {"kernel_name": name, "thetas_per_launch": args.batch}
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                try:
                    k[m] = float(row[i]) * (UNIT.get(units[i], 1.0) if m.startswith(("dram__", "gpu__time")) else 1.0)
                except ValueError:
                    pass
        k["dram_bytes_per_launch"] = k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
        k["ncu_duration_ms"] = k.get("gpu__time_duration.sum")
        kernels[stage] = k

    for stage, k in kernels.items():
        short = k["kernel_name"].split("(")[0].split("<")[0].split()[-1].split("::")[-1]
        text = ncu("-i", args.report, "--page", "source", "--print-source", "sass", "--csv", "-k", f"regex:{short}")
        cols, ops, warp_ops, seen = None, collections.Counter(), collections.Counter(), 0
        for r in csv.reader(io.StringIO(text)):
            if not r:
                continue
            if r[0] == "Kernel Name":
                seen += 1
                if seen > 1:
                    break
                continue
            if r[0] == "Address":
                cols = r
                continue
            if cols is None or not r[0].startswith("0x"):
                continue
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[1].strip())
            op = m.group(2) if m else "?"
            try:
                ops[op] += int(r[cols.index("Predicated-On Thread Instructions Executed")] or 0)
                warp_ops[op] += int(r[cols.index("Instructions Executed")] or 0)
            except ValueError:
                pass
        n = float(args.batch)
        tn = 128
        mt = re.search(r"pix_contract_kernel<\(?(?:int\))?(\d+)", k["kernel_name"])
        if mt:
            tn = int(mt.group(1))
        k["thread_inst"] = {op: ops[op] for op in ("FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2", "MUFU", "DFMA", "DMUL",
                                                  "DADD", "UTCHMMA") if ops[op]}
        k["fp32_flop_per_eval"] = (2 * ops["FFMA"] + ops["FMUL"] + ops["FADD"] + 2 * (2 * ops["FFMA2"] + ops["FMUL2"] + ops["FADD2"])) / n
        k["mufu_per_eval"] = ops["MUFU"] / n
        # every instruction of the XU pipe (MUFU and the int <-> float conversions that share it): the pipe's busy
        # fraction over the launch x its peak rate x the SMs of the capture (sm__inst_executed_pipe_xu)
        if "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed" in k and "sm__cycles_elapsed.avg" in k:
            k["xu_ops_per_eval"] = (k["sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed"] / 100.0 *
                                    k["sm__cycles_elapsed.avg"] * XU_LANES_PER_CLK_SM * args.sms) / n
        k["fp64_flop_per_eval"] = (2 * ops["DFMA"] + ops["DMUL"] + ops["DADD"]) / n
        k["tensor_flop_per_eval"] = warp_ops["UTCHMMA"] * 2.0 * 128 * tn * 8 / n
        k["warp_inst_per_eval"] = sum(warp_ops.values()) / n
        top = warp_ops.most_common(12)
        tot = sum(warp_ops.values()) or 1
        k["warp_inst_mix_pct"] = {op: round(100.0 * v / tot, 1) for op, v in top}
    out = {"workload": args.workload, "report": args.report.split("/")[-1], "command": args.command, "kernels": kernels,
           "method": "thread-level SASS opcode counts of the ncu source page (see tools/ncu_counters.py)"}
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    for stage, k in kernels.items():
        print(f"{stage:9s} {k['ncu_duration_ms']:.4f} ms  fp32 {k['fp32_flop_per_eval']:.0f} flop/eval  mufu {k['mufu_per_eval']:.0f}  "
              f"fp64 {k['fp64_flop_per_eval']:.0f}  tensor {k['tensor_flop_per_eval']:.0f}  warp-inst {k['warp_inst_per_eval']:.0f}")


if __name__ == "__main__":
    main()
