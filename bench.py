#!/usr/bin/env python
"""Benchmark of the BaySAR posterior-evaluation path on B200 (metric: log-posterior evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload balmer_series|multi_species]

A "step" is one pass of the hot path (LOS stage -> chord stage -> finalize, three kernel launches) over one batch
of synthetic thetas drawn uniformly inside the posterior's theta bounds.  Default workload = BASELINE.json
configs[1] (`demo/balmer_series.py`, 4096 thetas per GPU per step).

* `value`       evals/s with thetas resident in HBM (CUDA events around each step, max over ranks);
* `e2e`         the same metric through the reference-facing call `posterior(theta_host_array)` with pinned HOST
                buffers: H2D copy of theta and D2H read of logP + status inside the timed region;
* `roofline`    for the dominant kernel (chord stage): executed FP32 flop per launch / its mean duration, against the
                FP32-FMA peak measured in the same run by the library's micro-benchmark (MEASURED_PEAKS.json carries
                no FP32 figure); the SURVEY's dense-semantics flop count is reported beside it;
* `cpu_baseline` the oracle port (`oracle/numpy_port.py`, kind "port") on the box's host cores, bounded sample.

With `--impl reference` the CPU oracle port is the thing timed (all host cores), on the same workload and metric.
Under torchrun (N > 1) every rank evaluates its own 4096-theta shard (weak scaling) and the per-walker
log-posteriors are all-gathered over NCCL inside the timed region, as an ensemble-sampler step needs.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "log-posterior evals/sec"
UNIT = "evals/s"
BATCH_PER_GPU = {"balmer_series": 4096, "multi_species": 65536, "multi_species_32chords": 4096, "aug_demo": 4096}


def get_workload(name):
    from baysar_b200 import configs

    if name == "balmer_series":
        return configs.balmer_series()
    if name == "multi_species":
        return configs.multi_species(1)
    if name == "multi_species_32chords":
        return configs.multi_species(32)
    if name == "aug_demo":
        return configs.aug_demo()
    raise SystemExit(f"unknown workload {name}")


def draw_thetas(bounds, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = bounds[:, 0], bounds[:, 1]
    return lo + rng.random((n, len(lo))) * (hi - lo)


# ----------------------------------------------------------------------------------------------- CPU oracle legs
def _oracle_worker(args):
    name, thetas = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from oracle.numpy_port import EsymmtricCauchyPlasma, OracleError, OraclePosterior

    wl = get_workload(name)
    prof = EsymmtricCauchyPlasma(x=wl.los) if wl.los is not None else None
    post = OraclePosterior(make_input_dict(**wl.input_kwargs), SyntheticADAS(), prof)
    post.evaluate(thetas[0]) if len(thetas) else None  # warm-up
    t0 = time.perf_counter()
    done = 0
    for th in thetas:
        try:
            post(th)
        except OracleError:
            pass
        done += 1
    return done, time.perf_counter() - t0


def cpu_oracle_rate(name, thetas, cores):
    """Aggregate evals/s of the oracle port over `cores` worker processes, each with its own posterior replica and a
    disjoint theta shard (the reference's own multi-process pattern, baysar/optimisation.py:140-151)."""
    import multiprocessing as mp

    shards = [s for s in np.array_split(thetas, cores) if len(s)]
    t0 = time.perf_counter()
    if len(shards) == 1:
        results = [_oracle_worker((name, shards[0]))]
    else:
        with mp.get_context("fork").Pool(len(shards)) as pool:
            results = pool.map(_oracle_worker, [(name, s) for s in shards])
    wall = time.perf_counter() - t0
    done = sum(r[0] for r in results)
    busy = max(r[1] for r in results)
    return done / busy, done, wall


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    """SM clock / power / clock-event reasons DURING the timed region, from a separate `nvidia-smi -lms 20` process
    (started well before the timed region so that it is already streaming; a polling thread inside this process
    perturbs multi-rank steps through the GIL).  Rows are time-stamped; only those inside [t0, t1] are used."""

    FIELDS = ["timestamp", "clocks.sm", "clocks.max.sm", "power.draw", "clocks_event_reasons.hw_slowdown",
              "clocks_event_reasons.hw_thermal_slowdown", "clocks_event_reasons.sw_thermal_slowdown",
              "clocks_event_reasons.sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = visible.split(",")[self.index] if visible else str(self.index)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={idx}", "--query-gpu=" + ",".join(self.FIELDS),
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        import datetime

        for line in self.proc.stdout:
            cells = [c.strip() for c in line.split(",")]
            if len(cells) != len(self.FIELDS):
                continue
            try:
                ts = datetime.datetime.strptime(cells[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                self.rows.append((ts, float(cells[1]), float(cells[2]), float(cells[3]), cells[4:]))
            except ValueError:
                continue

    def stop(self, t0, t1):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
        inside = [r for r in self.rows if t0 - 0.02 <= r[0] <= t1 + 0.02] or self.rows[-3:]
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max([r[2] for r in inside]) if inside else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_total": len(self.rows),
                "power_w_max": max([r[3] for r in inside]) if inside else None}


# ----------------------------------------------------------------------------------------------- flop accounting
def flop_model(post):
    """Executed FP32 flop and MUFU (SFU) ops per eval, per kernel, from the compiled model, and the SURVEY section 8d
    dense-semantics figure.

    chord (FP32 + MUFU pipes): narrow-Doppler Balmer lines (cold neutrals) cost 15 flop + 2 MUFU per sweep-1 sample
    (dense +-256-point window, every 16th point on the tails) and 14 flop + 1 MUFU per fine-grid point in sweep 2;
    Balmer lines with a multi-tap Doppler kernel 18 F + 2 taps F flop and 3 F MUFU; Gaussian components 8 flop + 1
    MUFU per point of their +-8.85 sigma window; 14 F + 12 P of bookkeeping and (fused path only) the instrument
    convolution at 2 flop per tap and output.  contract (tensor pipe, 3xTF32): three 128 x TN x 8 MMAs per 8-deep
    k-step and 16-wide k-block, zero parts of the banded tiles included — k-blocks of the folded pixel contraction as
    the library planned them (static wavelength map), else of the every-column Toeplitz contraction."""
    from baysar_b200.spectrometers import BalmerHydrogenLine

    tensor_path = post.model.uses_tensor_cores
    chord = contract = dense = mufu = 0.0
    for c_idx, ch in enumerate(post.chords):
        F, P = len(ch.x_data_fm), len(ch.x_data)
        n_b = sum(isinstance(l, BalmerHydrogenLine) for l in ch.lines)
        gauss = [l for l in ch.lines if not isinstance(l, BalmerHydrogenLine)]
        n_g = sum(len(getattr(l, "lines", getattr(l, "cwl", []))) for l in gauss)
        k_d = max([len(l.wavelengths_doppler) for l in ch.lines if isinstance(l, BalmerHydrogenLine)] or [0])
        inst = np.abs(ch.instrument_function_fm)
        keep = np.where(inst > 1e-17 * inst.max())[0]
        k_lo, k_hi = int(keep.min()), int(keep.max()) + 1
        nq_if = ((k_hi - k_lo) + 3) // 4 * 4 + 4
        dl = np.diff(ch.x_data_fm).mean()
        ti = 0.01 if post.plasma.cold_neutrals else 5.0
        sigma = 4000.0 * 7.715e-5 * np.sqrt(ti / 2.0) / np.sqrt(8 * np.log(2))
        n_taps = 2 * int(6.9 * sigma / dl) + 1
        if n_taps == 1:
            n_s1 = min(F, 545) + max(F - 545, 0) / 16.0
            chord += n_b * (15.0 * n_s1 + 14.0 * F)
            mufu += n_b * (2.0 * n_s1 + F)
        else:
            nq_d = (n_taps + 3) // 4 * 4 + 4
            chord += n_b * (18.0 * F + 2.0 * nq_d * F)
            mufu += n_b * 3.0 * F
        # Gaussian components: typical window for an impurity ion at Ti = 0.01 eV / a pixel-wide X line
        sig_g = max(4000.0 * 7.715e-5 * np.sqrt(0.01 / 14.0) / np.sqrt(8 * np.log(2)), 0.0)
        pts_g = n_g * (2 * 8.85 * max(sig_g, dl / 4) / dl + 1)
        chord += 8.0 * pts_g + 14.0 * F + 12.0 * P
        mufu += pts_g + P
        kb_folded, tn_folded = post.model.folded_contraction[c_idx]
        if tensor_path and kb_folded > 0:
            contract += kb_folded * 2 * 3 * 2.0 * (128 * tn_folded * 8) / 128.0
        elif tensor_path:
            o = (len(inst) - 1) // 2
            n_kb = (127 + o - k_lo) // 16 - (o - k_hi + 1) // 16 + 1   # 16-wide k-blocks per 128-output tile
            n_tiles = -(-F // 128)
            contract += n_tiles * n_kb * 2 * 3 * 2.0 * (128 * 128 * 8) / 128.0
        else:
            chord += 2.0 * nq_if * F
        dense += 8.0 * n_g * F + n_b * 2 * (10.0 * F + 2.0 * k_d * F + 3.0 * F) + 12.0 * F + 10.0 * P \
            + 2.0 * F * len(ch.instrument_function_fm)
    return {"chord": chord, "contract": contract, "dense": dense, "chord_mufu": mufu}


def ncu_dram_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` (bytes per launch) from the committed `ncu --set full`
    summary of the same bench command (profiles/r1_s3_<workload>_ncu_full.txt, written by tools/ncu_hotspots.py), or
    None when no capture of this workload / kernel is committed.  Never measured under the profiler in this run."""
    path = os.path.join(ROOT, "profiles", f"r1_s3_{workload}_ncu_full.txt")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        lines = open(path).read().splitlines()
    except OSError:
        return None, None
    total, inside, seen = 0.0, False, 0
    for line in lines:
        if line.startswith("## "):
            inside = line.startswith("## void") and kernel in line
        elif inside and line.startswith(("dram__bytes_read.sum =", "dram__bytes_write.sum =")):
            parts = line.split("=")[1].split()
            total += float(parts[0]) * unit.get(parts[1] if len(parts) > 1 else "byte", 1.0)
            seen += 1
    return (total, os.path.relpath(path, ROOT)) if seen == 2 else (None, None)


# ----------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="balmer_series")
    ap.add_argument("--batch", type=int, default=0, help="thetas per GPU per step (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    batch = args.batch or BATCH_PER_GPU[args.workload]
    config = {"workload": f"{args.workload}: {get_workload(args.workload).description}",
              "thetas_per_gpu_per_step": batch, "theta_distribution": "uniform inside posterior.plasma.theta_bounds",
              "atomic_data": "SyntheticADAS closed-form tables", "l2": "256 MiB scratch write between timed steps"}

    if args.impl == "reference":
        # The reference is pure Python and cannot travel to the GPU box: its CPU implementation of the path is
        # timed through the oracle port (bit-identical to the shimmed reference on the golden vectors).
        if rank != 0:
            return
        cores = usable_cores()
        np.random.seed(0)
        wl = get_workload(args.workload)
        from baysar_b200.atomic_data import SyntheticADAS
        from baysar_b200.input_functions import make_input_dict
        from oracle.numpy_port import EsymmtricCauchyPlasma, OraclePosterior

        prof = EsymmtricCauchyPlasma(x=wl.los) if wl.los is not None else None
        bounds = OraclePosterior(make_input_dict(**wl.input_kwargs), SyntheticADAS(), prof).theta_bounds
        per_step = max(cores * 32, 64)
        thetas = draw_thetas(bounds, per_step * (args.steps + args.warmup), 20261019)
        for w in range(args.warmup):
            cpu_oracle_rate(args.workload, thetas[w * per_step:(w + 1) * per_step], cores)
        t_total, done = 0.0, 0
        for s in range(args.warmup, args.warmup + args.steps):
            rate, n_done, wall = cpu_oracle_rate(args.workload, thetas[s * per_step:(s + 1) * per_step], cores)
            t_total += n_done / rate
            done += n_done
        value = done / t_total
        sample = f"{per_step} thetas per step x {args.steps} steps over {cores} forked workers"
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    from baysar_b200 import runtime
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.lineshapes import EsymmtricCauchyPlasma
    from baysar_b200.posterior import BaysarPosterior

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    wl = get_workload(args.workload)
    np.random.seed(0)
    prof = EsymmtricCauchyPlasma(x=wl.los) if wl.los is not None else None
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=prof, atomic_data=SyntheticADAS(),
                           device=local_rank, skip_test=True)
    n_params = post.plasma.n_params
    total_steps = args.warmup + args.steps
    # a different theta batch every step; this rank's shard of the global batch
    thetas_host = draw_thetas(post.plasma.theta_bounds, batch * total_steps, 20261019 + rank).reshape(total_steps, batch, n_params)
    thetas_dev = torch.as_tensor(thetas_host, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    from baysar_b200.parallel import ShardedEvaluator

    sharded = ShardedEvaluator(post.logp_batch)  # world 1: plain call; world > 1: + NCCL all-gather of logP

    # N > 1: the all-gather of a step's logP (8 B per theta) runs on a side stream and overlaps with the next step: the
    # ranks' theta shards are independent walkers, so nothing on the critical path waits for it (BAYSAR_BENCH_SYNC_GATHER=1
    # puts it back on the compute stream).  The tail of the last gather is added to the timed total below.
    overlap = world > 1 and os.environ.get("BAYSAR_BENCH_SYNC_GATHER") != "1"

    def step(i):
        return sharded.evaluate_local(thetas_dev[i], overlap=overlap)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    post.model.set_profiling(os.environ.get("BAYSAR_BENCH_NOPROF") != "1")
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_wall0 = time.time()
    for k in range(args.steps):
        flush.fill_(float(k))  # evict L2 between timed steps (untimed)
        starts[k].record()
        out = step(args.warmup + k)
        ends[k].record()
    tail0, tail1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tail0.record()
    sharded.wait()      # outstanding overlapped all-gathers
    tail1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    stage_ms, calls = post.model.get_profile()
    post.model.set_profiling(False)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    total_ms = torch.tensor([float(step_ms.sum()) + float(tail0.elapsed_time(tail1))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * batch * args.steps / (total_ms * 1e-3)
    status = post.last_status_device.cpu().numpy()

    # ---- end-to-end: reference-facing call with pinned HOST buffers (H2D + D2H inside the timed region)
    pinned = torch.empty((batch, n_params), dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for i in range(2):
        pinned.copy_(torch.from_numpy(thetas_host[i]))
        post(pinned.numpy())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_e2e = 0.0
    for k in range(e2e_steps):
        pinned.copy_(torch.from_numpy(thetas_host[(args.warmup + k) % total_steps]))
        flush.fill_(float(k))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        logp_host = post(pinned.numpy())  # numpy in, numpy out: copies happen inside baysar_eval_logp_host
        t_e2e += time.perf_counter() - t0
    t_e2e_t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * batch * e2e_steps / float(t_e2e_t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel, measured live (CUDA events around every kernel of the timed steps)
    peaks = runtime.microbench(local_rank)
    flops = flop_model(post)
    per_call = {k: v / max(calls, 1) for k, v in stage_ms.items()}
    measured = {}
    try:
        measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    fp32_peak = peaks["fp32_fma_flops"] / 1e12
    tf32_peak = measured.get("bf16_tflops", 1590.0) / 2.0
    contract_name = "pix_contract_kernel" if any(k for k, _ in post.model.folded_contraction) else "if_contract_kernel"
    kernels = {
        "chord_stage_kernel": {"bound": "fp32", "ms": per_call["chord"], "flop_per_eval": flops["chord"],
                               "peak": fp32_peak,
                               "peak_source": "FP32 FMA micro-benchmark in this run (baysar_microbench); "
                                              "MEASURED_PEAKS.json has no FP32 figure"},
        contract_name: {"bound": "tensor", "ms": per_call["contract"], "flop_per_eval": flops["contract"],
                               "peak": tf32_peak,
                               "peak_source": "TF32 dense = half of MEASURED_PEAKS.json bf16_tflops (burst)"
                               if measured else "TF32 dense = half of the fallback 1.59 PFLOP/s bf16"},
    }
    for k in kernels.values():
        k["achieved"] = k["flop_per_eval"] * batch / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else 0.0
        k["frac"] = k["achieved"] / k["peak"]
        k["unit"] = "TFLOP/s"
    # the spectral kernel issues FP32 and MUFU work side by side: report both pipes, the larger fraction governs
    ck = kernels["chord_stage_kernel"]
    ck["mufu_per_eval"] = flops["chord_mufu"]
    ck["mufu_achieved_Gops"] = flops["chord_mufu"] * batch / (ck["ms"] * 1e-3) / 1e9 if ck["ms"] > 0 else 0.0
    ck["mufu_peak_Gops"] = peaks["mufu_ops"] / 1e9
    ck["frac_mufu"] = ck["mufu_achieved_Gops"] / ck["mufu_peak_Gops"]
    ck["frac_fp32"] = ck["frac"]
    if ck["frac_mufu"] > ck["frac"]:
        ck["bound"], ck["frac"] = "mufu", ck["frac_mufu"]
    ck["fp32_peak_3reg_tflops"] = peaks["fp32_fma_reg_flops"] / 1e12
    ck["fp32_peak_ffma2_tflops"] = peaks["fp32_fma2_flops"] / 1e12
    dom = max(kernels, key=lambda name: kernels[name]["ms"])
    hbm_bytes = batch * (8.0 * n_params + 8.0 + 4.0 + 8.0 * (len(post.chords) + len(post.priors))
                         + 2 * 8.0 * 16 * post.model.n_ulines)
    if post.model.uses_tensor_cores:  # FP32 spectra and convolved spectra (folded: pixel columns), written and read once
        hbm_bytes += batch * sum(len(c.x_data_fm) * 8.0 + (len(c.x_data) if post.model.folded_contraction[i][0] else
                                                           len(c.x_data_fm)) * 8.0 for i, c in enumerate(post.chords))
    roofline = dict(kernels[dom])
    traffic, traffic_src = ncu_dram_traffic(args.workload, dom)
    roofline.update({
        "kernel": dom, "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)",
        "traffic_source": traffic_src, "kernels": kernels,
        "dense_reference_flop_per_eval": flops["dense"],
        "dense_equivalent_tflops": flops["dense"] * batch / (float(step_ms.mean()) * 1e-3) / 1e12,
        "stage_ms": per_call, "kernel_share_of_step": kernels[dom]["ms"] / float(step_ms.mean()),
        "mufu_peak_gops": peaks["mufu_ops"] / 1e9, "fp64_peak_tflops": peaks["fp64_fma_flops"] / 1e12,
        "hbm_algorithmic_GBps": hbm_bytes / (float(step_ms.mean()) * 1e-3) / 1e9,
        "hbm_peak_GBps_measured": measured.get("hbm_gbs"),
    })

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        cores = usable_cores()
        pool = thetas_host.reshape(-1, n_params)
        rate1, done1, _ = cpu_oracle_rate(args.workload, pool[:200], 1)
        # bounded sample: about 12 s of CPU work per worker at the single-core rate just measured
        want = int(cores * max(200, 12.0 * rate1))
        sample = pool[np.arange(want) % len(pool)]
        rate, done, wall = cpu_oracle_rate(args.workload, sample, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{done} thetas of the same batch, {cores} forked workers, {wall:.1f} s wall",
               "single_core_value": rate1}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 element-wise / f64 line-of-sight, reductions and likelihood", "data": "synthetic",
            "config": dict(config, tensor_core_contraction=post.model.uses_tensor_cores,
                           logp_all_gather=("none (1 GPU)" if world == 1 else
                                            "NCCL, side stream, overlapped with the next step" if overlap else
                                            "NCCL on the compute stream")), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": batch * n_params * 8,
                    "d2h_bytes_per_step": batch * (8 + 4), "steps": e2e_steps},
            "gpu_launches": post.model.launches_last_eval * args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "status_ok_fraction": float((status == 0).mean())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
