#!/usr/bin/env python
"""Benchmark of the BaySAR posterior-evaluation path on B200 (metric: log-posterior evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--scaling weak|strong]

A "step" is one pass of the hot path (LOS stage -> spectral stage -> contraction -> pixel stage / finalize) over one
batch of synthetic thetas drawn uniformly inside the posterior's theta bounds.  Default workload = BASELINE.json
configs[1] (`demo/balmer_series.py`, 4096 thetas per GPU per step).

* `value`       evals/s with thetas resident in HBM (CUDA events around each step, max over ranks);
* `e2e`         the same metric through the reference-facing call `posterior(theta_host_array)` with pinned HOST
                buffers: H2D copy of theta and D2H read of logP + status inside the timed region;
* `parity`      the log-posteriors of one TIMED theta batch (and the forward-model spectra of 16 of its thetas)
                against the oracle port evaluated in this same process on the same thetas; the run exits non-zero
                when logP differs by more than 1e-6 relative or a pixel by more than 1e-5 relative;
* `roofline`    for the dominant kernel: executed FP32 flop / MUFU ops per launch (counts derived from the ncu
                instruction counters committed under profiles/, see `work_counts`) / its mean duration measured live
                with CUDA events, against the FP32-FMA / MUFU peaks measured in the same run by the tools library
                (MEASURED_PEAKS.json carries no FP32 figure); the SURVEY's dense-semantics flop count beside it;
* `cpu_baseline` the oracle port (`oracle/numpy_port.py`, kind "port") on the box's host cores, bounded sample;
* `extra_workloads` further lines of the same shape (without the CPU baseline): BASELINE configs[2] (multi-species
                chord, 65 536 thetas) and the hot-neutral Balmer chord (multi-tap Doppler kernels); under torchrun
                the multi-species chord in its weak (65 536 per GPU) and strong (65 536 in total) forms, and the
                main workload with the logP all-gather on the compute stream next to the overlapped one.

With `--impl reference` the CPU oracle port is the thing timed (all host cores), on the same workload and metric.
Under torchrun (N > 1) every rank evaluates its own shard and the per-walker log-posteriors are all-gathered over
NCCL inside the timed region, as an ensemble-sampler step needs.
"""
import argparse
import gc
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
BATCH_PER_GPU = {"balmer_series": 4096, "balmer_hot": 4096, "multi_species": 65536, "multi_hot": 16384,
                 "multi_species_32chords": 4096, "aug_demo": 4096}
LOGP_RTOL, SPECTRA_RTOL = 1e-6, 1e-5   # BASELINE.json north_star tolerances


def get_workload(name):
    from baysar_b200 import configs

    if name == "balmer_series":
        return configs.balmer_series()
    if name == "balmer_hot":
        return configs.balmer_series(flags=dict(cold_neutrals=False))
    if name == "multi_species":
        return configs.multi_species(1)
    if name == "multi_hot":
        return configs.multi_species(1, flags=dict(cold_neutrals=False, cold_ions=False))
    if name == "multi_species_32chords":
        return configs.multi_species(32)
    if name == "aug_demo":
        return configs.aug_demo()
    raise SystemExit(f"unknown workload {name}")


def workload_config(name, batch):
    """The `config` object of a bench line: identical for the b200 and the reference arm."""
    return {"workload": f"{name}: {get_workload(name).description}", "thetas_per_gpu_per_step": batch,
            "theta_distribution": "uniform inside posterior.plasma.theta_bounds",
            "atomic_data": "SyntheticADAS closed-form tables", "l2": "256 MiB scratch write between timed steps"}


def draw_thetas(bounds, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = bounds[:, 0], bounds[:, 1]
    return lo + rng.random((n, len(lo))) * (hi - lo)


# ----------------------------------------------------------------------------------------------- CPU oracle legs
def _oracle_posterior(name):
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from oracle.numpy_port import EsymmtricCauchyPlasma, OraclePosterior

    wl = get_workload(name)
    prof = EsymmtricCauchyPlasma(x=wl.los) if wl.los is not None else None
    return OraclePosterior(make_input_dict(**wl.input_kwargs), SyntheticADAS(), prof)


def _oracle_worker(args):
    """(workload, thetas, n_spectra) -> (evaluations done, busy seconds, logP [n] (NaN where the oracle raised), spectra)."""
    name, thetas, n_spectra = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle.numpy_port import OracleError

    post = _oracle_posterior(name)
    if len(thetas):
        try:
            post.evaluate(thetas[0])  # warm-up
        except OracleError:
            pass
    logp = np.full(len(thetas), np.nan)
    spectra = []
    t0 = time.perf_counter()
    for i, th in enumerate(thetas):
        try:
            if i < n_spectra:
                res = post.evaluate(th)
                logp[i] = res["logp"]
                spectra.append(np.asarray(res["spectra"][0], dtype=float))
            else:
                logp[i] = post(th)
        except OracleError:
            if i < n_spectra:
                spectra.append(None)
    return len(thetas), time.perf_counter() - t0, logp, spectra


def cpu_oracle_rate(name, thetas, cores, n_spectra=0):
    """Aggregate evals/s of the oracle port over `cores` worker processes, each with its own posterior replica and a
    disjoint theta shard (the reference's own multi-process pattern, baysar/optimisation.py:140-151).
    -> (evals/s, evaluations, wall seconds, logP per theta, spectra of the first n_spectra thetas)."""
    import multiprocessing as mp

    shards = [s for s in np.array_split(thetas, cores) if len(s)]
    n_spec = [max(0, min(len(s), n_spectra - sum(len(x) for x in shards[:k]))) for k, s in enumerate(shards)]
    t0 = time.perf_counter()
    if len(shards) == 1:
        results = [_oracle_worker((name, shards[0], n_spec[0]))]
    else:
        with mp.get_context("fork").Pool(len(shards)) as pool:
            results = pool.map(_oracle_worker, [(name, s, k) for s, k in zip(shards, n_spec)])
    wall = time.perf_counter() - t0
    done = sum(r[0] for r in results)
    busy = max(r[1] for r in results)
    spectra = [s for r in results for s in r[3]]
    return done / busy, done, wall, np.concatenate([r[2] for r in results]), spectra


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

    def window(self, t0, t1):
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

    def stop(self):
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()


# ----------------------------------------------------------------------------------------------- work accounting
def work_counts(name):
    """Executed work per evaluation of each kernel, from the ncu instruction counters of a committed capture:
    profiles/<round>_<workload>_counters.json (written by tools/ncu_counters.py from `ncu --set full`):
    thread-level FFMA/FMUL/FADD (packed forms counted twice), MUFU and FP64 instructions and DRAM bytes per launch,
    divided by the thetas of the captured launch.  None when no capture of this workload is committed."""
    import glob

    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{name}_counters.json")))
    if not paths:
        return None, None
    with open(paths[-1]) as fh:
        return json.load(fh), os.path.relpath(paths[-1], ROOT)


def dense_reference_flop(post):
    """SURVEY section 8d dense-semantics FP32 flop per evaluation (every Doppler tap and instrument tap multiplied out)."""
    from baysar_b200.spectrometers import BalmerHydrogenLine

    dense = 0.0
    for ch in post.chords:
        F, P = len(ch.x_data_fm), len(ch.x_data)
        n_b = sum(isinstance(l, BalmerHydrogenLine) for l in ch.lines)
        gauss = [l for l in ch.lines if not isinstance(l, BalmerHydrogenLine)]
        n_g = sum(len(getattr(l, "lines", getattr(l, "cwl", []))) for l in gauss)
        k_d = max([len(l.wavelengths_doppler) for l in ch.lines if isinstance(l, BalmerHydrogenLine)] or [0])
        dense += 8.0 * n_g * F + n_b * 2 * (10.0 * F + 2.0 * k_d * F + 3.0 * F) + 12.0 * F + 10.0 * P \
            + 2.0 * F * len(ch.instrument_function_fm)
    return dense


KERNEL_OF_STAGE = {"los": "los_stage_kernel", "chord": "chord_stage_kernel", "contract": "contraction",
                   "pixel": "pixel_stage", "finalize": "finalize_kernel"}


def build_roofline(name, post, batch, per_call, step_ms_mean, peaks, measured):
    """Roofline object of the dominant kernel; every kernel's achieved rates under `kernels`."""
    counts, src = work_counts(name)
    fp32_peak, mufu_peak, fp64_peak = peaks["fp32_fma_flops"], peaks["mufu_ops"], peaks["fp64_fma_flops"]
    tf32_peak = measured.get("bf16_tflops", 1590.0) / 2.0 * 1e12
    kernels = {}
    for stage, ms in per_call.items():
        if ms <= 0:
            continue
        k = {"ms": ms, "share_of_step": ms / step_ms_mean}
        c = (counts or {}).get("kernels", {}).get(stage)
        if c:
            sec = ms * 1e-3
            k["counters"] = c
            rates = {"fp32": (c.get("fp32_flop_per_eval", 0.0) * batch / sec, fp32_peak),
                     "mufu": (c.get("mufu_per_eval", 0.0) * batch / sec, mufu_peak),
                     "xu": (c.get("xu_ops_per_eval", 0.0) * batch / sec, mufu_peak),  # all XU-pipe instructions
                     "fp64": (c.get("fp64_flop_per_eval", 0.0) * batch / sec, fp64_peak),
                     "tensor": (c.get("tensor_flop_per_eval", 0.0) * batch / sec, tf32_peak)}
            k["pipes"] = {p: {"achieved": a, "peak": pk, "frac": a / pk} for p, (a, pk) in rates.items() if a > 0}
            if k["pipes"]:
                bound = max(k["pipes"], key=lambda p: k["pipes"][p]["frac"])
                k.update(bound=bound, achieved=k["pipes"][bound]["achieved"], peak=k["pipes"][bound]["peak"],
                         frac=k["pipes"][bound]["frac"], unit="op/s" if bound in ("mufu", "xu") else "flop/s")
            if c.get("dram_bytes_per_launch") is not None:
                k["traffic"] = c["dram_bytes_per_launch"] * batch / max(c.get("thetas_per_launch", batch), 1)
        kernels[stage] = k
    dom = max(kernels, key=lambda s: kernels[s]["ms"])
    d = kernels[dom]
    dense = dense_reference_flop(post)
    n_params = post.plasma.n_params
    hbm_bytes = batch * (8.0 * n_params + 8.0 + 4.0 + 8.0 * (len(post.chords) + len(post.priors))
                         + 2 * 8.0 * 16 * post.model.n_ulines)
    if post.model.uses_tensor_cores:  # FP32 spectra and pixel columns, written and read once
        hbm_bytes += batch * sum(len(c.x_data_fm) * 8.0 + (len(c.x_data) if post.model.folded_contraction[i][0] else
                                                           len(c.x_data_fm)) * 8.0 for i, c in enumerate(post.chords))
    roof = {"kernel": KERNEL_OF_STAGE.get(dom, dom), "stage": dom,
            "bound": {"fp32": "fp32", "mufu": "mufu", "xu": "xu", "fp64": "fp64", "tensor": "tensor"}.get(d.get("bound"), None),
            "achieved": d.get("achieved"), "peak": d.get("peak"), "unit": d.get("unit"), "frac": d.get("frac"),
            "traffic": d.get("traffic"), "traffic_unit": "DRAM bytes read + written per launch (ncu --set full capture)",
            "counters_source": src,
            "peak_source": "FP32 FMA / MUFU / FP64 FMA micro-benchmarks of the tools library in this run; tensor = half of "
                           "MEASURED_PEAKS.json bf16_tflops (TF32 dense)",
            "kernels": kernels, "stage_ms": per_call, "kernel_share_of_step": d["share_of_step"],
            "dense_reference_flop_per_eval": dense,
            "dense_equivalent_tflops": dense * batch / (step_ms_mean * 1e-3) / 1e12,
            "fp32_peak_tflops": fp32_peak / 1e12, "mufu_peak_gops": mufu_peak / 1e9, "fp64_peak_tflops": fp64_peak / 1e12,
            "hbm_algorithmic_GBps": hbm_bytes / (step_ms_mean * 1e-3) / 1e9,
            "hbm_peak_GBps_measured": measured.get("hbm_gbs")}
    return roof


# ----------------------------------------------------------------------------------------------- one workload
class Env:
    """Process-wide state of a bench run (ranks, clocks sampler, L2 flush buffer, measured peaks)."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.sampler = None
        self.flush = None
        self.peaks = None


def parity_check(name, post, thetas, logp_dev, status_dev, n_logp, n_spectra, cores):
    """GPU log-posteriors (and spectra) of a timed batch against the oracle port on the same thetas, same process."""
    n_logp = min(n_logp, len(thetas))
    n_spectra = min(n_spectra, n_logp)
    _, _, _, ref_logp, ref_spec = cpu_oracle_rate(name, thetas[:n_logp], min(cores, 16), n_spectra)
    gpu_logp = logp_dev[:n_logp].cpu().numpy()
    status = status_dev[:n_logp].cpu().numpy()
    gpu_spec, _ = post.spectra_batch(thetas[:n_spectra], chord=0)
    ok = (status == 0) & np.isfinite(ref_logp)
    mismatch = int(((status == 0) != np.isfinite(ref_logp)).sum())  # oracle raised <-> status flag, must agree
    rel = np.abs(gpu_logp[ok] - ref_logp[ok]) / np.abs(ref_logp[ok])
    spec_rel, n_spec = 0.0, 0
    for i in range(n_spectra):
        if ok[i] and ref_spec[i] is not None:
            spec_rel = max(spec_rel, float(np.max(np.abs(gpu_spec[i] - ref_spec[i]) / np.abs(ref_spec[i]))))
            n_spec += 1
    out = {"n": int(ok.sum()), "n_spectra": n_spec, "logp_max_rel": float(rel.max()) if len(rel) else None,
           "spectra_max_rel": spec_rel if n_spec else None, "status_mismatch": mismatch,
           "logp_rtol": LOGP_RTOL, "spectra_rtol": SPECTRA_RTOL, "oracle": "oracle/numpy_port.py, same thetas, same process"}
    out["ok"] = bool(len(rel) >= n_logp // 2 and rel.max() <= LOGP_RTOL and mismatch == 0 and
                     (n_spec == 0 or spec_rel <= SPECTRA_RTOL))
    return out


def run_workload(env, name, batch, steps, warmup, sync_gather=False, scaling="weak", with_e2e=True, parity=(256, 16),
                 with_roofline=True, options=None):
    """One bench line (dict) of workload `name`: device-timed value, e2e, per-stage times, parity, roofline."""
    import torch
    import torch.distributed as dist

    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.lineshapes import EsymmtricCauchyPlasma
    from baysar_b200.parallel import ShardedEvaluator
    from baysar_b200.posterior import BaysarPosterior

    rank, world = env.rank, env.world
    wl = get_workload(name)
    np.random.seed(0)
    prof = EsymmtricCauchyPlasma(x=wl.los) if wl.los is not None else None
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=prof, atomic_data=SyntheticADAS(),
                           device=env.local_rank, skip_test=True, options=options)
    n_params = post.plasma.n_params
    total_steps = warmup + steps
    # a different theta batch every step; this rank's shard of the global batch
    thetas_host = draw_thetas(post.plasma.theta_bounds, batch * total_steps, 20261019 + rank).reshape(total_steps, batch, n_params)
    thetas_dev = torch.as_tensor(thetas_host, device="cuda")
    sharded = ShardedEvaluator(post.logp_batch)  # world 1: plain call; world > 1: + NCCL all-gather of logP
    # N > 1: the all-gather of a step's logP (8 B per theta) runs on a side stream and overlaps with the next step: the
    # ranks' theta shards are independent walkers, so nothing on the critical path waits for it (sync_gather puts it back
    # on the compute stream).  The tail of the last gather is added to the timed total below.
    overlap = world > 1 and not sync_gather

    def step(i):
        return sharded.evaluate_local(thetas_dev[i], overlap=overlap)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    # no interpreter garbage collection inside the timed loop: a pause of one rank's host thread stalls that rank's
    # launches and, through the all-gather, every rank's next step.  Collected BEFORE the barrier: rank 0 holds the larger
    # heap (oracle, parity arrays), and a rank that enters the timed loop late makes every other rank's gathers wait for it
    gc.collect()
    gc.disable()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall0 = time.time()
    for k in range(steps):
        env.flush.fill_(float(k))  # evict L2 between timed steps (untimed)
        starts[k].record()
        step(warmup + k)
        ends[k].record()
    gc.enable()
    tail0, tail1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tail0.record()
    sharded.wait()      # outstanding overlapped all-gathers
    tail1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    # per-stage times from a separate, untimed pass over the same batches: the events recorded between the kernels of a
    # profiled evaluation are stream operations of their own and keep the kernels from overlapping their launch
    # (programmatic dependent launch), so the timed steps above run without them
    post.model.set_profiling(True)
    for i in range(min(warmup, 2)):  # the event pool fills
        step(i)
    torch.cuda.synchronize()
    post.model.get_profile()
    for k in range(min(steps, 20)):
        env.flush.fill_(float(k))
        step(warmup + k)
    sharded.wait()
    torch.cuda.synchronize()
    stage_ms, calls = post.model.get_profile()
    post.model.set_profiling(False)
    if world > 1:
        dist.barrier()
    launches = post.model.launches_last_eval
    step_ms = np.array([s.elapsed_time(e) for s, e in zip(starts, ends)])
    if os.environ.get("BAYSAR_BENCH_RANK_STATS"):  # diagnostics: every rank's slowest steps
        worst = np.argsort(step_ms)[-3:]
        print(f"[rank {rank}] {name} steps {steps}: median {np.median(step_ms):.4f} ms, slowest "
              + ", ".join(f"#{i} {step_ms[i]:.3f}" for i in worst), file=sys.stderr, flush=True)
    total_ms = torch.tensor([float(step_ms.sum()) + float(tail0.elapsed_time(tail1))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * batch * steps / (total_ms * 1e-3)

    # ---- end-to-end: reference-facing call with pinned HOST buffers (H2D + D2H inside the timed region)
    e2e = None
    if with_e2e:
        pinned = torch.empty((batch, n_params), dtype=torch.float64).pin_memory()
        e2e_steps = max(3, min(steps, 10))
        for i in range(2):
            pinned.copy_(torch.from_numpy(thetas_host[i]))
            post(pinned.numpy())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_e2e = 0.0
        gc.collect()
        gc.disable()
        for k in range(e2e_steps):
            pinned.copy_(torch.from_numpy(thetas_host[(warmup + k) % total_steps]))
            env.flush.fill_(float(k))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            post(pinned.numpy())  # numpy in, numpy out: copies happen inside baysar_eval_logp_host
            t_e2e += time.perf_counter() - t0
        gc.enable()
        t_e2e_t = torch.tensor([t_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e2e_t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * batch * e2e_steps / float(t_e2e_t.item()), "unit": UNIT,
               "h2d_bytes_per_step": batch * n_params * 8, "d2h_bytes_per_step": batch * (8 + 4), "steps": e2e_steps}

    # ---- the first timed batch once more, for the parity block and the status statistics (same kernels, same stream)
    logp_dev, status_dev = post.logp_batch(thetas_dev[warmup], return_status=True)
    torch.cuda.synchronize()
    if rank != 0:
        return None
    status = status_dev.cpu().numpy()
    per_call = {k: v / max(calls, 1) for k, v in stage_ms.items()}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32 element-wise / f64 line-of-sight, reductions and likelihood", "data": "synthetic",
            "config": workload_config(name, batch),
            "implementation": {"tensor_core_contraction": post.model.uses_tensor_cores,
                               "logp_all_gather": ("none (1 GPU)" if world == 1 else
                                                   "NCCL, side stream, overlapped with the next step" if overlap else
                                                   "NCCL on the compute stream"),
                               "options": post.model.options},
            "clocks": env.sampler.window(t_wall0, t_wall1) if env.sampler else None, "e2e": e2e,
            "gpu_launches": launches * steps, "stage_ms": per_call, "status_ok_fraction": float((status == 0).mean()),
            "step_ms": {"median": float(np.median(step_ms)), "min": float(step_ms.min()), "max": float(step_ms.max()),
                        "gather_tail_ms": float(tail0.elapsed_time(tail1))}}
    if parity and parity[0] > 0:
        line["parity"] = parity_check(name, post, thetas_host[warmup], logp_dev, status_dev, parity[0], parity[1], usable_cores())
    if with_roofline:
        measured = {}
        try:
            measured = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        line["roofline"] = build_roofline(name, post, batch, per_call, float(step_ms.mean()), env.peaks, measured)
    line["_thetas_pool"] = thetas_host.reshape(-1, n_params)
    return line


def run_ensemble(env, walkers_per_gpu, steps, warmup, chords=32):
    """BASELINE configs[3]: the 32-chord divertor set under the on-device affine-invariant ensemble (stretch move), the
    walkers sharded over the ranks as ONE ensemble: every half-step all-gathers the complementary half's positions over
    NCCL, proposes, evaluates the batched posterior on this rank's half and accepts.  One step = two half-steps = one
    posterior evaluation per walker; value = posterior evals/s over all ranks (device time, max over ranks)."""
    import torch
    import torch.distributed as dist

    import baysar_b200.lineshapes as host_ns
    from baysar_b200 import configs
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.ensemble import StretchEnsemble
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.posterior import BaysarPosterior

    rank, world = env.rank, env.world
    wl = configs.multi_species(chords)
    np.random.seed(0)
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                           atomic_data=SyntheticADAS(), device=env.local_rank, skip_test=True)
    centre = wl.theta_from_tags(post.plasma.slices, post.random_start())
    tb = np.asarray(post.plasma.theta_bounds, dtype=float)
    rng = np.random.default_rng(20261019 + 4 + rank)
    start = centre + 1e-3 * (tb[:, 1] - tb[:, 0]) * rng.standard_normal((walkers_per_gpu, len(centre)))
    ens = StretchEnsemble(post, start, seed=1234)  # one seed: the draws are counted by global walker index
    for _ in range(warmup):
        ens.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(steps):
        ens.step()
    allp = ens.gather_logp()
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # the stored log-posteriors are those of the stored walkers (same kernels: bit-identical), a cheap in-run consistency check
    again = post.logp_batch(ens.walkers[:4096])
    consistent = bool(torch.equal(again, ens.logp[:4096]))
    if rank != 0:
        return None
    total_ms = float(ms.item())
    n_params = int(post.plasma.n_params)
    return {"metric": METRIC, "value": world * walkers_per_gpu * steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 element-wise / f64 line-of-sight, reductions and likelihood", "data": "synthetic",
            "config": {"workload": "ensemble_32chords", "chords": chords, "walkers_per_gpu": walkers_per_gpu,
                       "walkers": world * walkers_per_gpu, "n_params": n_params,
                       "l2_between_steps": f"inputs larger than L2 ({walkers_per_gpu // 2 * chords * 12 // 1024} MB of spectra per half-step)"},
            "implementation": {"ensemble": "one global ensemble: complementary half all-gathered over NCCL every half-step"
                               if world > 1 else "one ensemble on one GPU",
                               "complement_all_gather_bytes_per_half_step": world * (walkers_per_gpu // 2) * n_params * 8},
            "clocks": env.sampler.window(t_wall0, t_wall1) if env.sampler else None,
            "gpu_launches": steps * 2 * (2 + post.model.launches_last_eval),
            "acceptance_fraction": ens.acceptance_fraction, "stored_logp_bit_identical": consistent,
            "mean_logp": float(allp.mean().item())}


# ----------------------------------------------------------------------------------------------- main
def reference_arm(args, batch):
    """`--impl reference`: the reference is pure Python and cannot travel to the GPU box; its CPU implementation of the
    path is timed through the oracle port (bit-identical to the shimmed reference on the golden vectors)."""
    cores = usable_cores()
    np.random.seed(0)
    bounds = _oracle_posterior(args.workload).theta_bounds
    per_step = max(cores * 32, 64)
    thetas = draw_thetas(bounds, per_step * (args.steps + args.warmup), 20261019)
    for w in range(args.warmup):
        cpu_oracle_rate(args.workload, thetas[w * per_step:(w + 1) * per_step], cores)
    t_total, done = 0.0, 0
    for s in range(args.warmup, args.warmup + args.steps):
        rate, n_done, _, _, _ = cpu_oracle_rate(args.workload, thetas[s * per_step:(s + 1) * per_step], cores)
        t_total += n_done / rate
        done += n_done
    value = done / t_total
    sample = f"{per_step} thetas per step x {args.steps} steps over {cores} forked workers"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
                      "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": workload_config(args.workload, batch),
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="balmer_series")
    ap.add_argument("--batch", type=int, default=0, help="thetas per GPU per step (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: the workload's batch is the TOTAL over all GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="only the main workload's line")
    ap.add_argument("--ensemble-walkers", type=int, default=131072, help="walkers per GPU of the ensemble_32chords line")
    ap.add_argument("--sync-gather", action="store_true", help="N > 1: logP all-gather on the compute stream")
    ap.add_argument("--options", default="", help="baysar_options_t fields of the main workload's model, key=value,... "
                                                  "(kernel-selection experiments; default: the library's choices)")
    args = ap.parse_args()

    env = Env()
    batch = args.batch or BATCH_PER_GPU.get(args.workload, args.ensemble_walkers)
    if args.scaling == "strong":
        batch = max(batch // env.world, 1)

    if args.impl == "reference":
        if env.rank == 0:
            reference_arm(args, batch)
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(env.local_rank)
    if env.world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{env.local_rank}"))
    from baysar_b200 import runtime

    if env.rank == 0:
        env.sampler = ClockSampler(env.local_rank)
        env.sampler.start()
    env.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    env.peaks = runtime.microbench(env.local_rank) if env.rank == 0 else None
    parity = None if args.no_parity else (256, 16)

    if args.workload == "ensemble_32chords":
        line = run_ensemble(env, args.ensemble_walkers, args.steps, max(args.warmup, 1))
        if env.rank == 0:
            env.sampler.stop()
            print(json.dumps(line))
        if env.world > 1:
            dist.destroy_process_group()
        return

    options = {k: int(v) for k, v in (kv.split("=") for kv in args.options.split(",") if kv)} or None
    line = run_workload(env, args.workload, batch, args.steps, args.warmup, sync_gather=args.sync_gather,
                        scaling=args.scaling, parity=parity, options=options)

    # ---- further workloads (same shape of line, shorter runs, no CPU baseline)
    extras = []
    if not args.no_extra and args.workload == "balmer_series":
        short = max(5, min(args.steps, 20))
        small_parity = None if args.no_parity else (64, 4)
        plan = []
        if env.world == 1:
            plan = [("multi_species", BATCH_PER_GPU["multi_species"], "weak", False),
                    ("balmer_hot", BATCH_PER_GPU["balmer_hot"], "weak", False)]
        else:
            plan = [("balmer_series", batch, args.scaling, True),                       # gather on the compute stream
                    ("multi_species", BATCH_PER_GPU["multi_species"], "weak", False),
                    ("multi_species", BATCH_PER_GPU["multi_species"] // env.world, "strong", False),
                    ("multi_species", BATCH_PER_GPU["multi_species"] // env.world, "strong", True)]
        for name, b, scaling, sync in plan:
            res = run_workload(env, name, b, short, 3, sync_gather=sync, scaling=scaling, with_e2e=(env.world == 1),
                               parity=small_parity, with_roofline=(env.world == 1))
            if res is not None:
                res.pop("_thetas_pool", None)
                extras.append(res)
        # BASELINE configs[3]: 32 chords, 2^20 walkers over 8 GPUs = 131072 per GPU (weak), one global ensemble
        res = run_ensemble(env, args.ensemble_walkers, 3, 1)
        if res is not None:
            extras.append(res)

    if env.rank != 0:
        if env.world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (oracle port) on this box's host cores, bounded sample
    pool = line.pop("_thetas_pool")
    cpu = None
    if not args.no_cpu_baseline:
        cores = usable_cores()
        rate1, _, _, _, _ = cpu_oracle_rate(args.workload, pool[:200], 1)
        # bounded sample: about 12 s of CPU work per worker at the single-core rate just measured
        want = int(cores * max(200, 12.0 * rate1))
        sample = pool[np.arange(want) % len(pool)]
        rate, done, wall, _, _ = cpu_oracle_rate(args.workload, sample, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{done} thetas of the same batch, {cores} forked workers, {wall:.1f} s wall",
               "single_core_value": rate1,
               "port_vs_reference": "single core, build container (tools/port_vs_reference.py -> "
                                    "profiles/r2_port_vs_reference.jsonl): the port runs at 1.0-1.1x (Balmer demo) / 1.4-1.8x "
                                    "(multi-species) the unmodified reference's rate under the import shims"}
    line["cpu_baseline"] = cpu
    line["extra_workloads"] = extras
    env.sampler.stop()
    print(json.dumps(line))
    if env.world > 1:
        dist.destroy_process_group()
    bad = [l["config"]["workload"] for l in [line] + extras if l.get("parity") and not l["parity"]["ok"]]
    if bad:
        print(f"PARITY FAILURE against the oracle on: {bad}", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
