#!/usr/bin/env python
"""Shape sweep / roofline report (BASELINE.json configs[4], SURVEY.md section 8d cfg 5).

    python tools/shape_sweep.py [--full] [--out profiles/rNN_shape_sweep.jsonl]

Synthetic chords (`baysar_b200.configs.synthetic_sweep`): L line-of-sight points x n_G single-component impurity
lines (+ 2 Balmer lines) x P pixels at 0.15 A per pixel, refine 0.05 (F ~ 3 P).  By default one axis is varied at a
time around (L, n_G, P) = (100, 20, 1024); `--full` walks the whole 5 x 5 x 4 grid.  For every shape: log-posterior
evals/s on one GPU (thetas resident in HBM, CUDA events, theta batches sized so that a step takes >= 50 ms), the time
per stage (the pipe fractions of the stage kernels come from ncu opcode counters, tools/ncu_counters.py, for the two
bench workloads).  One JSON line per shape.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_shape(L, G, P, balmer, peaks, target_ms=50.0):
    import torch

    import baysar_b200.lineshapes as host_ns
    from baysar_b200 import configs
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.posterior import BaysarPosterior
    from bench import draw_thetas

    wl = configs.synthetic_sweep(L, G, P, balmer=balmer)
    np.random.seed(0)
    t0 = time.time()
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                           atomic_data=SyntheticADAS(), device=0, skip_test=True)
    build_s = time.time() - t0
    n_params = post.plasma.n_params
    # size the batch from a probe step
    probe = torch.as_tensor(draw_thetas(post.plasma.theta_bounds, 1024, 1), device="cuda")
    post.logp_batch(probe)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    post.logp_batch(probe)
    e1.record()
    torch.cuda.synchronize()
    per_theta_ms = e0.elapsed_time(e1) / 1024
    n = int(min(max(1024, 2 ** int(np.ceil(np.log2(target_ms / per_theta_ms)))), 262144))
    steps = 4
    thetas = torch.as_tensor(draw_thetas(post.plasma.theta_bounds, n * (steps + 1), 2).reshape(steps + 1, n, n_params),
                             device="cuda")
    post.logp_batch(thetas[0])
    torch.cuda.synchronize()
    post.model.set_profiling(True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        ev[k][0].record()
        post.logp_batch(thetas[k + 1])
        ev[k][1].record()
    torch.cuda.synchronize()
    stage_ms, calls = post.model.get_profile()
    post.model.set_profiling(False)
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    status = post.last_status_device.cpu().numpy()
    out = {"L": L, "n_G": G, "P": P, "F": len(post.chords[0].x_data_fm), "n_B": 2 if balmer else 0, "n_params": n_params,
           "thetas_per_step": n, "ms_per_step": ms, "evals_per_s": n / (ms * 1e-3),
           "stage_ms": {k: v / max(calls, 1) for k, v in stage_ms.items()},
           "tensor_core_contraction": post.model.uses_tensor_cores,
           "status_ok_fraction": float((status == 0).mean()), "model_build_s": build_s}
    post.model.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    from baysar_b200 import runtime

    peaks = runtime.microbench(0)
    Ls, Gs, Ps = [50, 100, 200, 500, 1000], [10, 20, 50, 100, 200], [512, 1024, 2048, 4096]
    if args.full:
        shapes = [(L, G, P, True) for L in Ls for G in Gs for P in Ps]
    else:
        shapes = [(L, 20, 1024, True) for L in Ls] + [(100, G, 1024, True) for G in Gs if G != 20] + \
                 [(100, 20, P, True) for P in Ps if P != 1024] + [(100, 20, 1024, False), (1000, 200, 4096, True)]
    sink = open(args.out, "w") if args.out else None
    for L, G, P, b in shapes:
        rec = run_shape(L, G, P, b, peaks)
        line = json.dumps(rec)
        print(line, flush=True)
        if sink:
            sink.write(line + "\n")
            sink.flush()


if __name__ == "__main__":
    main()
