#!/usr/bin/env python
"""BASELINE.json configs[3]: 32-chord divertor spectrometer set, affine-invariant ensemble sharded over the GPUs.

    python tools/ensemble_bench.py [--walkers-per-gpu 131072] [--steps 3] [--chords 32]
    torchrun --nproc-per-node N tools/ensemble_bench.py ...        (one rank per GPU, 2^20 walkers over 8 GPUs)

Walkers start in a Gaussian ball (sigma = 1e-3 of each parameter's range) around a reference theta and advance with the
on-device stretch move (a = 2, red/blue halves): per half-step one proposal kernel, the batched posterior and one accept
kernel; after every step the per-walker log-posteriors are all-gathered over NCCL (what a convergence monitor reads).
Prints one JSON line: walker-steps per second over all ranks (max over ranks of the device time), acceptance fraction.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers-per-gpu", type=int, default=131072)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--chords", type=int, default=32)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import baysar_b200.lineshapes as host_ns
    from baysar_b200 import configs
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.ensemble import StretchEnsemble
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.posterior import BaysarPosterior

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = configs.multi_species(args.chords)
    np.random.seed(0)
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                           atomic_data=SyntheticADAS(), device=local, skip_test=True)
    centre = wl.theta_from_tags(post.plasma.slices, post.random_start())
    bounds = post.plasma.theta_bounds
    rng = np.random.default_rng(20261019 + 4 + rank)
    walkers = centre + 1e-3 * (bounds[:, 1] - bounds[:, 0]) * rng.standard_normal((args.walkers_per_gpu, len(centre)))
    ens = StretchEnsemble(post, walkers, seed=1234 + rank)
    for _ in range(args.warmup):
        ens.step()
        ens.gather_logp()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ens.step()
        allp = ens.gather_logp()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        total = world * args.walkers_per_gpu
        print(json.dumps({"metric": "ensemble walker-steps/sec (stretch move, 2 posterior evals of half the walkers per step)",
                          "value": total * args.steps / (float(ms.item()) * 1e-3), "unit": "walker-steps/s",
                          "n_gpus": world, "walkers": total, "chords": args.chords, "n_params": int(post.plasma.n_params),
                          "steps": args.steps, "ms_per_step": float(ms.item()) / args.steps,
                          "acceptance_fraction": ens.acceptance_fraction, "gathered_logp": int(allp.numel()),
                          "mean_logp": float(allp.mean().item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
