#!/usr/bin/env python
"""LOS-stage time per theta for line-of-sight lengths around the warp-per-theta / team switch, default choice against
the forced team form (options.los_team_points = 1).   python tools/los_team_probe.py [L ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import baysar_b200.lineshapes as host_ns
    from baysar_b200 import configs
    from baysar_b200.atomic_data import SyntheticADAS
    from baysar_b200.input_functions import make_input_dict
    from baysar_b200.posterior import BaysarPosterior
    from bench import draw_thetas

    for L in [int(a) for a in sys.argv[1:]] or [200, 300, 500, 700, 1000]:
        wl = configs.synthetic_sweep(L, 20, 1024)
        variants = [None, {"los_team_warps": 2}, {"los_team_warps": 4, "los_min_blocks": 4}, {"los_team_warps": 4, "los_min_blocks": 5},
                    {"los_team_warps": 8, "los_min_blocks": 2}]
        for opts in variants:
            np.random.seed(0)
            post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=wl.make_profile(host_ns),
                                   atomic_data=SyntheticADAS(), device=0, skip_test=True, options=opts)
            n = 32768
            th = torch.as_tensor(draw_thetas(post.plasma.theta_bounds, n, 2), device="cuda")
            post.logp_batch(th)
            post.model.set_profiling(True)
            for _ in range(3):
                post.logp_batch(th)
            torch.cuda.synchronize()
            stage_ms, calls = post.model.get_profile()
            print(json.dumps({"L": L, "options": opts, "los_ms_per_32768": stage_ms["los"] / calls}), flush=True)
            post.model.close()


if __name__ == "__main__":
    main()
