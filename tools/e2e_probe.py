#!/usr/bin/env python
"""Where the end-to-end call spends its time beyond the kernels: `posterior(theta_host)` against the bare C call with
preallocated arrays and against the device-resident step (CUDA events).  python tools/e2e_probe.py [workload] [batch]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import draw_thetas, get_workload  # noqa: E402

from baysar_b200.atomic_data import SyntheticADAS  # noqa: E402
from baysar_b200.input_functions import make_input_dict  # noqa: E402
from baysar_b200.lineshapes import EsymmtricCauchyPlasma  # noqa: E402
from baysar_b200.posterior import BaysarPosterior  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "balmer_series"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    wl = get_workload(name)
    np.random.seed(0)
    prof = EsymmtricCauchyPlasma(x=wl.los) if wl.los is not None else None
    post = BaysarPosterior(make_input_dict(**wl.input_kwargs), profile_function=prof, atomic_data=SyntheticADAS(), device=0,
                           skip_test=True)
    n_params = post.plasma.n_params
    th = draw_thetas(post.plasma.theta_bounds, batch, 1).reshape(batch, n_params)
    pinned = torch.empty((batch, n_params), dtype=torch.float64).pin_memory()
    pinned.copy_(torch.from_numpy(th))
    arr = pinned.numpy()
    dev = torch.as_tensor(th, device="cuda")
    reps = 200 if batch <= 8192 else 30

    def timed(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6

    m = post.model
    logp = np.empty(batch)
    status = np.empty(batch, dtype=np.uint32)
    a_ptr, l_ptr, s_ptr = arr.ctypes.data, logp.ctypes.data, status.ctypes.data
    lib, handle = m.lib, m.handle
    out = {
        "posterior(theta_host) us": timed(lambda: post(arr)),
        "model.eval_logp_host us": timed(lambda: m.eval_logp_host(arr)),
        "bare C call us": timed(lambda: lib.baysar_eval_logp_host(handle, a_ptr, batch, l_ptr, s_ptr)),
        "device-resident logp_batch us (back to back, no host sync)": timed(lambda: post.logp_batch(dev)),
    }
    host, devv = post(arr), post.logp_batch(dev).cpu().numpy()
    same = np.array_equal(host, devv, equal_nan=True)
    print(f"host path == device path bit for bit on {batch} thetas: {same}")
    if not same:
        raise SystemExit(1)
    for k, v in out.items():
        print(f"{k:62s} {v:9.1f}")


if __name__ == "__main__":
    main()
