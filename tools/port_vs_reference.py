#!/usr/bin/env python
"""Single-core speed of the oracle port next to the unmodified reference (under the import shims), on one machine.

    python tools/port_vs_reference.py [balmer_series|multi_species ...]        # build container only (/root/reference)

bench.py's CPU baseline and `--impl reference` arm time the PORT (`oracle/numpy_port.py`), because the reference cannot
travel to the GPU box.  This script states how the two relate: same thetas, same tables, one process, one thread.
"""
import json
import os
import pathlib
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")
import numpy as np  # noqa: E402

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
from oracle import ref_harness, ref_shim  # noqa: E402


def main():
    names = sys.argv[1:] or ["balmer_series", "multi_species"]
    for name in names:
        wl = bench.get_workload(name)
        ref = ref_harness.build_reference_posterior(wl)
        port = bench._oracle_posterior(name)
        thetas = bench.draw_thetas(np.asarray(port.theta_bounds, dtype=float), 120, 20261019)
        rates = {}
        for label, fn in (("reference", lambda th: ref(list(th), True)), ("port", lambda th: port.evaluate(th))):
            done, t0 = 0, None
            for i, th in enumerate(thetas):
                if i == 20:
                    t0 = time.perf_counter()  # first 20: warm-up
                try:
                    with ref_shim.quiet():
                        fn(th)
                except Exception:
                    pass
                done += i >= 20
            rates[label] = done / (time.perf_counter() - t0)
        print(json.dumps({"workload": name, "reference_evals_per_s": rates["reference"], "port_evals_per_s": rates["port"],
                          "port_over_reference": rates["port"] / rates["reference"], "thetas": 100, "threads": 1}))


if __name__ == "__main__":
    main()
