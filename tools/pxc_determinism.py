"""GPU check: the folded contraction must be bit-reproducible run to run (same schedule, same accumulation order)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from baysar_b200 import runtime  # noqa: E402
from tests.pxc_cases import make_case  # noqa: E402


def main(n=4096, F=8002, K=8001, P=1024, sigma=25.0, reps=12):
    s, taps, idx, frac = make_case(F, K, sigma, P)
    s = np.tile(s, (n // 256, 1)) * (1.0 + 1e-3 * np.arange(n)[:, None] / n)
    sd = torch.as_tensor(s, dtype=torch.float32, device="cuda")
    for tn in (128, 64):
        ref = runtime.debug_pix_contract(sd, taps, idx, frac, tn)
        for rep in range(reps):
            res = runtime.debug_pix_contract(sd, taps, idx, frac, tn)
            bad = (res["out"] != ref["out"]).nonzero()
            bad_t = (res["tsum"] != ref["tsum"]).nonzero()
            msg = f"TN={tn} rep {rep}: {len(bad)} output elements differ, {len(bad_t)} tile sums differ"
            if len(bad):
                rows, cols = bad[:, 0], bad[:, 1]
                rel = ((res["out"][rows, cols] - ref["out"][rows, cols]) / ref["out"][rows, cols]).abs().max().item()
                msg += (f"; rows {sorted(set((rows // 128).tolist()))[:8]} (theta blocks), row-in-block "
                        f"{sorted(set((rows % 128).tolist()))[:8]}..., column tiles {sorted(set((cols // tn).tolist()))[:8]}, "
                        f"col-in-tile n={len(set((cols % tn).tolist()))}, max rel {rel:.2e}")
            if len(bad_t):
                r_, t_ = bad_t[:, 0], bad_t[:, 1]
                relt = ((res["tsum"][r_, t_] - ref["tsum"][r_, t_]) / ref["tsum"][r_, t_]).abs()
                g_new, g_ref = res["tsum_groups"][r_, t_], ref["tsum_groups"][r_, t_]
                msg += f"; groups new {g_new.tolist()[:3]} ref {g_ref.tolist()[:3]}"
                msg += (f"; tsum: theta blocks {(r_ // 128).tolist()[:8]} rows-in-block {(r_ % 128).tolist()[:8]} tiles "
                        f"{t_.tolist()[:8]} rel {[f'{v:.1e}' for v in relt.tolist()[:8]]}")
            if len(bad) or len(bad_t) or rep == reps - 1:
                print(msg, flush=True)


if __name__ == "__main__":
    main()
