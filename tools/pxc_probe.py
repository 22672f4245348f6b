"""GPU probe of the folded pixel contraction at bench size: accuracy statistics and kernel time per tile width."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from baysar_b200 import runtime  # noqa: E402
from tests.pxc_cases import make_case  # noqa: E402
from tests.test_gpu_if_contract import reference_same_conv  # noqa: E402


def main(n=4096, F=8002, K=8001, P=1024, sigma=25.0):
    s, taps, idx, frac = make_case(F, K, sigma, P)
    s = np.tile(s, (n // 256, 1))
    sd = torch.as_tensor(s, dtype=torch.float32, device="cuda")
    conv = reference_same_conv(sd[:256].double().cpu().numpy(), taps.astype(np.float32).astype(np.float64))
    it = torch.as_tensor(idx.astype(np.int64), device="cuda")
    ft = torch.as_tensor(frac.astype(np.float32).astype(np.float64), device="cuda")
    pix_ref = conv[:, it] * (1 - ft) + conv[:, it + 1] * ft
    for tn in (128, 64):
        res = runtime.debug_pix_contract(sd, taps, idx, frac, tn)
        col = torch.as_tensor(res["pix_col"].astype(np.int64), device="cuda")
        sel = col >= 0
        out = res["out"][:256].double()
        err = (out[:, col[sel]] - pix_ref[:, sel]) / pix_ref[:, sel].abs()
        n_fold = res["col_left0"] // tn
        tiles = out[:, :n_fold * tn].reshape(256, n_fold, tn)
        kappa = res["tsum"][:256, :n_fold] / tiles.sum(dim=2)
        err_k = ((tiles * kappa[:, :, None]).reshape(256, -1)[:, col[sel]] - pix_ref[:, sel]) / pix_ref[:, sel].abs()
        print(f"TN={tn}: kernel {res['ms']:.4f} ms for {n} rows ({res['k_blocks']} k-blocks per theta block, ring {res['ring']}); "
              f"signed rel err mean {err.mean().item():.3e} std {err.std().item():.3e} min {err.min().item():.3e} "
              f"max {err.max().item():.3e}; with kappa mean {err_k.mean().item():.3e} std {err_k.std().item():.3e} "
              f"max|.| {err_k.abs().max().item():.3e}", flush=True)


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:5]])
