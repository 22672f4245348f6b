"""GPU unit test of the folded pixel contraction (csrc/pix_contract.cuh): instrument convolution + static wavelength
map as one tcgen05 contraction, against a float64 direct convolution (fftconvolve "same" alignment) followed by the
two-tap interpolation of interp1d (spectrometers.py:269-310)."""
import numpy as np
import pytest

from tests.pxc_cases import CASES, make_case
from tests.test_gpu_if_contract import reference_same_conv

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("F,K,sigma,P", CASES)
@pytest.mark.parametrize("tile_n", [128, 64])
def test_pix_contract_matches_float64(F, K, sigma, P, tile_n):
    import torch

    from baysar_b200 import runtime

    s, taps, idx, frac = make_case(F, K, sigma, P)
    sd = torch.as_tensor(s, dtype=torch.float32, device="cuda")
    res = runtime.debug_pix_contract(sd, taps, idx, frac, tile_n)
    out = res["out"].double()
    conv = reference_same_conv(sd.double().cpu().numpy(), taps.astype(np.float32).astype(np.float64))
    it = torch.as_tensor(idx.astype(np.int64), device="cuda")
    ft = torch.as_tensor(frac.astype(np.float32).astype(np.float64), device="cuda")
    pix_ref = conv[:, it] * (1 - ft) + conv[:, it + 1] * ft
    col = torch.as_tensor(res["pix_col"].astype(np.int64), device="cuda")
    sel = col >= 0
    err = (out[:, col[sel]] - pix_ref[:, sel]) / pix_ref[:, sel].abs()
    xl, xr = res["x_l"], res["x_r"]
    err_l = (out[:, res["col_left0"]:res["col_left0"] + xl] - conv[:, :xl]) / conv[:, :xl].abs()
    err_r = (out[:, res["col_right0"]:res["col_right0"] + xr] - conv[:, F - xr:]) / conv[:, F - xr:].abs()
    # per-tile bias correction of pixel_fold_kernel: kappa = exact tile sum (converter threads) / sum of the tile's columns
    n_fold = res["col_left0"] // tile_n
    tiles = out[:, :n_fold * tile_n].reshape(out.shape[0], n_fold, tile_n)
    kappa = res["tsum"][:, :n_fold] / tiles.sum(dim=2)
    fixed = (tiles * kappa[:, :, None]).reshape(out.shape[0], -1)
    err_k = (fixed[:, col[sel]] - pix_ref[:, sel]) / pix_ref[:, sel].abs()
    print(f"F={F} K={K} P={P} TN={tile_n}: pixels max rel {err.abs().max().item():.2e} mean signed {err.mean().item():.2e}; "
          f"with kappa: max {err_k.abs().max().item():.2e} mean {err_k.mean().item():.2e} std {err_k.std().item():.2e}; "
          f"edges {max(err_l.abs().max().item(), err_r.abs().max().item()):.2e}; kernel {res['ms']:.3f} ms, "
          f"{res['k_blocks']} k-blocks per theta block, folded pixels {int(sel.sum())}/{P}")
    assert int(sel.sum()) > 0.8 * P
    # TMEM accumulation truncates: the raw error is one-sided and shrinks with the number of hi*hi accumulators
    assert err.abs().max().item() < 6e-6 and abs(err.mean().item()) < (2e-6 if tile_n == 128 else 1e-6)
    assert err_l.abs().max().item() < 6e-6 and err_r.abs().max().item() < 6e-6
    assert (kappa - 1).abs().max().item() < 1e-5
    assert err_k.abs().max().item() < 4e-6 and abs(err_k.mean().item()) < 2e-7


def test_pix_contract_is_bit_reproducible():
    """Same input, same schedule, same accumulation order: outputs and exact tile sums must not change run to run
    (a timing-dependent hazard in the A pipeline shows up here as stale rows)."""
    import torch

    from baysar_b200 import runtime

    F, K, sigma, P = CASES[0]
    s, taps, idx, frac = make_case(F, K, sigma, P)
    s = np.tile(s, (8, 1)) * (1.0 + 1e-3 * np.arange(2048)[:, None] / 2048)
    sd = torch.as_tensor(s, dtype=torch.float32, device="cuda")
    for tile_n in (128, 64):
        ref = runtime.debug_pix_contract(sd, taps, idx, frac, tile_n)
        for _ in range(4):
            res = runtime.debug_pix_contract(sd, taps, idx, frac, tile_n)
            assert torch.equal(res["out"], ref["out"])
            assert torch.equal(res["tsum_groups"], ref["tsum_groups"])
