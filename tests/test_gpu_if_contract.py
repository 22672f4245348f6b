"""GPU unit test of the tcgen05 instrument-function contraction (csrc/if_contract.cuh) against a float64 direct
convolution with scipy.signal.fftconvolve's "same" alignment, out[j] = full[j + (K - 1) // 2]."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def reference_same_conv(s, w):
    import torch

    K = len(w)
    o = (K - 1) // 2
    x = torch.as_tensor(s, dtype=torch.float64, device="cuda")[:, None, :]
    k = torch.as_tensor(np.asarray(w, dtype=np.float64)[::-1].copy(), device="cuda")[None, None, :]
    full = torch.nn.functional.conv1d(x, k, padding=K - 1)[:, 0]  # full[i] = sum_k w[k] s[i - k]
    return full[:, o:o + s.shape[1]]


@pytest.mark.parametrize("F,K,sigma", [(8002, 8001, 25.0), (3002, 88, 9.6), (2984, 180, 12.0), (1000, 31, 2.0)])
def test_if_contract_matches_float64_convolution(F, K, sigma):
    import torch

    from baysar_b200 import runtime

    rng = np.random.default_rng(F + K)
    n = 256
    # spectra with the dynamic range of the real path: smooth wings + narrow strong lines, all positive
    x = np.arange(F)
    s = 1e9 * (1 + rng.random((n, F)))
    for _ in range(6):
        c, wdt, amp = rng.uniform(0, F, n), rng.uniform(0.3, 6, n), 10 ** rng.uniform(12, 16, n)
        s += amp[:, None] / (1 + np.abs((x[None] - c[:, None]) / wdt[:, None]) ** 2.5)
    taps = np.exp(-0.5 * ((np.arange(K) - (K - 1) / 2 + 0.3) / sigma) ** 2)
    taps[taps < 1e-300] = 0.0
    taps /= taps.sum()
    sd = torch.as_tensor(s, dtype=torch.float32, device="cuda")
    row_min = sd.min(dim=1).values.contiguous()
    out, part, ms = runtime.debug_if_contract(sd, taps, row_min)
    ref = torch.maximum(reference_same_conv(sd.double().cpu().numpy(), taps), row_min.double()[:, None])
    err = (out.double() - ref) / ref.abs()
    rel = err.abs().max().item()
    print(f"F={F} K={K}: max rel err {rel:.2e}, mean signed {err.mean().item():.2e}, kernel {ms:.3f} ms for {n} rows")
    # TMEM accumulation truncates, so the error is one-sided; the forward model divides by the convolved area
    # (spectrometers.py:286-288), which removes the common part of the bias
    assert rel < 6e-6
    w = torch.ones(F, dtype=torch.float64, device="cuda")
    w[0] = w[-1] = 0.5
    area = (out.double() * w).sum(dim=1)
    renorm = out.double() * ((ref * w).sum(dim=1) / area)[:, None]
    assert ((renorm - ref).abs() / ref.abs()).max().item() < 5e-6
    np.testing.assert_allclose(part.cpu().numpy(), area.cpu().numpy(), rtol=1e-7)
