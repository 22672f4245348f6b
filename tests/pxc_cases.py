"""Shared inputs of the folded pixel-contraction tests: spectra with the dynamic range of the real path, Gaussian
instrument functions, and a static wavelength map (pixel -> fine-grid index / weight)."""
import numpy as np

CASES = [  # F, K, sigma, P
    (8002, 8001, 25.0, 1024),
    (3002, 455, 20.0, 1024),
    (2984, 180, 12.0, 512),
    (6000, 301, 8.0, 4096),
]


def make_case(F, K, sigma, P, n=256, seed=None):
    rng = np.random.default_rng(F + K + P if seed is None else seed)
    x = np.arange(F)
    s = 1e9 * (1 + rng.random((n, F)))
    for _ in range(6):
        c, wdt, amp = rng.uniform(0, F, n), rng.uniform(0.3, 6, n), 10 ** rng.uniform(12, 16, n)
        s += amp[:, None] / (1 + np.abs((x[None] - c[:, None]) / wdt[:, None]) ** 2.5)
    taps = np.exp(-0.5 * ((np.arange(K) - (K - 1) / 2 + 0.3) / sigma) ** 2)
    taps[taps < 1e-300] = 0.0
    taps /= taps.sum()
    pos = np.linspace(0.0, F - 1.0001, P) + rng.uniform(-0.2, 0.2, P).clip(0, None) * 0.0
    idx = np.floor(pos).astype(np.int32)
    frac = pos - idx
    return s, taps, idx, frac
