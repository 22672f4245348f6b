"""Design check: the chord kernel's float32 algorithm (tests/kernel_model.py) holds the parity gates on CPU.

Spectra within 1e-5 relative per pixel and the chord log-likelihood within 1e-6 of the log-posterior of the reference-generated
fixtures (BASELINE.json north_star tolerances)."""
import numpy as np
import pytest

from .conftest import GOLDEN
from .kernel_model import chord_forward_model
from .test_oracle_golden import build
from .workloads import WORKLOADS


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_fp32_algorithm_holds_tolerance(name):
    z = np.load(GOLDEN / f"{name}.npz")
    post = build(name)
    worst_s, worst_l = 0.0, 0.0
    for i, theta in enumerate(z["theta"][:8]):
        r = post.evaluate(theta)
        for c, ch in enumerate(post.chords):
            fm, logl = chord_forward_model(post, ch, r["state"], r["line_scalars"][c])
            ref = z[f"c{c}_spectra"][i]
            worst_s = max(worst_s, np.max(np.abs(fm - ref) / np.abs(ref)))
            # the north-star tolerance is on the log-posterior the chord's likelihood is a term of (a well-fitted chord's
            # own value can be a small part of it, and float32 pixel rounding does not shrink with it)
            worst_l = max(worst_l, abs(logl - z["components"][i, c]) / abs(z["logp"][i]))
    print(f"{name}: spectra rel {worst_s:.2e}  logL rel {worst_l:.2e}")
    assert worst_s < 1e-5
    assert worst_l < 1e-6


def test_closed_form_tail_integral_matches_dense_trapz():
    """Sweep 1 of the narrow-Doppler Balmer passes: dense window + closed-form tails + Euler-Maclaurin end terms equals
    the dense trapz over the whole fine grid to < 1e-8 relative, for every Stark width and line position."""
    from .kernel_model import stark_integral_closed_tails

    rng = np.random.default_rng(5)
    worst = 0.0
    for trial in range(600):
        F = int(rng.integers(700, 9000))
        delta = float(rng.choice([0.01, 0.05, 0.1]))
        gamma = 10 ** rng.uniform(-2.5, 1.2)  # Angstrom
        jc = int(rng.integers(-400, F + 400)) if trial % 4 == 0 else int(rng.integers(0, F))
        r = float(rng.uniform(-0.5, 0.5)) * delta
        x = (np.arange(F) - jc) * delta + r
        w = np.ones(F)
        w[0] = w[-1] = 0.5
        dense = np.sum(w / (np.abs(x) ** 2.5 + gamma ** 2.5))
        worst = max(worst, abs(stark_integral_closed_tails(F, delta, jc, r, gamma ** 2.5) - dense) / dense)
    print(f"closed-form tail integral: worst relative error {worst:.2e}")
    assert worst < 1e-8


def test_far_field_series_of_stark_pair():
    from .kernel_model import FAR_RATIO, stark_pair_far_field

    rng = np.random.default_rng(6)
    worst = 0.0
    for _ in range(2000):
        g_r, g_e = 10 ** rng.uniform(-6, 3, size=2)
        s_r, s_e = 10 ** rng.uniform(8, 14, size=2)
        d = (FAR_RATIO * max(g_r, g_e)) ** 0.4 * 10 ** rng.uniform(0, 2)
        exact = s_r / (d ** 2.5 + g_r) + s_e / (d ** 2.5 + g_e)
        worst = max(worst, abs(stark_pair_far_field(d, s_r, s_e, g_r, g_e) - exact) / exact)
    print(f"far-field series: worst relative error {worst:.2e}")
    assert worst < 2e-8


def test_far_field_moment_series_of_the_doppler_convolution():
    """Multi-tap Balmer lines: beyond BAYSAR_FAR_SIGMAS Doppler widths (and the far ratio of the Stark series) the
    windowed discrete convolution sum_u g_u m(x - u delta) equals G_0 ia (c0 P_2.5(y) - ia (c1 P_5(y) - ia c2 P_7.5(y)))
    with the polynomials of the exact tap moments (csrc/chord_stage.cuh, far pass of the multi-tap path): relative error
    < 2e-8 for every width, sub-grid tap offset and pair of Stark widths."""
    rng = np.random.default_rng(0)
    r25 = [2.5, 4.375, 6.5625, 9.0234375, 11.73046875, 14.6630859375]
    r50, r75 = [5.0, 15.0, 35.0, 70.0], [7.5, 31.875]
    worst = 0.0
    for _ in range(200):
        delta = float(rng.choice([0.05, 0.1]))
        F = 6000
        sigma = 10 ** rng.uniform(-1.2, 0.0)
        cwl = 4000 + rng.uniform(0, 300)
        x = 3900.0 + np.arange(F) * delta
        KD = 2 * int(8 * sigma / delta + 5) + 1
        o = (KD - 1) // 2
        u = np.arange(KD) - o
        z = (u * delta + rng.uniform(-0.5, 0.5) * delta) / sigma
        g = np.where(np.abs(z) <= 6.9, np.exp(-0.5 * z * z), 0.0)
        g_r, g_e = 10 ** rng.uniform(-6, 1.5, 2)
        s_r, s_e = 10 ** rng.uniform(8, 12, 2)
        d = x - cwl
        a = np.abs(d) ** 2.5
        m = s_r / (a + g_r) + s_e / (a + g_e)
        exact = np.convolve(m, g, mode="full")[o:o + F]  # out[j] = sum_kk g[kk] m[j + o - kk]
        G0 = g.sum()
        mu = [(g * (u * delta) ** n).sum() / G0 for n in range(7)]
        y, ia = 1 / d, np.abs(d) ** -2.5
        p1 = 1 + sum(mu[n + 1] * r25[n] * y ** (n + 1) for n in range(6))
        p2 = 1 + sum(mu[n + 1] * r50[n] * y ** (n + 1) for n in range(4))
        p3 = 1 + sum(mu[n + 1] * r75[n] * y ** (n + 1) for n in range(2))
        c0, c1, c2 = s_r + s_e, s_r * g_r + s_e * g_e, s_r * g_r ** 2 + s_e * g_e ** 2
        far = G0 * ia * (c0 * p1 - ia * (c1 * p2 - ia * c2 * p3))
        T = int((np.abs(z) <= 6.9).sum())
        W = max(40 * sigma, (500 * max(g_r, g_e)) ** 0.4)
        sel = (np.abs(d) >= W + 2 * delta) & (np.arange(F) >= T) & (np.arange(F) < F - T)
        if sel.any():
            worst = max(worst, float(np.max(np.abs(far[sel] - exact[sel]) / exact[sel])))
    print(f"far-field moment series: worst relative error {worst:.2e}")
    assert worst < 2e-8
