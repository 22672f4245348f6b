"""float32 numpy model of the chord kernel's ALGORITHM (not its code) — a design check that runs on CPU.

The CUDA chord kernel departs from the reference's arithmetic in five deliberate ways; this model applies exactly
those departures to the oracle's float64 line scalars so that their combined effect on spectra / logL can be
checked against the reference-generated fixtures without a GPU (tests/test_kernel_model.py):

1. element-wise spectral work in float32 (Stark profile, Gaussians, IF convolution), float64 only for the
   wavelength differences, the tap tables, the reductions and the pixel/likelihood epilogue;
2. Stark (x) Doppler: the two `trapz` normalisations collapse into one integral of the convolved profile, obtained
   WITHOUT forming the convolution as  sum_i u[i] * G[i]  with  G[i] = sum_j w_j g[j + o - i]  (constant in the
   interior, prefix sums of the taps near the two ends);
3. Doppler taps, Gaussian components and instrument-function taps are truncated where exp(-z^2/2) < 1e-17
   (resp. |tap| < 1e-17 max|tap|);
4. direct (not FFT) convolutions with the same "same" alignment, out[j] = full[j + (K - 1) // 2];
5. uniform-grid trapz weights (the fine grid is `arange`, spacing jitter ~1e-13 relative);
6. the four convolution outputs at the ends of the fine grid are accumulated in float64 (extrapolation lever arm).
"""
import numpy as np
from scipy.constants import pi, speed_of_light

Z_CUT = 8.85
Z_CUT_DOPPLER = 6.9  # Doppler taps: exp(-z^2/2) >= 4.6e-11 of the central tap
FWHM_TO_SIGMA = 1 / np.sqrt(8 * np.log(2))
f32 = np.float32


def _taps_doppler(line, cwl_new, cwl0, ti):
    sigma = cwl_new * 7.715e-5 * np.sqrt(ti / line["mass"]) * FWHM_TO_SIGMA
    xk = line["doppler_x"] + (cwl_new - cwl0)
    z = (xk - cwl_new) / sigma
    g = np.exp(-0.5 * z * z)
    keep = np.abs(z) <= Z_CUT_DOPPLER
    k = np.where(keep)[0]
    return g, int(k.min()), int(k.max()) + 1


def _conv_same_direct(s, taps, k_lo, k_hi, K):
    """out[j] = sum_{k in [k_lo,k_hi)} taps[k] * s[j + o - k], o = (K-1)//2, zero padded; float32 accumulate."""
    o = (K - 1) // 2
    F = len(s)
    out = np.zeros(F, dtype=f32)
    pad = max(o, K)
    sp = np.zeros(F + 2 * pad, dtype=f32)
    sp[pad:pad + F] = s
    for k in range(k_lo, k_hi):
        out += f32(taps[k]) * sp[pad + o - k: pad + o - k + F]
    return out


def stark_doppler_weights(g, k_lo, k_hi, K, F):
    """G[i] / dl for every i (float64): trapz weights of out[j] = sum_k g[k] u[j + o - k] pulled back onto u."""
    o = (K - 1) // 2
    gs = np.zeros(K + 1)
    gs[1:] = np.cumsum(np.where((np.arange(K) >= k_lo) & (np.arange(K) < k_hi), g, 0.0))
    i = np.arange(F)
    # valid j in [0, F-1]  <->  k = j + o - i in [o - i, F - 1 + o - i]
    lo = np.clip(o - i, 0, K)
    hi = np.clip(F - 1 + o - i + 1, 0, K)
    G = gs[hi] - gs[lo]
    gm = np.where((np.arange(K) >= k_lo) & (np.arange(K) < k_hi), g, 0.0)
    k0 = o - i  # tap hitting j = 0
    k1 = F - 1 + o - i  # tap hitting j = F-1
    G -= 0.5 * np.where((k0 >= 0) & (k0 < K), gm[np.clip(k0, 0, K - 1)], 0.0)
    G -= 0.5 * np.where((k1 >= 0) & (k1 < K), gm[np.clip(k1, 0, K - 1)], 0.0)
    return G


def chord_forward_model(oracle, ch, st, scalars):
    """fp32 algorithm model of one (theta, chord) evaluation -> (fm[P] float64, logL float64)."""
    x = ch.x_data_fm
    F = len(x)
    dl = np.diff(x).mean()
    S = np.zeros(F, dtype=f32)
    c = ch.number
    cal = st[f"cal_{c}"][0] if oracle.calibrate_intensity else 1.0
    for line, o in zip(ch.lines, scalars):
        if line["kind"] == "balmer":
            sp = line["species"]
            cwl0 = line["cwl"]
            cwl = cwl0 * (st[sp + "_velocity"][0] / speed_of_light + 1) if sp + "_velocity" in st else cwl0
            g, k_lo, k_hi = _taps_doppler(line, cwl, cwl0, o["exc_ti"])
            K = len(g)
            Gw = stark_doppler_weights(g, k_lo, k_hi, K, F)
            a_ij, b_ij, c_ij = line["loman"]
            d = (x - cwl).astype(f32)
            a = np.abs(d)
            a25 = a * a * np.sqrt(a)
            m = np.zeros(F, dtype=f32)
            for which in ("rec", "exc"):
                gamma = 10.0 * c_ij * (1e6 * o[which + "_ne"]) ** a_ij / o[which + "_te"] ** b_ij / 2.0
                u = f32(1.0) / (a25 + f32(gamma**2.5))
                A = dl * np.sum(u.astype(np.float64) * Gw)
                m += f32(o[which + "_sum"] / A / cal) * u
            S += _conv_same_direct(m, g, k_lo, k_hi, K)
        else:
            if line["kind"] == "adas":
                sp = line["species"]
                cw = line["lines"]
                if sp + "_velocity" in st:
                    cw = line["cwls"] * (st[sp + "_velocity"][0] / speed_of_light + 1)
                fwhm = cw * 7.715e-5 * np.sqrt(o["ti"] / line["mass"])
                fr, ems = line["fractions"], o["emission_fitted"]
            else:
                cw, fr, ems = line["cwl"], line["fractions"], o["emission_fitted"]
                fwhm = np.full(len(cw), line["fwhm"])
            for ck, fk, fw in zip(cw, fr, fwhm):
                sigma = fw * FWHM_TO_SIGMA
                amp = ems * fk / (np.sqrt(2 * np.pi) * sigma) / cal
                z = (x - ck) / sigma
                w = np.abs(z) <= Z_CUT
                S[w] += (amp * np.exp(-0.5 * z[w] ** 2)).astype(f32)
    S64 = S.astype(np.float64)
    pre = dl * (S64.sum() - 0.5 * (S64[0] + S64[-1]))
    mn = S.min()
    if oracle.calibrate_instrument_function:
        fw = st[f"calint_func_{c}"][0]
        sigma = fw * FWHM_TO_SIGMA
        z = (x - x.mean()) / sigma
        taps = np.exp(-0.5 * z * z)
        keep = np.where(np.abs(z) <= Z_CUT)[0]
        taps = taps / taps[keep.min():keep.max() + 1].sum()
        k_lo, k_hi = int(keep.min()), int(keep.max()) + 1
    else:
        taps = ch.instrument_function_fm
        keep = np.where(np.abs(taps) > 1e-17 * np.abs(taps).max())[0]
        k_lo, k_hi = int(keep.min()), int(keep.max()) + 1
    Sc = np.maximum(_conv_same_direct(S, taps, k_lo, k_hi, len(taps)), mn)
    Sc64 = Sc.astype(np.float64)
    # the two points at either end drive interp1d's linear EXTRAPOLATION when the wavelength map is sampled
    # (pixels up to ~100 fine-grid steps outside the grid): accumulate those four outputs in float64
    o = (len(taps) - 1) // 2
    for j in (0, 1, F - 2, F - 1):
        k = np.arange(max(k_lo, j + o - (F - 1)), min(k_hi, j + o + 1))
        Sc64[j] = max(np.sum(np.asarray(taps, dtype=np.float64)[k] * S64[j + o - k]), float(mn))
    post = dl * (Sc64.sum() - 0.5 * (Sc64[0] + Sc64[-1]))
    ratio = post / pre
    bg = st[f"background_{c}"][0]
    cal_theta = st[f"calwave_{c}"] if oracle.calibrate_wavelength else ch.calwave_default
    pix = np.arange(len(ch.x_data), dtype=float)
    pix -= np.mean(pix)
    pix *= cal_theta[1]
    pix += cal_theta[0]
    i = np.clip(np.searchsorted(x, pix) - 1, 0, F - 2)
    a = (pix - x[i]) / (x[i + 1] - x[i])
    fm = np.maximum(bg + (Sc64[i] + a * (Sc64[i + 1] - Sc64[i])) / ratio, bg)
    r2 = ((ch.y_data - fm) / ch.error) ** 2
    return fm, -np.log(1 + r2).sum()


# ---------------------------------------------------------------------------------------------------------------
# Narrow-Doppler Balmer lines: closed-form tail integral (Euler-Maclaurin) and far-field series of the fused sweeps
WIN_GAMMAS, WIN_MIN, TAIL_TERMS, FAR_RATIO = 10.0, 48, 4, 500.0  # keep in sync with csrc/chord_stage.cuh


def stark_integral_closed_tails(F, delta, jc, r, g):
    """trapz (unit spacing, half weights at the two ends of the grid) of 1 / (|x_j|^2.5 + g), x_j = (j - jc) delta + r,
    the way sweep 1 of the chord kernel forms it: point by point inside +-max(WIN_MIN, WIN_GAMMAS g^0.4 / delta) points
    of the centre; on each tail the integral of the four-term series in g / |x|^2.5 plus the first two Euler-Maclaurin
    end corrections.  float64 throughout (the kernel's float32 window samples are covered by the fixture comparisons)."""
    f = lambda x: 1.0 / (np.abs(x) ** 2.5 + g)  # noqa: E731
    x = (np.arange(F) - jc) * delta + r
    n_w = int(min(max(np.ceil(WIN_GAMMAS * g ** 0.4 / delta), WIN_MIN), F))
    jL, jR = int(np.clip(jc - n_w, 0, F - 1)), int(np.clip(jc + n_w, 0, F - 1))
    tot = 0.0
    if jR > jL:
        w = np.ones(jR - jL + 1)
        w[0] = w[-1] = 0.5
        tot += np.sum(w * f(x[jL:jR + 1]))
    for a, b in ((0, jL), (jR, F - 1)):
        if b <= a:
            continue
        sgn = 1.0 if x[a] > 0 else -1.0
        A, d1, d3 = [], [], []
        for xe in (abs(x[a]), abs(x[b])):
            sq = np.sqrt(xe)
            u = -g / (xe * xe * sq)
            A.append(-sum(u ** k / (2.5 * k + 1.5) for k in range(TAIL_TERMS)) / (xe * sq))
            D, D1, D2, D3 = xe * xe * sq + g, 2.5 * xe * sq, 3.75 * sq, 1.875 / sq
            d1.append(-D1 / D ** 2)
            d3.append(-D3 / D ** 2 + 6 * D1 * D2 / D ** 3 - 6 * D1 ** 3 / D ** 4)
        tot += sgn * ((A[1] - A[0]) / delta + delta * (d1[1] - d1[0]) / 12 - delta ** 3 * (d3[1] - d3[0]) / 720)
    return tot


def stark_pair_far_field(d, s_r, s_e, g_r, g_e):
    """s_r / (d^2.5 + g_r) + s_e / (d^2.5 + g_e) by the three-term series in 1 / d^2.5 that sweep 2 uses where
    d^2.5 >= FAR_RATIO max(g_r, g_e)."""
    ia = d ** -2.5
    return ia * ((s_r + s_e) - ia * ((s_r * g_r + s_e * g_e) - ia * (s_r * g_r ** 2 + s_e * g_e ** 2)))
