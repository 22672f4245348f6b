"""CPU check of the CUDA kernel SOURCE: csrc/*.cuh compiled with g++ against tests/emu/cpu_emu.h (one std::thread per
CUDA thread, barriers for __syncthreads / shuffles) and run on the descriptor blob the model compiler produces.

This validates the host set-up (tags, slices, bounds, fine grid, instrument function), the model compiler and the
kernels' logic against the reference-generated fixtures without a GPU.  The GPU parity tests (-m gpu) repeat the same
comparisons through the real C ABI.
"""
import numpy as np
import pytest

from .conftest import GOLDEN
from .emu_harness import emu_eval, emu_eval_split, host_model
from .workloads import WORKLOADS

N_THETA = 3


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_host_setup_matches_reference(name):
    z = np.load(GOLDEN / f"{name}.npz")
    plasma, chords, priors, blob, info = host_model(WORKLOADS[name]())
    assert list(plasma.slices.keys()) == [str(s) for s in z["setup_tags"]]
    assert [s.start for s in plasma.slices.values()] == list(z["setup_slice_start"])
    assert [s.stop for s in plasma.slices.values()] == list(z["setup_slice_stop"])
    assert plasma.n_params == int(z["setup_n_params"])
    np.testing.assert_allclose(plasma.theta_bounds, z["setup_theta_bounds"], rtol=1e-14)
    names = ["SpectrometerChord"] * len(chords) + [type(p).__name__ for p in priors]
    late = [n for n, kw in WORKLOADS[name]().extra_priors if any(isinstance(v, str) and v.startswith("@") for v in kw.values())]
    assert names + late == [str(s) for s in z["setup_components"]]  # (host-evaluated priors are appended after construction)
    for c, ch in enumerate(chords):
        np.testing.assert_array_equal(ch.x_data_fm, z[f"setup_c{c}_x_data_fm"])
        np.testing.assert_allclose(ch.error, z[f"setup_c{c}_error"], rtol=1e-14)
        np.testing.assert_allclose(ch.instrument_function_fm, z[f"setup_c{c}_instrument_function_fm"], rtol=1e-13,
                                   atol=1e-300)
        assert [type(l).__name__ for l in ch.lines] == [str(s) for s in z[f"setup_c{c}_line_types"]]


@pytest.mark.parametrize("name", sorted(WORKLOADS))
def test_emulated_kernels_match_reference(name):
    z = np.load(GOLDEN / f"{name}.npz")
    plasma, chords, priors, blob, info = host_model(WORKLOADS[name]())
    th = z["theta"][:N_THETA]
    for c in range(len(chords)):
        o = emu_eval(blob, th, chord=c, n_chords=len(chords), n_priors=len(priors), n_ul=info["n_ulines"],
                     L=len(plasma.los), P=len(chords[c].x_data), F=len(chords[c].x_data_fm))
        assert (o["status"] == 0).all()
        # host-evaluated optional priors (baysar_b200/host_priors.py) are the trailing columns of the reference's
        # components and are not part of the device model: the kernels' logP is the total without them
        ncol = len(chords) + len(priors)
        ref_logp = z["logp"][:N_THETA] - z["components"][:N_THETA, ncol:].sum(axis=1)
        np.testing.assert_allclose(o["logp"], ref_logp, rtol=1e-6)  # north_star: 1e-6 relative
        np.testing.assert_allclose(o["spectra"], z[f"c{c}_spectra"][:N_THETA], rtol=1e-5)  # 1e-5 per pixel
        # components: the same 1e-6, measured against the posterior they add up to (a well-fitted chord's likelihood
        # can be a small part of it, and float32 pixel rounding does not shrink with it)
        for i in range(N_THETA):
            np.testing.assert_allclose(o["comps"][i], z["components"][i, :ncol], rtol=1e-6, atol=1e-6 * abs(z["logp"][i]))
        np.testing.assert_allclose(o["profiles"][:, 0], z["ne"][:N_THETA], rtol=1e-12)
        np.testing.assert_allclose(o["profiles"][:, 2], z["main_ion_density"][:N_THETA], rtol=1e-12)
        ref = z[f"c{c}_line_scalars"][:N_THETA]
        for k, u in enumerate(info["chord_line_uline"][c]):
            kind = str(z[f"setup_c{c}_line_types"][k])
            rec = o["lines"][:, u]
            if kind == "BalmerHydrogenLine":
                mine = rec[:, [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 6]]
                np.testing.assert_allclose(mine, ref[:, k, :11], rtol=1e-11)
            elif kind == "ADAS406Lines":
                np.testing.assert_allclose(rec[:, :5], ref[:, k, :5], rtol=1e-11)
            else:
                np.testing.assert_allclose(rec[:, :1], ref[:, k, :1], rtol=1e-12)


def test_emulated_gaussian_likelihood():
    """likelihood(cauchy=False) (spectrometers.py:187-188) as a model flag: the chord column of a model compiled with it
    equals the reference's value (fixture multi_recombining carries c0_gauss_likelihood)."""
    name = "multi_recombining"
    z = np.load(GOLDEN / f"{name}.npz")
    plasma, chords, priors, blob, info = host_model(WORKLOADS[name](), gaussian_likelihood=True)
    th = z["theta"][:N_THETA]
    o = emu_eval(blob, th, chord=0, n_chords=len(chords), n_priors=len(priors), n_ul=info["n_ulines"],
                 L=len(plasma.los), P=len(chords[0].x_data), F=len(chords[0].x_data_fm))
    np.testing.assert_allclose(o["comps"][:, 0], z["c0_gauss_likelihood"][:N_THETA], rtol=1e-5)
    np.testing.assert_allclose(o["comps"][:, 1:], z["components"][:N_THETA, 1:], rtol=1e-6, atol=1e-9)


def test_status_flags_in_emulation():
    """NaN / inf theta and an out-of-grid impurity tau are reported as flags, not as crashes."""
    plasma, chords, priors, blob, info = host_model(WORKLOADS["multi_species"]())
    z = np.load(GOLDEN / "multi_species.npz")
    th = np.repeat(z["theta"][:1], 3, axis=0)
    th[0, 3] = np.nan
    th[1, 2] = np.inf
    th[2, plasma.slices["N_1_tau"].start] = -7.5
    o = emu_eval(blob, th, n_chords=1, n_priors=len(priors), n_ul=info["n_ulines"], L=len(plasma.los),
                 P=len(chords[0].x_data), F=len(chords[0].x_data_fm))
    assert o["status"][0] & 1
    assert o["status"][1] & 2
    assert o["status"][2] & 4


@pytest.mark.parametrize("name", ["balmer_series", "multi_species", "aug_demo"])
def test_split_path_stages_in_emulation(name):
    """Spectral stage (MODE 1) -> emulated 3xTF32 contraction -> pixel stage reproduces the fixtures."""
    z = np.load(GOLDEN / f"{name}.npz")
    plasma, chords, priors, blob, info = host_model(WORKLOADS[name]())
    th = z["theta"][:2]
    o = emu_eval_split(blob, th, chord=0, n_chords=len(chords), n_priors=len(priors), n_ul=info["n_ulines"],
                       P=len(chords[0].x_data))
    assert (o["status"] == 0).all()
    np.testing.assert_allclose(o["logp"], z["logp"][:2], rtol=1e-6)
    np.testing.assert_allclose(o["spectra"], z["c0_spectra"][:2], rtol=1e-5)


@pytest.mark.parametrize("name", ["multi_species", "multi_bowman", "multi_mesh_curv", "multi_priors", "balmer_series"])
def test_emulated_los_team_mode(name, monkeypatch):
    """K1 with one CTA (four warps) per theta — the form the product launches for long lines of sight, forced here
    for the fixtures' shorter ones: log-posterior, prior components and line scalars as in the warp-per-theta form."""
    monkeypatch.setenv("BAYSAR_EMU_LOS_TEAM", "1")
    z = np.load(GOLDEN / f"{name}.npz")
    plasma, chords, priors, blob, info = host_model(WORKLOADS[name]())
    th = z["theta"][:4]
    o = emu_eval(blob, th, chord=0, n_chords=len(chords), n_priors=len(priors), n_ul=info["n_ulines"],
                 L=len(plasma.los), P=len(chords[0].x_data), F=len(chords[0].x_data_fm))
    assert (o["status"] == 0).all()
    np.testing.assert_allclose(o["logp"], z["logp"][:4], rtol=1e-6)
    np.testing.assert_allclose(o["comps"], z["components"][:4], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(o["profiles"][:, 0], z["ne"][:4], rtol=1e-12)
    ref = z["c0_line_scalars"][:4]
    for k, u in enumerate(info["chord_line_uline"][0]):
        kind = str(z["setup_c0_line_types"][k])
        if kind == "ADAS406Lines":
            np.testing.assert_allclose(o["lines"][:, u][:, :5], ref[:, k, :5], rtol=1e-11)
