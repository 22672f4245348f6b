"""TEST INFRASTRUCTURE — generate committed fixtures from the reference itself (build container only).

    python -m oracle.make_golden            # writes baysar_b200/data/workload_inputs.npz and tests/golden/*.npz

Step 1 exports the measured demo spectra (reference `demo/` directory) and the self-consistent synthetic spectra
("forward model at a reference theta + noise", the second stage of `demo/balmer_series.py:145-160`).
Step 2 runs the shimmed reference posterior on seeded theta batches for every workload and records logP,
per-component values, forward-model spectra, line scalars and the constant setup state.
"""
import json
import pathlib
import sys

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from baysar_b200 import configs  # noqa: E402
from baysar_b200.atomic_data import SyntheticADAS  # noqa: E402
from oracle import ref_harness, ref_shim  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
DATA = ROOT / "baysar_b200" / "data" / "workload_inputs.npz"
REF = pathlib.Path(ref_shim.REFERENCE_ROOT)


def export_demo_inputs():
    cfg = json.loads((REF / "demo" / "zero_fit_config.json").read_text())
    out = {
        "balmer_wave": np.array(cfg["wavelength_axis"][0], dtype=float),
        "balmer_zero_fit_ems": np.array(cfg["experimental_emission"][0], dtype=float),
        "balmer_if": np.array(cfg["instrument_function"][0], dtype=float),
    }
    # demo/baysar_asdex_demo.py:317-324 — one flat list of numbers; > 1e5 are emission, the rest wavelength in nm
    tmp = (REF / "demo" / "AUG_ROV012_32244.txt").read_text().replace("\n", "")
    d = np.array([float(t) for t in tmp.split()])
    out["aug_ems"] = d[d > 1e5] * 1e-1 * 1e-4
    out["aug_wave"] = d[d < 1e5] * 10
    return out


def self_consistent_spectrum(workload, noise_sigma, seed):
    post = ref_harness.build_reference_posterior(workload)
    np.random.seed(seed)
    theta = workload.theta_from_tags(post.plasma.slices, post.random_start())
    post(list(theta), True)
    rng = np.random.default_rng(seed)
    out = []
    for c in range(post.plasma.num_chords):
        fm = post.posterior_components[c].forward_model()
        out.append(fm + rng.normal(0.0, noise_sigma, len(fm)))
    return out


def save(name, post, rec):
    setup = ref_harness.describe_setup(post)
    path = GOLDEN / f"{name}.npz"
    np.savez_compressed(path, **{f"setup_{k}": v for k, v in setup.items()}, **rec)
    ok = np.isfinite(rec["logp"]).sum()
    print(f"{name}: {len(rec['logp'])} thetas ({ok} finite), logP[0] = {rec['logp'][0]:.6f}, "
          f"{path.stat().st_size / 1024:.0f} KiB")
    for e in sorted(set(rec["error"])):
        if e:
            print("   raised:", e, "x", int((rec["error"] == e).sum()))


def main():
    GOLDEN.mkdir(parents=True, exist_ok=True)
    ref_shim.install(SyntheticADAS())

    data = export_demo_inputs()
    np.savez_compressed(DATA, **data)
    # self-consistent spectra need the first export to exist
    data["balmer_selfconsistent_ems"] = self_consistent_spectrum(configs.balmer_series("zero_fit"), 1e12, 11)[0]
    np.savez_compressed(DATA, **data)
    for nch in (1, 3):
        spec = self_consistent_spectrum(configs.multi_species(nch, data="flat"), 2e11, 13)
        for c, s in enumerate(spec):
            if nch == 1 or c > 0:
                data[f"multi_selfconsistent_ems_{c}"] = s
    np.savez_compressed(DATA, **data)
    print("wrote", DATA, sorted(data))

    jobs = [
        ("balmer_series", configs.balmer_series(), 24, 2, {}),
        ("balmer_zero_fit", configs.balmer_series("zero_fit"), 8, 0, {}),
        ("multi_species", configs.multi_species(1), 24, 2, {}),
        ("multi_species_3chords", configs.multi_species(3), 8, 0, {}),
        ("aug_demo", configs.aug_demo(), 16, 1, {}),
        ("multi_hot", configs.multi_species(1, flags=dict(cold_neutrals=False, cold_ions=False)), 12, 1, {}),
        ("multi_sampled_neutrals", configs.multi_species(1, flags=dict(no_sample_neutrals=False)), 8, 0, {}),
        ("multi_calibrated", configs.multi_species(1, flags=dict(calibrate_wavelength=True, calibrate_intensity=True,
                                                                 calibrate_instrument_function=True)), 8, 1, {}),
        ("multi_no_doppler", configs.multi_species(1, flags=dict(doppler_shifts=False)), 6, 0, {}),
        ("multi_bowman", configs.multi_species_bowman(), 10, 1, {}),
        ("multi_bowman_bg", configs.multi_species_bowman(background=True), 8, 0, {}),
        ("multi_mesh_curv", configs.multi_species_mesh(curvature=50.0), 10, 1, {}),
        ("multi_priors", configs.multi_species_priors(), 12, 0, {}),
        ("multi_bowman_priors", configs.multi_species_bowman_priors(), 8, 0, {}),
        ("balmer_hot", configs.balmer_series(flags=dict(cold_neutrals=False)), 10, 1, {}),
        ("multi_species_32chords", configs.multi_species(32), 3, 0, {}),
        ("multi_recombining", configs.multi_species_switched(recombining=True), 10, 0, {}),
        ("multi_reverse_electroneutrality", configs.multi_species_switched(reverse_electroneutrality=True), 10, 0, {}),
        ("multi_ne_tau", configs.multi_species_switched(ne_tau=True), 3, 0, {}),
        *[("multi_" + f, configs.multi_species_family(f), 8, 0, {}) for f in configs.CLOSED_FORM_FAMILIES],
        ("multi_esym_gaussian", configs.multi_species_family("esymmetric", ptype="gaussian"), 8, 0, {}),
        ("multi_host_priors", configs.multi_species_host_priors(), 8, 0, {}),
        ("multi_bolometry", configs.multi_species_bolometry(), 6, 0, {}),
        ("multi_mesh_host_priors", configs.multi_species_host_priors(mesh=True), 8, 0, {}),
        ("file_adas_multi", configs.multi_species_file_adas(GOLDEN / "adf"), 8, 0, {}),
    ]
    only = sys.argv[1:]
    for name, wl, n, nfine, attrs in jobs:
        if only and name not in only:
            continue
        post = ref_harness.build_reference_posterior(wl, provider=wl.make_provider(), plasma_attrs={**wl.plasma_attrs, **attrs})
        thetas = ref_harness.draw_thetas(post, n, wl, seed=20261019)
        rec = ref_harness.evaluate(post, thetas, full_fine_grid=nfine)
        save(name, post, rec)


if __name__ == "__main__":
    main()
