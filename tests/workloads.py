"""Fixture name -> workload builder, shared by the oracle and the GPU parity tests."""
import pathlib

from baysar_b200 import configs

ADF_DIR = pathlib.Path(__file__).resolve().parent / "golden" / "adf"

WORKLOADS = {
    "balmer_series": lambda: configs.balmer_series(),
    "balmer_zero_fit": lambda: configs.balmer_series("zero_fit"),
    "multi_species": lambda: configs.multi_species(1),
    "multi_species_3chords": lambda: configs.multi_species(3),
    "aug_demo": lambda: configs.aug_demo(),
    "multi_hot": lambda: configs.multi_species(1, flags=dict(cold_neutrals=False, cold_ions=False)),
    "multi_sampled_neutrals": lambda: configs.multi_species(1, flags=dict(no_sample_neutrals=False)),
    "multi_calibrated": lambda: configs.multi_species(
        1, flags=dict(calibrate_wavelength=True, calibrate_intensity=True, calibrate_instrument_function=True)),
    "multi_no_doppler": lambda: configs.multi_species(1, flags=dict(doppler_shifts=False)),
    "multi_bowman": lambda: configs.multi_species_bowman(),
    "multi_bowman_bg": lambda: configs.multi_species_bowman(background=True),
    "multi_mesh_curv": lambda: configs.multi_species_mesh(curvature=50.0),
    "multi_priors": lambda: configs.multi_species_priors(),
    "multi_bowman_priors": lambda: configs.multi_species_bowman_priors(),
    "balmer_hot": lambda: configs.balmer_series(flags=dict(cold_neutrals=False)),
    "multi_species_32chords": lambda: configs.multi_species(32),
    **{"multi_" + f: (lambda f=f: configs.multi_species_family(f)) for f in configs.CLOSED_FORM_FAMILIES},
    "multi_esym_gaussian": lambda: configs.multi_species_family("esymmetric", ptype="gaussian"),
    "multi_host_priors": lambda: configs.multi_species_host_priors(),
    "multi_bolometry": lambda: configs.multi_species_bolometry(),
    "multi_mesh_host_priors": lambda: configs.multi_species_host_priors(mesh=True),
    "file_adas_multi": lambda: configs.multi_species_file_adas(ADF_DIR),
    "multi_recombining": lambda: configs.multi_species_switched(recombining=True),
    "multi_reverse_electroneutrality": lambda: configs.multi_species_switched(reverse_electroneutrality=True),
}
