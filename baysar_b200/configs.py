"""Named workloads (BASELINE.json `configs`, restated as concrete inputs in SURVEY.md §8d).

Each builder returns a :class:`Workload` holding the keyword arguments for ``make_input_dict`` (identical for
the reference's and this package's ``make_input_dict``), the line-of-sight grid of the default
``EsymmtricCauchyPlasma`` profile, and a reference theta given *by tag* so that each side assembles the vector
through its own ``plasma.slices``.

Measured spectra for the two demo configs come from the reference's demo directory
(`/root/reference/demo/zero_fit_config.json`, `/root/reference/demo/AUG_ROV012_32244.txt`, parsed as in
`demo/baysar_asdex_demo.py:317-324`) and are stored, together with the self-consistent synthetic spectra, in
``data/workload_inputs.npz`` (written by ``oracle/make_golden.py``).
"""
import pathlib

import numpy as np

_DATA = pathlib.Path(__file__).resolve().parent / "data" / "workload_inputs.npz"

X_TRIPLET = [[[4024.78, 4025.33, 4026.33]], [[0.541, 0.373, 0.086]]]


class Workload:
    def __init__(self, name, input_kwargs, los, reference_theta, description, profile=None, curvature=None,
                 extra_priors=(), plasma_attrs=None, provider_factory=None):
        self.name = name
        self.input_kwargs = input_kwargs
        self.los = los  # None -> profile default linspace(-10, 25, 50)
        self.reference_theta = reference_theta  # {tag: list of values}
        self.description = description
        self.profile = profile      # None -> EsymmtricCauchyPlasma; {"family": "bowman_tee" | "mesh", ...kwargs}
        self.curvature = curvature  # BaysarPosterior(curvature=...) (CurvatureCost needs the mesh family)
        # optional priors appended after the default components: [(class name in priors.py, keyword arguments)]
        self.extra_priors = list(extra_priors)
        # attributes assigned to posterior.plasma after construction, the way the reference's switches are used
        # (recombining, reverse_electroneutrality, ne_tau)
        self.plasma_attrs = dict(plasma_attrs or {})
        self.provider_factory = provider_factory  # None -> SyntheticADAS(); else a callable returning the atomic-data provider

    def make_provider(self):
        from .atomic_data import SyntheticADAS

        return self.provider_factory() if self.provider_factory is not None else SyntheticADAS()

    def make_priors(self, ns, plasma=None, plasma_state=None, posterior=None, late=None):
        """Instances of the optional priors from the namespace ``ns`` (the reference's ``baysar.priors`` or
        ``baysar_b200.priors``: same class names and signatures).  Entries whose keyword arguments hold the markers
        ``"@plasma"`` / ``"@plasma_state"`` / ``"@posterior"`` are built from keywords alone, with the markers replaced by
        the live objects -- they can only be made once the posterior exists (``late=True`` selects them, ``late=False``
        the others); the older entries take the plasma (TeMinPrior: the plasma_state) as their first argument."""
        live = {"@plasma": plasma, "@plasma_state": plasma_state, "@posterior": posterior}
        out = []
        for name, kw in self.extra_priors:
            marked = any(isinstance(v, str) and v in live for v in kw.values())
            if late is not None and marked != late:
                continue
            if marked:
                out.append(getattr(ns, name)(**{k: live[v] if isinstance(v, str) and v in live else v for k, v in kw.items()}))
            else:
                out.append(getattr(ns, name)(plasma_state if name == "TeMinPrior" else plasma, **kw))
        return out

    def make_profile(self, ns):
        """Profile-function object from the namespace ``ns`` — the reference's ``baysar.lineshapes`` module,
        ``oracle.numpy_port`` or ``baysar_b200.lineshapes`` all expose the same class names and signatures."""
        spec = dict(self.profile or {"family": "esymmetric"})
        family = spec.pop("family")
        if family == "esymmetric":
            if spec:  # e.g. ptype="gaussian"
                return ns.EsymmtricCauchyPlasma(x=np.array(self.los, dtype=float), **spec)
            return ns.EsymmtricCauchyPlasma(x=np.array(self.los, dtype=float)) if self.los is not None else None
        if family == "PoissonPlasma" and not hasattr(ns, family):
            # the reference's Poisson-density SimplePlasma (lineshapes.py:443-453) is shadowed by the second definition:
            # compose it the way a reference user has to
            import types

            x = np.array(self.los, dtype=float)
            return types.SimpleNamespace(x=x, electron_density=ns.Poisson(x), electron_temperature=ns.ExpDecay(x))
        if family in CLOSED_FORM_FAMILIES:
            return getattr(ns, family)(x=np.array(self.los, dtype=float), **spec)
        if family == "bowman_tee":
            return ns.BowmanTeePlasma(x=np.array(self.los, dtype=float), **spec)
        if family == "mesh":
            return ns.MeshPlasma(x=np.array(spec.pop("knots"), dtype=float), **spec)
        raise ValueError(f"unknown profile family {family}")

    def theta_from_tags(self, slices, start):
        """Overwrite ``start`` (a vector inside the bounds) with the tagged reference values."""
        theta = np.array(start, dtype=float)
        for tag, vals in self.reference_theta.items():
            if tag in slices:
                theta[slices[tag]] = vals
        return theta


def _stored(key, default=None):
    if _DATA.exists():
        with np.load(_DATA) as z:
            if key in z.files:
                return z[key]
    if default is None:
        raise FileNotFoundError(f"{key} not in {_DATA}; run oracle/make_golden.py in the build container")
    return default


def gaussian_if(npix=31, sigma_px=1.2):
    x = np.arange(float(npix))
    return np.exp(-0.5 * ((x - (npix - 1) / 2.0) / sigma_px) ** 2)


def balmer_series(data="selfconsistent", flags=None):
    """configs[1]: `demo/balmer_series.py` — H Balmer n=5..9->2, 1024 px, refine 0.1, L=50, n=10.
    ``flags``: extra `make_input_dict` keywords (e.g. ``cold_neutrals=False``: thermalised neutrals, whose Doppler kernels
    span 30-200 taps of the 0.1 A grid — the multi-tap Stark (x) Doppler path)."""
    wave = _stored("balmer_wave")
    if data == "zero_fit":
        ems = _stored("balmer_zero_fit_ems")
    else:
        ems = _stored("balmer_selfconsistent_ems", default=_stored("balmer_zero_fit_ems"))
    kw = dict(wavelength_axis=[wave], experimental_emission=[ems], instrument_function=[_stored("balmer_if")],
              emission_constant=[1e11], noise_region=[[4150.0, 4250.0]], species=["H"], ions=[["0"]],
              mystery_lines=None, refine=0.1, no_sample_neutrals=False, ion_resolved_temperatures=False,
              ion_resolved_tau=False)
    if flags:
        kw.update(flags)
    ref = {"electron_density": [np.log10(8e13), 0.0, 3.0], "electron_temperature": [np.log10(5.0), 3.0, 1.0],
           "H_0_dens": [np.log10(8e13)], "H_0_tau": [-8.0], "H_0_velocity": [0.0], "background_0": [12.05]}
    hot = bool(flags) and flags.get("cold_neutrals") is False
    return Workload("balmer_hot" if hot else "balmer_series", kw, None, ref,
                    "H Balmer series with Stark broadening (demo/balmer_series.py)" +
                    (", thermalised neutrals (multi-tap Doppler kernels)" if hot else ""))


def balmer_two_chords():
    """Two copies of the Balmer-series chord (second one with a shifted background level): exercises the per-chord
    loop of the tensor-core path and its shared workspace."""
    wl = balmer_series()
    kw = dict(wl.input_kwargs)
    wave, ems, inst = kw["wavelength_axis"][0], kw["experimental_emission"][0], kw["instrument_function"][0]
    kw.update(wavelength_axis=[wave, wave.copy()], experimental_emission=[ems, ems * 1.02 + 3e10],
              instrument_function=[inst, inst.copy()], emission_constant=[1e11, 1e11],
              noise_region=[[4150.0, 4250.0], [4150.0, 4250.0]], refine=[0.1, 0.1])
    ref = dict(wl.reference_theta)
    ref["background_1"] = [12.06]
    return Workload("balmer_two_chords", kw, None, ref, "two H Balmer-series chords sharing one plasma")


def multi_species(n_chords=1, data="selfconsistent", flags=None):
    """configs[2]/[3]: D Balmer + N II/N III + C III + X triplet, 1024 px, refine 0.05, L=100."""
    npix = 1024
    wave = np.linspace(3960.0, 4110.0, npix)
    rng = np.random.default_rng(20261019 + 3)
    waves, emss, ifs = [], [], []
    for c in range(n_chords):
        key = f"multi_selfconsistent_ems_{c}"
        if data == "selfconsistent" and _DATA.exists() and key in np.load(_DATA).files:
            ems = _stored(key)
        else:
            ems = 1e12 + rng.normal(0.0, 2e10, npix)
        waves.append(wave.copy())
        emss.append(ems)
        ifs.append(gaussian_if(31, 1.2))
    kw = dict(wavelength_axis=waves, experimental_emission=emss, instrument_function=ifs,
              emission_constant=[1e11] * n_chords, noise_region=[[4008.0, 4020.0]] * n_chords,
              species=["D", "N", "C"], ions=[["0"], ["1", "2"], ["2"]], mystery_lines=X_TRIPLET,
              refine=[0.05] * n_chords, ion_resolved_temperatures=False, ion_resolved_tau=True)
    if flags:
        kw.update(flags)
    ref = {"electron_density": [14.6, 0.5, 3.0], "electron_temperature": [0.9, 3.0, 0.8],
           "N_1_dens": [12.6], "C_2_dens": [12.2], "D_0_dens": [12.9], "D_0_tau": [-4.0], "N_1_tau": [-1.0],
           "N_2_tau": [-3.0], "C_2_tau": [-2.0], "D_0_velocity": [2.0], "N_1_velocity": [-3.0],
           "C_2_velocity": [1.0], "X_4024.78_4025.33_4026.33": [12.3]}
    for c in range(n_chords):
        ref[f"background_{c}"] = [12.0 + 0.01 * c]
    name = "multi_species" if n_chords == 1 else f"multi_species_{n_chords}chords"
    return Workload(name, kw, np.linspace(-10.0, 25.0, 100), ref,
                    f"D Balmer + N II/N III + C III + X triplet, {n_chords} chord(s)")


def multi_species_switched(**plasma_attrs):
    """The multi-species chord with post-construction plasma switches (`plasma.recombining = True`, ...)."""
    wl = multi_species(1)
    tag = "_".join(k for k, v in sorted(plasma_attrs.items()) if v)
    return Workload(f"multi_{tag}", wl.input_kwargs, wl.los, wl.reference_theta, wl.description + f", plasma.{tag} = True",
                    plasma_attrs=plasma_attrs)


CLOSED_FORM_FAMILIES = {
    # family: (line of sight, reference theta of the density, of the temperature)
    "SimplePlasma": (np.linspace(0.0, 30.0, 100), [14.3, -1.0], [1.0, -1.3]),
    "LessSimplePlasma": (np.linspace(0.0, 30.0, 100), [14.3, -1.0, 10.0, -1.0, 0.5], [1.0, -1.3]),
    "SimpleGaussianPlasma": (np.linspace(0.0, 30.0, 100), [14.3, 0.9, 12.0], [1.0, 1.1]),
    "SlabPlasma": (np.linspace(0.0, 20.0, 100), [14.3, 1.0], [0.8]),
    "DoubleSlabPlasma": (np.linspace(0.0, 20.0, 100), [14.3, 13.8, 5.0], [0.3, 1.2]),
    "PoissonPlasma": (np.linspace(0.0, 30.0, 100), [14.3, 1.0], [1.0, -1.3]),
}


def multi_species_family(family, **spec):
    """The multi-species chord on one of the reference's other profile families (lineshapes.py:456-606, :855-948), or on
    the default family with ``ptype="gaussian"``."""
    wl = multi_species(1)
    ref = dict(wl.reference_theta)
    if family == "esymmetric":
        los, name = wl.los, "multi_esym_" + "_".join(str(v) for v in spec.values())
    else:
        los, ref["electron_density"], ref["electron_temperature"] = CLOSED_FORM_FAMILIES[family]
        name = "multi_" + family
    out = Workload(name, wl.input_kwargs, los, ref, wl.description + f", profile family {family}",
                   profile=dict(family=family, **spec))
    if family == "LessSimplePlasma":  # uniform draws from its bounds overflow (base^(x0 - x)): fixtures stay near the reference theta
        out.theta_jitter = 0.03
    return out


def file_adas_provider(adf_dir):
    """`FileADAS` over the synthetic ADF11 / ADF15 text files of ``adf_dir`` (written by oracle/make_adf_golden.py):
    hydrogen tables, the N and C rate coefficients and the PEC files the multi-species lines name -- no other provider."""
    from .adas_files import FileADAS

    adf_dir = pathlib.Path(adf_dir)
    kinds = ("scd", "acd", "plt", "prb")
    return FileADAS(adf15=adf_dir / "pec12_h0_synthetic.dat", adf11={k: adf_dir / f"{k}12_h_synthetic.dat" for k in kinds},
                    impurity_adf11={e: {k: adf_dir / f"{k}96_{e.lower()}_synthetic.dat" for k in kinds} for e in ("N", "C")},
                    impurity_adf15={p.name[:-len(".synthetic")]: p for p in adf_dir.glob("*.pass.synthetic")})


def multi_species_file_adas(adf_dir):
    """The multi-species chord with every atomic table read from ADAS-format text files."""
    wl = multi_species(1, data="flat")
    return Workload("file_adas_multi", wl.input_kwargs, wl.los, wl.reference_theta, wl.description + ", tables from ADF files",
                    provider_factory=lambda: file_adas_provider(adf_dir))


def multi_species_bowman(background=False):
    """The multi-species chord on the BowmanTeePlasma profile family (reference lineshapes.py:332-440), L = 120."""
    wl = multi_species(1)
    ref = dict(wl.reference_theta)
    ref["electron_density"] = [14.6, 0.5, 2.5, 2.0, 3.0, 5.0, 0.5] + ([12.0] if background else [])
    ref["electron_temperature"] = [0.9, 3.0, 2.0, 3.0, 5.0, 0.5] + ([-0.5] if background else [])
    return Workload("multi_bowman" + ("_bg" if background else ""), wl.input_kwargs, np.linspace(-15.0, 25.0, 120), ref,
                    "multi-species chord, BowmanTee ne/Te profiles", profile=dict(family="bowman_tee", background=background))


OPTIONAL_PRIORS = [("ElectronDensityTDVCost", dict(threshold=2.0, sigma=0.3)), ("PeakTePrior", dict(te_min=8.0, te_err=1.5)),
                   ("TauPrior", dict(mean=0.5, sigma=0.4)), ("WallConditionsPrior", dict(te_min=0.6, ne_min=3e12)),
                   ("SimpleWallPrior", dict(te_min=0.6, ne_min=3e12)), ("TeMinPrior", dict(ratio=12.0, sigma=2.5)),
                   ("TeTauPrior", dict(sigma=0.4)), ("ImpurityTauPrior", dict(sigma=0.6))]


def multi_species_priors():
    """The multi-species chord with the optional priors of priors.py that only read the plasma state (TDV cost, peak /
    wall / minimum temperature, tau ordering priors), appended after the default components."""
    wl = multi_species(1)
    return Workload("multi_priors", wl.input_kwargs, wl.los, wl.reference_theta, wl.description + ", optional priors",
                    extra_priors=OPTIONAL_PRIORS)


def multi_species_bowman_priors():
    """BowmanTee profiles with BowmanTeePrior (priors.py:484-507) and two of the state priors."""
    wl = multi_species_bowman()
    return Workload("multi_bowman_priors", wl.input_kwargs, wl.los, wl.reference_theta, wl.description + ", optional priors",
                    profile=wl.profile, extra_priors=[("BowmanTeePrior", dict(sigma_err=0.3)), ("PeakTePrior", {}),
                                                      ("TeMinPrior", {})])


HOST_PRIORS = [  # evaluated on the host (baysar_b200/host_priors.py); built once the posterior exists
    ("FlatnessPriorBasic", dict(scale=-15.5, tag="electron_density", plasma="@plasma")),
    ("FlatnessPriorBasic", dict(scale=-1.5, tag="electron_temperature", plasma="@plasma")),
    ("NIIITauPrior", dict(plasma_state="@plasma_state", mean=1, sigma=0.3)),
    ("CiiiPrior", dict(plasma="@plasma", mean=0, sigma=0.5)),
    ("EmsTeOrderPrior", dict(posterior="@posterior", mean=2.0, sigma=0.5, species=["N_1", "N_2"])),
]
MESH_HOST_PRIORS = [  # the two that refresh the knot vectors first: PlasmaGradientPrior reads what they leave there
    ("GaussianProcessPrior", dict(mu=13.0, plasma="@plasma", A=2.0, L=0.5, profile="electron_density")),
    ("FlatnessPrior", dict(scale=-1.0, tag="electron_temperature", plasma="@plasma")),
    ("PlasmaGradientPrior", dict(plasma="@plasma", te_scale=0.05, ne_scale=0.05)),
    ("SeparatrixTePrior", dict(scale=0.2, plasma="@plasma", index=2, ne_scale=0.1)),
]


def multi_species_host_priors(mesh=False):
    wl = multi_species_mesh() if mesh else multi_species(1)
    name = "multi_mesh_host_priors" if mesh else "multi_host_priors"
    return Workload(name, wl.input_kwargs, wl.los, wl.reference_theta, wl.description + ", host-evaluated optional priors",
                    profile=wl.profile, extra_priors=MESH_HOST_PRIORS if mesh else HOST_PRIORS)


def multi_species_bolometry():
    """The multi-species chord with BolometryPrior (priors.py:595-643), which reads the radiated-power side outputs."""
    wl = multi_species(1)
    return Workload("multi_bolometry", wl.input_kwargs, wl.los, wl.reference_theta, wl.description + ", BolometryPrior",
                    extra_priors=[("BolometryPrior", dict(plasma="@plasma", mean=1500.0, sigma=400.0))])


def multi_species_mesh(curvature=None):
    """The multi-species chord on the MeshPlasma profile family (reference lineshapes.py:235-329): 10^theta at 7
    knots, linear interpolation onto arange(-10, 25, 0.1) (L = 350); optional CurvatureCost (priors.py:104-127)."""
    wl = multi_species(1)
    # The reference's MeshLine counts the two fixed end points as variables when ``zero_bounds`` is given, which its
    # own bounds check then rejects (lineshapes.py:262, plasmas.py:419-425): the usable form has no fixed ends, so the
    # knots span the whole line of sight.
    knots = [-10.0, -6.0, -3.0, -1.0, 0.5, 2.0, 6.0, 12.0, 25.0]
    ref = dict(wl.reference_theta)
    ref["electron_density"] = [12.4, 13.0, 13.6, 14.2, 14.6, 14.3, 13.5, 12.6, 12.0]
    ref["electron_temperature"] = [-0.3, 0.0, 0.4, 0.8, 0.9, 0.7, 0.3, -0.2, -0.5]
    return Workload("multi_mesh" + ("_curv" if curvature else ""), wl.input_kwargs, None, ref,
                    "multi-species chord, MeshLine ne/Te profiles",
                    profile=dict(family="mesh", knots=knots, bounds=[-10.0, 25.0], zero_bounds_ne=None,
                                 zero_bounds_te=None), curvature=curvature)


def aug_demo():
    """configs[0]: AUG divertor chord ROV012 #32244 (`demo/baysar_asdex_demo.py`), 512 px, refine 0.05."""
    wave = _stored("aug_wave")
    ems = _stored("aug_ems")
    x = np.arange(31.0)
    fwhm_px = 0.4 / np.mean(np.diff(wave))
    sigma = fwhm_px / np.sqrt(8.0 * np.log(2.0))
    inst = np.exp(-0.5 * ((x - 15.0) / sigma) ** 2) / (np.sqrt(2.0 * np.pi) * sigma)
    kw = dict(wavelength_axis=[wave], experimental_emission=[ems], instrument_function=[inst],
              emission_constant=[1e11], noise_region=[[3975.0, 3990.0]], species=["D", "N"],
              ions=[["0"], ["1", "2", "3"]], mystery_lines=X_TRIPLET, refine=[0.05],
              ion_resolved_temperatures=False, ion_resolved_tau=True)
    ref = {"background_0": [12.0], "electron_density": [14.8, 0.0, 3.0], "electron_temperature": [1.0, 3.0, 1.0],
           "N_1_dens": [12.7], "D_0_tau": [-4.0], "N_1_tau": [-1.0], "N_2_tau": [-4.0], "N_3_tau": [-4.0],
           "D_0_velocity": [0.0], "N_1_velocity": [0.0], "X_4024.78_4025.33_4026.33": [12.3]}
    return Workload("aug_demo", kw, None, ref, "AUG_ROV012_32244 divertor chord, D + N II/III/IV + X triplet")


def synthetic_sweep(n_los=100, n_lines=20, n_pix=1024, balmer=True):
    """configs[4]: shape sweep — single-component impurity lines at evenly spaced cwls (custom line_data)."""
    wave = 3960.0 + 0.15 * np.arange(n_pix)
    rng = np.random.default_rng(20261019 + 5)
    ems = 1e12 + rng.normal(0.0, 2e10, n_pix)
    line_data = {"N": {"atomic_mass": 14, "atomic_charge": 7, "ionisation_balance_year": 96, "ions": ["1"], "1": {}}}
    lo, hi = wave.min() + 2.0, wave.max() - 2.0
    for k, cwl in enumerate(np.linspace(lo, hi, n_lines)):
        cwl = float(np.round(cwl, 3))
        line_data["N"]["1"][cwl] = {"wavelenght": cwl, "jj_frac": [1], "pec": "sweep_n1.dat", "exc_block": 1 + (k % 28),
                                    "rec_block": 31 + (k % 28)}
    species, ions = ["N"], [["1"]]
    if balmer:
        from .line_data import fresh_line_data

        line_data["D"] = fresh_line_data()["D"]
        species, ions = ["D", "N"], [["0"], ["1"]]
    kw = dict(wavelength_axis=[wave], experimental_emission=[ems], instrument_function=[gaussian_if(31, 1.2)],
              emission_constant=[1e11], noise_region=[[float(wave[5]), float(wave[60])]], species=species, ions=ions,
              line_data=line_data, mystery_lines=None, refine=[0.05], ion_resolved_temperatures=False,
              ion_resolved_tau=True)
    ref = {"electron_density": [14.5, 0.0, 3.0], "electron_temperature": [1.0, 3.0, 1.0], "N_1_dens": [12.5],
           "N_1_tau": [-2.0], "D_0_tau": [-4.0], "background_0": [12.0]}
    return Workload(f"sweep_L{n_los}_G{n_lines}_P{n_pix}", kw, np.linspace(-10.0, 25.0, n_los), ref,
                    "synthetic shape sweep")
