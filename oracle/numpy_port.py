"""TEST INFRASTRUCTURE — CPU restatement (numpy/scipy) of BaySAR's posterior-evaluation path.

This file is the *oracle*: a plain, stateless re-statement of what the reference computes in
``BaysarPosterior.__call__`` (`/root/reference/baysar/posterior.py:219-235`), written as a pure function of
theta.  It is NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of
``bench.py`` import it.  The product (``baysar_b200``) never does, and fails loudly without its CUDA library.

Pinning: the reference has no golden vectors for this path, so parity is anchored on outputs of
the reference itself, run unmodified under import shims (``oracle/ref_shim.py``) and committed as fixtures in
``tests/golden/*.npz`` by ``oracle/make_golden.py``.  ``tests/test_oracle_golden.py`` checks this port against
every fixture (logP, per-component values, forward-model spectra, line scalars, slices, bounds, setup arrays).

Each block cites the reference lines it follows.  Third-party arithmetic (scipy ``fftconvolve``, ``interp1d``,
``RectBivariateSpline``) is called directly, as the reference does.
"""
import collections

import numpy as np
from scipy.constants import pi, speed_of_light
from scipy.interpolate import RectBivariateSpline, interp1d
from scipy.ndimage import center_of_mass
from scipy.signal import fftconvolve

HYDROGEN_ISOTOPES = ["H", "D", "T"]
ATOMIC_MASS = {"H": 1, "D": 2, "He": 4, "Be": 9, "B": 10.8, "C": 12, "N": 14, "O": 16, "Ne": 20.2}  # linemodels.py:305-310
ATOMIC_NUMBER = {"He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Ne": 10}  # plasmas.py:161-171
LOMAN = {  # linemodels.py:533-551
    "32": [0.7665, 0.064, 3.710e-18], "42": [0.7803, 0.050, 8.425e-18], "52": [0.6796, 0.030, 1.310e-15],
    "62": [0.7149, 0.028, 3.954e-16], "72": [0.7120, 0.029, 6.258e-16], "82": [0.7159, 0.032, 7.378e-16],
    "92": [0.7177, 0.033, 8.947e-16], "102": [0.7158, 0.032, 1.239e-15], "112": [0.7146, 0.028, 1.632e-15],
    "122": [0.7388, 0.026, 6.459e-16], "132": [0.7356, 0.020, 9.012e-16], "43": [0.7449, 0.045, 1.330e-16],
    "53": [0.7356, 0.044, 6.640e-16], "63": [0.7118, 0.016, 2.481e-15], "73": [0.7137, 0.029, 3.270e-15],
    "83": [0.7133, 0.032, 4.343e-15], "93": [0.7165, 0.033, 5.588e-15],
}
FWHM_TO_SIGMA = 1 / np.sqrt(8 * np.log(2))  # lineshapes.py:106
ROOT_2PI = np.sqrt(2 * np.pi)  # lineshapes.py:123
trapz = getattr(np, "trapezoid", None) or np.trapz


class OracleError(Exception):
    """A numerical failure the reference reports by raising (kind = reference exception type)."""

    def __init__(self, kind, msg):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind


# ---------------------------------------------------------------------------------------------- primitives
def gaussian_norm(x, cwl, fwhm, intensity):
    """lineshapes.py:82-103,126-131."""
    if np.any(np.asarray(fwhm) <= 0):
        raise OracleError("ValueError", "fwhm must be a positive scalar")
    if np.any(np.asarray(intensity) <= 0):
        raise OracleError("ValueError", "fwhm must be a positive scalar")
    sigma = fwhm * FWHM_TO_SIGMA
    k = intensity / (ROOT_2PI * sigma)
    return k * np.exp(-0.5 * ((x - cwl) / sigma) ** 2)


def gaussian_multiplet(x, cwls, fwhms, fractions, intensity):
    """``Gaussian.make_peak`` with ``normalise=True`` over the full grid (lineshapes.py:206-217)."""
    fwhms = np.atleast_1d(fwhms)
    if len(fwhms) < len(cwls) and len(fwhms) == 1:
        fwhms = [fwhms[0] for _ in cwls]
    peak = np.zeros(len(x))
    for f, c, fw in zip(fractions, cwls, fwhms):
        peak += gaussian_norm(x, c, fw, f)
    return intensity * peak


def bilinear_extrap(g0, g1, table, x0, x1):
    """``scipy.interpolate.interpn(method="linear", bounds_error=False, fill_value=None)`` restated
    (what ``DataArray.interp(..., kwargs=dict(bounds_error=False, fill_value=None))`` resolves to;
    plasmas.py:723-731, linemodels.py:782-788): edge-cell linear extrapolation, unclamped weights."""
    i = np.clip(np.searchsorted(g0, x0) - 1, 0, len(g0) - 2)
    j = np.clip(np.searchsorted(g1, x1) - 1, 0, len(g1) - 2)
    a = (x0 - g0[i]) / (g0[i + 1] - g0[i])
    b = (x1 - g1[j]) / (g1[j + 1] - g1[j])
    return (table[i, j] * (1 - a) * (1 - b) + table[i, j + 1] * (1 - a) * b
            + table[i + 1, j] * a * (1 - b) + table[i + 1, j + 1] * a * b)


def centre_peak(peak, places=12):
    """tools.py:84-100 — iterate linear-interpolation shifts until the centre of mass is central."""
    x = np.arange(len(peak))
    centre = np.mean(x)
    while not np.round(abs(centre - center_of_mass(peak))[0], places) == 0:
        shift = center_of_mass(peak) - centre
        peak = interp1d(x, peak, bounds_error=False, fill_value=0.0)(x + shift)
    return peak


def within(stuff, boxes):
    """tools.py:103-127 (the variant `spectrometers.get_lines` uses)."""
    if type(stuff) in (tuple, list):
        return any([any([(box.min() < float(s)) and (float(s) < box.max()) for s in stuff]) for box in boxes])
    stuff = float(stuff)
    return any((box.min() < stuff) and (stuff < box.max()) for box in boxes)


def clip_data(x, y, x_range):
    """tools.py:26-47."""
    idx = np.where((min(x_range) < x) & (x < max(x_range)))
    return x[idx], y[idx]


def high_pass(tmp, threshold, error):  # priors.py:4-6
    return -0.5 * (max([0, threshold - tmp]) / error) ** 2


def low_pass(tmp, threshold, error):  # priors.py:9-11
    return -0.5 * (max([0, tmp - threshold]) / error) ** 2


class EsymmtricProfile:
    """lineshapes.py:779-826 (``ptype``: :769-784)."""

    def __init__(self, x, log_peak_bounds, centre=None, ptype="cauchy"):
        self.x = x
        self.ptype = ptype
        self.centre = centre
        self.bounds = [log_peak_bounds]
        if centre is None:
            self.bounds.append([-5, 2])
        self.bounds.append([1, 6])
        if centre is not None:
            self.bounds.append([0.2, 3])
        self.number_of_variables = len(self.bounds)

    def __call__(self, theta):
        if self.centre is None:
            A, c, f = theta
            floor = 1e-3
        else:
            A, f, floor = theta
            c = self.centre
        dx = self.x - c
        sigma = 0.2 * f + f / (1 + np.exp(-dx))
        ms = np.square(dx / sigma)
        peak = np.power(10, A) * np.exp(-0.5 * ms) if self.ptype == "gaussian" else np.power(10, A) / (1 + ms)
        return peak.clip(floor)


class EsymmtricCauchyPlasma:
    """lineshapes.py:829-852."""

    def __init__(self, x=None, bounds_ne=(13, 15), bounds_te=(0.0, 1.7), ptype="cauchy"):
        self.x = np.linspace(-10, 25, 50) if x is None else np.asarray(x, dtype=float)
        self.electron_density = EsymmtricProfile(self.x, list(bounds_ne), centre=None, ptype=ptype)
        self.electron_temperature = EsymmtricProfile(self.x, list(bounds_te), centre=0, ptype=ptype)


def _pf(x, bounds, fn):
    """A closed-form profile object: bounds, number_of_variables and the callable theta -> profile."""
    obj = type("Profile", (), {})()
    obj.x, obj.bounds, obj.number_of_variables, obj.fn = x, [list(b) for b in bounds], len(bounds), fn
    type(obj).__call__ = lambda self, theta: self.fn(np.asarray(theta, dtype=float))
    return obj


def _exp_decay(x, peak_bounds=(-1, 2)):  # lineshapes.py:483-495
    return _pf(x, [peak_bounds, [-2, 1]], lambda t: (np.power(10.0, t[0]) * np.power(np.ones(len(x)) + np.power(10.0, t[1]), -x)).clip(0.01))


def _double_exp_decay(x):  # lineshapes.py:498-518
    def fn(t):
        base = (np.ones(len(x)) + np.power(10.0, t[1])) * np.exp(np.power(10.0, t[3]) * np.tanh(t[4] * (x - t[2])))
        return (np.power(10.0, t[0]) * np.power(base, t[2] - x)).clip(0.01)
    return _pf(x, [[12, 16], [-2, 1], [0, 50], [-3, 1], [-5, 5]], fn)


def _gaussian_plasma(x, cwl=None):  # lineshapes.py:556-591 (height-normalised gaussian :109-120)
    def fn(t):
        c = t[2] if cwl is None else cwl
        sigma = np.power(10.0, t[1]) / np.sqrt(8 * np.log(2))
        return (np.power(10.0, t[0]) * np.exp(-0.5 * ((x - c) / sigma) ** 2)).clip(0.01)
    bounds = [[12, 16], [-1, np.log10(x.max() / 2)]] + ([[x.min(), x.max()]] if cwl is None else [])
    return _pf(x, bounds, fn)


def _poisson(x):  # lineshapes.py:537-550
    from scipy.special import factorial

    def fn(t):
        a, nu = np.power(10.0, t[0]), np.power(10.0, t[1])
        peak = np.power(np.zeros(len(x)) + nu, x)
        peak *= np.exp(-nu) / factorial(x)
        peak *= a / peak.max()
        return peak
    return _pf(x, [[12, 16], [-3, 2]], fn)


class PoissonPlasma:  # lineshapes.py:443-453 (the SimplePlasma definition the second one shadows)
    def __init__(self, x=None):
        self.x = np.linspace(0, 60, 100) if x is None else np.asarray(x, dtype=float)
        self.electron_density, self.electron_temperature = _poisson(self.x), _exp_decay(self.x)


class SimplePlasma:  # lineshapes.py:469-480
    def __init__(self, x=None):
        self.x = np.linspace(0, 60, 100) if x is None else np.asarray(x, dtype=float)
        self.electron_density, self.electron_temperature = _exp_decay(self.x, (12, 16)), _exp_decay(self.x)


class LessSimplePlasma:  # lineshapes.py:521-532
    def __init__(self, x=None):
        self.x = np.linspace(0, 60, 100) if x is None else np.asarray(x, dtype=float)
        self.electron_density, self.electron_temperature = _double_exp_decay(self.x), _exp_decay(self.x)


class SimpleGaussianPlasma:  # lineshapes.py:594-606
    def __init__(self, x=None):
        self.x = np.linspace(0, 60, 100) if x is None else np.asarray(x, dtype=float)
        self.electron_density, self.electron_temperature = _gaussian_plasma(self.x), _gaussian_plasma(self.x, cwl=0)


class SlabPlasma:  # lineshapes.py:855-901
    def __init__(self, x=None, dr_bounds=(-5, 2), bounds_ne=(13, 15), bounds_te=(0.0, 2.0)):
        x = self.x = np.linspace(0, 20, 200) if x is None else np.asarray(x, dtype=float)
        self.electron_density = _pf(x, [bounds_ne, dr_bounds],
                                    lambda t: np.where(x < t[1] + (x[1] - x[0]), np.power(10.0, t[0]), np.zeros(x.shape) + 1e10))
        self.electron_temperature = _pf(x, [bounds_te], lambda t: np.power(10.0, t[0]) + np.zeros(x.shape))


class DoubleSlabPlasma:  # lineshapes.py:906-948: the density carries the centre and hands it to the temperature
    def __init__(self, x=None, centre=4):
        x = self.x = np.linspace(0, 20, 200) if x is None else np.asarray(x, dtype=float)
        state = {"centre": centre}

        def hat(t):
            c = state["centre"]
            return np.where(x < c, np.power(10.0, t[0]), np.where(c < x, np.power(10.0, t[1]), 0.0))

        def density(t):
            state["centre"] = t[2]
            return hat(t)

        self.electron_density = _pf(x, [[13, 15], [13, 14.7], [2, 8]], density)
        self.electron_temperature = _pf(x, [[-1, 0.7], [0, 2.0]], hat)


class MeshLine:
    """lineshapes.py:235-301 (``kind="linear"``): theta at the knots (plus fixed end values ``zero_bounds`` at
    ``x_ends``), ``10**`` if ``log``, linear interpolation with linear extrapolation onto ``arange(x_ends)``."""

    def __init__(self, x, x_ends=(-10, 10), zero_bounds=None, resolution=0.1, log=False, bounds=None):
        x = np.asarray(x, dtype=float)
        self.log, self.zero_bounds = log, zero_bounds
        if zero_bounds is None:
            self.empty_theta = np.zeros(len(x))
            self.slice = slice(0, len(x))
            self.x_points = x
        else:
            self.slice = slice(1, -1)
            self.empty_theta = np.zeros(len(x) + 2)
            self.empty_theta[0] = self.empty_theta[-1] = zero_bounds
            self.x_points = np.concatenate([np.array([min(x_ends)]), x, np.array([max(x_ends)])])
        self.x_points = self.x_points.astype(float)
        self.x = np.arange(min(x_ends), max(x_ends), resolution)
        self.number_of_variables = len(self.x_points)  # sic: counts the fixed end points too (lineshapes.py:262)
        self.bounds = bounds

    def __call__(self, theta):
        from scipy.interpolate import interp1d

        values = self.empty_theta.copy()
        values[self.slice] = theta
        if self.log:
            values = np.power(10, values)
        return interp1d(self.x_points, values, "linear", bounds_error=False, fill_value="extrapolate")(self.x)


class MeshPlasma:
    """lineshapes.py:304-329."""

    def __init__(self, x=None, bounds=None, zero_bounds_ne=11, zero_bounds_te=-2, bounds_ne=(11, 16), bounds_te=(-1, 2)):
        self.electron_density = MeshLine(x=x, zero_bounds=zero_bounds_ne, x_ends=bounds, log=True,
                                         bounds=[list(bounds_ne) for _ in x])
        self.electron_temperature = MeshLine(x=x, zero_bounds=zero_bounds_te, x_ends=bounds, log=True,
                                             bounds=[list(bounds_te) for _ in x])


def bowman_tee_distribution(x, theta):
    """lineshapes.py:332-343."""
    A, x0, sigma0, q, nu, k, f, b = theta
    sigma = sigma0 * np.exp(f * np.tanh(k * (x - x0)))
    z = np.power(abs((x - x0) / sigma), q)
    return A * np.power(1 + z / nu, -(nu + 1) * 0.5) + b


class BowmanTeeProfile:
    """BowmanTeeNe / BowmanTeeTe, lineshapes.py:354-376: ``10**`` of the peak (and of the background level when
    ``background``), the temperature profile is centred on 0."""

    def __init__(self, x, background, centred):
        self.x, self.background, self.centred = x, background, centred

    def __call__(self, theta):
        theta = np.array(theta, dtype=float)
        if self.background:
            theta[0] = np.power(10, theta[0])
            theta[-1] = np.power(10, theta[-1])
        else:
            theta[0] = np.power(10, theta[0])
            theta = np.concatenate((theta, np.zeros(1)))
        if self.centred:
            A, sigma0, q, nu, k, f, b = theta
            theta = [A, 0, sigma0, q, nu, k, f, b]
        return bowman_tee_distribution(self.x, theta)


class BowmanTeePlasma:
    """lineshapes.py:381-440."""

    def __init__(self, x=None, bounds=None, dr_bounds=(-2, 2), bounds_ne=(11, 16), bounds_te=(-1, 2), background=False):
        self.x = np.linspace(-15, 25, 500) if x is None else x
        shape = [[1e-1, 10], [1.2, 3], [1, 5], [1, 50], [0, 2]] if bounds is None else bounds
        self.background = background
        self.electron_density = BowmanTeeProfile(self.x, background, centred=False)
        self.electron_temperature = BowmanTeeProfile(self.x, background, centred=True)
        b_ne = [list(bounds_ne), list(dr_bounds)] + [list(b) for b in shape]
        b_te = [list(bounds_te)] + [list(b) for b in shape]
        if background:
            b_ne.append([bounds_ne[0], 13.0])
            b_te.append([bounds_te[0], 0])
        self.electron_density.bounds, self.electron_temperature.bounds = b_ne, b_te
        self.electron_density.number_of_variables = len(b_ne)
        self.electron_temperature.number_of_variables = len(b_te)


def curvature(profile):
    """priors.py:104-107."""
    grad1 = np.gradient(profile / max(profile))
    grad2 = np.gradient(grad1)
    return -sum(np.square(grad2) / np.power((1 + np.square(grad1)), 3))


# ---------------------------------------------------------------------------------------------- the oracle
class OraclePosterior:
    """Stateless restatement of ``BaysarPosterior`` for the default component set."""

    def __init__(self, input_dict, provider, profile_function=None, curvature=None, extra_priors=(), plasma_attrs=None):
        # post-construction switches of the reference's PlasmaLine (read with hasattr on every evaluation):
        # recombining (plasmas.py:735-737), reverse_electroneutrality (:663-671), ne_tau (linemodels.py:398-399)
        for k in ("recombining", "reverse_electroneutrality", "ne_tau"):
            setattr(self, k, (plasma_attrs or {}).get(k, False))
        self.input_dict = input_dict
        self.curvature = curvature
        self.extra_priors = list(extra_priors)  # [(class name, kwargs)] appended after the default components
        self.provider = provider
        d = input_dict
        # ---- PlasmaLine.__init__ (plasmas.py:295-406)
        self.num_chords = len(d["wavelength_axis"])
        self.species = d["species"]
        self.impurity_species = [s for s in self.species if not ("H_" in s or "D_" in s or "T_" in s)]
        self.hydrogen_species = [s for s in self.species if s not in self.impurity_species]
        self.contains_hydrogen = self.species != self.impurity_species
        tau_exc = np.logspace(-7, 2, 10 + 12)
        tau_rec = np.logspace(1, -4, 12 - 4)
        self.magical_tau = np.concatenate((np.log10(tau_exc), -np.log10(tau_rec) + 4))  # plasmas.py:326-328
        self.adas_te = np.logspace(-1, 2, 48)
        self.adas_ne = np.logspace(12, 15, 15)
        self.tau_exc, self.tau_rec = tau_exc, tau_rec
        self.big_ne = np.array([self.adas_ne for _ in self.adas_te]).T
        flag = lambda k, default: d[k] if k in d else default  # noqa: E731  (plasmas.py:340-385)
        self.include_doppler_shifts = flag("doppler_shifts", True)
        self.calibrate_wavelength = flag("calibrate_wavelength", False)
        self.calibrate_intensity = flag("calibrate_intensity", False)
        self.calibrate_instrument_function = flag("calibrate_instrument_function", False)
        self.zeeman = flag("zeeman", False)
        self.no_sample_neutrals = flag("no_sample_neutrals", True)
        self.thermalised = flag("thermalised", True)
        self.cold_neutrals = flag("cold_neutrals", True)
        self.cold_ions = flag("cold_ions", True)
        if self.zeeman:
            raise NotImplementedError("Zeeman splitting is outside the hot-path scope (SURVEY.md §2 row 3)")
        if len(self.hydrogen_species) > 1:
            raise NotImplementedError("one hydrogen-isotope species per posterior")
        self.concstar = True  # plasmas.py:948-950
        self.impurities = []
        for s in self.impurity_species:  # plasmas.py:411-417
            elem = s[: s.index("_")]
            if elem not in self.impurities:
                self.impurities.append(elem)
        self._build_ion_balance()
        self._build_tecs()
        if self.contains_hydrogen:
            self._build_hydrogen_tables()
        self.profile_function = profile_function or EsymmtricCauchyPlasma()
        self.los = self.profile_function.electron_density.x
        self._build_tags_slices_and_bounds()
        # ---- BaysarPosterior.build_posterior_components / get_priors / set_sensible_bounds
        self.chords = [self._build_chord(c, refine) for c, refine in enumerate(d["refine"])]
        self._build_prior_list()
        self._set_sensible_bounds()

    # ------------------------------------------------------------------------------------ atomic tables
    def _build_ion_balance(self):
        """plasmas.py:939-1079 (fractional abundances only; the radiated-power tables are side outputs)."""
        self.impurity_ion_bal = {}
        for elem in self.impurities:
            nion = ATOMIC_NUMBER[elem] + 1
            bal = np.zeros((len(self.tau_exc) + len(self.tau_rec), len(self.adas_ne), len(self.adas_te), nion))
            for half, (taus, min_frac0, meta_index) in enumerate(
                    [(self.tau_exc, 1e-8, 0), (self.tau_rec, 1e-4, -1)]):
                meta = np.zeros(nion)
                meta[meta_index] = 1
                for t_counter, t in enumerate(taus):
                    ion = np.array(self.provider.ionisation_balance(elem, self.adas_te, self.adas_ne, t, meta))
                    ion[ion < min_frac0] = 1e-15
                    ion = ion / np.arange(nion).clip(1)[None, None, :]  # concstar correction, :993-996
                    bal[half * len(self.tau_exc) + t_counter] = ion
            self.impurity_ion_bal[elem] = bal

    def _radiated_power_splines(self):
        """plasmas.py:997-1020, :1052-1075, :1083-1096: log radiated-power coefficient splines per (ion, tau).  The
        reference's second (tau_rec) loop re-writes power[t_counter] of the first half, so the rec slices stay zero."""
        if getattr(self, "_rad_splines", None) is None:
            big_ne = np.array([self.adas_ne for _ in self.adas_te]).T
            out = {}
            for elem in self.impurities:
                bal = self.impurity_ion_bal[elem]
                n_pow = ATOMIC_NUMBER[elem]
                power = np.zeros((len(self.magical_tau), len(self.adas_ne), len(self.adas_te), n_pow))
                for t in range(len(self.tau_exc)):
                    for ion in range(n_pow):
                        plt = self.provider.adf11_rates(elem, "plt", ion + 1, self.adas_te, self.adas_ne)
                        prb = self.provider.adf11_rates(elem, "prb", ion + 1, self.adas_te, self.adas_ne)
                        power[t, :, :, ion] = (bal[t, :, :, ion] * plt.T + bal[t, :, :, ion + 1] * prb.T) * big_ne
                for ion in range(n_pow):
                    out[f"{elem}_{ion}"] = {tau: RectBivariateSpline(self.adas_ne, self.adas_te, np.log(power[t, :, :, ion].clip(1e-30)))
                                            for t, tau in enumerate(self.magical_tau)}
            self._rad_splines = out
        return self._rad_splines

    def _build_tecs(self):
        """plasmas.py:1207-1266 — 30 bicubic splines of log(ne*TEC) per transition."""
        self.impurity_tecs = {}
        for species in self.impurity_species:
            elem, ion = species.split("_")[0], int(species.split("_")[1])
            for line in self.input_dict[species].keys():
                line_str = str(line).replace(", ", "_")
                for bad in ["[", "]", "(", ")"]:
                    line_str = line_str.replace(bad, "")
                rec = self.input_dict[species][line]
                pecs_exc = self.provider.adf15_pecs(rec["pec"], rec["exc_block"], self.adas_te, self.adas_ne)
                pecs_rec = self.provider.adf15_pecs(rec["pec"], rec["rec_block"], self.adas_te, self.adas_ne)
                splines = []
                for t in range(len(self.magical_tau)):
                    z_bal = self.impurity_ion_bal[elem][t, :, :, ion]
                    zp1_bal = self.impurity_ion_bal[elem][t, :, :, ion + 1]
                    rates = self.big_ne * (pecs_exc.T * z_bal + pecs_rec.T * zp1_bal)
                    splines.append(RectBivariateSpline(self.adas_ne, self.adas_te, np.log(rates.clip(1e-30))))
                self.impurity_tecs[species + "_" + line_str] = splines

    def _build_hydrogen_tables(self):
        """plasmas.py:1277-1316 (log PEC tables) and :706-721 (adf11 scd/acd/plt/prb, block 1)."""
        species = self.hydrogen_species[0]
        element, charge = species.split("_")
        tab = self.provider.hydrogen_adf15(element.lower(), int(charge))
        self.h_ne, self.h_te = tab.ne, tab.te
        self.hydrogen_pecs = {}
        for sp in self.hydrogen_species:
            for line in self.input_dict[sp].keys():
                tag = sp + "_" + str(line).replace(", ", "_")
                rec = self.input_dict[sp][line]
                self.hydrogen_pecs[tag + "_exc"] = np.log(tab.select(rec["exc_block"]))
                self.hydrogen_pecs[tag + "_rec"] = np.log(tab.select(rec["rec_block"]))
        self.neutral_adf11 = {}
        for k in ["scd", "acd", "plt", "prb"]:
            t = self.provider.hydrogen_adf11(k)
            self.neutral_adf11[k] = (t.ne, t.te, t.select(1))

    # ------------------------------------------------------------------------------------ parameter layout
    def _build_tags_slices_and_bounds(self):
        """plasmas.py:457-634 (incl. the first-character slice sharing of non-resolved tags, :602-616)."""
        d = self.input_dict
        tags, bounds, lengths = [], [], []
        cal = ("cal", 1, [[-5, 20]], self.calibrate_intensity)
        calwave = ("calwave", 2, [[1e2, 1e5], [0.01, 1.0]], self.calibrate_wavelength)
        calint = ("calint_func", 1, [[-1, 1]], self.calibrate_instrument_function)
        background = ("background", 1, [[-5, 20]], True)
        for tag, nvar, b, on in (cal, calwave, calint, background):
            for chord in range(self.num_chords):
                if on:
                    tags.append(f"{tag}_{chord}")
                    lengths.append(nvar)
                    bounds.extend(b)
        for fn, tag in ((self.profile_function.electron_density, "electron_density"),
                        (self.profile_function.electron_temperature, "electron_temperature")):
            tags.append(tag)
            lengths.append(fn.number_of_variables)
            bounds.extend(fn.bounds)
        if self.thermalised:
            impurity_tags, impurity_bounds = ["_dens", "_tau"], [[11, 14], [-6, 4]]
        else:
            impurity_tags, impurity_bounds = ["_dens", "_Ti", "_tau"], [[11, 14], [-1, 2], [-6, 4]]
        if self.include_doppler_shifts:
            impurity_tags.append("_velocity")
            impurity_bounds.append([-10, 10])
        for tag in impurity_tags:
            for s in self.species:
                is_h = any([s[0:2] == h + "_" for h in HYDROGEN_ISOTOPES])
                if tag == "_tau" and is_h:
                    tags.append(s + tag)
                    lengths.append(1)
                    bounds.append([-10, -2])
                elif tag == "_dens" and is_h and self.no_sample_neutrals:
                    pass
                elif tag == "_Ti" and is_h and self.cold_neutrals:
                    pass
                elif tag == "_Ti" and not is_h and self.cold_ions:
                    pass
                else:
                    sion = s + tag
                    tags.append(sion)
                    lengths.append(1)
                    bounds.append(impurity_bounds[np.where([t in sion for t in impurity_tags])[0][0]])
        if "X_lines" in d:
            for line in d["X_lines"]:
                tags.append("X_" + str(line)[1:-1].replace(", ", "_"))
                lengths.append(1)
                bounds.append([0, 20])
        resolved = {}
        for tag in tags:
            applies = any(tag.endswith(t) for t in impurity_tags)
            resolved[tag] = ((tag.endswith("_velocity") and d.get("ion_resolved_velocity", False))
                             or (tag.endswith("_Ti") and d["ion_resolved_temperatures"])
                             or (tag.endswith("_tau") and d["ion_resolved_tau"]) or not applies)
        slices, keep = [], []
        for p, L in zip(tags, lengths):
            keep.extend([resolved[p]] * L)
            if len(slices) == 0:
                slices.append((p, slice(0, L)))
            elif resolved[p]:
                last = slices[-1][1].stop
                slices.append((p, slice(last, last + L)))
            else:
                suffix = [t for t in impurity_tags if p.endswith(t)][0]
                shared = [s for tag, s in slices if tag.startswith(p[:1]) and tag.endswith(suffix)]
                if shared:
                    slices.append((p, shared[0]))
                else:
                    last = slices[-1][1].stop
                    slices.append((p, slice(last, last + L)))
                    keep[-1] = True
        self.tags = tags
        self.n_params = slices[-1][1].stop
        self.slices = collections.OrderedDict(slices)
        self.theta_bounds = np.array([np.array(b) for n, b in enumerate(bounds) if keep[n]], dtype=float)
        if self.n_params != len(self.theta_bounds):
            raise ValueError("self.n_params!=len(self.theta_bounds)")

    # ------------------------------------------------------------------------------------ chord set-up
    def _build_chord(self, c, refine):
        """spectrometers.py:45-97 (+ get_error :115-152, get_lines :337-382, int_func_sparce_matrix :384-430)."""
        d = self.input_dict
        ch = type("Chord", (), {})()
        ch.number = c
        ch.y_data = np.asarray(d["experimental_emission"][c], dtype=float)
        ch.x_data = np.asarray(d["wavelength_axis"][c], dtype=float)
        if refine is not None:
            low, up = ch.x_data.min(), ch.x_data.max() + refine
            ch.x_data_fm = np.arange(low, up, refine)
        else:
            ch.x_data_fm = ch.x_data
        ch.dispersion = np.diff(ch.x_data)[0]
        ch.dispersion_fm = np.diff(ch.x_data_fm)[0]
        ch.dispersion_ratios = ch.dispersion_fm / ch.dispersion
        ch.calwave_default = np.array([ch.x_data.mean(), np.diff(ch.x_data).mean()])
        # error model
        _, ch.y_data_continuum = clip_data(ch.x_data, ch.y_data, d["noise_region"][c])
        a_cal = d["emission_constant"][c]
        ch.error = np.sqrt(np.square(np.std(ch.y_data_continuum)) + ch.y_data * a_cal + 0.0 * ch.y_data)
        if np.isnan(ch.error).any():
            raise TypeError("NaNs in self.error!")
        # lines
        ch.lines = []
        for species in d["species"]:
            elem = species[: species.index("_")]
            for l in d[species]:
                if within(l, [ch.x_data]):
                    if elem in HYDROGEN_ISOTOPES:
                        ch.lines.append(self._make_balmer(species, l, ch.x_data_fm))
                    else:
                        ch.lines.append(self._make_adas406(species, l, ch.x_data_fm))
        if "X_lines" in d:
            for l, f in zip(d["X_lines"], d["X_fractions"]):
                if within(l, [ch.x_data]):
                    tag = "X" + "".join("_" + str(cw) for cw in l)
                    ch.lines.append(dict(kind="x", tag=tag, cwl=np.array(l, dtype=float), fractions=f,
                                         fwhm=np.diff(ch.x_data)[0]))
        # instrument function on the fine grid
        inst = centre_peak(np.asarray(d["instrument_function"][c], dtype=float))
        inst = np.true_divide(inst, inst.sum())
        n = len(inst)
        if n % 2 == 0:
            if_x = np.linspace(-(n // 2 - 0.5), n // 2 - 0.5, n)
        else:
            if_x = np.linspace(-n // 2, n // 2, n)  # NB -n // 2 == -(n + 1) / 2 for odd n (Q1)
        interp = interp1d(if_x, inst, bounds_error=False, fill_value=0.0)
        res = ch.dispersion_ratios * np.diff(if_x)[0]
        if_x_fm = np.arange(if_x.min(), if_x.max(), res)
        inst_fm = interp(if_x_fm)
        inst_fm = np.true_divide(inst_fm, inst_fm.sum())
        ch.instrument_function = inst
        ch.instrument_function_fm = centre_peak(inst_fm)
        return ch

    def _make_balmer(self, species, cwl, wavelengths):
        """linemodels.py:729-756 and HydrogenLineShape.__init__ :610-686."""
        rec = self.input_dict[species][cwl]
        key = str(rec["n_upper"]) + str(rec["n_lower"])
        if key not in LOMAN:
            raise KeyError(key)
        dlambda = np.diff(wavelengths).mean()
        name = species + "_" + str(cwl)
        return dict(kind="balmer", species=species, name=name, cwl=cwl, loman=LOMAN[key],
                    mass=ATOMIC_MASS[species.split("_")[0]],
                    doppler_x=np.arange(cwl - 10, cwl + 10 + dlambda, dlambda),
                    exc_pec=self.hydrogen_pecs[name + "_exc"], rec_pec=self.hydrogen_pecs[name + "_rec"])

    def _make_adas406(self, species, cwls, wavelengths):
        """linemodels.py:334-384."""
        rec = self.input_dict[species][cwls]
        cw = list(cwls) if type(cwls) in (tuple, list, np.ndarray) else [cwls]
        name = species + "_" + str(cw)[1:-1].replace(", ", "_")
        lines, fracs = [], []
        for c, f in zip(cw, rec["jj_frac"] if type(rec["jj_frac"]) in (list, tuple) else [rec["jj_frac"]]):
            if not (c < min(wavelengths) or c > max(wavelengths)):
                lines.append(c)
                fracs.append(f)
        return dict(kind="adas", species=species, name=name, cwls=np.array(cw, dtype=float),
                    lines=np.array(lines, dtype=float), fractions=fracs,
                    mass=ATOMIC_MASS[species[: species.index("_")]], tec=self.impurity_tecs[name])

    # ------------------------------------------------------------------------------------ priors / bounds
    def _build_prior_list(self):
        """posterior.py:108-165 — the default component order after the chords."""
        p = []
        if self.curvature is not None:  # posterior.py:109-110
            p.append(("CurvatureCost", dict(scale=self.curvature)))
        if self.calibrate_wavelength:
            for num in range(self.num_chords):
                inmean = self.chords[num].x_data.mean()
                indisp = np.diff(self.chords[0].x_data).mean()
                p.append(("WavelengthCalibrationPrior", dict(mean=np.array([inmean, indisp]),
                                                             std=np.array([2 * indisp, indisp * 0.05]))))
        if self.calibrate_intensity:
            p.append(("CalibrationPrior", {}))
        if self.calibrate_instrument_function:
            p.append(("GaussianInstrumentFunctionPrior", {}))
        p.append(("StaticElectronPressureCost", {}))
        if self.contains_hydrogen:
            p.append(("NeutralFractionCost", {}))
            first = [(c, k) for c, ch in enumerate(self.chords) for k, l in enumerate(ch.lines)
                     if l["kind"] == "balmer"]
            if first:
                p.append(("StarkToPeakPrior", dict(chord=first[0][0], line=first[0][1])))
        if len(self.impurities) > 0:
            p.append(("AntiprotonCost", {}))
            p.append(("MainIonFractionCost", {}))
            index = {}
            for c, ch in enumerate(self.chords):  # priors.py:510-524
                for k, l in enumerate(ch.lines):
                    if "species" in l and l["species"] in self.impurity_species and l["species"] not in index:
                        index[l["species"]] = (c, k)
            p.append(("ChargeStateOrderPrior", dict(index=index)))
        # posterior.py:163-165: passed priors come last (`got_p` compares a type with an instance, so they are
        # appended even when a default component of the same class exists)
        p.extend((name, dict(args)) for name, args in self.extra_priors)
        self.priors = p
        self.component_names = ["SpectrometerChord"] * self.num_chords + [n for n, _ in p]

    def _set_sensible_bounds(self, dcwl=2, ddisp=0.005):
        """posterior.py:271-356."""
        tb, sl = self.theta_bounds, self.slices
        for c, ch in enumerate(self.chords):
            if f"cal_{c}" in sl:
                tb[sl[f"cal_{c}"]] = [-0.3, 0.3] if ch.y_data.max() > 1e6 else [8, 15]
            if f"calwave_{c}" in sl:
                cwl, disp = ch.x_data.mean(), np.diff(ch.x_data).mean()
                tb[sl[f"calwave_{c}"]][0] = [cwl - dcwl, cwl + dcwl]
                tb[sl[f"calwave_{c}"]][1] = [disp - ddisp, disp + ddisp]
            est, dbg = ch.y_data_continuum.mean(), ch.y_data_continuum.std()
            bb = [max(est - dbg, est / 10), min(est + dbg, est * 10)]
            tb[sl[f"background_{c}"]] = [np.log10(b) for b in bb]
        for s in sl:
            if s.endswith("_dens"):
                charge = float(s.split("_")[-2])
                new_limit = tb[sl["electron_density"]][0, 1]
                if not charge == 0:
                    new_limit -= np.log10(charge)
                    nb_range = 2
                else:
                    nb_range = 3
                tb[sl[s]][0] = [new_limit - nb_range, new_limit]
            if not self.thermalised and s.endswith("_Ti"):
                tb[sl[s]][0] = [0, 1.7]
        for ch in self.chords:
            for l in ch.lines:
                if l["kind"] == "x":  # linemodels.py:189-207
                    half_width = 2 * np.diff(ch.x_data).mean()
                    est = []
                    for cwl in l["cwl"]:
                        nx, ny = clip_data(ch.x_data, ch.y_data, [cwl - half_width, cwl + half_width])
                        est.append(trapz(ny, nx))
                    e = np.log10(sum(est))
                    tb[sl[l["tag"]]] = [e - 1.0, e + 0.5]
        if np.isnan(tb).any():
            raise ValueError("NaNs in theta bounds!")

    def random_start(self, rng=np.random):
        """posterior.py:358-395 (``order=1, flat=False``), drawing from the global numpy RNG like the reference."""
        start = [np.mean(rng.uniform(b[0], b[1], size=1)) for b in self.theta_bounds]
        for ch in self.chords:
            mean = np.log10(np.mean(ch.y_data_continuum))
            start[self.slices[f"background_{ch.number}"].start] = rng.uniform(mean - 0.1, mean + 0.1)
        return np.array(start)

    # ------------------------------------------------------------------------------------ evaluation
    def plasma_state(self, theta):
        """plasmas.py:642-747 as a pure function: theta -> dict."""
        theta = np.asarray(theta, dtype=float)
        if len(theta) != self.n_params:
            raise ValueError("len(theta)!=self.n_params")
        st = {}
        for p, slc in self.slices.items():
            v = theta[slc].flatten()
            if p == "electron_density":
                st["_theta_electron_density"] = v.copy()
                v = self.profile_function.electron_density(v)
            elif p == "electron_temperature":
                v = self.profile_function.electron_temperature(v)
            elif p.startswith("calwave"):
                pass
            elif p.endswith("_velocity"):
                v = 1e3 * v
            elif p in ("b-field", "viewangle"):
                pass
            else:  # cal, calint_func, background, species dens/Ti/tau, X lines -> 10**theta
                v = np.power(10, v)
            st[p] = v
            st.setdefault("_plasma_theta", {})[p] = theta[slc].flatten()  # plasma.plasma_theta (raw slices)
        ne = st["electron_density"]
        total = 0
        for elem in self.impurities:  # plasmas.py:876-889 (concstar)
            sp = [s for s in self.impurity_species if s.startswith(elem + "_")][0]
            conc = (st[sp + "_dens"] / ne.max()) * ne
            total = total + (conc / ne.max()) * ne
        st["total_impurity_electrons"] = np.nan_to_num(total)
        if self.reverse_electroneutrality:  # plasmas.py:667-671
            st["main_ion_density"] = ne.copy()
            ne = st["electron_density"] = ne + st["total_impurity_electrons"]
        else:
            st["main_ion_density"] = ne - st["total_impurity_electrons"]
        for sp in self.impurity_species:  # get_impurity_electrons / impurity_power index the tau grid (Q16)
            lt = np.log10(st[sp + "_tau"][0])
            if not (lt > self.magical_tau[0] and lt <= self.magical_tau[-1]):
                raise OracleError("ValueError", f"tau of {sp} outside the ADAS tau grid")
        if self.contains_hydrogen:
            sp = self.hydrogen_species[0]
            te = st["electron_temperature"]
            n1 = st["main_ion_density"]
            n0 = st[sp + "_dens"][0] if (sp + "_dens") in self.slices else n1.copy()
            rates = {k: bilinear_extrap(g_ne, g_te, tab, ne, te) for k, (g_ne, g_te, tab) in self.neutral_adf11.items()}
            tau = st[sp + "_tau"][0]
            if self.recombining is True:  # plasmas.py:735-737
                n0 = n0 + n1 * ne * rates["acd"] * tau
            n0 = n0 - n0 * ne * rates["scd"] * tau
            st[sp + "_dens"] = n0.clip(1)
            st["adf11"] = rates
        return st

    def _balmer(self, line, st, wavelengths):
        """linemodels.py:758-865 with HydrogenLineShape.__call__ :688-723 and stehle_param :554-566."""
        sp = line["species"]
        n1, ne, te, n0 = st["main_ion_density"], st["electron_density"], st["electron_temperature"], st[sp + "_dens"]
        if np.isnan(n0).any():
            raise OracleError("ValueError", "Negative numbers in neutral density profile!")
        cwl = line["cwl"]
        doppler_x = line["doppler_x"]
        if sp + "_velocity" in st:
            new_cwl = cwl * (st[sp + "_velocity"][0] / speed_of_light + 1)
            doppler_x = doppler_x + (new_cwl - cwl)  # reducedx[0] += cwl - old_cwl (Q8)
            cwl = new_cwl
        rec_pec = np.exp(bilinear_extrap(self.h_ne, self.h_te, line["rec_pec"], ne, te))
        exc_pec = np.exp(bilinear_extrap(self.h_ne, self.h_te, line["exc_pec"], ne, te))
        if np.isnan(exc_pec).any() or np.isnan(rec_pec).any():
            raise OracleError("ValueError", "NaNs in the PECs!")
        rec_profile = n1.clip(1) * ne * rec_pec
        exc_profile = n0 * ne * exc_pec
        rec_w = rec_profile / rec_profile.sum()
        exc_w = exc_profile / exc_profile.sum()
        o = dict(rec_sum=trapz(rec_profile, x=self.los) / (4 * pi), exc_sum=trapz(exc_profile, x=self.los) / (4 * pi))
        ems_profile = rec_profile + exc_profile
        o["emission_fitted"] = trapz(ems_profile, x=self.los) / (4 * pi)
        o["exc_ne"], o["exc_te"] = np.dot(exc_w, ne), np.dot(exc_w, te)
        o["rec_ne"], o["rec_te"] = np.dot(rec_w, ne), np.dot(rec_w, te)
        o["f_rec"] = o["rec_sum"] / o["emission_fitted"]
        o["ems_ne"] = np.dot(ems_profile, ne) / ems_profile.sum()
        o["ems_te"] = np.dot(ems_profile, te) / ems_profile.sum()
        if self.cold_neutrals:
            ti = 0.01
        elif self.thermalised:
            ti = (1 - o["f_rec"]) * o["exc_te"] + o["f_rec"] * o["rec_te"]
        else:
            ti = st[sp + "_Ti"][0]
        o["exc_ti"] = ti

        def shape(ne_w, te_w):
            fwhm = cwl * 7.715e-5 * np.sqrt(ti / line["mass"])
            doppler = gaussian_norm(doppler_x, cwl, fwhm, 1.0) * 1
            a_ij, b_ij, c_ij = line["loman"]
            gamma = 10.0 * c_ij * np.divide((1e6 * ne_w) ** a_ij, te_w**b_ij) / 2.0
            ls = 1 / (abs(wavelengths - cwl) ** 2.5 + gamma**2.5)
            stark = ls / trapz(ls, wavelengths)
            peak = fftconvolve(stark, doppler, "same")
            return peak / trapz(peak, wavelengths)

        exc_peak = np.nan_to_num(shape(o["exc_ne"], o["exc_te"])) * o["exc_sum"]
        rec_peak = np.nan_to_num(shape(o["rec_ne"], o["rec_te"])) * o["rec_sum"]
        peak = rec_peak + exc_peak
        if np.isnan(peak).any():
            raise OracleError("ValueError", f"NaNs in {line['name']} line peak")
        return peak, o

    def _adas406(self, line, st, wavelengths):
        """linemodels.py:386-443 with get_tec :445-494."""
        sp = line["species"]
        ne, te = st["electron_density"], st["electron_temperature"]
        n0 = st[sp + "_dens"][0] * (ne / ne.max())
        tau = st[sp + "_tau"][0]
        if self.ne_tau:  # linemodels.py:398-399: tau becomes an array over the line of sight ...
            tau = 1e3 * tau / (ne * 1e-13)
            if np.size(tau) > 1:  # ... and get_tec (:454) takes the truth value of an array comparison
                raise OracleError("ValueError", "The truth value of an array with more than one element is ambiguous")
        cwls = line["lines"]
        if sp + "_velocity" in st:
            cwls = line["cwls"] * (st[sp + "_velocity"][0] / speed_of_light + 1)
            if len(cwls) != len(line["lines"]):  # linemodels.py:244-249 (Q6)
                raise OracleError("ValueError", "Transition has a multiplet member outside the chord")
        j = self.magical_tau.searchsorted(np.log10(tau))
        i = j - 1
        if not (i < len(self.magical_tau) and j < len(self.magical_tau)):
            raise OracleError("ValueError", "tau indices out of bounds")
        da = [self.magical_tau[i], np.log10(tau), self.magical_tau[j]]
        i_w, j_w = 1 + np.diff(da) / (da[0] - da[-1])
        tec = np.nan_to_num(np.exp(line["tec"][i].ev(ne, te)) * i_w + np.exp(line["tec"][j].ev(ne, te)) * j_w)
        profile = (n0 * tec).clip(1)
        o = dict(emission_fitted=trapz(profile, x=self.los) / (4 * pi))
        o["ems_ne"] = np.dot(profile, ne) / profile.sum()
        o["ems_conc"] = np.dot(profile, n0 / ne) / profile.sum()
        o["ems_te"] = np.dot(profile, te) / profile.sum()
        if self.cold_ions:
            ti = 0.01
        elif self.thermalised:
            ti = o["ems_te"]
        else:
            ti = st[sp + "_Ti"][0]
        o["ti"] = ti
        fwhm = cwls * 7.715e-5 * np.sqrt(ti / line["mass"])
        peak = gaussian_multiplet(wavelengths, cwls, fwhm, line["fractions"], o["emission_fitted"])
        if np.isnan(peak).any() or np.isinf(peak).any():
            raise OracleError("TypeError", f"NaNs/infs in peaks of {line['name']}")
        return peak, o

    def forward_model(self, ch, st, want=None):
        """spectrometers.py:197-335."""
        c = ch.number
        x_fm = ch.x_data_fm
        spectra, scalars = 0, []
        for l in ch.lines:
            if l["kind"] == "balmer":
                pk, o = self._balmer(l, st, x_fm)
            elif l["kind"] == "adas":
                pk, o = self._adas406(l, st, x_fm)
            else:
                ems = st[l["tag"]].mean()
                pk = gaussian_multiplet(x_fm, l["cwl"], [l["fwhm"]], l["fractions"], ems)
                o = dict(emission_fitted=ems)
            spectra = spectra + pk.flatten()
            scalars.append(o)
        background = st[f"background_{c}"]
        cal = st[f"cal_{c}"] if self.calibrate_intensity else np.array([1.0])
        if want is not None:
            want["lines_sum"] = spectra.copy()
        spectra = spectra / cal[0]
        if self.calibrate_instrument_function:
            fwhm = st[f"calint_func_{c}"][0]
            inst = gaussian_norm(x_fm, x_fm.mean(), fwhm, 1)
            inst = inst / inst.sum()
            if not np.isclose(inst.sum(), 1.0):
                raise OracleError("ValueError", "Instrument function is not pixel normalised!")
        else:
            inst = ch.instrument_function_fm
        pre = trapz(spectra, x_fm)
        spectra = fftconvolve(spectra, inst, mode="same").clip(spectra.min())
        if np.isinf(spectra).any() or np.isnan(spectra).any():
            raise OracleError("TypeError", "spectra contains infs/NaNs")
        post = trapz(spectra, x_fm)
        spectra = spectra / (post / pre)
        ratio2 = trapz(spectra, x_fm) / pre
        if not np.isclose(ratio2, 1, rtol=1e-2):
            raise OracleError("ValueError", "Area not conserved in convolution!")
        spectra = spectra + background
        if want is not None:
            want.update(prewavecal=spectra.copy(), preconv_integral=pre, postconv_integral=post)
        cal_theta = st[f"calwave_{c}"] if self.calibrate_wavelength else ch.calwave_default
        pix = np.arange(len(ch.x_data), dtype=float)
        pix -= np.mean(pix)
        pix *= cal_theta[1]
        pix += cal_theta[0]
        out = interp1d(x_fm, spectra, bounds_error=False, fill_value="extrapolate")(pix).clip(background)
        if np.isnan(out).any() or np.isinf(out).any():
            raise OracleError("TypeError", "spectra contains NaNs/infs")
        return out, scalars

    def likelihood(self, ch, fm, cauchy=True):
        """spectrometers.py:154-195 (Cauchy by default, :187-188 the Gaussian form)."""
        r2 = ((ch.y_data - fm) / ch.error) ** 2
        if (r2 == 0).any():
            raise OracleError("ValueError", "Some points of square_normalised_residuals are zero")
        if np.isinf(r2).any() or np.isnan(r2).any():
            raise OracleError("ValueError", "Some points of square_normalised_residuals are inf/nan")
        return -np.log(1 + r2).sum() if cauchy else -0.5 * r2.sum()

    def prior(self, name, args, st, line_scalars):
        """priors.py — the default components (`posterior.py:108-161`)."""
        ne, te = st["electron_density"], st["electron_temperature"]
        if name == "CurvatureCost":  # priors.py:110-127 (raw theta at the density knots between the fixed end values)
            fn = self.profile_function.electron_density
            values = fn.empty_theta.copy()
            values[fn.slice] = st["_theta_electron_density"]
            return curvature(values) * args["scale"]
        if name == "StaticElectronPressureCost":  # priors.py:51-64
            pe = te.clip(0.01) * ne.clip(1)
            return sum([low_pass(f, 1, 0.1) for f in pe / 5e15])
        if name == "NeutralFractionCost":  # priors.py:83-101
            f0 = st[self.hydrogen_species[0] + "_dens"].clip(1) / ne
            return sum([high_pass(f, 1, 0.1) for f in f0 / 5e-3])
        if name == "StarkToPeakPrior":  # priors.py:646-667
            o = line_scalars[args["chord"]][args["line"]]
            ne_peak = ne.max()
            exc = high_pass(o["exc_ne"] / ne_peak, 0.65, 0.075) * (1 - o["f_rec"])
            rec = high_pass(o["rec_ne"] / ne_peak, 0.65, 0.075) * o["f_rec"]
            return exc + rec
        if name == "AntiprotonCost":  # priors.py:14-24
            n_i = st["main_ion_density"]
            if any(n_i < 0):
                return -0.5 * np.square(n_i.clip(max=0) / 1e11).sum()
            return 0.0
        if name == "MainIonFractionCost":  # priors.py:27-48
            nec = ne.clip(1)
            n_ion = st["main_ion_density"]
            factor = nec.copy()
            factor /= factor[np.argmax(n_ion)]
            factor = factor.clip(0.3)
            return sum([high_pass(f, k * 0.8, 0.1) for f, k in zip(n_ion / nec, factor)])
        if name == "ChargeStateOrderPrior":  # priors.py:509-556
            cost = 0
            idx = args["index"]
            for elem in self.impurities:
                hist = []
                for ion in sorted([ion for ion in idx if ion.startswith(elem)], key=lambda x: idx[x]):
                    c, k = idx[ion]
                    hist.append((ion, line_scalars[c][k]["ems_te"]))
                dts = np.diff([v for _, v in sorted(hist, key=lambda x: float(x[0].split("_")[1]))])
                for dt in dts:
                    cost += high_pass(dt, 2.0, 0.2)
            return cost
        if name == "WavelengthCalibrationPrior":  # priors.py:1046-1060 (always reads calwave_0, Q11)
            return (-0.5 * np.square((args["mean"] - st["calwave_0"]) / args["std"])).sum()
        if name == "CalibrationPrior":  # priors.py:1063-1076
            return (-0.5 * np.square((1 - st["cal_0"]) / 0.2)).sum()
        if name == "GaussianInstrumentFunctionPrior":  # priors.py:1079-1089
            return -0.5 * np.square((st["calint_func_0"][0] - 0.8) / 0.1)
        # ---- optional priors a user passes in `priors=[...]` (posterior.py:163-165); defaults from their signatures
        raw = st["_plasma_theta"]
        if name == "ElectronDensityTDVCost":  # priors.py:67-80
            ne_tdv = abs(np.diff(ne))
            ni_tdv = abs(np.diff(st["main_ion_density"]))
            mean = (ne_tdv - ni_tdv) / ni_tdv
            return low_pass(mean.sum() / args.get("threshold", 1), 1, args.get("sigma", 0.2))
        if name == "PeakTePrior":  # priors.py:241-251
            return low_pass(te.max(), args.get("te_min", 1), args.get("te_err", 1))
        if name == "TauPrior":  # priors.py:254-301 (tau_difference: -diff(log10 tau) between consecutive ions of an element)
            taus, diffs = [], []
            for imp in self.impurities:
                ions = [ion for ion in self.species if imp + "_" in ion]
                for ion in ions:
                    charge = ion.split("_")[1]
                    taus.append(st[imp + "_" + str(charge) + "_tau"][0])
                for dtau in -np.diff(np.log10(taus[-len(ions):])):
                    diffs.append(dtau)
            return np.array([low_pass(d, args.get("mean", 0), args.get("sigma", 0.2)) for d in diffs]).sum()
        if name == "WallConditionsPrior":  # priors.py:432-456
            te_err = 0.2 if args.get("te_err") is None else args["te_err"]
            ne_err = 0.2 if args.get("ne_err") is None else args["ne_err"]
            cost = 0
            for t in [te[0], te[-1]]:
                cost += low_pass(t / args.get("te_min", 0.5), 1, te_err)
            for nn in [ne[0], ne[-1]]:
                cost += low_pass(nn / args.get("ne_min", 2e12), 1, ne_err)
            return cost
        if name == "SimpleWallPrior":  # priors.py:459-481
            te_min, ne_min = args.get("te_min", 0.5), args.get("ne_min", 2e12)
            te_err = 0.5 * te_min if args.get("te_err") is None else args["te_err"]
            ne_err = 0.5 * ne_min if args.get("ne_err") is None else args["ne_err"]
            return 0 + low_pass(te[-1], te_min, te_err) + low_pass(ne[-1], ne_min, ne_err)
        if name == "TeMinPrior":  # priors.py:670-680
            return high_pass(te.max() / te.min(), args.get("ratio", 2.0), args.get("sigma", 0.05))
        if name == "TeTauPrior":  # priors.py:696-727
            te_max = te.max()
            bounds = [-5, 0] if te_max > 5 else ([3, 7] if te_max < 3 else [-3, 7])
            sig = args.get("sigma", 0.3)
            cost = 0
            for z in self.impurity_species:
                log_tau = np.log10(st[z + "_tau"][0])
                cost += 0 + high_pass(log_tau, bounds[0], sig) + low_pass(log_tau, bounds[1], sig)
            return cost
        if name == "ImpurityTauPrior":  # priors.py:395-429 (raw tau parameters of equal charge states across elements)
            groups = {}
            for sp in self.impurity_species:
                z0, z = sp.split("_")
                groups.setdefault(z, []).append(z0)
            cost = 0
            for charge, elems in groups.items():
                if len(elems) > 1:
                    vals = [raw[z0 + f"_{charge}" + "_tau"][0] for z0 in elems]
                    cost += -0.5 * (np.square(np.diff(vals) / args.get("sigma", 0.3))).sum()
            return cost
        if name == "BowmanTeePrior":  # priors.py:484-506
            sigma_diff = raw["electron_temperature"][1] - raw["electron_density"][2]
            return 0.0 + low_pass(sigma_diff, 0.0, args.get("sigma_err", 0.2))
        # ---- the optional priors the product evaluates on the host (baysar_b200/host_priors.py)
        fns = {"electron_density": self.profile_function.electron_density,
               "electron_temperature": self.profile_function.electron_temperature}

        def knot_vector(tag):  # priors.py:187-208: the mesh parameter vector with the sampled knots filled in
            fn = fns[tag]
            v = fn.empty_theta.copy()
            v[fn.slice] = raw[tag]
            return v

        if name == "FlatnessPriorBasic":  # priors.py:133-152
            return -np.power(10, args["scale"]) * abs(np.gradient(st[args["tag"]])).sum()
        if name == "FlatnessPrior":  # priors.py:180-208
            return -np.power(10, args["scale"]) * abs(np.gradient(knot_vector(args["tag"]))).sum()
        if name == "PlasmaGradientPrior":  # priors.py:155-177 (knot vectors as the preceding priors left them: current theta)
            cost = 0
            for tag, scale in (("electron_density", args["ne_scale"]), ("electron_temperature", args["te_scale"])):
                cost += -scale * abs(np.gradient(knot_vector(tag), fns[tag].x_points)).sum()
            return cost
        if name == "SeparatrixTePrior":  # priors.py:211-238
            parts = []
            for tag, scale in (("electron_temperature", args["scale"]), ("electron_density", args.get("ne_scale", 0))):
                v = np.array(raw[tag])
                parts.append(scale * abs(fns[tag].x[np.where(v == v.max())]).max() * (v - v.max())[args["index"]])
            return parts[0] + parts[1]
        if name == "NIIITauPrior":  # priors.py:325-344
            return high_pass(st["N_1_tau"][0] / st["N_2_tau"][0], args.get("mean", 1), args.get("sigma", 0.3))
        if name == "NIVTauPrior":  # priors.py:304-322
            r, mean, sigma, ratio = st["N_3_tau"][0] / st["N_2_tau"][0], args.get("mean", 1), args.get("sigma", 0.2), args.get("ratio", 5)
            return high_pass(r, mean, sigma) + low_pass(r, ratio * mean, ratio * sigma)
        if name == "CiiiPrior":  # priors.py:559-573
            if "C_2_tau" not in st:
                return 0
            r_tau = np.log10(st["C_2_tau"].mean() / st["N_2_tau"].mean())
            return -0.5 * np.square((r_tau - args.get("mean", 0)) / args.get("sigma", 0.2))
        if name == "EmsTeOrderPrior":  # priors.py:347-391: first line of each wanted species, highest charge first
            found = {}
            for c, ch in enumerate(self.chords):
                for k, l in enumerate(ch.lines):
                    sp = l.get("species")
                    if sp in self.impurity_species and sp in args["species"] and sp not in found:
                        found[sp] = (sp, c, k)
            order = sorted(found.values(), key=lambda rec: -int(rec[0].split("_")[1]))
            ems_te = [line_scalars[c][k]["ems_te"] for _, c, k in order]
            return high_pass(ems_te[0] - ems_te[1], args.get("mean", 2.0), args.get("sigma", 0.2))
        if name == "BolometryPrior":  # priors.py:595-643 on the side outputs of plasmas.py:997-1020, :1083-1096, :1137-1192
            trapz_ = trapz
            total = []
            if self.contains_hydrogen:
                n1 = st["main_ion_density"]
                for sp in self.hydrogen_species:  # (exc, rec) per species; only the first species has a "_rec" entry
                    total.append(trapz_(ne * st[sp + "_dens"] * st["adf11"]["plt"], self.los).sum())
                    total.append(trapz_(ne * n1 * st["adf11"]["prb"], self.los).sum())
            splines = self._radiated_power_splines()
            for elem in self.impurities:
                own = [s_ for s_ in self.impurity_species if s_.startswith(f"{elem}_")]
                ref = elem + "_2" if elem + "_2" in self.impurity_species else own[-1]
                tau = np.log10(st[ref + "_tau"])
                upper = self.magical_tau.searchsorted(tau)
                upper_tau, lower_tau = self.magical_tau[upper][0], self.magical_tau[upper - 1][0]
                upper_weight = (tau - lower_tau) / (upper_tau - lower_tau)
                lower_weight = 1 - upper_weight
                power = []
                for charge in range(ATOMIC_NUMBER[elem]):
                    sp_ = splines[f"{elem}_{charge}"]
                    rates = upper_weight * np.exp(sp_[upper_tau].ev(ne, te)) + lower_weight * np.exp(sp_[lower_tau].ev(ne, te))
                    power.append(st[ref + "_dens"] * rates)
                total.append(trapz_(np.array(power), self.los).sum())
            res = (np.array(total).sum() - args.get("mean", 10)) / args.get("sigma", 0.5)
            return -0.5 * np.square(res)
        if name == "GaussianProcessPrior":  # priors.py:1129-1176
            x = fns["electron_density"].x_points
            cov = (args["A"] ** 2) * np.exp(-0.5 * np.subtract.outer(x, x) ** 2 / args["L"] ** 2)
            field = knot_vector(args["profile"]) - args["mu"]
            return -0.5 * field.T.dot(np.linalg.solve(cov, np.eye(len(cov))).dot(field))
        raise KeyError(name)

    def evaluate(self, theta, details=False):
        """posterior.py:219-258: returns a dict with logp, components, spectra, line scalars."""
        theta = list(theta)
        if any(np.isnan(theta)):
            raise OracleError("TypeError", "Theta contains NaNs")
        if any(np.isinf(theta)):
            raise OracleError("TypeError", "Theta contains infinities")
        st = self.plasma_state(theta)
        comps, spectra, scalars, extra = [], [], [], []
        for ch in self.chords:
            want = {} if details else None
            fm, sc = self.forward_model(ch, st, want)
            spectra.append(fm)
            scalars.append(sc)
            extra.append(want)
            comps.append(self.likelihood(ch, fm))
            if details:
                want["gauss_likelihood"] = self.likelihood(ch, fm, cauchy=False)
        for name, args in self.priors:
            comps.append(self.prior(name, args, st, scalars))
        prob = sum(comps)
        if np.isnan(prob):
            raise OracleError("ValueError", "logP is NaNs")
        if np.isinf(prob):
            raise OracleError("ValueError", "logP is inf")
        if prob >= 0.0:
            raise OracleError("ValueError", "lopP is positive")
        return dict(logp=prob, components=np.array(comps, dtype=float), spectra=spectra, line_scalars=scalars,
                    state=st, extra=extra)

    def __call__(self, theta):
        return self.evaluate(theta)["logp"]

    def cost(self, theta):
        return -self(theta)
