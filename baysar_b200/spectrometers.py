"""``SpectrometerChord`` and its line objects — host-side set-up and result views.

Set-up follows the reference's constructor (`/root/reference/baysar/spectrometers.py:45-97`): fine wavelength grid,
error model (:115-152), line construction (:337-382), instrument-function centring / resampling (:384-430).  The
per-theta work (``forward_model`` :197-335, ``likelihood`` :154-195) runs in the CUDA chord kernel
(csrc/chord_stage.cuh); ``forward_model()`` / ``likelihood()`` / ``__call__()`` here evaluate the posterior's *last
scalar theta* on the GPU and return host values, so reference-style scripts keep working.

Line objects (``BalmerHydrogenLine``, ``ADAS406Lines``, ``XLine``) are plain descriptors; after a scalar posterior
call they carry the diagnostic attributes the reference exposes (``exc_ne``, ``rec_ne``, ``f_rec``, ``ems_te`` ...,
`baysar/linemodels.py:800-836, :407-426`).
"""
import numpy as np
from scipy.interpolate import interp1d

from .tools import centre_peak, clip_data, within

ATOMIC_MASSES = {"H": 1, "D": 2, "He": 4, "Be": 9, "B": 10.8, "C": 12, "N": 14, "O": 16, "Ne": 20.2}

# Parameterised Stark widths (a_ij, b_ij, c_ij) per "n_upper n_lower" — constants tabulated in the reference
# (`baysar/linemodels.py:533-551`, from Lomanowski et al.); data, keyed as the reference keys them.
LOMAN_COEFF = {
    "32": (0.7665, 0.064, 3.710e-18), "42": (0.7803, 0.050, 8.425e-18), "52": (0.6796, 0.030, 1.310e-15),
    "62": (0.7149, 0.028, 3.954e-16), "72": (0.7120, 0.029, 6.258e-16), "82": (0.7159, 0.032, 7.378e-16),
    "92": (0.7177, 0.033, 8.947e-16), "102": (0.7158, 0.032, 1.239e-15), "112": (0.7146, 0.028, 1.632e-15),
    "122": (0.7388, 0.026, 6.459e-16), "132": (0.7356, 0.020, 9.012e-16), "43": (0.7449, 0.045, 1.330e-16),
    "53": (0.7356, 0.044, 6.640e-16), "63": (0.7118, 0.016, 2.481e-15), "73": (0.7137, 0.029, 3.270e-15),
    "83": (0.7133, 0.032, 4.343e-15), "93": (0.7165, 0.033, 5.588e-15),
}


class BalmerHydrogenLine:
    """Descriptor of one hydrogen-isotope line (reference `linemodels.py:729-756`, `HydrogenLineShape` :610-686)."""

    fields = ("rec_sum", "exc_sum", "rec_ne", "rec_te", "exc_ne", "exc_te", "exc_ti", "f_rec", "ems_ne", "ems_te",
              "emission_fitted")

    def __init__(self, plasma, species, cwl, wavelengths):
        rec = plasma.input_dict[species][cwl]
        self.species, self.cwl, self.wavelengths = species, cwl, wavelengths
        self.element = species.split("_")[0]
        self.line = species + "_" + str(cwl)
        self.n_upper, self.n_lower = rec["n_upper"], rec["n_lower"]
        self.atomic_mass = ATOMIC_MASSES[self.element]
        self.loman_ij_abc = LOMAN_COEFF[str(self.n_upper) + str(self.n_lower)]  # KeyError for Lyman, as the reference
        dlambda = np.diff(wavelengths).mean()
        self.wavelengths_doppler = np.arange(cwl - 10, cwl + 10 + dlambda, dlambda)
        self.exc_pec = plasma.hydrogen_pecs[self.line + "_exc"]
        self.rec_pec = plasma.hydrogen_pecs[self.line + "_rec"]

    def _publish(self, record):
        for name, value in zip(self.fields, record):
            setattr(self, name, float(value))
        self.rec_ti = self.exc_ti


class ADAS406Lines:
    """Descriptor of one impurity transition / multiplet (reference `linemodels.py:322-384`)."""

    fields = ("emission_fitted", "ems_ne", "ems_conc", "ems_te", "ti")

    def __init__(self, plasma, species, cwls, wavelengths):
        rec = plasma.input_dict[species][cwls]
        self.species = species
        self.cwls = list(cwls) if type(cwls) in (tuple, list, np.ndarray) else [cwls]
        self.line = species + "_" + str(self.cwls)[1:-1].replace(", ", "_")
        self.atomic_mass = ATOMIC_MASSES[species[: species.index("_")]]
        self.jj_frac_full = rec["jj_frac"] if type(rec["jj_frac"]) in (list, tuple) else [rec["jj_frac"]]
        self.wavelengths = wavelengths
        self.lines, self.jj_frac = [], []
        for cwl, frac in zip(self.cwls, self.jj_frac_full):
            if not (cwl < min(wavelengths) or cwl > max(wavelengths)):
                self.lines.append(cwl)
                self.jj_frac.append(frac)
        if plasma.include_doppler_shifts and len(self.lines) != len(self.cwls):
            # the reference Doppler-shifts ALL catalogue members and then length-checks against the in-range subset
            # (linemodels.py:244-259), i.e. it raises on every evaluation — fail at construction instead
            raise ValueError("Transition has {} line and {} where given".format(len(self.lines), len(self.cwls)))
        if any(f <= 0 for f in self.jj_frac):
            raise ValueError("multiplet fractions must be positive (lineshapes.py:102-103)")
        self.tec406 = plasma.impurity_tecs[self.line]

    def _publish(self, record):
        for name, value in zip(self.fields, record):
            setattr(self, name, float(value))


class XLine:
    """Mystery line: fixed-FWHM normalised Gaussian multiplet scaled by 10**theta (`linemodels.py:213-233`)."""

    fields = ("emission_fitted",)

    def __init__(self, cwl, wavelengths, plasma, fractions, fwhm=0.2, species="X"):
        self.species, self.fwhm, self.fractions = species, fwhm, fractions
        self.cwl = list(cwl)
        self.wavelengths = wavelengths
        self.line_tag = species + "".join("_" + str(c) for c in cwl)
        if fwhm <= 0 or any(f <= 0 for f in fractions):
            raise ValueError("fwhm must be a positive scalar")

    def estimate_ems_and_bounds(self, wavelength, spectra, half_range=0.5):
        """Data-driven amplitude estimate and bounds (`linemodels.py:189-207`)."""
        half_width = 2 * np.diff(wavelength).mean()
        trapz = getattr(np, "trapezoid", None) or np.trapz
        areas = []
        for cwl in self.cwl:
            nx, ny = clip_data(wavelength, spectra, [cwl - half_width, cwl + half_width])
            areas.append(trapz(ny, nx))
        self.estimate = np.log10(sum(areas))
        self.bounds = [self.estimate - 2 * half_range, self.estimate + half_range]

    def _publish(self, record):
        self.emission_fitted = float(record[0])


class SpectrometerChord:
    def __init__(self, plasma, refine=None, chord_number=None):
        self.plasma, self.chord_number, self.refine = plasma, chord_number, refine
        d = self.input_dict = plasma.input_dict
        c = chord_number
        self.y_data = np.asarray(d["experimental_emission"][c], dtype=float)
        self.x_data = np.asarray(d["wavelength_axis"][c], dtype=float)
        if refine is not None:
            if not np.isreal(refine):
                raise TypeError("self.refine is not a real number")
            self.x_data_fm = np.arange(self.x_data.min(), self.x_data.max() + refine, refine)
        else:
            self.x_data_fm = self.x_data
        self.dispersion = np.diff(self.x_data)[0]
        self.dispersion_fm = np.diff(self.x_data_fm)[0]
        self.dispersion_ratios = self.dispersion_fm / self.dispersion
        self.instrument_function = np.asarray(d["instrument_function"][c], dtype=float)
        self.noise_region = d["noise_region"][c]
        self.a_cal = d["emission_constant"][c]
        self.anomalous_error = 0.0
        self.calwave_default = np.array([self.x_data.mean(), np.diff(self.x_data).mean()])
        self.convolution_tolerence = 1e-2
        self.get_error()
        self.get_lines()
        self.int_func_sparce_matrix()
        self._posterior = None

    # -- set-up -------------------------------------------------------------------------------------------------
    def get_error(self):
        """sqrt(noise variance + shot noise + anomalous) (spectrometers.py:115-152)."""
        self.x_data_continuum, self.y_data_continuum = clip_data(self.x_data, self.y_data, self.noise_region)
        self.error_components = [np.square(np.std(self.y_data_continuum)), self.y_data * self.a_cal,
                                 self.anomalous_error * self.y_data]
        if any(np.isnan(ec).any() for ec in self.error_components):
            raise TypeError("NaNs in self.error!")
        self.error = np.sqrt(sum(self.error_components))
        if np.isnan(self.error).any():
            raise TypeError("NaNs in self.error!")

    def get_lines(self):
        d = self.input_dict
        self.lines = []
        for species in d["species"]:
            elem = species[: species.index("_")]
            for l in d[species]:
                if within(l, [self.x_data]):
                    if elem in self.plasma.hydrogen_isotopes:
                        self.lines.append(BalmerHydrogenLine(self.plasma, species, l, self.x_data_fm))
                    else:
                        self.lines.append(ADAS406Lines(self.plasma, species, l, self.x_data_fm))
        for l, f in zip(d.get("X_lines", []), d.get("X_fractions", [])):
            if within(l, [self.x_data]):
                self.lines.append(XLine(cwl=l, wavelengths=self.x_data_fm, plasma=self.plasma, fractions=f,
                                        fwhm=np.diff(self.x_data)[0]))

    def int_func_sparce_matrix(self):
        """Centre and pixel-normalise the instrument function, resample it on the fine grid
        (spectrometers.py:392-430; the dense pixel matrix of :432-445 is dead code behind ``self.dot = False``)."""
        inst = centre_peak(self.instrument_function)
        inst = np.true_divide(inst, inst.sum())
        n = len(inst)
        if n % 2 == 0:
            axis = np.linspace(-(n // 2 - 0.5), n // 2 - 0.5, n)
        else:
            axis = np.linspace(-n // 2, n // 2, n)  # asymmetric for odd n: -n // 2 == -(n + 1) / 2  (Q1)
        resample = interp1d(axis, inst, bounds_error=False, fill_value=0.0)
        step = self.dispersion_ratios * np.diff(axis)[0]
        self.instrument_function = inst
        self.instrument_function_x = axis
        self.instrument_function_x_fm = np.arange(axis.min(), axis.max(), step)
        fine = resample(self.instrument_function_x_fm)
        self.instrument_function_fm = centre_peak(np.true_divide(fine, fine.sum()))

    # -- evaluation (GPU) -----------------------------------------------------------------------------------------
    def _require(self):
        if self._posterior is None or self._posterior.last_proposal is None:
            raise RuntimeError("call the posterior with a theta first; the chord evaluates posterior.last_proposal")
        return self._posterior

    def forward_model(self):
        """Spectrum on the pixel axis for the posterior's last theta (ph/cm2/A/sr/s)."""
        post = self._require()
        spectra, status = post.spectra_batch(np.asarray(post.last_proposal, dtype=float)[None], self.chord_number)
        post._raise_from_status(int(status[0]), mask=0x7FF)
        return spectra[0]

    def likelihood(self, cauchy=True):
        """logP of the spectral fit for the posterior's last theta (spectrometers.py:154-195): Cauchy by default,
        ``cauchy=False`` the Gaussian form -0.5 sum r^2 (a second device model, compiled on first use)."""
        post = self._require()
        comps, status = post.components_batch(np.asarray(post.last_proposal, dtype=float)[None], gaussian=not cauchy)
        post._raise_from_status(int(status[0]), mask=0x1FFF)
        return float(comps[0, self.chord_number])

    def __call__(self, *args, **kwargs):
        return self.likelihood()
