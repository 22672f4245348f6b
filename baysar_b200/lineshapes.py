"""Host-side line-shape primitives and line-of-sight profile families.

`gaussian`, `gaussian_norm`, `Gaussian` keep the reference's signatures and argument checks
(`baysar/lineshapes.py:82-217`; exercised by the reference's own unit tests).  The profile-family objects carry what
the parameter handler needs (``x``, ``number_of_variables``, ``bounds``) plus a ``kind`` tag the model compiler maps
to the device implementation; their ``__call__`` is a numpy evaluation for plotting / inspection only — the
per-theta evaluation runs in the CUDA line-of-sight kernel (csrc/los_stage.cuh).
"""
import numpy as np

fwhm_to_sigma = 1 / np.sqrt(8 * np.log(2))
root_two_pi = np.sqrt(2 * np.pi)


def gaussian_check_input(x, cwl, fwhm, intensity):
    """Argument checks of the reference (lineshapes.py:82-103), including its exception types."""
    if type(x) is not np.ndarray:
        raise TypeError("type(x) is not np.ndarray")
    if not any(np.isreal(x)):
        raise TypeError("x must only contain real scalars")
    if not np.isreal(cwl):
        raise TypeError("cwl must be a real scalar")
    if not np.isreal(fwhm):
        raise TypeError("fwhm must be a real scalar")
    if fwhm <= 0:
        raise ValueError("fwhm must be a positive scalar")
    if not np.isreal(intensity):
        raise TypeError("intensity must be a real scalar")
    if intensity <= 0:
        raise ValueError("intensity must be a positive scalar")


def gaussian(x, cwl, fwhm, intensity):
    """Height-normalised Gaussian (lineshapes.py:109-120)."""
    gaussian_check_input(x, cwl, fwhm, intensity)
    sigma = fwhm * fwhm_to_sigma
    return intensity * np.exp(-0.5 * ((x - cwl) / sigma) ** 2)


def gaussian_norm(x, cwl, fwhm, intensity):
    """Area-normalised Gaussian (lineshapes.py:126-131)."""
    gaussian_check_input(x, cwl, fwhm, intensity)
    sigma = fwhm * fwhm_to_sigma
    return intensity / (root_two_pi * sigma) * np.exp(-0.5 * ((x - cwl) / sigma) ** 2)


def put_in_iterable(value):
    return value if type(value) in (tuple, list, np.ndarray) else [value]


class Gaussian:
    """Multiplet of Gaussians on a fixed axis (lineshapes.py:146-217, full-grid variant)."""

    def __init__(self, x=None, cwl=None, fwhm=None, fractions=None, normalise=True):
        self.x = x
        self.cwl = np.array(put_in_iterable(cwl))
        self.fwhm = fwhm
        self.fractions = [1 / len(self.cwl) for _ in self.cwl] if fractions is None else fractions
        self.func = gaussian_norm if normalise else gaussian

    def __call__(self, theta):
        if self.fwhm is not None:
            return self.make_peak(self.cwl, self.fwhm, theta)
        fwhm, intensity = theta
        return self.make_peak(self.cwl, fwhm, intensity)

    def make_peak(self, cwl, fwhm, intensity):
        fwhm = put_in_iterable(fwhm)
        if len(fwhm) < len(cwl) and len(fwhm) == 1:
            fwhm = [fwhm[0] for _ in cwl]
        peak = np.zeros(len(self.x))
        for frac, centre, width in zip(self.fractions, cwl, fwhm):
            peak += self.func(self.x, centre, width, frac)
        return intensity * peak


class EsymmtricProfile:
    """Asymmetric Cauchy profile: ``10^A / (1 + ((x-c)/s(x))^2)``, ``s = 0.2 f + f / (1 + exp(-(x-c)))``, floored
    (reference lineshapes.py:779-826).  Density: theta = [A, c, f]; temperature (fixed centre): [A, f, floor]."""

    kind = "esymmetric_cauchy"

    def __init__(self, x, log_peak_bounds, centre=None):
        self.x = np.asarray(x, dtype=float)
        self.centre = centre
        self.bounds = [list(log_peak_bounds)]
        if centre is None:
            self.bounds.append([-5, 2])
        self.bounds.append([1, 6])
        if centre is not None:
            self.bounds.append([0.2, 3])
        self.number_of_variables = len(self.bounds)

    def __call__(self, theta):
        if self.centre is None:
            amp, centre, factor = theta
            floor = 1e-3
        else:
            amp, factor, floor = theta
            centre = self.centre
        dx = self.x - centre
        width = 0.2 * factor + factor / (1 + np.exp(-dx))
        return (np.power(10, amp) / (1 + np.square(dx / width))).clip(floor)


class EsymmtricCauchyPlasma:
    """Default profile pair (reference lineshapes.py:829-852): LOS ``linspace(-10, 25, 50)`` unless given."""

    def __init__(self, x=None, dr_bounds=(-5, 2), bounds_ne=(13, 15), bounds_te=(0.0, 1.7), dr=None, **_):
        self.x = np.linspace(-10, 25, 50) if x is None else np.asarray(x, dtype=float)
        self.dr = dr
        self.electron_density = EsymmtricProfile(self.x, bounds_ne, centre=dr)
        self.electron_temperature = EsymmtricProfile(self.x, bounds_te, centre=0)
