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

    def __init__(self, x, log_peak_bounds, centre=None, ptype="cauchy"):
        if ptype not in ("cauchy", "gaussian"):
            raise ValueError("ptype not in ('cauchy', 'gaussian')")
        self.ptype = ptype  # "gaussian": exp(-0.5 ((x-c)/s)^2) instead of the Cauchy form (lineshapes.py:769-776)
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
        if self.ptype == "gaussian":
            return (np.power(10, amp) * np.exp(-0.5 * np.square(dx / width))).clip(floor)
        return (np.power(10, amp) / (1 + np.square(dx / width))).clip(floor)


class EsymmtricCauchyPlasma:
    """Default profile pair (reference lineshapes.py:829-852): LOS ``linspace(-10, 25, 50)`` unless given."""

    def __init__(self, x=None, dr_bounds=(-5, 2), bounds_ne=(13, 15), bounds_te=(0.0, 1.7), dr=None, ptype="cauchy", **_):
        self.x = np.linspace(-10, 25, 50) if x is None else np.asarray(x, dtype=float)
        self.dr = dr
        self.electron_density = EsymmtricProfile(self.x, bounds_ne, centre=dr, ptype=ptype)
        self.electron_temperature = EsymmtricProfile(self.x, bounds_te, centre=0, ptype=ptype)


class MeshLine:
    """Profile given by its values at knots, interpolated linearly onto ``arange(min(x_ends), max(x_ends),
    resolution)`` with linear extrapolation (reference lineshapes.py:235-301, ``kind="linear"``).  ``log``: the
    sampled parameters are log10 of the knot values.  With ``zero_bounds`` the reference adds two fixed end points
    at ``x_ends`` and counts them as variables (lineshapes.py:262), which its own bounds check then rejects — the
    same ValueError is raised here by the parameter handler."""

    kind = "mesh"

    def __init__(self, x, x_ends=(-10, 10), zero_bounds=None, resolution=0.1, log=False, kind="linear", **kwargs):
        self.__dict__.update(kwargs)
        if kind != "linear":
            raise NotImplementedError("only linear mesh interpolation is compiled for the GPU")
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
        if any(t == 0 for t in np.diff(self.x_points)):
            raise ValueError(f"Some of self.x_points are the same. Zero bounds = {zero_bounds}")
        self.x = np.arange(min(x_ends), max(x_ends), resolution)
        self.number_of_variables = len(self.x_points)
        self.dr = False

    def __call__(self, theta):
        values = self.empty_theta.copy()
        values[self.slice] = theta
        if self.log:
            values = np.power(10, values)
        idx = np.clip(np.searchsorted(self.x_points, self.x), 1, len(self.x_points) - 1)
        lo, hi = idx - 1, idx
        slope = (values[hi] - values[lo]) / (self.x_points[hi] - self.x_points[lo])
        return slope * (self.x - self.x_points[lo]) + values[lo]


class MeshPlasma:
    """Reference lineshapes.py:304-329."""

    def __init__(self, x=None, bounds=None, zero_bounds_ne=11, zero_bounds_te=-2, bounds_ne=(11, 16), bounds_te=(-1, 2)):
        self.electron_density = MeshLine(x=x, zero_bounds=zero_bounds_ne, x_ends=bounds, log=True,
                                         bounds=[list(bounds_ne) for _ in x])
        self.electron_temperature = MeshLine(x=x, zero_bounds=zero_bounds_te, x_ends=bounds, log=True,
                                             bounds=[list(bounds_te) for _ in x])


def bowman_tee_distribution(x, theta):
    """Reference lineshapes.py:332-343."""
    A, x0, sigma0, q, nu, k, f, b = theta
    sigma = sigma0 * np.exp(f * np.tanh(k * (x - x0)))
    z = np.power(abs((x - x0) / sigma), q)
    return A * np.power(1 + z / nu, -(nu + 1) * 0.5) + b


class BowmanTeeProfile:
    """BowmanTeeNe / BowmanTeeTe (reference lineshapes.py:354-376).  Density: theta = [log10 A, x0, sigma0, q, nu, k, f
    (, log10 b)]; temperature: centred on 0, theta = [log10 A, sigma0, q, nu, k, f (, log10 b)]."""

    kind = "bowman_tee"

    def __init__(self, x, background, centred):
        self.x, self.background, self.centred = np.asarray(x, dtype=float), background, centred

    def __call__(self, theta):
        theta = np.array(theta, dtype=float)
        theta[0] = np.power(10, theta[0])
        if self.background:
            theta[-1] = np.power(10, theta[-1])
        else:
            theta = np.concatenate((theta, np.zeros(1)))
        if self.centred:
            theta = np.concatenate((theta[:1], np.zeros(1), theta[1:]))
        return bowman_tee_distribution(self.x, theta)


class BowmanTeePlasma:
    """Reference lineshapes.py:381-440."""

    def __init__(self, x=None, bounds=None, dr_bounds=(-2, 2), bounds_ne=(11, 16), bounds_te=(-1, 2), background=False):
        self.x = np.linspace(-15, 25, 500) if x is None else np.asarray(x, dtype=float)
        self.bounds = [[1e-1, 10], [1.2, 3], [1, 5], [1, 50], [0, 2]] if bounds is None else bounds
        self.background = background
        self.electron_density = BowmanTeeProfile(self.x, background, centred=False)
        self.electron_temperature = BowmanTeeProfile(self.x, background, centred=True)
        self.bounds_ne = [list(bounds_ne), list(dr_bounds)] + [list(b) for b in self.bounds]
        self.bounds_te = [list(bounds_te)] + [list(b) for b in self.bounds]
        if background:
            self.bounds_ne.append([bounds_ne[0], 13.0])
            self.bounds_te.append([bounds_te[0], 0])
        self.electron_density.bounds, self.electron_temperature.bounds = self.bounds_ne, self.bounds_te
        self.electron_density.number_of_variables = len(self.bounds_ne)
        self.electron_temperature.number_of_variables = len(self.bounds_te)


# ---------------------------------------------------------------------------------------------------------------------
# Closed-form profile families (reference lineshapes.py:456-606, :855-948).  Every class carries ``kind = "closed_form"``
# and ``fn``, the name of the device function K1 evaluates per line-of-sight point (csrc/los_stage.cuh, layout.h
# BaysarProfileFn); ``__call__`` is the host statement of the same formula.
class _ClosedForm:
    kind = "closed_form"

    def __init__(self, x, bounds):
        self.x = np.asarray(x, dtype=float)
        self.x_len = len(self.x)
        self.bounds = [list(b) for b in bounds]
        self.number_of_variables = len(self.bounds)


class ExpDecay(_ClosedForm):
    """``10^a (1 + 10^b)^(-x)``, floored at 0.01 (lineshapes.py:483-495)."""

    fn = "PF_EXP_DECAY"

    def __init__(self, x):
        super().__init__(x, [[-1, 2], [-2, 1]])

    def __call__(self, theta):
        a, b = (np.power(10.0, t) for t in theta)
        return (a * np.power(1.0 + b, -self.x)).clip(0.01)


class ADoubleExpDecay(_ClosedForm):
    """Two-sided decay about ``x0`` with a tanh-switched base (lineshapes.py:498-518); theta = [a, b, x0, f, k]."""

    fn = "PF_DOUBLE_EXP_DECAY"

    def __init__(self, x):
        super().__init__(x, [[12, 16], [-2, 1], [0, 50], [-3, 1], [-5, 5]])

    def __call__(self, theta):
        a, b, x0, f, k = theta
        base = (1.0 + np.power(10.0, b)) * np.exp(np.power(10.0, f) * np.tanh(k * (self.x - x0)))
        return (np.power(10.0, a) * np.power(base, x0 - self.x)).clip(0.01)


class Poisson(_ClosedForm):
    """``a p / max(p)``, ``p = nu^x e^-nu / x!`` with ``a, nu = 10^theta`` (lineshapes.py:537-550)."""

    fn = "PF_POISSON"

    def __init__(self, x):
        super().__init__(x, [[12, 16], [-3, 2]])

    def __call__(self, theta):
        from scipy.special import factorial

        a, nu = (np.power(10.0, t) for t in theta)
        peak = np.power(np.zeros(self.x_len) + nu, self.x)
        peak *= np.exp(-nu) / factorial(self.x)
        return peak * (a / peak.max())


class GaussianPlasma(_ClosedForm):
    """Height-normalised Gaussian ``10^I exp(-0.5 ((x - cwl) / sigma)^2)``, ``sigma`` from ``fwhm = 10^w``, floored at
    0.01; the centre is sampled unless ``cwl`` is given (lineshapes.py:556-591)."""

    fn = "PF_GAUSSIAN"

    def __init__(self, x=None, cwl=None):
        x = np.asarray(x, dtype=float)
        self.cwl = cwl
        bounds = [[12, 16], [-1, np.log10(x.max() / 2)]]
        if cwl is None:
            bounds.append([x.min(), x.max()])
        super().__init__(x, bounds)

    def __call__(self, theta):
        if self.cwl is None:
            intensity, fwhm, cwl = theta
        else:
            (intensity, fwhm), cwl = theta, self.cwl
        return gaussian(self.x, cwl, np.power(10, fwhm), np.power(10, intensity)).clip(0.01)


class FlatProfile(_ClosedForm):
    """Constant ``10^theta`` (lineshapes.py:855-866)."""

    fn = "PF_FLAT"

    def __init__(self, x, bounds_te):
        super().__init__(x, [bounds_te])

    def __call__(self, theta):
        return np.power(10, theta[0]) + np.zeros(self.x.shape)


class LeftTopHatProfile(_ClosedForm):
    """``10^ne`` left of ``l`` plus one grid step, ``1e10`` elsewhere (lineshapes.py:869-883)."""

    fn = "PF_LEFT_TOP_HAT"

    def __init__(self, x, bounds_ne, dr_bounds):
        super().__init__(x, [bounds_ne, dr_bounds])

    def __call__(self, theta):
        ne, edge = theta
        profile = np.zeros(self.x.shape) + 1e10
        profile[np.where(self.x < edge + (self.x[1] - self.x[0]))] = np.power(10, ne)
        return profile


class LeftRightTopHatProfile(_ClosedForm):
    """``10^left`` left of the centre, ``10^right`` right of it, 0 on it (lineshapes.py:906-929).  With ``alt`` (the
    density of a DoubleSlabPlasma) the centre is a third parameter and is handed to ``alt`` -- the temperature, which is
    evaluated after the density (the order of ``plasma.slices``)."""

    fn = "PF_LEFT_RIGHT_TOP_HAT"

    def __init__(self, x, centre, bounds_left, bounds_right, alt=None):
        self.centre, self.alt = centre, alt
        super().__init__(x, [bounds_left, bounds_right] + ([[2, 8]] if alt is not None else []))

    def __call__(self, theta):
        if self.alt is not None:
            left, right, centre = theta
            self.centre = self.alt.centre = centre
        else:
            left, right = theta
        profile = np.zeros(self.x.shape)
        profile[np.where(self.x < self.centre)] = np.power(10, left)
        profile[np.where(self.centre < self.x)] = np.power(10, right)
        return profile


class SimplePlasma:
    """lineshapes.py:469-480 (the second definition, which shadows the Poisson one at :456-466)."""

    def __init__(self, x=None):
        self.x = np.linspace(0, 60, 100) if x is None else np.asarray(x, dtype=float)
        self.electron_density = ExpDecay(self.x)
        self.electron_density.bounds[0] = [12, 16]
        self.electron_temperature = ExpDecay(self.x)


class PoissonPlasma:
    """The reference's FIRST ``SimplePlasma`` (lineshapes.py:443-453: Poisson density, ExpDecay temperature).  The second
    definition at :456 shadows it there, so a reference user reaches it only by composing ``Poisson`` and ``ExpDecay`` by
    hand; here it has its own name."""

    def __init__(self, x=None):
        self.x = np.linspace(0, 60, 100) if x is None else np.asarray(x, dtype=float)
        self.electron_density = Poisson(self.x)
        self.electron_temperature = ExpDecay(self.x)


class LessSimplePlasma:
    """lineshapes.py:521-532."""

    def __init__(self, x=None):
        self.x = np.linspace(0, 60, 100) if x is None else np.asarray(x, dtype=float)
        self.electron_density = ADoubleExpDecay(self.x)
        self.electron_temperature = ExpDecay(self.x)


class SimpleGaussianPlasma:
    """lineshapes.py:594-606."""

    def __init__(self, x=None):
        self.x = np.linspace(0, 60, 100) if x is None else np.asarray(x, dtype=float)
        self.electron_density = GaussianPlasma(x=self.x)
        self.electron_temperature = GaussianPlasma(x=self.x, cwl=0)


class SlabPlasma:
    """lineshapes.py:886-901."""

    def __init__(self, x=None, dr_bounds=(-5, 2), bounds_ne=(13, 15), bounds_te=(0.0, 2.0), **_):
        self.x = np.linspace(0, 20, 200) if x is None else np.asarray(x, dtype=float)
        self.electron_density = LeftTopHatProfile(self.x, list(bounds_ne), list(dr_bounds))
        self.electron_temperature = FlatProfile(self.x, list(bounds_te))


class DoubleSlabPlasma:
    """lineshapes.py:932-948."""

    def __init__(self, x=None, centre=4, **_):
        self.x = np.linspace(0, 20, 200) if x is None else np.asarray(x, dtype=float)
        self.electron_temperature = LeftRightTopHatProfile(self.x, centre, bounds_left=[-1, 0.7], bounds_right=[0, 2.0])
        self.electron_density = LeftRightTopHatProfile(self.x, centre, bounds_left=[13, 15], bounds_right=[13, 14.7],
                                                       alt=self.electron_temperature)


class LinearSeparatrix:
    """``log10(clip(10^Te - grad x, 0.1))`` (lineshapes.py:723-734).  The separatrix profiles return log10 values and carry
    the parameters of the reference's 2-D wrapper (`PlasmaSimple2D`, plasmas.py:1468), which never hands them to a
    `PlasmaLine`; they are host callables here as they are there."""

    def __init__(self, x, peak_bounds=(0, 1.7)):
        self.x = np.asarray(x, dtype=float)
        self.bounds = [list(peak_bounds), [0.5, 1.5]]
        self.number_of_variables = len(self.bounds)

    def __call__(self, theta):
        te, grad = theta
        return np.log10((np.power(10, te) - grad * self.x).clip(0.1))


class CauchySeparatrix:
    """``log10(10^ne_min + 10^ne / (1 + ((centre - x) / sigma)^2))`` (lineshapes.py:737-757; the reference's
    ``profile.clip(1e11)`` discards its result, so there is no floor)."""

    def __init__(self, x, peak_bounds=(12.7, 16), centre=None):
        self.x = np.asarray(x, dtype=float)
        self.centre = centre
        self.bounds = [list(peak_bounds), [12.7, 14]] + ([[-5, 5]] if centre is None else []) + [[0.5, 1.5]]

    def __call__(self, theta):
        if self.centre is None:
            ne, ne_min, centre, sigma = theta
        else:
            (ne, ne_min, sigma), centre = theta, self.centre
        return np.log10(np.power(10, ne_min) + np.power(10, ne) / (1 + np.square((centre - self.x) / sigma)))


class SimpleSeparatrix:
    """lineshapes.py:760-767."""

    def __init__(self, chords, bounds_ne=(11, 16), bounds_te=(-1, 2)):
        self.x = np.array(chords)
        self.electron_density = CauchySeparatrix(self.x, list(bounds_ne))
        self.electron_temperature = LinearSeparatrix(self.x, list(bounds_te))
