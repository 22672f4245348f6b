"""Optional priors of the reference that are evaluated on the HOST (reference `baysar/priors.py`).

They read only theta, a handful of plasma-state scalars or published line attributes, so each is a few numpy operations
per evaluation; the posterior treats them like any user-supplied zero-argument callable (`posterior.py:163-165, :228`):
the device evaluates chords and device priors for the whole batch, then these are called once per theta after that
theta's state has been published (`BaysarPosterior._add_host_priors`).  The priors that cost real arithmetic per
line-of-sight point are rows of the device prior table instead (`baysar_b200/priors.py`).

Same class names, constructor signatures and values as the reference (fixtures `multi_host_priors`,
`multi_mesh_host_priors`, `multi_bolometry` are written by the reference itself).  ``BolometryPrior`` reads the
radiated-power side outputs, which `baysar_b200/side_outputs.py` evaluates on the host from the published
state; ``NeutralPowerPrior`` raises KeyError on every call, as in the reference.  Not built: ``FastStarkPrior`` (its
constructor runs an MCMC fit with inference-tools).
"""
import numpy as np

from .priors import gaussian_high_pass_cost, gaussian_low_pass_cost


def flatness_prior(profile, scale=1, x=None):
    """-scale * sum |gradient| (priors.py:133-139)."""
    d = np.gradient(profile, x) if x is not None else np.gradient(profile)
    return -scale * abs(d).sum()


def _mesh_view(plasma, tag="electron_density"):
    """Knot positions, the parameter vector with its fixed end values and the slice the sampled knots fill
    (priors.py:187-205, :1134-1140)."""
    fn = getattr(plasma.profile_function, tag)
    end = slice(1, -1) if fn.zero_bounds is not None else slice(0, len(fn.empty_theta))
    return fn.x_points, fn.empty_theta, end


class FlatnessPriorBasic:
    """priors.py:145-152: flatness of a plasma_state profile, scale given as log10."""

    def __init__(self, scale, tag, plasma):
        self.tag, self.scale, self.plasma = tag, np.power(10, scale), plasma

    def __call__(self):
        return flatness_prior(self.plasma.plasma_state.get(self.tag), self.scale)


class PlasmaGradientPrior:
    """priors.py:155-177: gradient of the two mesh profiles' parameter vectors over their knot positions."""

    def __init__(self, plasma, te_scale, ne_scale):
        self.plasma, self.te_scale, self.ne_scale = plasma, te_scale, ne_scale

    def __call__(self):
        cost = 0
        for tag, scale in (("electron_density", self.ne_scale), ("electron_temperature", self.te_scale)):
            fn = getattr(self.plasma.profile_function, tag)
            cost += flatness_prior(fn.empty_theta, scale, x=fn.x_points)
        return cost


class FlatnessPrior:
    """priors.py:180-208: flatness of the sampled knot parameters of one mesh profile."""

    def __init__(self, scale, tag, plasma):
        if tag not in ("electron_density", "electron_temperature"):
            raise ValueError("tag not in ('electron_density', 'electron_temperature')")
        self.tag, self.scale, self.plasma = tag, np.power(10, scale), plasma
        self.los, self.empty_array, self.slice = _mesh_view(plasma, tag)

    def __call__(self):
        self.empty_array[self.slice] = self.plasma.plasma_theta.get(self.tag)
        return flatness_prior(self.empty_array, self.scale)


class SeparatrixTePrior:
    """priors.py:211-238: the profile parameters at `index` should be the largest."""

    def __init__(self, scale, plasma, index, ne_scale=0):
        self.scale, self.ne_scale, self.plasma, self.separatrix_position = scale, ne_scale, plasma, index

    def __call__(self):
        cost = 0.0
        for tag, scale in (("electron_temperature", self.scale), ("electron_density", self.ne_scale)):
            v = np.array(self.plasma.plasma_theta[tag])
            x = getattr(self.plasma.profile_function, tag).x
            cost = cost + scale * abs(x[np.where(v == v.max())]).max() * (v - v.max())[self.separatrix_position]
        return cost


class NIVTauPrior:
    """priors.py:304-322: tau(N IV) / tau(N III) between `mean` and `ratio * mean`."""

    def __init__(self, plasma_state, mean=1, sigma=0.2, ratio=5):
        self.plasma, self.mean, self.sigma, self.ratio = plasma_state, mean, sigma, ratio

    def __call__(self):
        r = self.plasma["N_3_tau"][0] / self.plasma["N_2_tau"][0]
        return (gaussian_high_pass_cost(r, self.mean, self.sigma)
                + gaussian_low_pass_cost(r, self.ratio * self.mean, self.ratio * self.sigma))


class NIIITauPrior:
    """priors.py:325-344: tau(N II) / tau(N III) not below `mean`."""

    def __init__(self, plasma_state, mean=1, sigma=0.3, ratio=5):
        self.plasma, self.mean, self.sigma, self.ratio = plasma_state, mean, sigma, ratio

    def __call__(self):
        return gaussian_high_pass_cost(self.plasma["N_1_tau"][0] / self.plasma["N_2_tau"][0], self.mean, self.sigma)


class EmsTeOrderPrior:
    """priors.py:347-391: the emission-weighted Te of the higher charge state exceeds the lower one's by `mean` eV.
    The first line of each wanted species (chords in order), sorted by charge, highest first."""

    def __init__(self, posterior, mean=2.0, sigma=0.2, species=("B_1", "C_2")):
        self.posterior, self.mean, self.sigma, self.species = posterior, mean, sigma, list(species)
        found = {}
        for c, chord in enumerate(posterior.posterior_components[: posterior.plasma.num_chords]):
            for k, line in enumerate(chord.lines):
                sp = getattr(line, "species", None)
                if sp in posterior.plasma.impurity_species and sp in self.species and sp not in found:
                    found[sp] = (sp, c, k)
        self.impurity_indicies = sorted(found.values(), key=lambda rec: -int(rec[0].split("_")[1]))

    def __call__(self):
        comps = self.posterior.posterior_components
        ems_te = [comps[c].lines[k].ems_te for _, c, k in self.impurity_indicies]
        return gaussian_high_pass_cost(ems_te[0] - ems_te[1], self.mean, self.sigma)


class CiiiPrior:
    """priors.py:559-573: log10 of tau(C III) / tau(N III) is normal about `mean`."""

    def __init__(self, plasma, mean=0, sigma=0.2):
        self.plasma, self.mean, self.sigma = plasma, mean, sigma

    def __call__(self):
        st = self.plasma.plasma_state
        if "C_2_tau" not in st:
            return 0
        r_tau = np.log10(st["C_2_tau"].mean() / st["N_2_tau"].mean())
        return -0.5 * np.square((r_tau - self.mean) / self.sigma)


class GaussianProcessPrior:
    """priors.py:1129-1176: zero-mean Gaussian process (squared-exponential kernel over the density knots) on
    `plasma_theta[profile] - mu`.  The reference inverts the covariance through an LDL factorisation and raises when
    max|iK K - I| > 1e-6; the same check guards the direct inverse used here."""

    def __init__(self, mu, plasma=None, A=2.0, L=0.02, profile=None):
        self.plasma, self.tag, self.A, self.L, self.mu = plasma, profile, A, L, mu
        self.los, self.empty_array, self.slice = _mesh_view(plasma)
        cov = (A ** 2) * np.exp(-0.5 * np.subtract.outer(self.los, self.los) ** 2 / L ** 2)
        self.iK = np.linalg.solve(cov, np.eye(len(cov)))
        err = abs(self.iK.dot(cov) - np.eye(len(cov))).max()
        if err > 1e-6:
            raise ValueError("Error in construction of covariance matrix! \n max(|iKK-I|) = {} > {}".format(err, 1e-6))

    def __call__(self):
        self.empty_array[self.slice] = self.plasma.plasma_theta.get(self.tag)
        field = self.empty_array - self.mu
        return -0.5 * field.T.dot(self.iK.dot(field))


class BolometryPrior:
    """priors.py:595-643: the line-integrated radiated power (hydrogen excitation and recombination, and for every
    impurity element the charge-state-extrapolated power of its reference species) is normal about `mean`.  Reads the
    radiated-power side outputs, evaluated on the host from the published state (`baysar_b200/side_outputs.py`)."""

    def __init__(self, plasma, mean=10, sigma=0.5):
        self.plasma, self.mean, self.sigma = plasma, mean, sigma
        self.reference_species = []
        for elem in plasma.impurities:  # the <elem>_2 species if sampled, else the element's last species
            own = [s for s in plasma.impurity_species if s.startswith(f"{elem}_")]
            self.reference_species.append(elem + "_2" if elem + "_2" in plasma.impurity_species else own[-1])

    def syntetic_bolometer(self):
        trapz = getattr(np, "trapezoid", None) or np.trapz
        pl, estimate = self.plasma, []
        if pl.contains_hydrogen:
            power = pl.radiated_power.neutral_power()
            for species in pl.hydrogen_species:
                for reaction in ("exc", "rec"):
                    estimate.append(trapz(power[species + f"_{reaction}"], pl.los).sum())
        self.power_profiles = {}
        for species in self.reference_species:
            profile = pl.extrapolate_impurity_power(species)
            self.power_profiles[species.split("_")[0]] = profile.copy()
            estimate.append(trapz(profile, pl.los).sum())
        return np.array(estimate)

    def __call__(self):
        self.synthetic_measurement = self.syntetic_bolometer().sum()
        return -0.5 * np.square((self.synthetic_measurement - self.mean) / self.sigma)


class NeutralPowerPrior:
    """priors.py:576-588.  The reference reads plasma_state["power"]["D_ADAS_0_exc"], a key nothing writes (total_power
    stores "<species>_exc", plasmas.py:1187): every call raises KeyError, and so does this one."""

    def __init__(self, plasma, mean=1.0, sigma=0.5):
        self.plasma, self.mean, self.sigma = plasma, mean, sigma

    def __call__(self):
        power = self.plasma.radiated_power.neutral_power()
        self.exc_power = (getattr(np, "trapezoid", None) or np.trapz)(power["D_ADAS_0_exc"], self.plasma.los)
        return gaussian_low_pass_cost(self.exc_power, self.mean, self.sigma)


def _needs_side_outputs(name, why):
    def ctor(*args, **kwargs):
        raise NotImplementedError(f"{name}: {why}")
    ctor.__name__ = name
    return ctor


FastStarkPrior = _needs_side_outputs("FastStarkPrior", "its constructor runs an MCMC fit (inference-tools), outside the posterior path")
