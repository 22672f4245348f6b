"""Default prior / cost components as descriptors.

The reference evaluates each prior as a zero-argument Python callable reading the shared ``plasma_state``
(`/root/reference/baysar/priors.py`).  Here every default component (`posterior.py:108-161`) is a small descriptor
that the model compiler turns into one row of the device prior table; the arithmetic runs in the CUDA line-of-sight
kernel (csrc/los_stage.cuh, citing priors.py per branch).  ``prior()`` returns the value at the posterior's last
scalar theta so that reference-style introspection (``for p in posterior.posterior_components: p()``) still works.
"""
import numpy as np


def gaussian_high_pass_cost(tmp, threshold, error):
    """Everything above the threshold is good (priors.py:4-6)."""
    return -0.5 * (max([0, threshold - tmp]) / error) ** 2


def gaussian_low_pass_cost(tmp, threshold, error):
    """Everything below the threshold is good (priors.py:9-11)."""
    return -0.5 * (max([0, tmp - threshold]) / error) ** 2


class _DevicePrior:
    kind = None  # name of the BaysarPriorKind enumerator in csrc/layout.h

    def __init__(self):
        self._posterior = None
        self._column = None

    def __call__(self):
        post = self._posterior
        if post is None or post.last_proposal is None:
            raise RuntimeError("call the posterior with a theta first")
        comps, _ = post.components_batch(np.asarray(post.last_proposal, dtype=float)[None])
        return float(comps[0, self._column])


class CurvatureCost(_DevicePrior):
    kind = "PRIOR_CURVATURE"  # priors.py:104-127: curvature of the raw density-knot parameters, times `scale`

    def __init__(self, plasma, scale):
        super().__init__()
        fn = plasma.profile_function.electron_density
        if getattr(fn, "kind", None) != "mesh":
            raise AttributeError("CurvatureCost needs a MeshLine density profile (x_points / empty_theta)")
        self.scale = scale


class StaticElectronPressureCost(_DevicePrior):
    kind = "PRIOR_STATIC_PRESSURE"  # priors.py:51-64 (threshold 5e15, sigma 0.1)


class NeutralFractionCost(_DevicePrior):
    kind = "PRIOR_NEUTRAL_FRACTION"  # priors.py:83-101 (threshold 5e-3, sigma 0.1)

    def __init__(self, species):
        super().__init__()
        self.species = species


class StarkToPeakPrior(_DevicePrior):
    kind = "PRIOR_STARK_TO_PEAK"  # priors.py:646-667 (mean 0.65, sigma 0.075)

    def __init__(self, line):
        super().__init__()
        self.line = line


class AntiprotonCost(_DevicePrior):
    kind = "PRIOR_ANTIPROTON"  # priors.py:14-24 (sigma 1e11)


class MainIonFractionCost(_DevicePrior):
    kind = "PRIOR_MAIN_ION_FRACTION"  # priors.py:27-48 (threshold 0.8, sigma 0.1)


class ChargeStateOrderPrior(_DevicePrior):
    kind = "PRIOR_CHARGE_STATE_ORDER"  # priors.py:509-556 (mean 2.0, sigma 0.2)

    def __init__(self, chords, plasma):
        super().__init__()
        # first line of every impurity ion, in chord then line order (priors.py:515-524)
        self.impurity_indicies_dict = {}
        for c, chord in enumerate(chords):
            for i, l in enumerate(chord.lines):
                sp = getattr(l, "species", None)
                if sp in plasma.impurity_species and sp not in self.impurity_indicies_dict:
                    self.impurity_indicies_dict[sp] = (c, i)

    def ordered_lines(self, plasma):
        """Per element: (chord, line) pairs ordered by charge state — prefix match on the element name (Q10)."""
        idx = self.impurity_indicies_dict
        out = []
        for elem in plasma.impurities:
            ions = sorted([ion for ion in idx if ion.startswith(elem)], key=lambda x: idx[x])
            ions = sorted(ions, key=lambda x: float(x.split("_")[1]))
            out.append([idx[ion] for ion in ions])
        return out


class WavelengthCalibrationPrior(_DevicePrior):
    kind = "PRIOR_WAVELENGTH_CAL"  # priors.py:1046-1060 — always reads calwave_0 (Q11)

    def __init__(self, mean, std):
        super().__init__()
        self.mean, self.std = mean, std


class CalibrationPrior(_DevicePrior):
    kind = "PRIOR_INTENSITY_CAL"  # priors.py:1063-1076 (mean 1, std 0.2)


class GaussianInstrumentFunctionPrior(_DevicePrior):
    kind = "PRIOR_GAUSSIAN_IF"  # priors.py:1079-1089 (mean 0.8, std 0.1)


# ---- optional priors (reference priors.py), passed as ``BaysarPosterior(..., priors=[...])``.  Constructor signatures
# follow the reference (first argument: the plasma, or its plasma_state for TeMinPrior); the argument is not used here —
# the model compiler resolves parameter indices from the posterior's own plasma — so ``None`` may be passed when the
# prior is created before the posterior exists.
class ElectronDensityTDVCost(_DevicePrior):
    kind = "PRIOR_TDV"  # priors.py:67-80

    def __init__(self, plasma=None, threshold=1, sigma=0.2):
        super().__init__()
        self.threshold, self.sigma = threshold, sigma


class PeakTePrior(_DevicePrior):
    kind = "PRIOR_PEAK_TE"  # priors.py:241-251

    def __init__(self, plasma=None, te_min=1, te_err=1):
        super().__init__()
        self.te_min, self.te_err = te_min, te_err


class TauPrior(_DevicePrior):
    kind = "PRIOR_TAU"  # priors.py:254-301

    def __init__(self, plasma=None, mean=0, sigma=0.2):
        super().__init__()
        self.mean, self.sigma = mean, sigma

    @staticmethod
    def tau_pairs(plasma):
        """(tag_a, tag_b) of consecutive ions of every impurity element, in ``tau_difference`` order (priors.py:269-284)."""
        pairs = []
        for imp in plasma.impurities:
            tags = [f"{imp}_{ion.split('_')[1]}_tau" for ion in plasma.species if imp + "_" in ion]
            pairs += list(zip(tags[:-1], tags[1:]))
        return pairs


class WallConditionsPrior(_DevicePrior):
    kind = "PRIOR_WALL_CONDITIONS"  # priors.py:432-456

    def __init__(self, plasma=None, te_min=0.5, ne_min=2e12, te_err=None, ne_err=None):
        super().__init__()
        self.te_min, self.ne_min = te_min, ne_min
        self.te_err = 0.2 if te_err is None else te_err
        self.ne_err = 0.2 if ne_err is None else ne_err


class SimpleWallPrior(_DevicePrior):
    kind = "PRIOR_SIMPLE_WALL"  # priors.py:459-481

    def __init__(self, plasma=None, te_min=0.5, ne_min=2e12, te_err=None, ne_err=None):
        super().__init__()
        self.te_min, self.ne_min = te_min, ne_min
        self.te_err = 0.5 * te_min if te_err is None else te_err
        self.ne_err = 0.5 * ne_min if ne_err is None else ne_err


class TeMinPrior(_DevicePrior):
    kind = "PRIOR_TE_MIN"  # priors.py:670-680 (takes the plasma_state in the reference)

    def __init__(self, plasma_state=None, ratio=2.0, sigma=0.05):
        super().__init__()
        self.ratio, self.sigma = ratio, sigma


class TeTauPrior(_DevicePrior):
    kind = "PRIOR_TE_TAU"  # priors.py:696-727

    def __init__(self, plasma=None, sigma=0.3):
        super().__init__()
        self.sigma = sigma


class ImpurityTauPrior(_DevicePrior):
    kind = "PRIOR_IMPURITY_TAU"  # priors.py:395-429

    def __init__(self, plasma=None, sigma=0.3):
        super().__init__()
        self.sigma = sigma

    @staticmethod
    def tau_groups(plasma):
        """Per charge state seen in more than one element: the tau tags of those elements' ions, in species order."""
        groups = {}
        for sp in plasma.impurity_species:
            z0, z = sp.split("_")
            groups.setdefault(z, []).append(z0)
        return [[f"{z0}_{z}_tau" for z0 in elems] for z, elems in groups.items() if len(elems) > 1]


class BowmanTeePrior(_DevicePrior):
    kind = "PRIOR_BOWMAN_TEE"  # priors.py:484-506

    def __init__(self, plasma=None, sigma_err=0.2, nu_err=0.2):
        super().__init__()
        self.sigma_err, self.nu_err = sigma_err, nu_err


# the reference's remaining optional priors, evaluated on the host (baysar_b200/host_priors.py): same namespace as above
from .host_priors import (BolometryPrior, CiiiPrior, EmsTeOrderPrior, FastStarkPrior, FlatnessPrior,  # noqa: E402,F401
                          FlatnessPriorBasic, GaussianProcessPrior, NeutralPowerPrior, NIIITauPrior, NIVTauPrior,
                          PlasmaGradientPrior, SeparatrixTePrior, flatness_prior)
