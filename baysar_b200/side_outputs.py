"""Radiated-power side outputs of the reference's plasma state, evaluated on the HOST on demand.

The reference recomputes them inside every `update_plasma_state` (`plasmas.py:689-694, 1174-1205`) although no default
posterior component reads them; the device path skips them.  `BolometryPrior` (`priors.py:595-643`) does read them, so
this module restates what it needs from the state the posterior publishes after an evaluation (ne, Te, main-ion and
neutral density profiles from the GPU): the hydrogen excitation / recombination power profiles and the per-charge-state
impurity power of `extrapolate_impurity_power`.  numpy / scipy, a few hundred FLOP per line-of-sight point: it is a host
prior's arithmetic, like any user callable; the tables are built on first use.
"""
import numpy as np
from scipy.interpolate import RectBivariateSpline, interpn

from .atomic_data import number_of_ions


class RadiatedPower:
    def __init__(self, plasma):
        self.plasma = plasma
        self._splines = None

    # ---- tables (plasmas.py:997-1020, :1052-1075, :1083-1096)
    def splines(self):
        """{"<elem>_<ion>": {tau: RectBivariateSpline(ne, Te, log power coefficient)}}.  Faithful to the reference's
        loops: the downstream-transport half (tau_rec) re-writes the entries of the first half instead of its own
        (`power[t_counter]` at :1075), so its tau slices keep the initial zeros, i.e. log(1e-30) after the clip."""
        if self._splines is None:
            pl = self.plasma
            te, ne = pl.adas_plasma_inputs["te"], pl.adas_plasma_inputs["ne"]
            big_ne = pl.adas_plasma_inputs["big_ne"]
            taus = pl.adas_plasma_inputs["magical_tau"]
            n_exc = len(pl.adas_plasma_inputs["tau_exc"])
            self._splines = {}
            for elem in pl.impurities:
                bal = pl.impurity_ion_bal[elem]
                n_pow = number_of_ions(elem) - 1
                power = np.zeros((len(taus), len(ne), len(te), n_pow))
                for ion in range(n_pow):
                    plt = pl.atomic_data.adf11_rates(elem, "plt", ion + 1, te, ne)
                    prb = pl.atomic_data.adf11_rates(elem, "prb", ion + 1, te, ne)
                    for t in range(n_exc):
                        power[t, :, :, ion] = (bal[t, :, :, ion] * plt.T + bal[t, :, :, ion + 1] * prb.T) * big_ne
                for ion in range(n_pow):
                    self._splines[f"{elem}_{ion}"] = {tau: RectBivariateSpline(ne, te, np.log(power[t, :, :, ion].clip(1e-30)))
                                                      for t, tau in enumerate(taus)}
        return self._splines

    # ---- plasmas.py:1182-1192
    def neutral_power(self):
        """{"<species>_exc": ne n0 plt, "<first species>_rec": ne ni prb} profiles over the line of sight."""
        pl, st = self.plasma, self.plasma.plasma_state
        ne, te, ni = st["electron_density"], st["electron_temperature"], st["main_ion_density"]
        pts = np.stack([ne, te], axis=-1)
        rate = {}
        for kind in ("plt", "prb"):
            t = pl.neutral_adf11s[kind]
            rate[kind] = interpn((t.ne, t.te), t.select(1), pts, method="linear", bounds_error=False, fill_value=None)
        out = {}
        for species in pl.hydrogen_species:
            out[species + "_exc"] = ne * st[species + "_dens"] * rate["plt"]
        out[pl.hydrogen_species[0] + "_rec"] = ne * ni * rate["prb"]
        return out

    # ---- plasmas.py:1137-1172
    def extrapolate_impurity_power(self, species="N_2", proton_number=None):
        pl, st = self.plasma, self.plasma.plasma_state
        elem = species.split("_")[0]
        if proton_number is None:
            proton_number = number_of_ions(elem) - 1
        dens, tau = st[species + "_dens"], np.log10(st[species + "_tau"])
        ne, te = st["electron_density"], st["electron_temperature"]
        grid = pl.adas_plasma_inputs["magical_tau"]
        upper = grid.searchsorted(tau)
        upper_tau, lower_tau = grid[upper][0], grid[upper - 1][0]
        upper_weight = (tau - lower_tau) / (upper_tau - lower_tau)
        lower_weight = 1 - upper_weight
        power = []
        for charge in range(proton_number):
            sp = self.splines()[f"{elem}_{charge}"]
            rates = upper_weight * np.exp(sp[upper_tau].ev(ne, te)) + lower_weight * np.exp(sp[lower_tau].ev(ne, te))
            power.append(dens * rates)
        return np.array(power)
