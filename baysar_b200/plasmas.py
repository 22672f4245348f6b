"""``PlasmaLine`` — the parameter handler and owner of the atomic-data tables (host side).

Mirrors the public surface of the reference's ``PlasmaLine`` (`/root/reference/baysar/plasmas.py:286-874`):
``tags``, ``slices``, ``theta_bounds``, ``n_params``, ``is_resolved``, ``los``, ``species`` lists, option flags,
``plasma_state`` and ``__call__(theta)``.  Differences in kind, not in results:

* the object is *immutable configuration*: evaluating a theta never mutates shared state that other evaluations
  read (the reference's ``plasma_state`` is a blackboard written by line objects during evaluation);
* ``__call__(theta)`` runs the CUDA line-of-sight kernel and then *publishes* the resulting profiles into
  ``plasma_state`` for inspection, so reference-style code that reads ``plasma.plasma_state[...]`` keeps working;
* atomic data comes from an explicit provider (``baysar_b200.atomic_data``), not from files/network.

Table algebra follows the reference: ionisation-balance tables and the concstar correction (plasmas.py:939-1079),
``ne*TEC`` bicubic splines (plasmas.py:1207-1243), log-PEC tables for hydrogen (plasmas.py:1277-1316).
"""
import collections

import numpy as np
from scipy.interpolate import RectBivariateSpline

from .atomic_data import SyntheticADAS, number_of_ions
from .lineshapes import EsymmtricCauchyPlasma

HYDROGEN_ISOTOPES = ["H", "D", "T"]


def power10(var):
    return np.power(10, var)


def kms_to_ms(velocity):
    return 1e3 * velocity


def calibrate(num_pixels, theta):
    """Pixel wavelengths from (centre wavelength, dispersion) — reference plasmas.py:72-80."""
    cwl, dispersion = theta
    pixel_array = np.arange(num_pixels, dtype=float)
    pixel_array -= np.mean(pixel_array)
    pixel_array *= dispersion
    pixel_array += cwl
    return pixel_array


class _Functional:
    """Common shape of the per-chord calibration functionals (plasmas.py:88-158)."""

    def __init__(self, number_of_variables, bounds):
        self.number_of_variables = number_of_variables
        self.bounds = bounds
        if number_of_variables != len(bounds):
            raise ValueError("self.number_of_variables!=len(self.bounds)")


class default_cal_function(_Functional):
    def __init__(self, number_of_variables=1, bounds=None):
        super().__init__(number_of_variables, bounds or [[-5, 20]])

    def __call__(self, *args):
        return np.power(10, args[0])

    def inverse_calibrate(self, y, theta):
        return y / theta[0]


class default_background_function(_Functional):
    def __init__(self, number_of_variables=1, bounds=None):
        super().__init__(number_of_variables, bounds or [[-5, 20]])

    def __call__(self, *args):
        return np.power(10, args[0])

    def calculate_background(self, theta):
        return theta


class default_calwave_function(_Functional):
    def __init__(self, number_of_variables=2, bounds=None):
        super().__init__(number_of_variables, bounds or [[1e2, 1e5], [0.01, 1.0]])

    def __call__(self, *args):
        return args[0]

    def calibrate(self, x, theta):
        return calibrate(len(x), theta)


class GaussianInstrumentFunctionCalibratior(_Functional):
    def __init__(self, bounds=None):
        super().__init__(1, bounds or [[-1, 1]])

    def __call__(self, theta):
        return [np.power(10, t) for t in theta]


class PlasmaLine:
    # post-construction switches the reference reads with hasattr() on every evaluation (plasmas.py:663, :735-737,
    # linemodels.py:398): here they are compiled into the device model, so assigning one makes the posterior recompile
    _RECOMPILE_ATTRS = ("recombining", "ne_tau", "reverse_electroneutrality")

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name in self._RECOMPILE_ATTRS:
            callback = self.__dict__.get("_on_option_change")
            if callback is not None:
                callback()

    def __init__(self, input_dict, profile_function=None, atomic_data=None):
        self.input_dict = d = input_dict
        self.atomic_data = atomic_data if atomic_data is not None else SyntheticADAS()
        self.num_chords = len(d["wavelength_axis"])
        self.species = d["species"]
        self.hydrogen_isotopes = list(HYDROGEN_ISOTOPES)
        self.impurity_species = [s for s in self.species if not ("H_" in s or "D_" in s or "T_" in s)]
        self.hydrogen_species = [s for s in self.species if s not in self.impurity_species]
        self.contains_hydrogen = len(self.hydrogen_species) > 0

        # ADAS evaluation grids (plasmas.py:326-338); magical_tau is two log10 grids glued together (Q17)
        tau_exc = np.logspace(-7, 2, 22)
        tau_rec = np.logspace(1, -4, 8)
        self.adas_plasma_inputs = {
            "te": np.logspace(-1, 2, 48), "ne": np.logspace(12, 15, 15), "tau_exc": tau_exc, "tau_rec": tau_rec,
            "magical_tau": np.concatenate((np.log10(tau_exc), -np.log10(tau_rec) + 4)),
        }
        self.adas_plasma_inputs["big_ne"] = np.array(
            [self.adas_plasma_inputs["ne"] for _ in self.adas_plasma_inputs["te"]]).T

        # option flags and their defaults (plasmas.py:340-385)
        for key, attr, default in (("doppler_shifts", "include_doppler_shifts", True),
                                   ("calibrate_wavelength", "calibrate_wavelength", False),
                                   ("calibrate_intensity", "calibrate_intensity", False),
                                   ("calibrate_instrument_function", "calibrate_instrument_function", False),
                                   ("zeeman", "zeeman", False), ("no_sample_neutrals", "no_sample_neutrals", True),
                                   ("thermalised", "thermalised", True), ("cold_neutrals", "cold_neutrals", True),
                                   ("cold_ions", "cold_ions", True)):
            setattr(self, attr, d[key] if key in d else default)
        if self.zeeman:
            raise NotImplementedError("zeeman=True is outside the B200 hot path (north_star: Zeeman-free profiles)")
        if len(self.hydrogen_species) > 1:
            # The reference cannot evaluate a second hydrogen isotope either, and its constructor already runs one evaluation
            # (posterior.py:197-217): with no_sample_neutrals the isotope has no `<species>_dens` entry and total_power
            # raises KeyError (plasmas.py:1184-1185); with sampled neutrals only the first isotope gets a neutral profile
            # (plasmas.py:678-686), the second stays a one-element array and BalmerHydrogenLine raises ValueError in
            # dot(weights, n0_profile) (linemodels.py:817).  Same exception types here, at construction.
            second = self.hydrogen_species[1]
            if self.no_sample_neutrals:
                raise KeyError(second + "_dens")
            n_los = len(profile_function.electron_density.x) if profile_function is not None else 50
            raise ValueError(f"shapes ({n_los},) and (1,) not aligned: only the first hydrogen isotope has a neutral "
                             "density profile")
        self.concstar = True  # plasmas.py:948-950; the non-concstar TEC path raises in the reference too

        self.impurities = []
        for s in self.impurity_species:
            elem = s[: s.index("_")]
            if elem not in self.impurities:
                self.impurities.append(elem)

        self._build_impurity_tables()
        if self.contains_hydrogen:
            self._build_hydrogen_tables()

        self.profile_function = profile_function if profile_function is not None else EsymmtricCauchyPlasma()
        n = self.num_chords
        self.cal_functions = [default_cal_function() for _ in range(n)]
        self.background_functions = [default_background_function() for _ in range(n)]
        self.calwave_functions = [default_calwave_function() for _ in range(n)]
        self.calintfun_functions = [GaussianInstrumentFunctionCalibratior() for _ in range(n)]
        self.build_tags_slices_and_bounds()
        self.los = self.profile_function.electron_density.x
        self.plasma_state = collections.OrderedDict()
        self._evaluator = None  # set by BaysarPosterior: callable(theta) -> dict of device results

    # ------------------------------------------------------------------------------------------ atomic tables
    @staticmethod
    def species_to_elem_and_ion(species):
        elem, ion = species.split("_")
        return elem, int(ion)

    def _build_impurity_tables(self):
        """Fractional abundances over (tau, ne, Te) per element, then one bicubic spline of log(ne*TEC) per tau for
        every transition: the (tx, ty, c) triplets the CUDA kernel evaluates (FITPACK `bispeu` restated)."""
        te, ne = self.adas_plasma_inputs["te"], self.adas_plasma_inputs["ne"]
        halves = ((self.adas_plasma_inputs["tau_exc"], 1e-8, 0), (self.adas_plasma_inputs["tau_rec"], 1e-4, -1))
        self.impurity_ion_bal = {}
        for elem in self.impurities:
            nion = number_of_ions(elem)
            slabs = []
            for taus, min_frac, meta_index in halves:
                meta = np.zeros(nion)
                meta[meta_index] = 1
                for t in taus:
                    frac = np.array(self.atomic_data.ionisation_balance(elem, te, ne, t, meta), dtype=float)
                    frac[frac < min_frac] = 1e-15
                    slabs.append(frac / np.arange(nion).clip(1)[None, None, :])  # concstar correction
            self.impurity_ion_bal[elem] = np.array(slabs)

        self.impurity_tecs = {}
        self.tec_knots = None
        big_ne = self.adas_plasma_inputs["big_ne"]
        for species in self.impurity_species:
            elem, ion = self.species_to_elem_and_ion(species)
            bal = self.impurity_ion_bal[elem]
            for line, rec in self.input_dict[species].items():
                label = str(line).replace(", ", "_")
                for ch in "[]()":
                    label = label.replace(ch, "")
                exc = self.atomic_data.adf15_pecs(rec["pec"], rec["exc_block"], te, ne)
                rcm = self.atomic_data.adf15_pecs(rec["pec"], rec["rec_block"], te, ne)
                coeffs = []
                for t in range(bal.shape[0]):
                    rates = big_ne * (exc.T * bal[t, :, :, ion] + rcm.T * bal[t, :, :, ion + 1])
                    spline = RectBivariateSpline(ne, te, np.log(rates.clip(1e-30)))
                    tx, ty = spline.get_knots()
                    if self.tec_knots is None:
                        self.tec_knots = (np.array(tx), np.array(ty))
                    coeffs.append(np.array(spline.get_coeffs()))
                self.impurity_tecs[species + "_" + label] = np.array(coeffs)

    def _build_hydrogen_tables(self):
        species = self.hydrogen_species[0]
        element, charge = species.split("_")
        table = self.atomic_data.hydrogen_adf15(element.lower(), int(charge))
        self.hydrogen_pec_grid = (table.ne, table.te)
        self.hydrogen_pecs = {}
        for sp in self.hydrogen_species:
            for line, rec in self.input_dict[sp].items():
                tag = sp + "_" + str(line).replace(", ", "_")
                self.hydrogen_pecs[tag + "_exc"] = np.log(table.select(rec["exc_block"]))
                self.hydrogen_pecs[tag + "_rec"] = np.log(table.select(rec["rec_block"]))
        self.neutral_adf11s = {k: self.atomic_data.hydrogen_adf11(k) for k in ("scd", "acd", "plt", "prb")}

    # ------------------------------------------------------------------------------------------ parameter layout
    def build_tags_slices_and_bounds(self):
        """Tag order, slice sharing and bounds exactly as the reference lays them out (plasmas.py:457-634).

        A tag that is not "resolved" shares the slice of the first earlier tag with the same FIRST CHARACTER and
        the same suffix (plasmas.py:602-616) — e.g. ``N_1_dens``/``N_2_dens`` share one parameter."""
        d = self.input_dict
        entries = []  # (tag, number of variables, bounds rows)

        per_chord = (("cal", self.cal_functions, self.calibrate_intensity),
                     ("calwave", self.calwave_functions, self.calibrate_wavelength),
                     ("calint_func", self.calintfun_functions, self.calibrate_instrument_function),
                     ("background", self.background_functions, True))
        for name, functionals, enabled in per_chord:
            if enabled:
                for chord, fn in enumerate(functionals):
                    entries.append((f"{name}_{chord}", fn.number_of_variables, [list(b) for b in fn.bounds]))
        for tag, fn in (("electron_density", self.profile_function.electron_density),
                        ("electron_temperature", self.profile_function.electron_temperature)):
            rows = fn.bounds if type(fn.bounds[0]) in (list, tuple) else [fn.bounds] * fn.number_of_variables
            if len(rows) != fn.number_of_variables:
                raise ValueError(type(fn), "len(functional.bounds)!=functional.number_of_variables")
            entries.append((tag, fn.number_of_variables, [list(b) for b in rows]))

        suffix_bounds = collections.OrderedDict([("_dens", [11, 14])])
        if not self.thermalised:
            suffix_bounds["_Ti"] = [-1, 2]
        suffix_bounds["_tau"] = [-6, 4]
        if self.include_doppler_shifts:
            suffix_bounds["_velocity"] = [-10, 10]
        suffixes = list(suffix_bounds)
        for suffix in suffixes:
            for s in self.species:
                is_h = any(s[0:2] == h + "_" for h in self.hydrogen_isotopes)
                if suffix == "_tau" and is_h:
                    entries.append((s + suffix, 1, [[-10, -2]]))
                elif (suffix == "_dens" and is_h and self.no_sample_neutrals) or \
                        (suffix == "_Ti" and ((is_h and self.cold_neutrals) or (not is_h and self.cold_ions))):
                    continue
                else:
                    first = [t for t in suffixes if t in s + suffix][0]
                    entries.append((s + suffix, 1, [list(suffix_bounds[first])]))
        for line in d.get("X_lines", []):
            entries.append(("X_" + str(line)[1:-1].replace(", ", "_"), 1, [[0, 20]]))

        resolved_flag = {"_velocity": d.get("ion_resolved_velocity", False), "_Ti": d["ion_resolved_temperatures"],
                         "_dens": False, "_tau": d["ion_resolved_tau"]}
        self.tags = [e[0] for e in entries]
        self.is_resolved = collections.OrderedDict()
        for tag in self.tags:
            hit = [sfx for sfx in suffixes if tag.endswith(sfx)]
            self.is_resolved[tag] = True if not hit else any(resolved_flag[sfx] and tag.endswith(sfx) for sfx in suffixes)

        slices = collections.OrderedDict()
        bounds = []
        cursor = 0
        for tag, nvar, rows in entries:
            share = None
            if slices and not self.is_resolved[tag]:
                sfx = [t for t in suffixes if tag.endswith(t)][0]
                for other, slc in slices.items():
                    if other.startswith(tag[:1]) and other.endswith(sfx):
                        share = slc
                        break
            if share is not None:
                slices[tag] = share
            else:
                slices[tag] = slice(cursor, cursor + nvar)
                cursor += nvar
                bounds.extend(rows)
        self.slices = slices
        self.n_params = cursor
        self.theta_bounds = np.array(bounds, dtype=float)
        if np.isnan(self.theta_bounds).any():
            raise ValueError("Bounds contain NaNs!")
        if self.n_params != len(self.theta_bounds):
            raise ValueError("self.n_params!=len(self.theta_bounds)")
        self._assign_theta_functions()

    def _assign_theta_functions(self):
        """tag -> transform (plasmas.py:803-874): 10**theta everywhere except profiles, calwave, velocities."""
        fns = collections.OrderedDict()
        for tag in self.tags:
            if tag.startswith("cal_") or tag.startswith("calint_func_") or tag.startswith("background_"):
                fns[tag] = power10
            elif tag == "electron_density":
                fns[tag] = self.profile_function.electron_density
            elif tag == "electron_temperature":
                fns[tag] = self.profile_function.electron_temperature
            elif tag.endswith("_velocity"):
                fns[tag] = kms_to_ms
            elif tag.startswith("calwave_"):
                continue
            else:
                fns[tag] = power10
        self.theta_functions = fns

    # ------------------------------------------------------------------------------------------ evaluation
    def is_theta_within_bounds(self, theta):
        return [((b[0] < r) and (r < b[1])) or any(x == r for x in b) for b, r in zip(self.theta_bounds, theta)]

    def __call__(self, theta):
        self.update_plasma_state(theta)

    # ------------------------------------------------------------------------------------------ side outputs (P12)
    @property
    def radiated_power(self):
        """Host-side radiated-power side outputs (`baysar_b200/side_outputs.py`), built on first use."""
        if self.__dict__.get("_radiated_power") is None:
            from .side_outputs import RadiatedPower

            self._radiated_power = RadiatedPower(self)
        return self._radiated_power

    def extrapolate_impurity_power(self, species="N_2", proton_number=None):
        """Per-charge-state radiated power of an element along the line of sight (reference plasmas.py:1137-1172)."""
        return self.radiated_power.extrapolate_impurity_power(species, proton_number)

    def update_plasma_state(self, theta, out=None):
        """Evaluate the line-of-sight stage on the GPU for one theta (or take its result ``out`` from a batched
        evaluation) and publish it in ``plasma_state``."""
        theta = np.asarray(theta, dtype=float).ravel()
        if len(theta) != self.n_params:
            raise ValueError("len(theta)!=self.n_params")
        if self._evaluator is None:
            raise RuntimeError("PlasmaLine is not attached to a BaysarPosterior (no CUDA model to evaluate with)")
        self.plasma_theta = collections.OrderedDict((p, theta[s]) for p, s in self.slices.items())
        if out is None:
            out = self._evaluator(theta)
        for p, s in self.slices.items():
            v = np.array(theta[s], dtype=float).flatten()
            if p in ("electron_density", "electron_temperature"):
                continue
            self.plasma_state[p] = self.theta_functions[p](v) if p in self.theta_functions else v
        self.plasma_state["electron_density"] = out["ne"]
        self.plasma_state["electron_temperature"] = out["te"]
        self.plasma_state["main_ion_density"] = out["main_ion_density"]
        if self.contains_hydrogen:
            self.plasma_state[self.hydrogen_species[0] + "_dens"] = out["n0"]
        return out
