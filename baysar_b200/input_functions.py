"""``make_input_dict`` — the drop-in input boundary (host side only; no GPU work).

Same signature, validation rules, exception types and resulting dictionary as the reference
(`/root/reference/baysar/input_functions.py:52-339`), with one deliberate difference: the catalogue passed as
``line_data`` is deep-copied before out-of-range transitions are removed, so the module-level catalogue is not
shrunk as a side effect of building an input dictionary (the reference pops from the shared dict in place,
input_functions.py:306-311).
"""
import copy

import numpy as np

from .line_data import adas_line_data


def within(stuff, boxes):
    """True if the value (or any member of a list/tuple) lies strictly inside any box
    (reference `input_functions.py:17-49`)."""
    boxes = boxes if type(boxes) is np.ndarray else np.array(boxes)
    if len(boxes.shape) != 2:
        raise ValueError("Incorrect boxes dimensions! Should be structure array([box0, box1, ...])")
    if type(stuff) in (tuple, list):
        return any(any((box.min() < float(s)) and (float(s) < box.max()) for s in stuff) for box in boxes)
    if type(stuff) is str:
        stuff = float(stuff)
    elif type(stuff) not in (int, float):
        raise TypeError("Input stuff is not of an appropriate type.")
    return any((box.min() < stuff) and (stuff < box.max()) for box in boxes)


def check_input_dict_input(wavelength_axis, experimental_emission, instrument_function, emission_constant,
                           noise_region, species, ions, line_data=adas_line_data, mystery_lines=None, refine=0.01,
                           ion_resolved_temperatures=False, ion_resolved_tau=False):
    """Input validation, rule for rule as the reference (`input_functions.py:52-213`)."""
    per_chord = [wavelength_axis, experimental_emission, instrument_function, noise_region, emission_constant]
    if not all(len(e) == len(per_chord[0]) for e in per_chord):
        raise ValueError("[wavelength_axis, experimental_emission, instrument_function, noise_region, "
                         "emission_constant] are not all the same length")
    if not all(type(e) in (list, np.ndarray) for e in per_chord):
        raise TypeError("[wavelength_axis, experimental_emission, instrument_function, noise_region, "
                        "emission_constant] are not all lists or arrays")
    if not all(type(e) in (list, np.ndarray) for group in per_chord[:-1] for e in group):
        raise TypeError("Contents of [wavelength_axis, experimental_emission, instrument_function, noise_region] "
                        "are not all lists or arrays")
    if not all(np.isreal(group).all() for group in per_chord[:-1]):
        raise TypeError("Data given in [wavelength_axis, experimental_emission, instrument_function, noise_region] "
                        "are not all ints or floats")
    if not all(np.isreal(e) for e in emission_constant):
        raise TypeError("Emission constants type must be in (int, float, np.float64, np.int64)")
    for region, axis in zip(noise_region, wavelength_axis):
        if not all(within(r, [axis]) for r in region):
            raise ValueError("Noise regions are not in the wavelength axes!")
        if np.diff(region) < np.diff(axis).min():
            raise ValueError("Noise regions is subpixel!")
    if len(species) != len(ions):
        raise ValueError("len(species)!=len(ions)")
    if not (type(species) == list and all(type(s) == str for s in species)):
        raise TypeError("species must be a list strings")
    if not (type(ions) == list and all(type(i) == list for i in ions)
            and all(type(i) == str for s in ions for i in s)):
        raise TypeError("All ions must be a list of list of strings")
    if type(line_data) != dict:
        raise TypeError("type(line_data)!=dict. Input line_data must be a dict")
    missing = [s + "_" + i for s, group in zip(species, ions) for i in group
               if not (s in line_data and i in line_data[s])]
    if missing:
        raise KeyError(missing, "not in given line_data")
    if mystery_lines is not None:
        if type(mystery_lines) != list:
            raise TypeError("type(mystery_lines)!=list")
        if len(mystery_lines[0]) != len(mystery_lines[1]):
            raise ValueError("len(mystery_lines[0])!=len(mystery_lines[1]) each mystery line needs a branching "
                             "ratio even is it is a singlet ([1])")
        if not (len(mystery_lines) == 2 and all(type(ml) == list for ml in mystery_lines)
                and all(type(l) == list for ml in mystery_lines for l in ml)):
            raise TypeError("mystery_lines must be a list of two lists than only contain lists")
        if not all(type(m) in (int, float) for ml in mystery_lines for l in ml for m in l):
            raise TypeError("Some of the give mystery line central wavelengths and/or branching ratios are not "
                            "ints or floats")
    if type(refine) in (float, int):
        refine = [refine for _ in wavelength_axis]
    elif type(refine) in (list, tuple, np.ndarray):
        if len(refine) != len(wavelength_axis):
            raise ValueError("len(refine)!=len(wavelength_axis)")
    else:
        raise TypeError("type(refine) must be in (float, int, list, tuple, np.ndarray).")
    if not all(type(v) == bool for v in [ion_resolved_temperatures, ion_resolved_tau]):
        raise TypeError("Not all ion resolutions are booleans")


def make_input_dict(wavelength_axis, experimental_emission, instrument_function, emission_constant, noise_region,
                    species, line_data=adas_line_data, mystery_lines=None, refine=0.01,
                    ion_resolved_temperatures=False, ion_resolved_tau=False, **kwargs):
    """Assemble the dictionary ``BaysarPosterior`` is built from (reference `input_functions.py:216-339`).

    ``ions`` is a required keyword (the reference reads it from ``kwargs``); every extra keyword (the option
    flags ``doppler_shifts``, ``calibrate_*``, ``no_sample_neutrals``, ``cold_neutrals`` ...) is copied through.
    """
    ions = kwargs["ions"]
    as_arrays = lambda x: [d if type(d) is np.ndarray else np.array(d) for d in x]  # noqa: E731
    wavelength_axis = as_arrays(wavelength_axis)
    experimental_emission = as_arrays(experimental_emission)
    instrument_function = as_arrays(instrument_function)
    refine = refine if type(refine) is list else [refine]

    check_input_dict_input(wavelength_axis, experimental_emission, instrument_function, emission_constant,
                           noise_region, species, ions, line_data, mystery_lines, refine,
                           ion_resolved_temperatures, ion_resolved_tau)

    line_data = copy.deepcopy(line_data)
    input_dict = {**kwargs}
    input_dict["species"] = []
    for s, group in zip(species, ions):
        if s not in line_data:
            print(s + " is not in adas_line_data")
            continue
        for ion in group:
            sion = s + "_" + ion
            if ion not in line_data[s]:
                print(sion + " is not in adas_line_data")
                continue
            input_dict[sion] = line_data[s][ion]
            input_dict[sion].pop("default_pecs", None)
            input_dict["species"].append(sion)
            for line in list(input_dict[sion].keys()):
                if not within(line, wavelength_axis):
                    input_dict[sion].pop(line)

    for key, value in (("wavelength_axis", wavelength_axis), ("experimental_emission", experimental_emission),
                       ("instrument_function", instrument_function), ("emission_constant", emission_constant),
                       ("noise_region", noise_region), ("refine", refine)):
        input_dict[key] = value
    input_dict["ion_resolved_temperatures"] = ion_resolved_temperatures
    input_dict["ion_resolved_tau"] = ion_resolved_tau
    if mystery_lines is not None:
        input_dict["X_lines"] = mystery_lines[0]
        input_dict["X_fractions"] = mystery_lines[1]
    return input_dict
