"""Set-up helpers used once per chord (host side): `clip_data`, `centre_peak`, `within`.

Same behaviour as the reference's `baysar/tools.py:26-47,84-127`; none of them runs per theta.
"""
import numpy as np
from scipy.interpolate import interp1d
from scipy.ndimage import center_of_mass


def clip_data(x_data, y_data, x_range):
    """Samples strictly inside ``x_range`` (reference tools.py:26-47)."""
    if len(x_data) != len(y_data):
        raise AssertionError("len(x_data)!=len(y_data)")
    keep = np.where((min(x_range) < x_data) & (x_data < max(x_range)))
    return x_data[keep], y_data[keep]


def centre_peak(peak, places=12, centre=None):
    """Shift ``peak`` by linear interpolation until its centre of mass sits at the middle sample.

    The reference iterates single shifts (tools.py:84-100); each pass slightly smooths the kernel, so the
    iteration itself is part of the instrument function's definition."""
    peak = np.asarray(peak, dtype=float)
    grid = np.arange(len(peak))
    target = np.mean(grid) if centre is None else centre
    while np.round(abs(target - center_of_mass(peak))[0], places) != 0:
        offset = center_of_mass(peak) - target
        peak = interp1d(grid, peak, bounds_error=False, fill_value=0.0)(grid + offset)
    return peak


def within(stuff, boxes):
    """True if the value (any member of a list/tuple) lies strictly inside any box (tools.py:103-127)."""
    if type(stuff) in (tuple, list):
        return any(any((box.min() < float(s)) and (float(s) < box.max()) for s in stuff) for box in boxes)
    if type(stuff) is str:
        stuff = float(stuff)
    elif type(stuff) not in (int, float):
        raise TypeError("Input stuff is not of an appropriate type.")
    return any((box.min() < stuff) and (stuff < box.max()) for box in boxes)
