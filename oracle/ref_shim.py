"""TEST INFRASTRUCTURE — import the UNMODIFIED reference (`/root/reference`) under import shims.

Only usable in the build container (the reference tree does not travel to the GPU box).  Nothing in the
product package imports this module; it is used by `oracle/make_golden.py` to generate the committed
fixtures under `tests/golden/` and by `tests/test_oracle_vs_reference.py` (skipped when the reference is
absent) to pin the numpy restatement (`oracle/numpy_port.py`) against the reference itself.

Why shims are needed (none of them touches reference arithmetic on the posterior path):

1. ``scipy.integrate.trapz`` was removed in SciPy >= 1.14 (`baysar/lineshapes.py:7`) -> alias to ``numpy.trapz``.
2. ``xarray`` is not installed (`OpenADAS/read_adf1*.py:10`, `plasmas.py:1282`).  A minimal ``DataArray``
   stand-in provides ``.data``, ``.sel(block=)``, numpy-ufunc support and ``.interp(ne=("pecs", a),
   Te=("pecs", b), kwargs=...)``.  xarray's joint interpolation of two indexers sharing one dimension
   resolves to ``scipy.interpolate.interpn((ne, Te), data, xi, method="linear", **kwargs)``; the kwargs the
   reference passes (``bounds_error=False, fill_value=None`` at `plasmas.py:723-727`, `linemodels.py:782-786`)
   are ``interpn`` keywords, i.e. bilinear with linear extrapolation.  (Residual risk: xarray itself could not
   be executed here.)
3. ``inference`` (inference-tools), ``matplotlib`` -> ``MagicMock`` modules (`priors.py:851`, `output_tools.py:1`).
4. ``backend`` opens ``<repo>/config.json`` at import (`backend/logger.py:16-19`) -> stub package.
5. Atomic data: ``OpenADAS.{get,load}_adf1{1,5}`` (imported lazily at `plasmas.py:708,1284`) and the undefined /
   old-signature names ``baysar.plasmas.{run_adas406, read_adf11, read_adf15, get_adf11}`` (`plasmas.py:983,
   999,1215,962`) are bound to a `baysar_b200.atomic_data.AtomicDataProvider` instance.
"""
import contextlib
import copy
import io
import os
import sys
import types
import warnings
from unittest.mock import MagicMock

import numpy as np

REFERENCE_ROOT = os.environ.get("BAYSAR_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "baysar"))


class DataArray:
    """The few xarray.DataArray behaviours the reference's posterior path relies on."""

    def __init__(self, data, dims=None, coords=None, name=None, attrs=None):
        self.data = np.asarray(data, dtype=float)
        self.dims = tuple(dims) if dims is not None else None
        self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
        self.name = name
        self.attrs = attrs or {}

    def __getattr__(self, name):
        # coordinate access as an attribute (`block_array.ne.data`, OpenADAS/read_adf15.py:96-97)
        coords = self.__dict__.get("coords", {})
        if name in coords:
            return DataArray(coords[name])
        raise AttributeError(name)

    @property
    def shape(self):
        return self.data.shape

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def __rmul__(self, other):  # `tau * scd` (gcr/ionisation_balance.py:37)
        return DataArray(other * self.data, dims=self.dims, coords=self.coords, name=self.name, attrs=self.attrs)

    __mul__ = __rmul__

    def sel(self, block):
        i = int(np.where(self.coords["block"] == block)[0][0])
        rest = {k: v for k, v in self.coords.items() if k != "block"}
        return DataArray(self.data[i], dims=self.dims[1:], coords=rest)

    def interp(self, ne=None, Te=None, kwargs=None):
        from scipy.interpolate import interpn

        ne = ne[1] if isinstance(ne, tuple) else ne
        Te = Te[1] if isinstance(Te, tuple) else Te
        xi = np.stack([np.ravel(ne), np.ravel(Te)], axis=-1)
        out = interpn((self.coords["ne"], self.coords["Te"]), self.data, xi, method="linear", **(kwargs or {}))
        return DataArray(out, dims=("pecs",))

    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        args = [a.data if isinstance(a, DataArray) else a for a in inputs]
        return DataArray(getattr(ufunc, method)(*args, **kw), dims=self.dims, coords=self.coords)


_installed = {}


def install(provider):
    """Install the shims, bind atomic data to ``provider`` and return the reference's modules."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    warnings.simplefilter("ignore")
    if "modules" not in _installed:
        import scipy.integrate

        if not hasattr(scipy.integrate, "trapz"):
            scipy.integrate.trapz = np.trapz
        if not hasattr(np, "trapz"):  # numpy >= 2.4 may drop the alias
            np.trapz = np.trapezoid
        if "xarray" not in sys.modules:
            xr = types.ModuleType("xarray")
            xr.DataArray = DataArray
            xr.concat = None
            xr.load_dataarray = None
            sys.modules["xarray"] = xr
        for name in ["inference", "inference.mcmc", "matplotlib", "matplotlib.pyplot", "matplotlib.pylab",
                     "matplotlib.colors", "matplotlib.gridspec"]:
            sys.modules.setdefault(name, MagicMock())
        if "backend" not in sys.modules:
            b = types.ModuleType("backend")
            bt = types.ModuleType("backend.time")
            bt.TimeIt = lambda f: f
            b.TimeIt = bt.TimeIt
            b.get_logger = lambda *a, **k: MagicMock()
            sys.modules["backend"] = b
            sys.modules["backend.time"] = bt
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        with contextlib.redirect_stdout(io.StringIO()):
            import OpenADAS
            import baysar.input_functions as ref_inputs
            import baysar.line_data as ref_line_data
            import baysar.lineshapes as ref_lineshapes
            import baysar.linemodels as ref_linemodels
            import baysar.plasmas as ref_plasmas
            import baysar.posterior as ref_posterior
            import baysar.priors as ref_priors
            import baysar.spectrometers as ref_spectrometers
            import baysar.tools as ref_tools
        _installed["modules"] = dict(OpenADAS=OpenADAS, input_functions=ref_inputs, line_data=ref_line_data,
                                     lineshapes=ref_lineshapes, linemodels=ref_linemodels, plasmas=ref_plasmas,
                                     posterior=ref_posterior, priors=ref_priors, spectrometers=ref_spectrometers,
                                     tools=ref_tools)
        _installed["pristine_line_data"] = copy.deepcopy(ref_line_data.adas_line_data)
    mods = _installed["modules"]
    bind_provider(provider)
    return mods


def bind_provider(provider):
    mods = _installed["modules"]
    OpenADAS, P = mods["OpenADAS"], mods["plasmas"]

    def _as_dataarray(tab):
        return DataArray(tab.data, dims=("block", "ne", "Te"), coords=dict(block=tab.block, ne=tab.ne, Te=tab.te),
                         name="pecs")

    OpenADAS.get_adf15 = lambda element, charge, year=96, **k: ("adf15", element, charge)
    OpenADAS.load_adf15 = lambda adf15, passed=False: _as_dataarray(provider.hydrogen_adf15(adf15[1], adf15[2]))
    OpenADAS.get_adf11 = lambda element, adf11type, year=96, **k: adf11type
    OpenADAS.load_adf11 = lambda adf11, passed=False: _as_dataarray(provider.hydrogen_adf11(adf11))

    def run_adas406(files=None, elem=None, te=None, dens=None, tint=None, meta=None, all=True, **k):
        return {"ion": provider.ionisation_balance(elem, te, dens, tint, meta)}, None

    def read_adf11(file=None, adf11type=None, is1=1, te=None, dens=None, all=True, **k):
        elem = str(file).split("_")[-1]
        return provider.adf11_rates(elem, adf11type, is1, te, dens)

    def read_adf15(file, block, te, dens, all=True, **k):
        return provider.adf15_pecs(file, block, te, dens), {"wavelength": 0.0}

    P.run_adas406 = run_adas406
    P.read_adf11 = read_adf11
    P.read_adf15 = read_adf15
    P.get_adf11 = lambda elem, yr, type, adf11_dir=None: f"{type}{yr}_{elem}"


def pristine_line_data():
    """A deep copy of the reference catalogue as it was at import (make_input_dict mutates the original)."""
    return copy.deepcopy(_installed["pristine_line_data"])


@contextlib.contextmanager
def quiet():
    """The reference prints hundreds of lines at construction (`plasmas.py:981`, `tools.py:63`)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
